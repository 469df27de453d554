# one ncu call per gpurun: plain run first (must exit 0), then the step kernels under ncu --set full
set -e
CMD="python bench.py --steps 24 --warmup 4 --pretrain-steps 64 --no-topk --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"update_kernel|stage_closed" -s 150 -c 4 \
      -f -o gpurun_out/r2_step_k3 $CMD > gpurun_out/ncu_k3.log 2>&1 || { tail -20 gpurun_out/ncu_k3.log; exit 1; }
ls -la gpurun_out/*.ncu-rep
