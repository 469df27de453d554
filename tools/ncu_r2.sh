# one gpurun call: plain run first (must exit 0), then (1) the launch list, (2) the step kernels under ncu --set full
set -e
CMD="python bench.py --steps 24 --warmup 4 --pretrain-steps 512 --no-graph --no-topk --no-cpu-baseline --no-e2e"
timeout 200 $CMD > gpurun_out/ncu_plain.log 2>&1 < /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"update_kernel|stage_closed" -s 1100 -c 48 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1 < /dev/null || { tail -5 gpurun_out/ncu_launches.log; exit 1; }
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"update_kernel|stage_closed" -s 1100 -c 8 \
      -f -o gpurun_out/r02_step $CMD > gpurun_out/ncu_full.log 2>&1 < /dev/null || { tail -20 gpurun_out/ncu_full.log; exit 1; }
ls -la gpurun_out/r02_step.ncu-rep gpurun_out/r02_launches.csv
