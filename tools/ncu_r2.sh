# one ncu call per gpurun: plain run first (must exit 0), then the step kernels under ncu --set full
# usage: bash tools/ncu_r2.sh <tag> [kernel regex]   (environment selects the variant)
set -e
TAG=${1:-k}
REGEX=${2:-"update_kernel|stage_closed"}
SKIP=${3:-150}
CMD="python bench.py --steps 24 --warmup 4 --pretrain-steps 64 --no-topk --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c 4 \
      -f -o gpurun_out/r2_step_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1 || { tail -20 gpurun_out/ncu_$TAG.log; exit 1; }
ls -la gpurun_out/r2_step_$TAG.ncu-rep
