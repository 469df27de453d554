timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for G in "" "--no-graph"; do timeout 200 python bench.py --workload text8 --batch 1024 --plan-steps 64 --steps 4096 --warmup 128 --pretrain-steps 1024 --no-topk --no-cpu-baseline --no-e2e $G < /dev/null > gpurun_out/r2_b1024$G.json 2>gpurun_out/r2_b1024$G.err; python tools/benchline.py "b1024$G" gpurun_out/r2_b1024$G.json; done
bash tools/ncu_r2.sh
