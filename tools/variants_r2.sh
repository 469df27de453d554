timeout 100 python tools/state_hash.py 20000 300 8192 40 Adam 2>&1 | tail -1
timeout 400 python -m pytest tests/test_train_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -3
B="python bench.py --steps 200 --warmup 20 --no-topk --no-cpu-baseline --no-e2e"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step']*1e3,1), 'us  stage', round(d['roofline']['kernels_ms']['stage']*1e3,1), 'update', round(d['roofline']['kernels_ms']['update']*1e3,1), 'frac', round(d['roofline']['frac'],3), 'loss', d['final_loss'])"; }
timeout 120 $B < /dev/null | pick atomic4
