timeout 100 python tools/state_hash.py 20000 300 8192 40 Adam 2>&1 | tail -1
timeout 100 python tools/state_hash.py 3000 100 1024 20 Adagrad 2>&1 | tail -1
timeout 400 python -m pytest tests/test_train_gpu.py tests/test_pipeline_gpu.py tests/test_contract_configs_gpu.py -x -q -k "not topk" 2>&1 | tail -3
B="python bench.py --steps 200 --warmup 20 --no-topk --no-cpu-baseline --no-e2e"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step']*1e3,1), 'us  stage', round(d['roofline']['kernels_ms']['stage']*1e3,1), 'update', round(d['roofline']['kernels_ms']['update']*1e3,1), 'frac', round(d['roofline']['frac'],3), 'loss', d['final_loss'])"; }
timeout 120 $B < /dev/null | pick pipe2
GLOVE_UPDATE_CTAS=3 timeout 120 $B < /dev/null | pick pipe2_c3
timeout 120 python bench.py --workload text8 --steps 192 --warmup 16 --no-topk --no-cpu-baseline --no-e2e < /dev/null | pick text8
