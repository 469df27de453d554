set -x
python -m pytest tests/test_train_gpu.py tests/test_pipeline_gpu.py tests/test_contract_configs_gpu.py -x -q -k "not topk" 2>&1 | tail -4
B="python bench.py --steps 200 --warmup 20 --no-topk --no-cpu-baseline --no-e2e"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step']*1e3,1), 'us  stage', round(d['roofline']['kernels_ms']['stage']*1e3,1), 'update', round(d['roofline']['kernels_ms']['update']*1e3,1), 'frac', round(d['roofline']['frac'],3), 'loss', d['final_loss'])"; }
GLOVE_UPDATE_KERNEL=1 $B | pick k1_diet
GLOVE_UPDATE_KERNEL=3 $B | pick k3_diet
