python -m pytest tests/test_train_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -3
B="python bench.py --steps 200 --warmup 20 --no-topk --no-cpu-baseline --no-e2e"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step']*1e3,1), 'us  stage', round(d['roofline']['kernels_ms']['stage']*1e3,1), 'update', round(d['roofline']['kernels_ms']['update']*1e3,1), 'frac', round(d['roofline']['frac'],3), 'loss', d['final_loss'])"; }
$B | pick keep_graph
$B --no-graph | pick keep_nograph
python bench.py --workload cc --steps 96 --warmup 16 --no-topk --no-cpu-baseline --no-e2e | pick keep_cc
python bench.py --workload text8 --steps 192 --warmup 16 --no-topk --no-cpu-baseline --no-e2e | pick text8_graph
