B="python bench.py --steps 200 --warmup 20 --no-topk --no-cpu-baseline --no-e2e"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step']*1e3,1), 'us  stage', round(d['roofline']['kernels_ms']['stage']*1e3,1), 'update', round(d['roofline']['kernels_ms']['update']*1e3,1), 'frac', round(d['roofline']['frac'],3), 'loss', d['final_loss'])"; }
$B | pick c5_l2pf
GLOVE_UPDATE_CTAS=4 $B | pick c4_l2pf
