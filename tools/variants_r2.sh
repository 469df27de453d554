set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
cp gpurun_out/parity_maxima.json gpurun_out/r02_parity_maxima.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err < /dev/null ) 2>&1 | grep real
python tools/benchline.py default gpurun_out/r02_bench_default.json
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2>&1 < /dev/null ) 2>&1 | grep real
tail -c 600 gpurun_out/r02_bench_reference.json
timeout 300 python bench.py --workload text8 --no-topk > gpurun_out/r02_bench_text8.json 2>/dev/null < /dev/null; python tools/benchline.py text8 gpurun_out/r02_bench_text8.json
