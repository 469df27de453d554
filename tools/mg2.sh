# usage: tools/mg2.sh N tag [bench args]   -- one bench.py run on N GPUs of this box, JSON line -> gpurun_out/<tag>.json
N=$1; TAG=$2; shift; shift
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/$TAG.json 2> gpurun_out/$TAG.err
echo "$TAG rc=$?"; python tools/benchline.py $TAG gpurun_out/$TAG.json 2>/dev/null || tail -3 gpurun_out/$TAG.err
