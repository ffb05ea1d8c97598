#!/bin/bash
# One multi-GPU bench line (run under `gpurun --gpus N`): usage: bash tools/scale_measure.sh <N> <tag>
n=$1; tag=${2:-scale}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
echo "rc=$?"; tail -3 gpurun_out/${tag}_n$n.err; python tools/bench_summary.py gpurun_out/${tag}_n$n.json
