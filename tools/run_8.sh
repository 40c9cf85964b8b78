#!/bin/bash
# 8-GPU lines: training step (native, NCCL all-reduce of the flat gradient buffer) and the default inference bench
mkdir -p gpurun_out/r8; O=gpurun_out/r8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --mode train --steps 10 --warmup 3 > $O/train8.json 2> $O/train8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench8.json 2> $O/bench8.err
grep -h "metric" $O/train8.json | head -c 600; echo; grep -h "metric" $O/bench8.json | head -c 300; echo; tail -3 $O/train8.err
