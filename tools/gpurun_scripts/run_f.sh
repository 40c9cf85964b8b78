#!/bin/bash
# 8-GPU run: scaling of the sharded inference (both gathers), configs[2] (v8x batch 64 = 8 per GPU), configs[3] (v11s train, 16 per GPU)
mkdir -p gpurun_out/rf; O=gpurun_out/rf
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --gather comm > $O/bench8_comm.json 2> $O/bench8_comm.err
timeout 300 $TR --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --gather nccl > $O/bench8_nccl.json 2> $O/bench8_nccl.err
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --model v8x --batch 8 --steps 20 --warmup 5 --gather comm > $O/bench8_v8x.json 2> $O/bench8_v8x.err
timeout 600 $TR --master-port 29524 bench.py --gpus 8 --mode train --steps 3 --warmup 2 > $O/train8.json 2> $O/train8.err
for f in $O/bench8_*.json; do echo $f; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['timing']['gather'])"; done; tail -n 1 $O/train8.json | cut -c1-250; tail -n 2 $O/bench8_comm.err $O/train8.err
