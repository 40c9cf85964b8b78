#!/bin/bash
mkdir -p gpurun_out/raj; O=gpurun_out/raj
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider > $O/multi.log 2>&1; echo "rc=$?" >> $O/multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 12 --warmup 3 > $O/train2.json 2> $O/train2.err
timeout 300 python bench.py --mode train --steps 12 --warmup 3 > $O/train1.json 2> $O/train1.err
tail -5 $O/multi.log; grep -h metric $O/train2.json | head -c 1500; echo; grep -h metric $O/train1.json | head -c 1500; echo; tail -2 $O/train2.err
