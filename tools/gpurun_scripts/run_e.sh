#!/bin/bash
# 2-GPU run: peer-memory exchange tests, bench N=2 with both gathers, training all-reduce
mkdir -p gpurun_out/re; O=gpurun_out/re
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider -x > $O/multi_tests.log 2>&1; echo "rc=$?" >> $O/multi_tests.log
for g in comm nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --gather $g > $O/bench2_$g.json 2> $O/bench2_$g.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode train --model v11n --steps 2 --warmup 1 > $O/train2.json 2> $O/train2.err
tail -n 5 $O/multi_tests.log; for f in $O/bench2_*.json; do echo $f; tail -n 1 $f | cut -c1-300; python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['timing']['gather'])"; done; tail -n 3 $O/bench2_comm.err; tail -n 1 $O/train2.json | cut -c1-300; tail -n 3 $O/train2.err
