#!/bin/bash
# GPU run C (1 GPU): restored kernel + layer chaining (env-gated): correctness with chaining on, bench off / on
mkdir -p gpurun_out/rc; O=gpurun_out/rc
YB_PRINT_LAYER_TABLE=1 python -m pytest tests/test_gpu_fp16_pinned.py -m gpu -q -s --no-header -p no:cacheprovider > $O/pinned.log 2>&1; echo "rc=$?" >> $O/pinned.log
YB_CHAIN=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fp16_pinned.py -m gpu -q --no-header -p no:cacheprovider -k "fp16 or batch or predict or v11 or seg or wide" > $O/chain_tests.log 2>&1; echo "rc=$?" >> $O/chain_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_off.json 2> $O/bench_off.err
YB_CHAIN=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-weights > $O/bench_chain.json 2> $O/bench_chain.err
YB_CHAIN=1 YB_CHAIN_GRID1=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-weights > $O/bench_chain_g2.json 2> $O/bench_chain_g2.err
YB_CHAIN=1 timeout 300 python bench.py --model v8s --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_chain_v8s.json 2> $O/bench_chain_v8s.err
python bench.py --model v8s --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_off_v8s.json 2> $O/bench_off_v8s.err
YB_CHAIN=1 timeout 300 python tools/profile_ops.py v8n 32 > $O/ops_chain_v8n.txt 2>&1
tail -3 $O/pinned.log $O/chain_tests.log; for f in $O/bench_*.json; do echo $f $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['roofline']['forward_ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" 2>&1 | tail -1); done
