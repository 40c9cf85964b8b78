#!/bin/bash
# GPU run A (1 GPU): new parity tests, whole gpu suite, e2e NUMA diagnosis, bench (both arms), launch list
mkdir -p gpurun_out/ra
python -m pytest tests/test_gpu_fp16_pinned.py -m gpu -q -s -x --no-header -p no:cacheprovider > gpurun_out/ra/pinned.log 2>&1; echo "pinned rc=$?" >> gpurun_out/ra/pinned.log
python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --deselect tests/test_gpu_fp16_pinned.py > gpurun_out/ra/suite.log 2>&1; echo "suite rc=$?" >> gpurun_out/ra/suite.log
python tools/exp_e2e_numa.py > gpurun_out/ra/numa.log 2>&1
OMP_NUM_THREADS=1 python tools/exp_e2e_numa.py > gpurun_out/ra/numa_omp1.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/ra/bench.json 2> gpurun_out/ra/bench.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/ra/bench_ref.json 2> gpurun_out/ra/bench_ref.err
tail -3 gpurun_out/ra/pinned.log gpurun_out/ra/suite.log; cat gpurun_out/ra/numa.log; cat gpurun_out/ra/bench.json | cut -c1-1500
