#!/bin/bash
# 1-GPU run: whole -m gpu suite (timed), plan dump + timelines, sanitizer, bench
mkdir -p gpurun_out/rg; O=gpurun_out/rg
( time python -m pytest tests -m gpu -q --no-header -p no:cacheprovider ) > $O/suite.log 2>&1; echo "rc=$?" >> $O/suite.log
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
python bench.py --model v8s-seg --batch 16 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_seg.json 2> $O/bench_seg.err
YB_DEBUG_PLANS=1 python tools/exp_timeline.py 0 1 2 3 4 5 6 11 > $O/timeline.txt 2> $O/plans.txt
python tools/profile_ops.py v8n 32 > $O/ops_v8n.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_racecheck.log
tail -n 6 $O/suite.log; tail -n 3 $O/sanitizer_memcheck.log $O/sanitizer_racecheck.log; python -c "import json; d=json.load(open('$O/bench.json')); print(d['value'], d['roofline']['forward_ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['other_kernels'])"; python -c "import json; d=json.load(open('$O/bench_seg.json')); print(d['value'], d['e2e'])"; tail -n 3 $O/bench_seg.err; head -30 $O/timeline.txt
