#!/bin/bash
# 1-GPU run: plan dump, timelines of the high-resolution layers, sanitizer, whole -m gpu suite timing
mkdir -p gpurun_out/rg; O=gpurun_out/rg
YB_DEBUG_PLANS=1 python tools/exp_timeline.py 0 1 2 3 4 5 6 11 > $O/timeline.txt 2> $O/plans.txt
( time python -m pytest tests -m gpu -q --no-header -p no:cacheprovider ) > $O/suite.log 2>&1; echo "rc=$?" >> $O/suite.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_racecheck.log
tail -n 4 $O/suite.log; tail -n 4 $O/sanitizer_memcheck.log $O/sanitizer_racecheck.log; head -40 $O/timeline.txt
