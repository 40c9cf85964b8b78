#!/bin/bash
# co-residency experiment (half-SM plans, several part-batch engines) + NMS re-check after the kept-list change
mkdir -p gpurun_out/rh; O=gpurun_out/rh
python tools/exp_dual2.py v8n 32 > $O/dual_default.txt 2> $O/dual_default.err
YB_PLAN_SMALL=1 python tools/exp_dual2.py v8n 32 > $O/dual_small.txt 2> $O/dual_small.err
YB_PLAN_SMALL=1 python tools/exp_dual2.py v8n 64 > $O/dual_small64.txt 2> $O/dual_small64.err
python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "nms or smoke or predict" > $O/nms_tests.log 2>&1; echo "rc=$?" >> $O/nms_tests.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/racecheck.log 2>&1; echo "rc=$?" >> $O/racecheck.log
cat $O/dual_default.txt $O/dual_small.txt $O/dual_small64.txt; tail -3 $O/dual_small.err; tail -4 $O/nms_tests.log; grep -c "nms.cu" $O/racecheck.log; tail -2 $O/racecheck.log
