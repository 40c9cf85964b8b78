#!/bin/bash
mkdir -p gpurun_out/rai; O=gpurun_out/rai
timeout 600 python -m pytest tests/test_trainer_native.py tests/test_gpu_conv_tc.py -m gpu -q --no-header -p no:cacheprovider > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
( time timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/memcheck.log 2>&1; echo "rc=$?" >> $O/memcheck.log
tail -4 $O/tests.log; tail -6 $O/smoke.log; grep -E "ERROR SUMMARY|rc=|real" $O/memcheck.log
