YB_TL_SIZE=x YB_TL_BATCH=8 python tools/exp_timeline.py 2 3 2>&1 | tail -36
