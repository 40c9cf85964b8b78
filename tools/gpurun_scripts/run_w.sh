#!/bin/bash
mkdir -p gpurun_out/rw; O=gpurun_out/rw
YB_STEP_TIMES=1 python bench.py --mode train --steps 8 --warmup 3 > $O/a.json 2> $O/a.err
YB_NO_SAMPLER=1 YB_STEP_TIMES=1 python bench.py --mode train --steps 8 --warmup 3 > $O/b.json 2> $O/b.err
grep "step " $O/a.err | tr '\n' ' '; echo; grep "step " $O/b.err | tr '\n' ' '; echo; head -c 200 $O/a.json; echo; head -c 200 $O/b.json; echo
