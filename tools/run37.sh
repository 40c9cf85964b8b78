./tools/bin/exp_mma_issue 2>&1 | tail -9
