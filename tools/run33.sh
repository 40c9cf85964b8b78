./tools/bin/exp_mma_issue 2>&1 | tail -30
timeout 300 python -m pytest tests -q -m gpu -x -k "nms or NMS" 2>&1 | tail -2
python tools/exp_nms_time.py v8n 32 2>&1 | tail -1
