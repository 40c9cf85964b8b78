mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "nms or NMS" 2>&1 | tail -5
timeout 300 python tools/dbg_stem.py 2>&1 | tail -16
