YB_TL_SIZE=x YB_TL_BATCH=8 python tools/exp_timeline.py 11 12 2>&1 | tail -12
python - <<'PY'
import torch, time
x = torch.empty(32*3*640*640, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): d.copy_(x, non_blocking=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"H2D pinned 39.3 MB: {ms:.3f} ms = {x.numel()/ms/1e6:.1f} GB/s")
PY
