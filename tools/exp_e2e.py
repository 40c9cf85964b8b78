"""e2e (yb_predict_u8_submit/_wait, 2 slots) under one configuration; prints one line.  Used to find out why the two
slots serialised at N=1 (15 k img/s) while the same code reached 24.6 k per GPU under torchrun (round-1 SCALE)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import yolosharp_b200 as y  # noqa: E402
from tests.util import oracle_model, synth_image  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
if "--nccl" in sys.argv:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
B = 32
m = oracle_model("v8", "detect", "n")
eng = y.Engine("v8", "n", "detect", 80, "f16", 0, B, 640, 640)
eng.load_state_dict(m.state_dict())
eng.finalize()
u8 = [synth_image(B, 640, 640, seed=200 + i, dtype=torch.uint8).pin_memory() for i in range(2)]
dh = [torch.empty((B, 300, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
ch = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(2)]
for i in range(6):
    eng.predict_u8_submit(i & 1, u8[i & 1], dh[i & 1], ch[i & 1], 0.25, 0.45, 300)
    eng.predict_u8_wait(i & 1)
torch.cuda.synchronize()
n = 30
t0 = time.perf_counter()
for i in range(n):
    if i >= 2:
        eng.predict_u8_wait(i & 1)
    eng.predict_u8_submit(i & 1, u8[i & 1], dh[i & 1], ch[i & 1], 0.25, 0.45, 300)
eng.predict_u8_wait(0)
eng.predict_u8_wait(1)
dt = time.perf_counter() - t0
t1 = time.perf_counter()
for i in range(10):  # serial reference: one slot only
    eng.predict_u8_submit(0, u8[0], dh[0], ch[0], 0.25, 0.45, 300)
    eng.predict_u8_wait(0)
ds = time.perf_counter() - t1
print(f"[{tag}] pipelined {B * n / dt:9.1f} img/s ({dt / n * 1e3:.3f} ms/step)   one slot {B * 10 / ds:9.1f} img/s ({ds / 10 * 1e3:.3f} ms/step)"
      f"   omp={torch.get_num_threads()}", flush=True)
