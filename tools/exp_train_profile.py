"""Where the training step goes: one YOLOv11s (or argv[1]) step at batch argv[2] under torch.profiler, kernels grouped by
name.  python tools/exp_train_profile.py [v11s] [16] [tc|f32]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import MODELS, synth_targets  # noqa: E402
from tests.util import oracle_model, synth_image  # noqa: E402
from yolosharp_b200.train_v11 import KernelOpsV11, TrainStepV11  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "v11s"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
mode = sys.argv[3] if len(sys.argv) > 3 else "tc"   # tc | f32 (Python graph walk) | native (csrc/train_step.cu)
tc = mode != "f32"
arch, size, task, _ = MODELS[model]
dev = torch.device("cuda", 0)
m = oracle_model(arch, task, size)
if mode == "native":
    from yolosharp_b200.train_native import NativeTrainer
    st = NativeTrainer({k: v.detach().clone() for k, v in m.state_dict().items()}, "v11", size, 80, device=dev, max_batch=B)
else:
    st = TrainStepV11({k: v.detach().clone() for k, v in m.state_dict().items()}, size, 80, device=dev, ops=KernelOpsV11(tensor_cores=tc))
x, t = synth_image(B, 640, 640, seed=1).to(dev), synth_targets(B, 2)
for _ in range(2):
    st.step(x, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
st.step(x, t)
e1.record()
torch.cuda.synchronize()
print(f"# {model} batch {B} {mode}: step {e0.elapsed_time(e1):.2f} ms (CUDA events)")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    st.step(x, t)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"# device time {tot:.2f} ms in {sum(r[2] for r in rows)} kernels")
for k, ms, n in rows[:30]:
    print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}% {n:5d}x  {k[:110]}")
# YB_PROF_DUMP=substr[,substr]: every launch of the matching kernels in timeline order (duration, grid, block) - which layers
# a kernel family spends its time on
dump = [d for d in os.environ.get("YB_PROF_DUMP", "").split(",") if d]
if dump:
    evs = [e for e in prof.events() if e.device_type.name == "CUDA" and any(d in e.name for d in dump)]
    evs.sort(key=lambda e: e.time_range.start)
    for e in evs:
        print(f"{e.time_range.start / 1e3:10.3f} {e.device_time / 1:8.1f} us  {e.name[:60]}")
