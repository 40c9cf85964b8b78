"""Experiment: do half-SM plans (YB_PLAN_SMALL=1: every conv_tc plan <= 104 KiB / 256 TMEM columns, one CTA per SM)
let the kernels of several concurrent part-batch engines share every SM, each filling the other's pipeline bubbles?
python tools/exp_dual2.py [v8n] [total_batch]   (run once with YB_PLAN_SMALL unset, once with YB_PLAN_SMALL=1)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import yolosharp_b200 as y  # noqa: E402
from bench import MODELS  # noqa: E402
from tests.util import oracle_model, synth_image  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "v8n"
TOTAL = int(sys.argv[2]) if len(sys.argv) > 2 else 32
arch, size, task, _ = MODELS[model]
m = oracle_model(arch, task, size)
sd = m.state_dict()
print(f"# {model} total batch {TOTAL} YB_PLAN_SMALL={os.environ.get('YB_PLAN_SMALL', '')}", flush=True)


def make(B):
    e = y.Engine(arch, size, task, 80, "f16", 0, B, 640, 640)
    e.load_state_dict(sd)
    e.finalize()
    return e


def timeit(fn, iters=40, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def split(total, parts):
    q, r = divmod(total, parts)
    return [q + (1 if i < r else 0) for i in range(parts)]


e = make(TOTAL)
x = synth_image(TOTAL, 640, 640, dtype=torch.float16).cuda()
out = torch.empty((TOTAL, e.pred_channels, e.anchors), dtype=torch.float32, device="cuda")
t = timeit(lambda: e.forward(x, out_pred=out))
print(f"1 x B={TOTAL}: forward {t:.4f} ms -> {TOTAL / t * 1e3:8.0f} img/s", flush=True)
del e

for parts in (2, 3, 4):
    bs = split(TOTAL, parts)
    engs = [make(b) for b in bs]
    xs = [synth_image(b, 640, 640, dtype=torch.float16, seed=i).cuda() for i, b in enumerate(bs)]
    outs = [torch.empty((b, engs[0].pred_channels, engs[0].anchors), dtype=torch.float32, device="cuda") for b in bs]
    streams = [torch.cuda.Stream(priority=-1) for _ in range(parts)]
    main = torch.cuda.current_stream()

    def step():
        ev = torch.cuda.Event()
        ev.record(main)
        for e_, x_, o_, s_ in zip(engs, xs, outs, streams):
            s_.wait_event(ev)
            e_.forward(x_, out_pred=o_, stream=s_)
            d = torch.cuda.Event()
            d.record(s_)
            main.wait_event(d)

    t = timeit(step)
    print(f"{parts} x B={bs} on {parts} streams: {t:.4f} ms -> {TOTAL / t * 1e3:8.0f} img/s", flush=True)
    # staggered start: the engines' small layers line up with the others' large layers less often
    del engs
