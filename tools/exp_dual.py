"""Experiment: (1) forward time vs batch (fixed cost of the ~60-kernel chain), (2) two half-batch engines on two
streams vs one full-batch engine.  python tools/exp_dual.py [v8n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import yolosharp_b200 as y  # noqa: E402
from bench import MODELS  # noqa: E402
from tests.util import oracle_model, synth_image  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "v8n"
arch, size, task, _ = MODELS[model]
m = oracle_model(arch, task, size)
sd = m.state_dict()


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


for B in (1, 2, 4, 8, 16, 32, 64):
    e = make(B)
    x = synth_image(B, 640, 640, dtype=torch.float16).cuda()
    out = torch.empty((B, e.pred_channels, e.anchors), dtype=torch.float32, device="cuda")
    t = timeit(lambda: e.forward(x, out_pred=out))
    print(f"B={B:3d} forward {t:.4f} ms  -> {B / t * 1e3:8.0f} img/s", flush=True)
    del e

for parts, Bp in ((2, 16), (4, 8), (2, 32)):
    engs = [make(Bp) for _ in range(parts)]
    xs = [synth_image(Bp, 640, 640, dtype=torch.float16, seed=i).cuda() for i in range(parts)]
    outs = [torch.empty((Bp, engs[0].pred_channels, engs[0].anchors), dtype=torch.float32, device="cuda") for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    main = torch.cuda.current_stream()

    def step():
        ev = torch.cuda.Event()
        ev.record(main)
        for e, x, o, s in zip(engs, xs, outs, streams):
            s.wait_event(ev)
            e.forward(x, out_pred=o, stream=s)
            d = torch.cuda.Event()
            d.record(s)
            main.wait_event(d)

    t = timeit(step)
    print(f"{parts} x B={Bp} on {parts} streams: {t:.4f} ms -> {parts * Bp / t * 1e3:8.0f} img/s", flush=True)
    del engs
