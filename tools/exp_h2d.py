"""Host->device copy rate of one pinned uint8 batch (39.3 MB): one copy vs chunks on several streams."""
import torch
n = 32 * 3 * 640 * 640
x = torch.empty(n, dtype=torch.uint8).pin_memory()
x.random_(0, 255)
d = torch.empty(n, dtype=torch.uint8, device="cuda")


def run(chunks):
    streams = [torch.cuda.Stream() for _ in range(chunks)]
    step = n // chunks
    main = torch.cuda.current_stream()

    def once():
        ev = torch.cuda.Event()
        ev.record(main)
        for i, s in enumerate(streams):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                d[i * step:(i + 1) * step].copy_(x[i * step:(i + 1) * step], non_blocking=True)
            e = torch.cuda.Event()
            e.record(s)
            main.wait_event(e)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        once()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"{chunks} chunk(s): {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s")


for c in (1, 2, 4, 8):
    run(c)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -6", shell=True, capture_output=True, text=True).stdout)
