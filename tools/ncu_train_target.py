"""Target for ncu captures of the training kernels: three YOLOv11s tensor-core steps at batch argv[1] (default 16).
ncu -k regex:<kernel> picks the launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import MODELS, synth_targets  # noqa: E402
from tests.util import oracle_model, synth_image  # noqa: E402
from yolosharp_b200.train_native import NativeTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
arch, size, task, _ = MODELS["v11s"]
dev = torch.device("cuda", 0)
m = oracle_model(arch, task, size)
st = NativeTrainer({k: v.detach().clone() for k, v in m.state_dict().items()}, "v11", size, 80, device=dev, max_batch=B)
x, t = synth_image(B, 640, 640, seed=1).to(dev), synth_targets(B, 2)
for _ in range(2):
    st.step(x, t)
torch.cuda.synchronize()
