"""The native training step (csrc/train_step.cu, one C-ABI call per step) against its executable specification, the
Python step of yolosharp_b200/train.py / train_v11.py over the same kernels (which in turn is pinned to autograd through
the oracle, tests/test_train_step.py).  CPU: the parameter layout (names, shapes, bias group) against the oracle model's
state_dict.  GPU: one step on the same weights and batch - loss items, every parameter gradient, updated weights,
BatchNorm running statistics."""
import pytest
import torch

from tests.test_train_step import _targets
from tests.util import oracle_model, synth_image


@pytest.mark.parametrize("arch,size", [("v8", "n"), ("v8", "s"), ("v8", "m"), ("v8", "x"), ("v11", "n"), ("v11", "s"), ("v11", "m"),
                                       ("v11", "x")])
def test_native_trainer_layout_matches_reference_state_dict(arch, size):
    from yolosharp_b200.train_native import NativeTrainer
    sd = oracle_model(arch, "detect", size).state_dict()
    t = NativeTrainer(None, arch, size, 80, device="cpu")  # dry run: layout only
    want_p = {k: tuple(v.shape) for k, v in sd.items() if v.dtype.is_floating_point and v.numel() > 0 and ".dfl." not in k and
              not k.endswith(("running_mean", "running_var"))}
    want_s = {k: tuple(v.shape) for k, v in sd.items() if k.endswith(("running_mean", "running_var"))}
    assert {k: s for k, (_, _, s) in t.params.items()} == want_p
    assert {k: s for k, (_, _, s) in t.stats.items()} == want_s
    # flat layout: contiguous, the "bias" group first (YoloBaseTaskModel.cs:144-153)
    offs = sorted((o, c, k) for k, (o, c, _) in t.params.items())
    pos = 0
    for o, c, k in offs:
        assert o == pos
        pos += c
        assert ("bias" in k) == (o < t.n_bias)
    t.close()


def _run_pair(arch, B, H, W, py_cls, ops_cls):
    from yolosharp_b200.train_native import NativeTrainer
    torch.manual_seed(0)
    m = oracle_model(arch, "detect", "n")
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = synth_image(B, H, W).cuda()
    targets = _targets(B)
    a = py_cls(sd0, "n", 80, device="cuda", ops=ops_cls(tensor_cores=True), lr=1e-3)
    b = NativeTrainer(sd0, arch, "n", 80, device="cuda", max_batch=B, height=H, width=W, lr=1e-3)
    ia = a.step(x, targets).cpu()
    ib = b.step(x, targets)
    torch.cuda.synchronize()
    return a, b, ia, ib


def _check(a, b, ia, ib):
    torch.testing.assert_close(ib, ia, rtol=1e-5, atol=1e-6)
    gmax = float(a.P.grad.abs().max())
    errs = []
    for k in a.P.names:
        ga, gb = a.P.g(k), b.g(k)
        # same kernels, but the native walk writes through channel-slice views where the Python walk copies, so a few sums
        # are formed in another order: tensors whose gradient is mathematically ~0 (the BatchNorm bias of SPPF.cv1, see
        # test_train_step._compare) differ by their rounding noise - judged against the step's gradient scale
        errs.append((float((gb - ga).abs().max()) / max(float(ga.abs().max()), 1e-3 * gmax), k))
        torch.testing.assert_close(b.p(k), a.P.p(k), rtol=1e-5, atol=2.1e-3)  # Adam's first step: +-lr on a sign flip of a ~0 gradient
    errs.sort(reverse=True)
    ga, gb = a.P.grad.double(), torch.cat([b.g(k).reshape(-1) for k in a.P.names]).double()
    l2 = float((gb - ga).norm() / ga.norm())
    print("worst parameter gradients vs the Python step: " + ", ".join(f"{k} {e:.2e}" for e, k in errs[:3]) + f"; flat rel L2 {l2:.2e}")
    # same kernels in the same order: bit-level agreement (the nearest-2x upsample backward sums its 2 x 2 block in ATen's
    # order for that reason).  Input sizes are the ones at which the Python walk itself is run-to-run deterministic - its
    # max-pool backward is ATen's atomicAdd kernel, and at 128 x 160 two Python steps differ from each other by 1e-3 of the
    # flat gradient (tools/dbg_native_determinism.py); the native step is bitwise reproducible at every size.
    assert l2 < 1e-6 and errs[0][0] < 1e-4, (l2, errs[:4])
    for k, v in a.P.buffers.items():
        torch.testing.assert_close(b.p(k), v, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_native_train_step_v8_matches_python_step():
    from yolosharp_b200.train import KernelOps, TrainStepV8
    _check(*_run_pair("v8", 2, 64, 96, TrainStepV8, KernelOps))


@pytest.mark.gpu
def test_native_train_step_v11_matches_python_step():
    from yolosharp_b200.train_v11 import KernelOpsV11, TrainStepV11
    _check(*_run_pair("v11", 2, 128, 128, TrainStepV11, KernelOpsV11))


@pytest.mark.gpu
def test_native_train_step_u8_images_and_second_step():
    """uint8 images (scaled by 1/255 inside the library, YoloDataset.cs:140) and a second step on the updated weights stay finite and move."""
    from yolosharp_b200.train_native import NativeTrainer
    torch.manual_seed(0)
    m = oracle_model("v8", "detect", "n")
    t = NativeTrainer(m.state_dict(), "v8", "n", 80, device="cuda", max_batch=2, height=64, width=96, lr=1e-3)
    u8 = synth_image(2, 64, 96, dtype=torch.uint8).cuda()
    i1 = t.step(u8, _targets(2))
    w1 = t.flat.clone()
    i2 = t.step(u8, _targets(2))
    assert torch.isfinite(i1).all() and torch.isfinite(i2).all() and not torch.equal(w1, t.flat)
    f = NativeTrainer(m.state_dict(), "v8", "n", 80, device="cuda", max_batch=2, height=64, width=96, lr=1e-3)
    j1 = f.step(u8.float().mul(1 / 255.0), _targets(2))
    torch.testing.assert_close(j1, i1, rtol=1e-6, atol=1e-7)


def test_load_state_dict_skip_nc_not_equal_layers_cpu():
    """LoadModel(path, skipNcNotEqualLayers: true) (YoloBaseTaskModel.cs:82-92) on the trainer's flat buffers: an 80-class
    checkpoint into a 3-class trainer leaves every model.22.cv3 tensor untouched and loads the rest."""
    from yolosharp_b200.train_native import NativeTrainer, nc_skip_list
    sd80 = oracle_model("v8", "detect", "n").state_dict()
    assert nc_skip_list(sd80, 80) == []
    skip = nc_skip_list(sd80, 3)
    assert len(skip) == 42 and all(k.startswith("model.22.cv3.") for k in skip) and skip[-1] == "model.22.cv3.2.2.bias"
    t = NativeTrainer(None, "v8", "n", 3, device="cpu")  # dry run: layout only; give it host buffers
    t.flat = torch.full((sum(c for _, c, _ in t.params.values()),), 7.0)
    t.running = torch.zeros(sum(c for _, c, _ in t.stats.values()))
    with pytest.raises(ValueError):
        t.load_state_dict(sd80)
    assert t.load_state_dict(sd80, skipNcNotEqualLayers=True) == skip
    assert float(t.p("model.22.cv3.1.2.weight").min()) == 7.0 and torch.equal(t.p("model.0.conv.weight"), sd80["model.0.conv.weight"])
    assert torch.equal(t.p("model.22.cv2.0.0.conv.weight"), sd80["model.22.cv2.0.0.conv.weight"])
