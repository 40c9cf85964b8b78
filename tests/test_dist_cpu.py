"""world_size-2/3 gloo tests (CPU) of the multi-GPU host logic: block sharding, padding, and the
all-gather of the fixed-capacity detection buffers restore the global image order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolosharp_b200 import dist as ydist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        images = torch.rand(n_images, 3, 4, 4, generator=g)  # every rank builds the same global batch
        local, real = ydist.pad_shard(images, rank, world)
        # stand-in for forward+NMS: detections that encode the image content
        max_det = 5
        dets = torch.zeros(local.shape[0], max_det, 6)
        counts = torch.zeros(local.shape[0], dtype=torch.int32)
        for i in range(real):
            k = 1 + int(local[i].sum().item() * 7) % max_det
            counts[i] = k
            dets[i, :k, 4] = local[i].mean()
            dets[i, :k, 0] = torch.arange(k)
        all_dets, all_counts = ydist.gather_detections(dets, counts)
        d, c = ydist.unpad_gathered(all_dets, all_counts, n_images, world)
        ok = d.shape[0] == n_images
        for i in range(n_images):
            k = 1 + int(images[i].sum().item() * 7) % max_det
            ok &= int(c[i]) == k and bool(torch.allclose(d[i, :k, 4], images[i].mean().expand(k)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _run(world, n_images):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def test_shard_range_partitions():
    for n in (1, 7, 8, 32, 64):
        for w in (1, 2, 3, 8):
            spans = [ydist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gather_world2_even():
    _run(2, 8)


def test_gather_world3_ragged():
    _run(3, 7)


def _grad_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # a tiny conv + BN "model": every rank computes gradients on its shard of one global batch; after the flat
        # all-reduce every rank must hold the gradient of the whole batch (sum of per-sample losses, as Loss.cs:473)
        torch.manual_seed(0)
        w = torch.randn(8, 3, 3, 3)
        b = torch.randn(8)
        x = torch.randn(6, 3, 10, 10)
        bucket = ydist.GradBucket([w.shape, b.shape])

        def grads(xs):
            ww, bb = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
            loss = (torch.nn.functional.conv2d(xs, ww, bb, padding=1) ** 2).sum()
            return torch.autograd.grad(loss, (ww, bb))
        s, e = ydist.shard_range(6, rank, world)
        gw, gb = grads(x[s:e])
        bucket.view(0).copy_(gw)
        bucket.view(1).copy_(gb)
        bucket.all_reduce()
        rw, rb = grads(x)
        ok = torch.allclose(bucket.view(0), rw, rtol=1e-4, atol=1e-4) and torch.allclose(bucket.view(1), rb, rtol=1e-4, atol=1e-4)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_grad_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _ddp_step_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.test_train_step import _targets
        from tests.torch_train_ops import TorchOps
        from tests.util import oracle_model, synth_image
        from yolosharp_b200.train import TrainStepV8
        torch.manual_seed(0)
        m = oracle_model("v8", "detect", "n")
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        x = synth_image(2, 64, 64, seed=10 + rank)       # every rank has its own shard of the global batch
        t = _targets(2, seed=20 + rank)
        solo = TrainStepV8(sd, "n", 80, device="cpu", ops=TorchOps(), lr=1e-3)
        solo.group = False                                 # this rank's gradient alone
        solo.step(x, t)
        g_local = solo.P.grad.clone()
        ddp = TrainStepV8(sd, "n", 80, device="cpu", ops=TorchOps(), lr=1e-3)
        ddp.step(x, t)
        gathered = [torch.zeros_like(g_local) for _ in range(world)]
        dist.all_gather(gathered, g_local)
        ok = torch.allclose(ddp.P.grad, sum(gathered), rtol=1e-5, atol=1e-6)
        weights = [torch.zeros_like(ddp.P.flat) for _ in range(world)]
        dist.all_gather(weights, ddp.P.flat)
        ok = ok and all(torch.equal(weights[0], w) for w in weights)   # replicas stay bit-identical
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_train_step_world2():
    """Two ranks, each a shard of the batch: after the flat all-reduce every rank holds the SUM of the ranks' gradients
    and the AdamW step leaves the replicas bit-identical (train.py, SURVEY.md section 8(e) for the training path)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_step_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _fit_empty_shard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.test_train_step import _targets
        from tests.torch_train_ops import TorchOps
        from tests.util import oracle_model, synth_image
        from yolosharp_b200.train import TrainStepV8, fit
        torch.manual_seed(0)
        m = oracle_model("v8", "detect", "n")
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        st = TrainStepV8(sd, "n", 80, device="cpu", ops=TorchOps(), lr=1e-3)
        x = synth_image(2, 64, 64, seed=30 + rank)
        empty = torch.zeros(0, 6)
        # iteration 0: rank 1's shard has no targets (it must still enter the collective);
        # iteration 1: NO rank has targets (skipped everywhere); iteration 2: both have targets
        batches = [(x, _targets(2, seed=40) if rank == 0 else empty), (x, empty), (x, _targets(2, seed=41 + rank))]
        calls = []
        fit(st, batches, epochs=1, on_iteration=lambda e, i, lrs, items: calls.append((e, i)))
        weights = [torch.zeros_like(st.P.flat) for _ in range(world)]
        dist.all_gather(weights, st.P.flat)
        ok = calls == [(1, 0), (1, 1)] and st.step_count == 2 and all(torch.equal(weights[0], w) for w in weights)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_fit_with_an_empty_shard_world2():
    """a rank whose shard has no targets must not skip the step alone (the gradient all-reduce would
    mispair): a batch is skipped only when every rank is empty; replicas and step counts stay identical."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fit_empty_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


class _FakeStep:
    lr = 1e-3

    def step(self, images, targets, lrs=None):
        return torch.tensor([1.0, 1.0, 1.0])


def _fit_early_stop_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from yolosharp_b200.train import fit
        # alone, rank 0 (best epoch 2) would stop in epoch 4 and rank 1 (best epoch 3) in epoch 5; summed over the ranks the
        # best epoch is 3 and patience 2 ends the run in epoch 5 on BOTH
        val = {0: {1: 5.0, 2: 4.0, 3: 4.5, 4: 4.6, 5: 4.7, 6: 1.0}, 1: {1: 5.0, 2: 4.6, 3: 4.0, 4: 4.3, 5: 4.4, 6: 1.0}}[rank]
        best, ends = [], []
        hist = fit(_FakeStep(), [(None, torch.zeros(1, 6))] * 2, epochs=6, validate=lambda e: [val[e]], patience=2, on_best=best.append,
                   on_epoch_end=ends.append)
        q.put((rank, best, ends, len(hist)))
    finally:
        dist.destroy_process_group()


def test_fit_early_stopping_is_collective_world2():
    """Each rank validates its own shard; the fitness is summed over the ranks before EarlyStopping sees it, so every rank
    leaves the epoch loop in the same epoch (a rank that stopped alone would leave the others hanging in the next all-reduce)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fit_early_stop_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # summed losses: 10.0, 8.6, 8.5, 8.9, 9.1 -> best at 1, 2, 3; epochs 4 and 5 do not improve: stop in epoch 5, no last.bin for it
    assert res[0][1:] == res[1][1:]
    assert res[0][1] == [1, 2, 3] and res[0][2] == [1, 2, 3, 4] and res[0][3] == 5


def _gather_packed_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = ydist.DetectionGather(3, 5, 6, torch.device("cpu"), mode="nccl", slots=2)  # "nccl" = the process group's all-gather
        ok = True
        for step in range(4):
            d, c = g.local_buffers(step & 1)
            d.copy_(torch.arange(3 * 5 * 6, dtype=torch.float32).view(3, 5, 6) + 1000 * rank + step)
            c.copy_(torch.tensor([1, 2, 3], dtype=torch.int32) + 10 * rank + step)
            g.gather(step & 1)
            gd, gc = g.gathered(step & 1)
            for r in range(world):
                ok = ok and torch.equal(gd[3 * r:3 * r + 3], torch.arange(90, dtype=torch.float32).view(3, 5, 6) + 1000 * r + step)
                ok = ok and torch.equal(gc[3 * r:3 * r + 3], torch.tensor([1, 2, 3], dtype=torch.int32) + 10 * r + step)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_packed_detection_gather_world2():
    """One all-gather of the packed (dets + counts) payload: the layout the peer-memory exchange uses on GPUs."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_packed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
