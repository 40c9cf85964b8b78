"""Multi-GPU plumbing: one process per GPU (torchrun).
Inference: images sharded per rank in contiguous blocks, detections exchanged with ONE all-gather of the
fixed-capacity buffers.  Training: data-parallel gradient averaging over ONE flat buffer
(`GradBucket`) - the reference's `loss.sum().backward(); step()` (Utils/Amp.cs:260-286) is single-device
(Data/Config.cs:301), so this layer is net-new.
Works with the NCCL backend on GPUs and with gloo on CPU tensors (used by the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_images, rank, world):
    """Contiguous block partition of `n_images` over `world` ranks (first ranks take the remainder)."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_detections(dets, counts, out_dets=None, out_counts=None, group=None):
    """dets (B_local, max_det, W) float32, counts (B_local,) int32 with the same B_local on every rank
    -> (world*B_local, max_det, W), (world*B_local,) in rank order = global image order."""
    world = dist.get_world_size(group)
    if out_dets is None:
        out_dets = torch.empty((world * dets.shape[0],) + tuple(dets.shape[1:]), dtype=dets.dtype, device=dets.device)
    if out_counts is None:
        out_counts = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out_dets, dets.contiguous(), group=group)
    dist.all_gather_into_tensor(out_counts, counts.contiguous(), group=group)
    return out_dets, out_counts


def pad_shard(images, rank, world):
    """Slice this rank's block out of a global batch and pad it to the common per-rank size, so the
    fixed-capacity all-gather stays regular when world does not divide the batch.
    -> (local batch padded to ceil(N/world), number of real images in it)."""
    n = images.shape[0]
    per = -(-n // world)
    s, e = shard_range(n, rank, world)
    local = images[s:e]
    real = local.shape[0]
    if real < per:
        pad = torch.zeros((per - real,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
        local = torch.cat([local, pad], 0)
    return local, real


def unpad_gathered(all_dets, all_counts, n_images, world):
    """Inverse of pad_shard on the gathered buffers: drop the padding slots, restore global image order."""
    per = all_counts.shape[0] // world
    idx = []
    for r in range(world):
        s, e = shard_range(n_images, r, world)
        idx.extend(range(r * per, r * per + (e - s)))
    idx = torch.tensor(idx, dtype=torch.long, device=all_dets.device)
    return all_dets.index_select(0, idx), all_counts.index_select(0, idx)


class GradBucket:
    """All gradients of the model live in ONE flat float32 buffer (views per parameter), so data-parallel training
    needs a single all-reduce per step (NVSwitch: the cost is launch latency, not link count) and the optimizer
    (`yb_adamw_step`) runs over the same flat buffer.  The reference multiplies the loss by the LOCAL batch size
    (Loss.cs:473) and does not average over samples, so ranks' gradients are SUMMED (= one big batch), not meaned."""

    def __init__(self, shapes, device="cpu", dtype=torch.float32):
        self.shapes = [tuple(s) for s in shapes]
        self.offsets, n = [], 0
        for s in self.shapes:
            self.offsets.append(n)
            n += int(torch.Size(s).numel())
        self.flat = torch.zeros(n, dtype=dtype, device=device)

    def view(self, i):
        n = int(torch.Size(self.shapes[i]).numel())
        return self.flat[self.offsets[i]:self.offsets[i] + n].view(self.shapes[i])

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None, average=False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(dist.get_world_size(group))
        return self.flat
