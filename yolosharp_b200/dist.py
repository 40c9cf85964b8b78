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


def packed_detection_buffers(B, max_det, row_w, device, storage=None):
    """(dets (B,max_det,row_w) float32, counts (B,) int32) as views of ONE contiguous byte buffer laid out as the
    library's detection payload (dets, padded to 16 bytes, then counts) - so NMS writes straight into what is sent."""
    nd = B * max_det * row_w * 4
    off = (nd + 15) // 16 * 16
    total = off + (B * 4 + 15) // 16 * 16
    if storage is None:
        storage = torch.zeros(total, dtype=torch.uint8, device=device)
    assert storage.numel() >= total
    dets = storage[:nd].view(torch.float32).view(B, max_det, row_w)
    counts = storage[off:off + B * 4].view(torch.int32)
    return dets, counts, storage[:total]


class DetectionGather:
    """All-gather of the post-NMS detections of a batch sharded over the ranks.
      mode "comm": the library's peer-memory exchange (csrc/comm.cu): yb_nms writes into the send buffer, one push
                   kernel stores it into every rank's window over NVLink, a flag wait makes the window readable;
      mode "nccl": ONE ncclAllGather of the packed payload (dets + counts in one buffer) on the process group's
                   high-priority stream (round 1 used two gathers on a default-priority stream and starved).
    gathered(slot) -> (dets (world*B, max_det, row_w), counts (world*B,)) views in global image order."""

    def __init__(self, B, max_det, row_w, device, mode="comm", slots=2, group=None):
        from .engine import Comm, detection_payload_bytes
        self.B, self.max_det, self.row_w, self.device, self.mode, self.group = B, max_det, row_w, device, mode, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.bytes = detection_payload_bytes(B, max_det, row_w)
        self.nd = B * max_det * row_w * 4
        self.off = (self.nd + 15) // 16 * 16
        self.comm = None
        if mode == "comm":
            self.comm = Comm(self.rank, self.world, device.index, self.bytes, slots)
            self.send = [self.comm.send_buffer(k) for k in range(slots)]
            self.win = [self.comm.window(k) for k in range(slots)]
        else:
            self.send = [torch.zeros(self.bytes, dtype=torch.uint8, device=device) for _ in range(slots)]
            self.win = [torch.zeros((self.world, self.bytes), dtype=torch.uint8, device=device) for _ in range(slots)]

    def local_buffers(self, slot):
        d, c, _ = packed_detection_buffers(self.B, self.max_det, self.row_w, self.device, self.send[slot])
        return d, c

    def gather(self, slot, stream=None):
        """enqueue the exchange of slot `slot` on `stream` (default: current); afterwards (stream order) gathered(slot)
        is valid until the next gather of the same slot."""
        if self.comm is not None:
            self.comm.allgather(slot, stream)
            self.comm.release(slot, stream)  # the bench consumes nothing on the device; real consumers release later
        else:
            with torch.cuda.stream(stream) if stream is not None else _null():
                dist.all_gather_into_tensor(self.win[slot].view(-1), self.send[slot], group=self.group)

    def gathered(self, slot):
        w = self.win[slot]
        dets = w[:, :self.nd].contiguous().view(torch.float32).view(self.world * self.B, self.max_det, self.row_w)
        counts = w[:, self.off:self.off + self.B * 4].contiguous().view(torch.int32).view(self.world * self.B)
        return dets, counts

    # ---- end-to-end (host buffers) ----
    def predict_submit(self, eng, slot, images_host, all_dets_host, all_counts_host, conf, iou):
        if self.comm is not None:
            eng.predict_u8_submit_gather(self.comm, slot, images_host, all_dets_host, all_counts_host, conf, iou, self.max_det)
            return
        # NCCL variant: local predict into this rank's rows of the host buffers is not enough - gather on the device
        if not hasattr(self, "_h"):
            self._h = [(torch.empty((self.B, self.max_det, self.row_w), dtype=torch.float32).pin_memory(),
                        torch.empty((self.B,), dtype=torch.int32).pin_memory()) for _ in range(len(self.send))]
            self._s = [torch.cuda.Stream(self.device, priority=-1) for _ in range(len(self.send))]
        eng.predict_u8_submit(slot, images_host, self._h[slot][0], self._h[slot][1], conf, iou, self.max_det)
        self._pending = getattr(self, "_pending", {})
        self._pending[slot] = (all_dets_host, all_counts_host)

    def predict_wait(self, eng, slot):
        eng.predict_u8_wait(slot)
        if self.comm is not None:
            return
        all_d, all_c = self._pending.pop(slot)
        d, c = self.local_buffers(slot)
        with torch.cuda.stream(self._s[slot]):
            d.copy_(self._h[slot][0], non_blocking=True)
            c.copy_(self._h[slot][1], non_blocking=True)
            dist.all_gather_into_tensor(self.win[slot].view(-1), self.send[slot], group=self.group)
            gd, gc = self.gathered(slot)
            all_d.copy_(gd, non_blocking=True)
            all_c.copy_(gc, non_blocking=True)
        self._s[slot].synchronize()

    def describe(self):
        return ("yb_comm peer-memory push (NVLink stores + sequence flags, no collective kernel)" if self.comm is not None
                else "one packed ncclAllGather on a high-priority stream") + f", {self.bytes} B per rank"

    def close(self):
        if self.comm is not None:
            self.comm.close()


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


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
