"""Synthetic stand-in for the reference's `(img1, img2, flow_into_past, mask)` sample contract
(RC/datasets.py:100-146, RT/datasets.py): the reference ships no data and the SceneFlow / Videvo
loaders are outside the hot path.  Samples are deterministic (numpy PCG64 streams)
so every rank / box regenerates the same pairs; masks come from the repo's own `flow_warp_mask`
kernel when a device is given, batched on the GPU instead of per-sample in DataLoader workers."""
from __future__ import annotations

import torch

from . import ops, synth


class SyntheticPairs:
    """Iterable of `n_batches` batches: img1, img2 [B,3n,H,W] in 0..255, flow [B,2,H,W] px, mask [B,H,W] {0,1}."""

    def __init__(self, size_wh, input_frame_num: int = 1, batch_size: int = 2, n_batches: int = 8, seed: int = 7,
                 device=None, rank: int = 0):
        self.W, self.H = size_wh
        self.n, self.B, self.n_batches, self.seed, self.device, self.rank = input_frame_num, batch_size, n_batches, seed, device, rank

    def __len__(self):
        return self.n_batches

    def __iter__(self):
        H, W, B = self.H, self.W, self.B
        for i in range(self.n_batches):
            s = self.seed + 1000 * self.rank + i
            img1 = torch.cat([synth.smooth_frames(B, H, W, f"pairs:img1:{j}", s) for j in range(self.n)], 1)
            img2 = torch.cat([synth.smooth_frames(B, H, W, f"pairs:img2:{j}", s) for j in range(self.n)], 1)
            f01 = synth.smooth_flow(B, H, W, "pairs:f01", s, mag=3.0)
            f10 = -f01 + synth.flow(B, H, W, "pairs:noise", s, mag=0.8)
            if self.device is not None:
                dev = torch.device(self.device)
                img1, img2, f01, f10 = (t.to(dev, non_blocking=True) for t in (img1, img2, f01, f10))
                mask = ops.flow_warp_mask(f01, f10, 2.0)
            else:
                mask = synth.mask(B, H, W, "pairs:mask", s)
            yield img1, img2, f10, mask


class SceneFlowAdapter:
    """Device side of the SceneFlow / Monkaa sample contract (RC/datasets.py:100-146), batched.

    The host keeps what it does in the reference's DataLoader workers - PIL decode + resize of the frames and the
    motion-boundary image, PFM decode of the two flows (RC/flowlib.py:34-69) - and hands over NATIVE-resolution flows.
    On the GPU: bilinear resize of both flows (`F.interpolate(..., align_corners=False)`), the reference's per-channel
    rescale, `flow_warp_mask`, and the motion-boundary factor.  The rescale reproduces RC/datasets.py:131-134 literally:
    channel 0 (u) is multiplied by the HEIGHT ratio and channel 1 (v) by the WIDTH ratio (harmless at 960x540 -> 640x360
    where both are 2/3; SURVEY.md notes it as a quirk that defines parity)."""

    def __init__(self, resolution_wh=(640, 360), device="cuda"):
        self.W, self.H = resolution_wh
        self.device = torch.device(device)

    def __call__(self, img1, img2, flow_into_future, flow_into_past, motion):
        """img1, img2 [B,3n,H,W] 0..255 (already resized); flows [B,2,H0,W0] px at native size; motion [B,H,W] (any
        non-zero value marks a motion boundary) -> (img1, img2, flow_into_past [B,2,H,W], mask [B,H,W]) on the device."""
        dev = self.device
        ff, fp = flow_into_future.to(dev, torch.float32), flow_into_past.to(dev, torch.float32)
        if ff.dim() != 4 or ff.shape[1] != 2 or fp.shape != ff.shape:
            raise ValueError("SceneFlowAdapter: flows must be [B,2,H0,W0] and equal in shape")
        H0, W0 = ff.shape[2:]
        scale = (self.H / H0, self.W / W0)
        ff = ops.resize_bilinear(ff, (self.H, self.W), scale)
        fp = ops.resize_bilinear(fp, (self.H, self.W), scale)
        mask = ops.flow_warp_mask(ff, fp)
        ops.motion_mask_(mask, motion.to(dev, torch.float32))
        return img1.to(dev), img2.to(dev), fp, mask


class DevicePrefetcher:
    """Host batches -> device batches with the copy of batch i+1 hidden under the step of batch i.

    The reference moves every tensor of a batch with a blocking `.to(device)` at the top of the loop body
    (RC/train_single/train_starry-night.py:72-75, RT/train.py:101-104), so the step waits for ~32 MB of PCIe traffic.
    Here each batch (a tuple of host tensors, pinned or not) is copied into one of two device slots on a dedicated copy
    stream; the consumer's stream waits on the slot's `ready` event, and a slot is refilled only after the work the
    consumer enqueued while holding it has finished (`consumed` event).  Yields tuples of device tensors that stay valid
    until the next-but-one iteration."""

    def __init__(self, batches, device="cuda"):
        self.batches, self.device = batches, torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("DevicePrefetcher needs a CUDA device (no CPU fallback)")

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        dev = self.device
        copy = torch.cuda.Stream(dev)
        slots = [None, None]
        pins = [None, None]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def issue(k, host):
            host = tuple(host)
            if slots[k] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slots[k], host)):
                slots[k] = tuple(torch.empty(h.shape, dtype=h.dtype, device=dev) for h in host)
                pins[k] = tuple(None if (h.is_cuda or h.is_pinned()) else torch.empty(h.shape, dtype=h.dtype).pin_memory() for h in host)
            if any(h.is_cuda for h in host):            # device-resident sources (e.g. SyntheticPairs(device=...)): they were
                copy.wait_stream(torch.cuda.current_stream(dev))   # produced by kernels on the consumer's stream
            with torch.cuda.stream(copy):
                copy.wait_event(consumed[k])            # the previous tenant of this slot has been consumed
                for d, h, p in zip(slots[k], host, pins[k]):
                    if h.is_cuda:
                        h.record_stream(copy)           # the source may be freed by its producer before the copy has run
                    if p is not None:                   # pageable source: stage through pinned memory so the copy is async
                        ready[k].synchronize()          # the previous async copy out of this pinned buffer has finished
                        p.copy_(h)
                        h = p
                    d.copy_(h, non_blocking=True)
                ready[k].record(copy)

        it = iter(self.batches)
        first = next(it, None)
        if first is None:
            return
        issue(0, first)
        i = 0
        while True:
            k = i & 1
            nxt = next(it, None)
            if nxt is not None:
                issue(k ^ 1, nxt)
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ready[k])
            yield slots[k]
            consumed[k].record(torch.cuda.current_stream(dev))
            if nxt is None:
                return
            i += 1
