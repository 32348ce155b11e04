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
