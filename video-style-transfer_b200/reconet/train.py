"""Drop-in for the reference's ReCoNet training entry points (RC/train_single/train_*.py): module-level
constants, a no-argument `train()`, the same per-step loss terms in the progress postfix and the same
checkpoint naming.  The step itself is `vst_b200.train_core.PairTrainer` - hand-written forward,
adjoints and Adam, no autograd - and the data is synthetic unless `dataloader` is supplied
(the SceneFlow loader is outside the hot path)."""
from __future__ import annotations

import os
from collections import OrderedDict

import torch

from ..data import DevicePrefetcher, SyntheticPairs
from ..train_core import PairTrainer
from .network import ReCoNet, Vgg16

device = "cuda"
epoch_start = 1
epoch_end = 6
batch_size = 2
input_frame_num = 1
LR = 1e-3
ALPHA = 1e5
BETA = 1e11
GAMMA = 1e-2
LAMBDA_F = 1e12
LAMBDA_O = 1e7
IMG_SIZE = (640, 360)


def train(dataloader=None, style=None, model=None, vgg16=None, save_dir="./models", process_group=None, log=print,
          precision="fp32"):
    """RC/train_single/train_starry-night.py:31-171.  Returns the trained model."""
    if not torch.cuda.is_available():
        raise RuntimeError("train() needs a GPU: the product path has no CPU fallback")
    if dataloader is None:
        dataloader = SyntheticPairs(IMG_SIZE, input_frame_num, batch_size, device=device)
    model = (model or ReCoNet(input_frame_num)).to(device)
    vgg16 = (vgg16 or Vgg16()).ensure_weights("reconet.train.train()").to(device)
    if style is None:  # the reference loads ./styles/starry-night.jpg resized to IMG_SIZE (:49-51)
        from .. import synth

        style = synth.smooth_frames(1, IMG_SIZE[1], IMG_SIZE[0], "style")
    trainer = PairTrainer(model, vgg16, style, "reconet", lr=LR, alpha=ALPHA, beta=BETA, gamma=GAMMA,
                          lambda_f=LAMBDA_F, lambda_o=LAMBDA_O, process_group=process_group, precision=precision)
    for epoch in range(epoch_start, epoch_end + 1):
        # the reference's four blocking `.to(device)` calls become a double-buffered copy stream (data.DevicePrefetcher)
        for it, (img1, img2, flow, mask) in enumerate(DevicePrefetcher(dataloader, device)):
            terms = trainer.step(img1, img2, flow, mask).to_dict()
            postfix = OrderedDict((k, terms[k]) for k in ("loss", "CL", "SL", "FTL", "OTL", "RL"))
            if log:
                log(f"Epoch {epoch}/{epoch_end} it {it}: " + ", ".join(f"{k}={v:.4g}" for k, v in postfix.items()))
        if save_dir:
            os.makedirs(save_dir, exist_ok=True)
            torch.save(model.state_dict(), os.path.join(
                save_dir, f"Flow_input_{input_frame_num}_epoch_{epoch}_batchSize_{batch_size}.pth"))
    return model


if __name__ == "__main__":
    train()
