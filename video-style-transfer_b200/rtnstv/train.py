"""Drop-in for RT/train.py: module-level constants, `spatial_loss`-equivalent terms and a no-argument
`train()`; the step is `vst_b200.train_core.PairTrainer(family="rtnstv")` (RT/train.py:97-143)."""
from __future__ import annotations

import os
from collections import OrderedDict

import torch

from ..data import DevicePrefetcher, SyntheticPairs
from ..train_core import PairTrainer
from .network import StylizingNetwork
from .vgg19 import VGG19

device = "cuda"
epoch_start = 1
epoch_end = 10
batch_size = 2
LR = 1e-3
ALPHA = 1e7
BETA = 5e7
GAMMA = 5e-1
LAMBDA = 1e6
IMG_SIZE = (640, 360)


def train(dataloader=None, style=None, model=None, vgg19=None, save_dir="./models", process_group=None, log=print,
          precision="fp32"):
    if not torch.cuda.is_available():
        raise RuntimeError("train() needs a GPU: the product path has no CPU fallback")
    if dataloader is None:
        dataloader = SyntheticPairs(IMG_SIZE, 1, batch_size, device=device)
    model = (model or StylizingNetwork()).to(device)
    vgg19 = (vgg19 or VGG19()).ensure_weights("rtnstv.train.train()").to(device)
    if style is None:  # the reference loads ./styles/candy.jpg at its native size (RT/train.py:87-89)
        from .. import synth

        style = synth.smooth_frames(1, IMG_SIZE[1], IMG_SIZE[0], "style")
    trainer = PairTrainer(model, vgg19, style, "rtnstv", lr=LR, alpha=ALPHA, beta=BETA, gamma=GAMMA, lambda_o=LAMBDA,
                          process_group=process_group, precision=precision)
    for epoch in range(epoch_start, epoch_end + 1):
        # the reference's four blocking `.to(device)` calls become a double-buffered copy stream (data.DevicePrefetcher)
        for it, (img1, img2, flow, mask) in enumerate(DevicePrefetcher(dataloader, device)):
            terms = trainer.step(img1, img2, flow, mask).to_dict()
            postfix = OrderedDict((k, terms[k]) for k in ("loss", "CL", "SL", "RL", "TL"))
            if log:
                log(f"Epoch {epoch}/{epoch_end} it {it}: " + ", ".join(f"{k}={v:.4g}" for k, v in postfix.items()))
        if save_dir:
            os.makedirs(save_dir, exist_ok=True)
            torch.save(model.state_dict(), os.path.join(save_dir, f"epoch_{epoch}_batchSize_{batch_size}.pth"))
    return model


if __name__ == "__main__":
    train()
