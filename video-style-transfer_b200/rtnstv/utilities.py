"""Drop-in for the hot-path helpers of RT/utilities.py: warp, flow_warp_mask(threshold),
gram_matrix (divides by H*W only - SURVEY.md Q4), vgg_normalize (out of place)."""
from __future__ import annotations

import math
import struct
from typing import Iterable, Sequence, Union

import numpy as np
import torch

from .. import ops


def warp(x, flo, padding_mode="zeros"):
    """RT/utilities.py:59-77 (identical to RC's)."""
    if padding_mode != "zeros":
        raise NotImplementedError("only padding_mode='zeros' is used by the reference's hot path")
    return ops.warp(x, flo)


def flow_warp_mask(flo01, flo10, padding_mode="zeros", threshold=2):
    """RT/utilities.py:80-110."""
    if padding_mode != "zeros":
        raise NotImplementedError("only padding_mode='zeros' is used by the reference's hot path")
    return ops.flow_warp_mask(flo01, flo10, float(threshold))


def gram_matrix(y: torch.Tensor):
    """F F^T / (H*W) (RT/utilities.py:155-160)."""
    b, ch, h, w = y.shape
    return ops.gram(y, 1.0 / (h * w))


def vgg_normalize(batch: torch.Tensor):
    """RT/utilities.py:163-169 - does not modify its argument."""
    return ops.vgg_normalize(batch, inplace_div=False)


def read_sintel_flow(filename) -> np.ndarray:
    """Middlebury/Sintel `.flo` reader (RT/utilities.py:113-152): float32 [H, W, 2], same validation and errors."""
    TAG_FLOAT = 202021.25
    with open(filename, "rb") as stream:
        head = stream.read(12)
        if len(head) < 12 or struct.unpack("<f", head[:4])[0] != TAG_FLOAT:
            raise ValueError(f"ReadFlowFile({filename}): wrong tag (possibly due to big-endian machine?)")
        width, height = struct.unpack("<ii", head[4:])
        if width < 1 or width > 99999:
            raise ValueError(f"ReadFlowFile({filename}): illegal width {width}")
        if height < 1 or height > 99999:
            raise ValueError(f"ReadFlowFile({filename}): illegal height {height}")
        data = stream.read(width * height * 2 * 4)
        if len(data) != width * height * 2 * 4:
            raise ValueError(f"ReadFlowFile({filename}): file is too short")
        if stream.read(1):
            raise ValueError(f"ReadFlowFile({filename}): file is too long")
    return np.frombuffer(data, dtype="<f4").reshape(height, width, 2).astype(np.float32)


def temporal_error(styled: Sequence[torch.Tensor], flows: Sequence[torch.Tensor], masks: Sequence[torch.Tensor]) -> float:
    """The Sintel temporal-consistency metric on device tensors (RT/utilities.py:219-240):
    sqrt(mean over pairs of mean(mask * (styled[i] - warp(styled[i+1], flow[i]))^2)).
    styled[i]: [1,3,H,W]; flows[i]: [1,2,H,W] (frame i -> i+1 lookup); masks[i]: [1,H,W] with 1 = valid."""
    if len(styled) != len(flows) + 1 or len(flows) != len(masks) or not flows:
        raise ValueError("temporal_error: need n+1 styled frames for n flows / masks")
    sums = torch.zeros(2 * len(flows), dtype=torch.float32, device=styled[0].device)
    for i, (flow, mask) in enumerate(zip(flows, masks)):
        ops.output_temporal_sums(styled[i + 1], styled[i], None, None, flow, mask, luminance=False, out=sums[2 * i:2 * i + 2])
    per_pair = sums.view(-1, 2)[:, 0].cpu().double() / styled[0].numel()       # .mean() over [1,3,H,W]; one sync for all pairs
    return math.sqrt(float(per_pair.mean()))


def temporal_errors_sintel(model_class, model_path: str, scene: str, device: str = "cuda",
                           root: str = "../datasets/MPI-Sintel-complete/training") -> Union[float, int]:
    """RT/utilities.py:194-240 with the per-pair arithmetic on the GPU (frames / flows / occlusion PNGs are decoded on the
    host exactly as in the reference)."""
    import os

    import cv2

    def files(folder):
        return sorted(os.path.join(folder, f) for f in os.listdir(folder))

    frames_files, mask_files, flow_files = files(f"{root}/final/{scene}"), files(f"{root}/occlusions/{scene}"), files(f"{root}/flow/{scene}")
    model = model_class().to(device)
    model.load_state_dict(torch.load(model_path, weights_only=True), strict=True)
    styled, flows, masks = [], [], []
    for idx in range(len(flow_files) + 1):
        img = cv2.cvtColor(cv2.imread(frames_files[idx]), cv2.COLOR_BGR2RGB)
        x = torch.from_numpy(img).permute(2, 0, 1).float().unsqueeze(0).to(device)
        styled.append(model(x))
    for idx in range(len(flow_files)):
        flows.append(torch.from_numpy(read_sintel_flow(flow_files[idx])).permute(2, 0, 1).unsqueeze(0).contiguous().to(device))
        m = cv2.imread(mask_files[idx], cv2.IMREAD_GRAYSCALE)
        masks.append(torch.from_numpy((m == 0).astype(np.float32)).unsqueeze(0).to(device))
    return temporal_error(styled, flows, masks)


def cvframe_to_tensor(frame, resize=None):
    """BGR uint8 HxWx3 -> RGB float [3,H,W] in 0..255, optional (width, height) resize (RT/utilities.py:182-191)."""
    import cv2

    frame = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
    if resize is not None:
        frame = cv2.resize(frame, resize, interpolation=cv2.INTER_LINEAR)
    return torch.from_numpy(frame).permute(2, 0, 1).float()


class Inference:
    """Per-frame video stylisation iterator (RT/utilities.py:296-332): yields uint8 BGR 360x640x3 frames.  The network
    output is already in [0, 255] (tanh map), so the byte conversion is the reference's truncating `astype("uint8")`."""

    def __init__(self, model_class, model_path: str, video_path: str, device: str = "cuda", precision: str = "fp32"):
        import cv2

        self.model = model_class().to(device)
        self.model.load_state_dict(torch.load(model_path, weights_only=True), strict=True)
        self.model.set_precision(precision)          # "bf16": the captured tensor-core plan (infer.RtnstvStylizer)
        self.video_path, self.device = video_path, device
        self.cap = cv2.VideoCapture(video_path)

    def __del__(self):
        cap = getattr(self, "cap", None)
        if cap is not None:
            cap.release()

    def __iter__(self):
        from ..infer import RtnstvStylizer

        st = RtnstvStylizer(self.model, 360, 640)
        while True:
            ret, frame = self.cap.read()
            if not ret:
                break
            # the frame is [0, 255] by construction (tanh map); the byte conversion is the reference's truncating astype
            yield st.stylize_u8(cvframe_to_tensor(frame, resize=(640, 360)).unsqueeze(0))[0].copy()
