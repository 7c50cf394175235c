"""Image ingest for the hot path: the host half of the reference's `read_color_image`
(lib/datasets/utils.py:61-77, demo_inference.py:12-29) — decode + colour order + resize with cv2 — stopping BEFORE
`.float().permute(2, 0, 1) / 255`: the uint8 HWC RGB array goes to the GPU as it is (a quarter of the PCIe bytes) and
the division, the crop to multiples of 14 (mickey_extractor.py:46) and the patch gather run in one CUDA kernel
(`mk_forward_u8` / `mk_extract_u8`, csrc/io_ops.cu), bit-identical to feeding the reference's float tensor.
"""
from __future__ import annotations

import numpy as np
import torch


def read_color_image_u8(path, resize=(540, 720)) -> torch.Tensor:
    """cv2.imread -> BGR2RGB -> cv2.resize(resize = (w, h)): uint8 [h, w, 3], exactly the array the reference holds at
    lib/datasets/utils.py:71 before it is normalised."""
    import cv2
    image = cv2.imread(str(path), cv2.IMREAD_COLOR)
    if image is None:
        raise FileNotFoundError(str(path))
    image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
    if resize is not None:
        image = cv2.resize(image, resize)
    return torch.from_numpy(np.ascontiguousarray(image))


def read_color_image(path, resize=(540, 720)) -> torch.Tensor:
    """The reference's full function (lib/datasets/utils.py:61-77): float32 [3, h, w] in [0, 1]."""
    return to_float_chw(read_color_image_u8(path, resize))


def to_float_chw(img_u8: torch.Tensor) -> torch.Tensor:
    """uint8 [..., h, w, 3] -> float32 [..., 3, h, w] / 255 (lib/datasets/utils.py:74); what the u8 kernel fuses."""
    return img_u8.float().movedim(-1, -3) / 255


def from_float_chw(img: torch.Tensor) -> torch.Tensor:
    """Inverse of to_float_chw for tensors that came from 8-bit images (exact: v/255*255 rounds back to v)."""
    return (img * 255).round().clamp_(0, 255).to(torch.uint8).movedim(-3, -1).contiguous()
