"""Image / intrinsics helpers with the reference's names (lib/datasets/utils.py:61-99); the image reader is
mickey_b200.io (cv2 decode + resize on the host, normalisation on the host here or fused on the GPU for uint8 batches)."""
import torch

from mickey_b200.io import read_color_image_u8, to_float_chw


def read_color_image(path, resize=(640, 480), augment_fn=None):
    """float32 [3, h, w] in [0, 1]; resize = (w, h) (reference lib/datasets/utils.py:61-77)."""
    image = to_float_chw(read_color_image_u8(path, resize))
    return augment_fn(image) if augment_fn else image


def correct_intrinsic_scale(K, scale_x, scale_y):
    """Intrinsics of the image resized by (scale_x, scale_y) with pixel centres at integer coordinates
    (reference lib/datasets/utils.py:86-99): K' = diag-and-shift(scale) @ K."""
    S = torch.tensor([[scale_x, 0.0, scale_x / 2 - 0.5],
                      [0.0, scale_y, scale_y / 2 - 0.5],
                      [0.0, 0.0, 1.0]], dtype=torch.float32)
    return S @ torch.as_tensor(K, dtype=torch.float32)
