"""pad / crop / Interpolate / free_bits_kl as the reference's models/lvae.py:3-4 imports them."""
import torch.nn.functional as F
from torch import nn

from . import init  # noqa: F401


def _pad_crop(x, size, mode):
    size = tuple(int(s) for s in size)
    cur = tuple(x.shape[2:4])
    dr, dc = abs(cur[0] - size[0]), abs(cur[1] - size[1])
    dr1, dr2, dc1, dc2 = dr // 2, dr - dr // 2, dc // 2, dc - dc // 2
    if mode == "pad":
        return F.pad(x, [dc1, dc2, dr1, dr2, 0, 0, 0, 0])
    return x[:, :, dr1:cur[0] - dr2, dc1:cur[1] - dc2]


def pad_img_tensor(x, size):
    return _pad_crop(x, size, "pad")


def crop_img_tensor(x, size):
    return _pad_crop(x, size, "crop")


class Interpolate(nn.Module):
    def __init__(self, size=None, scale=None, mode="bilinear", align_corners=False):
        super().__init__()
        self.size, self.scale, self.mode, self.align_corners = size, scale, mode, align_corners

    def forward(self, x):
        return F.interpolate(x, size=self.size, scale_factor=self.scale, mode=self.mode, align_corners=self.align_corners)


def free_bits_kl(kl, free_bits, batch_average=False, eps=1e-6):
    if free_bits < eps:
        return kl.mean(0)
    if batch_average:
        return kl.mean(0).clamp(min=free_bits)
    return kl.clamp(min=free_bits).mean(0)
