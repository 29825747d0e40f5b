"""Tensor layouts shared by the weight exporter, the host wrapper and the tests.

All activations are NHWC fp16.  Conventions (see csrc/conv_gemm.h):
  * conv weights are K-major rows [cout][tap][cin] with taps in (ky, kx) order;
  * the 7x7/stride-2 stem reads a row-padded NHWC8 image (4 zero pixels left of every row, the
    next row's pad doubles as right pad) and sees each filter ROW as 8 pixels x 8 channels = 64
    K-values starting at pixel 2*ox-4, so its weights are [64][ky][8 px][8 ch] with px 0 and
    channels 6,7 zero;
  * stride-2 3x3 layers read a phase-split input [n][(y&1)*2+(x&1)][H/2][W/2][C].
"""
from __future__ import annotations

import torch

STEM_PAD = 4          # zero pixels stored left of every image row of the packed network input
STEM_PITCH = 160 + 4  # pixels per stored row
PACK_C = 8            # rgb(3) + xyz(3) + 2 zero channels


def pack_conv3x3(w: torch.Tensor) -> torch.Tensor:
    """[cout,cin,3,3] -> [cout, 9*cin] (tap-major, channel-minor)."""
    cout, cin, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin).contiguous()


def pack_stem7(w: torch.Tensor) -> torch.Tensor:
    """[cout,6,7,7] -> [cout, 7*64]: per filter row 8 pixels x 8 channels, pixel 0 = zero."""
    cout, cin, kh, kw = w.shape
    assert (cin, kh, kw) == (6, 7, 7)
    out = torch.zeros(cout, 7, 8, 8, dtype=w.dtype, device=w.device)
    out[:, :, 1:8, :6] = w.permute(0, 2, 3, 1)  # [cout, ky, kx, c]
    return out.reshape(cout, 7 * 64).contiguous()


def phase_split(x: torch.Tensor) -> torch.Tensor:
    """[n,H,W,C] -> [n,4,H/2,W/2,C] with plane = (y&1)*2 + (x&1)."""
    n, h, w, c = x.shape
    return x.reshape(n, h // 2, 2, w // 2, 2, c).permute(0, 2, 4, 1, 3, 5).reshape(n, 4, h // 2, w // 2, c).contiguous()


def phase_merge(x: torch.Tensor) -> torch.Tensor:
    """Inverse of phase_split."""
    n, _, h2, w2, c = x.shape
    return x.reshape(n, 2, 2, h2, w2, c).permute(0, 3, 1, 4, 2, 5).reshape(n, h2 * 2, w2 * 2, c).contiguous()


def pad_stem_input(x6: torch.Tensor) -> torch.Tensor:
    """[n,160,160,6] -> flat fp16 buffer holding [n,160,STEM_PITCH,8] plus one trailing pad row."""
    n, h, w, c = x6.shape
    buf = torch.zeros(n * h * STEM_PITCH * PACK_C + STEM_PITCH * PACK_C, dtype=torch.float16, device=x6.device)
    v = buf[: n * h * STEM_PITCH * PACK_C].view(n, h, STEM_PITCH, PACK_C)
    v[:, :, STEM_PAD:, :c] = x6.to(torch.float16)
    return buf
