"""Test helpers: layout conversion between NCHW fp32 torch tensors and the library's NHWC bf16 views."""
import ctypes as C

import torch
import torch.nn.functional as F

from mimo_unet_b200._lib import Act


def p8(c):
    return (c + 7) // 8 * 8


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bf16r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float()


def make_buffer(n, h, w, pad, cpitch, device="cuda", fill=float("nan")):
    """Allocates an NHWC bf16 buffer filled with NaN so that any read of unwritten memory shows up."""
    t = torch.full((n, h + 2 * pad, w + 2 * pad, cpitch), fill, dtype=torch.bfloat16, device=device)
    return t


def act_of(buf: torch.Tensor, pad: int, c_off: int, c: int) -> Act:
    n, hp, wp, cp = buf.shape
    return Act(buf.data_ptr(), n, hp - 2 * pad, wp - 2 * pad, pad, cp, c_off, c)


def put_nchw(buf: torch.Tensor, x: torch.Tensor, pad: int, c_off: int = 0, reflect: bool = True):
    """Writes x [N,C,H,W] fp32 into the view (interior + reflect halo)."""
    if pad:
        x = F.pad(x, (1, 1, 1, 1), mode="reflect") if reflect else F.pad(x, (1, 1, 1, 1))
    buf[..., c_off:c_off + x.shape[1]] = x.permute(0, 2, 3, 1).to(torch.bfloat16)


def get_nchw(buf: torch.Tensor, pad: int, c_off: int, c: int, with_halo: bool = False) -> torch.Tensor:
    t = buf
    if pad and not with_halo:
        t = t[:, 1:-1, 1:-1]
    return t[..., c_off:c_off + c].permute(0, 3, 1, 2).float().contiguous()


def stream():
    return torch.cuda.current_stream().cuda_stream
