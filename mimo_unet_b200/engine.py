"""UNetPlan: python handle of the C++ whole-network executor (csrc/engine.cu).

PyTorch is plumbing here: it owns the memory (one workspace tensor, one flat gradient tensor) and the
stream; every FLOP of the network runs in libmimo_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import Act, UnetConfig, check


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class UNetPlan:
    """Fixed-shape execution plan for MimoUNet (bilinear path) on the current CUDA device."""

    def __init__(self, in_channels: int, out_channels: int, num_subnetworks: int, filter_base_count: int,
                 batch: int, height: int, width: int, device: torch.device):
        self.lib = _lib.lib()
        check(self.lib.mimo_check_device(), "mimo_check_device")
        self.cfg = UnetConfig(in_channels, out_channels, num_subnetworks, filter_base_count, batch, height, width)
        self.key = (in_channels, out_channels, num_subnetworks, filter_base_count, batch, height, width)
        h = C.c_void_p()
        check(self.lib.mimo_unet_plan_create(C.byref(self.cfg), C.byref(h)), "mimo_unet_plan_create")
        self.handle = h
        self.device = device
        self.ws_bytes = int(self.lib.mimo_unet_workspace_bytes(h))
        # zero-initialised once: pad channels (C..cpitch) of the NHWC buffers are never written by the elementwise kernels
        # and are over-read by their 16-byte vector loads, so they must not hold NaN bit patterns
        self.workspace = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=device)
        self.n_state = int(self.lib.mimo_unet_num_state(h))
        self.n_dconv = int(self.lib.mimo_unet_num_double_convs(h))
        self.drop_channels = [int(self.lib.mimo_unet_dropout_channels(h, i)) for i in range(self.n_dconv)]
        self._state_keepalive: List[torch.Tensor] = []
        self._bound_sig = None

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None:
                self.lib.mimo_unet_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- binding -------------------------------------------------------------------------------
    def bind(self, state: Sequence[torch.Tensor], grads: Sequence[Optional[torch.Tensor]]):
        """state: tensors in MimoUNet.state_dict() order; grads: matching fp32 destinations or None."""
        assert len(state) == self.n_state and len(grads) == self.n_state
        sig = tuple(t.data_ptr() for t in state) + tuple(-1 if g is None else g.data_ptr() for g in grads)
        if sig == self._bound_sig:
            return
        for t in state:
            assert t.is_cuda and t.is_contiguous(), "parameters must be contiguous CUDA tensors"
        sp = (C.c_void_p * self.n_state)(*[t.data_ptr() for t in state])
        gp = (C.c_void_p * self.n_state)(*[_ptr(g) for g in grads])
        check(self.lib.mimo_unet_bind(self.handle, self.workspace.data_ptr(), self.ws_bytes, sp, gp, self.n_state), "mimo_unet_bind")
        self._state_keepalive = list(state) + [g for g in grads if g is not None]
        self._bound_sig = sig

    # -- execution -----------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, out: torch.Tensor, training: bool, gather: Optional[torch.Tensor] = None,
                drop_masks: Optional[Sequence[Optional[torch.Tensor]]] = None):
        assert x.is_contiguous() and x.dtype == torch.float32 and out.is_contiguous() and out.dtype == torch.float32
        mp = None
        if drop_masks is not None:
            assert len(drop_masks) == self.n_dconv
            mp = (C.c_void_p * self.n_dconv)(*[_ptr(m) for m in drop_masks])
            self._mask_keepalive = list(drop_masks)
        check(self.lib.mimo_unet_forward(self.handle, x.data_ptr(), _ptr(gather), int(training), mp, out.data_ptr(), stream_ptr()),
              "mimo_unet_forward")

    def backward(self, dout: torch.Tensor, dx: Optional[torch.Tensor] = None, grad_scale: Optional[torch.Tensor] = None,
                 accumulate: bool = False):
        assert dout.is_contiguous() and dout.dtype == torch.float32
        check(self.lib.mimo_unet_backward(self.handle, dout.data_ptr(), _ptr(grad_scale), _ptr(dx), int(accumulate), stream_ptr()),
              "mimo_unet_backward")

    def set_elementwise_dropout(self, center_keep: Optional[torch.Tensor], center_scale: float,
                                final_keep: Optional[Sequence[Optional[torch.Tensor]]], final_scale: float):
        """nn.Dropout keep masks (bf16 0/1, NHWC with the channel pitch rounded up to 8) for the next forward/backward pair;
        None disables. See mimo_unet_set_elementwise_dropout."""
        S = self.cfg.num_subnetworks
        fk = None
        if final_keep is not None:
            assert len(final_keep) == S
            fk = (C.c_void_p * S)(*[_ptr(m) for m in final_keep])
        self._elem_keepalive = (center_keep, list(final_keep) if final_keep is not None else None)
        check(self.lib.mimo_unet_set_elementwise_dropout(self.handle, _ptr(center_keep), float(center_scale), fk, float(final_scale)),
              "mimo_unet_set_elementwise_dropout")

    def set_inference_fusion(self, on: bool):
        """Fused inference epilogues for the following eval-mode forwards (no backward possible after them)."""
        check(self.lib.mimo_unet_set_inference_fusion(self.handle, int(bool(on))), "mimo_unet_set_inference_fusion")

    def set_backward_events(self, events: Optional[Sequence["torch.cuda.Event"]]):
        """events: 4 torch.cuda.Event objects (already recorded once so their handles exist) or None; see
        mimo_unet_set_backward_events."""
        if events is None:
            check(self.lib.mimo_unet_set_backward_events(self.handle, None), "mimo_unet_set_backward_events")
        else:
            arr = (C.c_void_p * 4)(*[e.cuda_event for e in events])
            check(self.lib.mimo_unet_set_backward_events(self.handle, arr), "mimo_unet_set_backward_events")

    def stage_first_state(self, stage: int) -> int:
        return int(self.lib.mimo_unet_backward_stage_first_state(self.handle, stage))

    @property
    def graph_state(self) -> int:
        """bit 0: forward body runs from a CUDA graph; bits 1..4: backward stages; bit 8: capture failed (eager for good)."""
        return int(self.lib.mimo_unet_graph_state(self.handle))

    @property
    def last_launches(self) -> int:
        return int(self.lib.mimo_unet_last_launches(self.handle))

    # -- test hook -----------------------------------------------------------------------------
    _STACKED = ("core.down2.in", "core.up3.in", "core.down2.c1.dpad", "core.up3.c1.dpad")

    def _logical_channels(self, name: str, t: torch.Tensor) -> torch.Tensor:
        """The buffers that start with the subnetwork stack keep its slices 8-channel aligned (mimo_unet_stack_layout): drop the gap
        channels so that callers see the reference's channel order."""
        if name not in self._STACKED:
            return t
        ln, st, n = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.mimo_unet_stack_layout(self.handle, C.byref(ln), C.byref(st), C.byref(n)), "mimo_unet_stack_layout")
        if n.value == 0:
            return t
        idx = [s * st.value + r for s in range(n.value) for r in range(ln.value)] + list(range(n.value * st.value, t.shape[1]))
        return t[:, torch.tensor(idx, device=t.device)].contiguous()

    def debug_tensor(self, name: str) -> torch.Tensor:
        """Returns an fp32 NCHW copy of a named intermediate (interior only) or an fp32 vector."""
        a = Act()
        kind = C.c_int()
        check(self.lib.mimo_unet_debug_view(self.handle, name.encode(), C.byref(a), C.byref(kind)), "mimo_unet_debug_view")
        base = self.workspace.data_ptr()
        off = a.ptr - base
        if kind.value == 1:
            return self.workspace[off: off + 4 * a.c].view(torch.float32).clone()
        # pad 1: halo all around (interior at (1,1)); pad 2: zero tail (interior at (0,0)); both are (h+2) x (w+2) buffers
        e, o = (2 if a.pad else 0), (1 if a.pad == 1 else 0)
        Hp, Wp = a.h + e, a.w + e
        nbytes = a.n * Hp * Wp * a.cpitch * 2
        t = self.workspace[off: off + nbytes].view(torch.bfloat16).view(a.n, Hp, Wp, a.cpitch)
        t = t[:, o: o + a.h, o: o + a.w, a.c_off: a.c_off + a.c]
        return self._logical_channels(name, t.permute(0, 3, 1, 2).float().contiguous())

    def debug_tensor_padded(self, name: str) -> torch.Tensor:
        """fp32 NCHW copy including the halo."""
        a = Act()
        kind = C.c_int()
        check(self.lib.mimo_unet_debug_view(self.handle, name.encode(), C.byref(a), C.byref(kind)), "mimo_unet_debug_view")
        off = a.ptr - self.workspace.data_ptr()
        e = 2 if a.pad else 0
        Hp, Wp = a.h + e, a.w + e
        t = self.workspace[off: off + a.n * Hp * Wp * a.cpitch * 2].view(torch.bfloat16).view(a.n, Hp, Wp, a.cpitch)
        return self._logical_channels(name, t[..., a.c_off: a.c_off + a.c].permute(0, 3, 1, 2).float().contiguous())
