"""NetworkRuntime: glue between the ``mimo.models...MimoUNet`` nn.Module (parameter owner, reference state_dict
layout) and the C++ executor.  Owns the per-shape plans, the flat gradient buffer and the autograd Function."""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Optional

import torch

from . import _lib
from .engine import UNetPlan


def _state_entries(net) -> List[torch.Tensor]:
    """Tensors in MimoUNet.state_dict() order: per conv (weight, bias, bn.weight, bn.bias, running_mean,
    running_var, num_batches_tracked), then per head (weight, bias)."""
    out = []
    for dc in net.double_convs():
        seq = dc.double_conv
        for ci, bi in ((0, 1), (3, 4)):
            conv, bn = seq[ci], seq[bi]
            out += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked]
    for oc in net.decoder.outcs:
        out += [oc.conv.weight, oc.conv.bias]
    return out


class _UNetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rt: "NetworkRuntime", x: torch.Tensor, gather, *params):
        out, plan = rt._forward_impl(x, gather)
        ctx.rt, ctx.plan = rt, plan
        ctx.x_shape = tuple(x.shape)
        ctx.need_dx = x.requires_grad
        ctx.n_params = len(params)
        ctx.token = rt._forward_token
        return out

    @staticmethod
    def backward(ctx, dout):
        rt: NetworkRuntime = ctx.rt
        if ctx.token != rt._forward_token:
            raise _lib.MimoError("MimoUNet backward called after another forward pass reused the executor's activation "
                                 "workspace; run backward before the next forward (one in-flight graph per module)")
        grads, dx = rt._backward_impl(ctx.plan, dout, ctx.need_dx, ctx.x_shape)
        return (None, dx, None) + tuple(grads)


class NetworkRuntime:
    MAX_PLANS = 3

    def __init__(self, net):
        self.net = net
        self.plans: "OrderedDict[tuple, UNetPlan]" = OrderedDict()
        self._forward_token = 0
        self._flat_grads: Optional[torch.Tensor] = None
        self._grad_views: List[Optional[torch.Tensor]] = []
        self._views_of: Optional[torch.Tensor] = None  # the flat buffer the cached views slice
        self._learnable_idx: List[int] = []
        self.last_launches = (0, 0)
        self.grad_sync = None  # OverlappedGradientSynchronizer when data-parallel training is enabled

    def enable_overlapped_allreduce(self, group=None):
        """Data-parallel training: average the flat gradient buffer over ranks, bucket by bucket, overlapped with
        backward. Call `grad_sync.wait()` before the optimizer step."""
        from .parallel import OverlappedGradientSynchronizer
        self.grad_sync = OverlappedGradientSynchronizer(group)
        return self.grad_sync

    def _stage_bounds(self, plan: UNetPlan, state):
        """Element ranges of the four backward stages inside the flat gradient buffer (state order)."""
        offs, off = {}, 0
        for i in self._learnable_idx:
            offs[i] = off
            off += state[i].numel()
        total = off

        def flat_off(state_index):
            for i in self._learnable_idx:
                if i >= state_index:
                    return offs[i]
            return total
        first = [flat_off(plan.stage_first_state(k)) for k in range(4)]   # stage 0 starts last
        return [(first[0], total), (first[1], first[0]), (first[2], first[1]), (first[3], first[2])]

    def __deepcopy__(self, memo):  # plans hold device workspaces and C handles: never copied with the module
        return None

    def __getstate__(self):
        return None

    # ------------------------------------------------------------------------------------------
    def _plan_for(self, B, H, W, device) -> UNetPlan:
        net = self.net
        key = (B, H, W, device.index)
        plan = self.plans.get(key)
        if plan is None:
            plan = UNetPlan(net.in_channels, net.out_channels, net.num_subnetworks, net.filter_base_count, B, H, W, device)
            self.plans[key] = plan
            while len(self.plans) > self.MAX_PLANS:
                self.plans.popitem(last=False)
        else:
            self.plans.move_to_end(key)
        return plan

    def _ensure_grad_buffer(self, state: List[torch.Tensor]):
        learn = [i for i, t in enumerate(state) if t.dtype == torch.float32 and t.requires_grad]
        total = sum(state[i].numel() for i in learn)
        dev = state[0].device
        if self._flat_grads is None or self._flat_grads.numel() != total or self._flat_grads.device != dev or learn != self._learnable_idx:
            self._flat_grads = torch.zeros(total, dtype=torch.float32, device=dev)
            self._learnable_idx = learn
        elif len(self._grad_views) == len(state) and self._views_of is self._flat_grads:
            return  # same buffer, same layout: the cached views (they only feed data_ptr() to bind) are still right
        self._views_of = self._flat_grads
        views: List[Optional[torch.Tensor]] = [None] * len(state)
        off = 0
        for i in learn:
            n = state[i].numel()
            views[i] = self._flat_grads[off: off + n].view(state[i].shape)
            off += n
        self._grad_views = views

    def _assert_grads_aliased(self):
        """Run by grad_sync.wait(): the overlapped all-reduce averages the flat buffer IN PLACE on a side stream, which is only
        correct if every parameter's .grad is a view of that buffer (autograd adopts the returned views when .grad was None).
        If autograd cloned or accumulated instead (a pre-existing non-aliased .grad, hooks), the optimizer would step on
        un-reduced gradients and the ranks would diverge silently: fail loudly instead."""
        if self._flat_grads is None:
            return
        lo = self._flat_grads.data_ptr()
        hi = lo + self._flat_grads.numel() * 4
        state = _state_entries(self.net)
        bad = [i for i in self._learnable_idx if state[i].grad is not None and not (lo <= state[i].grad.data_ptr() < hi)]
        if bad:
            raise _lib.MimoError(f"data-parallel gradient sync: {len(bad)} parameter gradients are not views of the flat gradient buffer "
                                 "(a .grad tensor existed before backward and autograd accumulated into it). Call "
                                 "optimizer.zero_grad(set_to_none=True) before the step (or keep the .grad tensors returned by this "
                                 "module), otherwise the all-reduced values never reach the optimizer.")

    @property
    def flat_grads(self) -> Optional[torch.Tensor]:
        """One contiguous fp32 tensor holding every parameter gradient (bucket for the data-parallel all-reduce)."""
        return self._flat_grads

    # ------------------------------------------------------------------------------------------
    def _dropout_masks(self, plan: UNetPlan, B: int, device):
        dcs = self.net.double_convs()
        active = [dc.dropout.training and dc.dropout.p > 0.0 for dc in dcs]
        if not any(active):
            return None
        # ONE random draw for all blocks (Dropout2d: whole (n, c) planes dropped, survivors scaled by 1/(1-p))
        sizes = [B * plan.drop_channels[i] if a else 0 for i, a in enumerate(active)]
        u = torch.rand(sum(sizes), device=device)
        masks, off = [], 0
        for i, dc in enumerate(dcs):
            if not active[i]:
                masks.append(None)
                continue
            p = dc.dropout.p
            m = u[off: off + sizes[i]]
            off += sizes[i]
            masks.append(((m >= p).float() / (1.0 - p)).contiguous())
        return masks

    def _elementwise_dropout(self, plan: UNetPlan, B: int, H: int, W: int, device):
        """nn.Dropout at the core centre (reference model.py:239) and in front of every head (model.py:294): bf16 0/1 keep
        masks in the executor's NHWC layout, drawn with the CUDA generator; the 1/(1-p) scale is applied in fp32 in-kernel."""
        net = self.net
        S, f = net.num_subnetworks, net.filter_base_count

        def keep(p, h, w, c):
            cp = (c + 7) // 8 * 8
            return (torch.rand(B, h, w, cp, device=device) >= p).to(torch.bfloat16).contiguous()

        cd = net.core.center_dropout
        center, cs = None, 1.0
        if cd.training and cd.p > 0.0:
            center, cs = keep(cd.p, H // 16, W // 16, 8 * f * S), 1.0 / (1.0 - cd.p) if cd.p < 1.0 else 0.0
        finals, fs, any_final = [], 1.0, False
        for d in net.decoder.final_dropouts:
            if d.training and d.p > 0.0:
                finals.append(keep(d.p, H, W, f))
                fs = 1.0 / (1.0 - d.p) if d.p < 1.0 else 0.0
                any_final = True
            else:
                finals.append(None)
        plan.set_elementwise_dropout(center, cs, finals if any_final else None, fs)

    def __call__(self, x: torch.Tensor, gather: Optional[torch.Tensor] = None):
        net = self.net
        if not x.is_cuda:
            raise _lib.MimoError("MimoUNet (B200) needs CUDA tensors: the forward/backward path is hand-written sm_100a "
                                 "CUDA and there is no CPU fallback. Move the module and inputs to a B200 device.")
        if gather is None:
            if x.dim() != 5 or x.shape[1] != net.num_subnetworks or x.shape[2] != net.in_channels:
                raise ValueError(f"expected x of shape [B, {net.num_subnetworks}, {net.in_channels}, H, W], got {tuple(x.shape)}")
        else:
            if x.dim() != 4 or x.shape[1] != net.in_channels or gather.dim() != 2 or gather.shape[0] != net.num_subnetworks:
                raise ValueError("with gather, x must be [B, C_in, H, W] and gather int64 [S, B * batch_repetitions]")
            if x.requires_grad:
                raise NotImplementedError("input gradients are not available together with a gather table")
        params = [p for p in net.parameters()]
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            return _UNetFunction.apply(self, x, gather, *params)
        out, _ = self._forward_impl(x, gather, inference=True)   # nothing records a graph: fused inference epilogues
        return out

    # ------------------------------------------------------------------------------------------
    def _forward_impl(self, x: torch.Tensor, gather, inference: bool = False):
        net = self.net
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        B, H, W = (x.shape[0] if gather is None else gather.shape[1]), x.shape[-2], x.shape[-1]
        dev = x.device
        plan = self._plan_for(B, H, W, dev)
        state = _state_entries(net)
        for t in state:
            if t.device != dev:
                raise _lib.MimoError("MimoUNet parameters and input live on different devices")
        self._ensure_grad_buffer(state)
        plan.bind([t.detach() for t in state], self._grad_views)
        out = torch.empty((B, net.num_subnetworks, net.out_channels, H, W), dtype=torch.float32, device=dev)
        # BatchNorm mode follows the BN modules (all share the module's mode; MC-dropout keeps BN in eval)
        bn_training = net.encoder.in_convs[0].double_conv[1].training
        masks = self._dropout_masks(plan, B, dev)
        with torch.cuda.device(dev):
            self._elementwise_dropout(plan, B, H, W, dev)
        g = None if gather is None else gather.to(device=dev, dtype=torch.int64).contiguous()
        with torch.cuda.device(dev):   # streams and launches follow the TENSORS' device, not whatever device is current
            plan.set_inference_fusion(inference and not bn_training)
            plan.forward(x, out, bn_training, gather=g, drop_masks=masks)
        self._forward_token += 1
        self.last_launches = (plan.last_launches, self.last_launches[1])
        return out, plan

    def _backward_impl(self, plan: UNetPlan, dout: torch.Tensor, need_dx: bool, x_shape):
        with torch.cuda.device(dout.device):
            return self._backward_on_device(plan, dout, need_dx, x_shape)

    def _backward_on_device(self, plan: UNetPlan, dout: torch.Tensor, need_dx: bool, x_shape):
        state = _state_entries(self.net)
        dout = dout.contiguous().float()
        dx = torch.empty(x_shape, dtype=torch.float32, device=dout.device) if need_dx else None
        lo = self._flat_grads.data_ptr()
        hi = lo + self._flat_grads.numel() * 4
        aliased = [i for i in self._learnable_idx if state[i].grad is not None and lo <= state[i].grad.data_ptr() < hi]
        sync = self.grad_sync
        if aliased and len(aliased) == len(self._learnable_idx):
            # the user kept .grad from the previous step (gradient accumulation, or zero_grad(set_to_none=False)): add in place,
            # hand nothing to autograd. Data parallel: the accumulated buffer is averaged again -- the part that was already
            # averaged is identical on every rank, so mean(previous mean + local) = previous mean + mean(local).
            plan.set_backward_events(sync.events if sync is not None else None)
            plan.backward(dout, dx=dx, accumulate=True)
            if sync is not None:
                sync.launch(self._flat_grads, self._stage_bounds(plan, state))
            grads = [None] * len(state)
        else:
            if aliased:  # mixed case: do not clobber live .grad tensors, use a private buffer for this pass
                self._flat_grads = None
                self._ensure_grad_buffer(state)
                plan._bound_sig = None
                plan.bind([t.detach() for t in state], self._grad_views)
            plan.set_backward_events(sync.events if sync is not None else None)
            plan.backward(dout, dx=dx, accumulate=False)
            if sync is not None:
                sync.launch(self._flat_grads, self._stage_bounds(plan, state))
                sync._after_wait = self._assert_grads_aliased
            # fresh view objects so autograd can adopt them as .grad without a copy
            grads, off = [None] * len(state), 0
            for i in self._learnable_idx:
                n = state[i].numel()
                grads[i] = self._flat_grads[off: off + n].view(state[i].shape)
                off += n
        self.last_launches = (self.last_launches[0], plan.last_launches)
        # autograd expects one entry per parameter passed to apply() == net.parameters() order
        by_id = {id(t): g for t, g in zip(state, grads)}
        return [by_id.get(id(p)) for p in self.net.parameters()], dx
