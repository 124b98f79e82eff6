"""Parity AT THE BENCHMARKED SHAPES (BASELINE.json configs C2, C3, C4), through the C ABI executor.

The kernel dispatch is shape dependent (flat / flatk / igemm fprop+dgrad, flat / stream-K wgrad, bulk-copy BN passes with
row segmenting, 148-CTA persistent grids), so the fixtures of tests/golden (B=2, 32x32) do not reach the code the
benchmark runs. Here every layer of the real configurations is checked "teacher-forced" (SURVEY 8c / App. F protocol): each
kernel's result is compared with the oracle op applied IN TRUE FP32 ON THE GPU to the executor's own stored inputs,
rounded to bf16 at the same storage point. Tolerance: rel-L2 <= 1e-3 per kernel (north star: bf16 compute / fp32
accumulate), fp32 vectors (BatchNorm statistics, dgamma/dbeta) <= 1e-4, max-pool routing bit-exact.
No environment override is used: the dispatch is the one bench.py measures.
Reference semantics: mimo/models/mimo_components/model.py:94-117, components.py:8-129.
"""
import pytest
import torch
import torch.nn.functional as F

from mimo_unet_b200.engine import UNetPlan
from oracle import mimo_oracle as O
from tests.util import bf16r, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-3

CONFIGS = {
    # BASELINE.json configs[1]: NYUv2 shape, the config the headline metric is quoted on
    "C2": dict(S=2, f=21, cin=3, B=64, H=128, W=160, training=True, drop=0.0, backward=True),
    # configs[2]: SEN12TP shape
    "C3": dict(S=2, f=30, cin=2, B=32, H=256, W=256, training=True, drop=0.0, backward=True),
    # configs[3]: M=4 inference (BatchNorm in eval mode), plain and with one MC-dropout pass (injected Dropout2d masks)
    "C4_eval": dict(S=4, f=21, cin=3, B=64, H=128, W=160, training=False, drop=0.0, backward=False),
    "C4_mc": dict(S=4, f=21, cin=3, B=64, H=128, W=160, training=False, drop=0.1, backward=False),
    # the same with the fused inference epilogues (mimo_unet_set_inference_fusion: eval BatchNorm + ReLU + Dropout2d applied in
    # the conv epilogue, written straight into the consumer's haloed buffer; what EnsembleModule.forward runs under no_grad)
    "C4_eval_fused": dict(S=4, f=21, cin=3, B=64, H=128, W=160, training=False, drop=0.0, backward=False, fused=True),
    "C4_mc_fused": dict(S=4, f=21, cin=3, B=64, H=128, W=160, training=False, drop=0.1, backward=False, fused=True),
    "C2_eval_fused": dict(S=2, f=21, cin=3, B=64, H=128, W=160, training=False, drop=0.1, backward=False, fused=True),
    # training with Dropout2d masks at the C2 shape (smaller batch: the eager path is the same code)
    "C2_drop": dict(S=2, f=21, cin=3, B=16, H=128, W=160, training=True, drop=0.1, backward=True),
}


def _nodes(S):
    enc_in = [(f"encoder.in_convs.{i}", f"encoder.in_convs.{i}.double_conv.") for i in range(S)]
    enc_dn = [(f"encoder.down1s.{i}", f"encoder.down1s.{i}.conv.double_conv.") for i in range(S)]
    core = [(f"core.{n}", f"core.{n}.conv.double_conv.") for n in ("down2", "down3", "down4", "up1", "up2", "up3")]
    dec = [(f"decoder.up4s.{i}", f"decoder.up4s.{i}.conv.double_conv.") for i in range(S)]
    return enc_in + enc_dn + core + dec   # canonical (state_dict) order == executor node order


def _fold(dpad):
    """adjoint of the reflect halo (F.pad reflect) in fp32"""
    N, C, Hp, Wp = dpad.shape
    t = torch.zeros(N, C, Hp - 2, Wp - 2, device=dpad.device, requires_grad=True)
    F.pad(t, (1, 1, 1, 1), mode="reflect").backward(dpad)
    return t.grad


def _pool_bwd(act, gp):
    t = act.clone().requires_grad_(True)
    F.max_pool2d(t, 2).backward(gp)
    return t.grad


def _up_adj(g, h, w):
    t = torch.zeros(g.shape[0], g.shape[1], h, w, device=g.device, requires_grad=True)
    O.pad_to(O.upsample_bilinear2x_ac(t), g.shape[2], g.shape[3]).backward(g)
    return t.grad


class _Checker:
    def __init__(self, tag):
        self.tag, self.worst, self.n = tag, {}, 0

    def __call__(self, what, got, ref, tol=TOL):
        e = rel_l2(got, ref)
        self.n += 1
        kind = what.split(":")[0]
        self.worst[kind] = max(self.worst.get(kind, 0.0), e)
        assert e <= tol, f"{self.tag} {what}: rel-L2 {e:.3e} > {tol:.1e}"

    def report(self):
        print(f"[{self.tag}] {self.n} checks; worst rel-L2 per kind: " + ", ".join(f"{k} {v:.2e}" for k, v in sorted(self.worst.items())))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_teacher_forced_fullshape(name):
    cfg = CONFIGS[name]
    S, f, cin, B, H, W = cfg["S"], cfg["f"], cfg["cin"], cfg["B"], cfg["H"], cfg["W"]
    training, p_drop = cfg["training"], cfg["drop"]
    dev = torch.device("cuda")
    torch.manual_seed(1)
    sd = {k: v.to(dev) for k, v in O.make_state_dict(cin, 2, S, f, seed=11).items()}
    if not training:  # non-trivial running statistics for the eval-mode affine
        for k in sd:
            if k.endswith("running_mean"):
                sd[k] = torch.randn_like(sd[k]) * 0.1
            elif k.endswith("running_var"):
                sd[k] = torch.rand_like(sd[k]) * 0.5 + 0.25
    names = [n for n, _, _ in O.state_dict_spec(cin, 2, S, f)]
    state = [sd[n].clone().contiguous() for n in names]
    grads = [torch.full_like(t, float("nan")) if t.dtype == torch.float32 and "running" not in n else None for n, t in zip(names, state)]
    G = dict(zip(names, grads))
    x = torch.rand(B, S, cin, H, W, device=dev)
    y_true = torch.rand(B, S, 1, H, W, device=dev)
    plan = UNetPlan(cin, 2, S, f, B, H, W, dev)
    plan.bind(state, grads)
    nodes = _nodes(S)
    assert plan.n_dconv == len(nodes)
    masks = None
    if p_drop > 0:
        masks = [((torch.rand(B * plan.drop_channels[i], device=dev) >= p_drop).float() / (1.0 - p_drop)).contiguous() for i in range(len(nodes))]
    out = torch.empty(B, S, 2, H, W, device=dev)
    fused = bool(cfg.get("fused"))
    plan.set_inference_fusion(fused)
    plan.forward(x, out, training, drop_masks=masks)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    chk = _Checker(name)
    dbg = plan.debug_tensor
    c = 2 * f * S

    def mask_of(i):
        return None if masks is None else masks[i].view(B, -1)

    def fused_layer(xin, w, pre, bi, ci, m):
        """conv -> eval BatchNorm -> ReLU -> Dropout2d with ONE bf16 rounding at the end (the fused epilogue works on the fp32 accumulator)"""
        gamma, beta, bias = sd[f"{pre}{bi}.weight"], sd[f"{pre}{bi}.bias"], sd[f"{pre}{ci}.bias"]
        z = O.batchnorm_eval(O.conv3x3_reflect(xin, w, None) + bias[None, :, None, None], gamma, beta, sd[f"{pre}{bi}.running_mean"],
                             sd[f"{pre}{bi}.running_var"])
        h = F.relu(z)
        if m is not None:
            h = h * m[:, :, None, None]
        return bf16r(h)

    def bn_relu(y, pre, bi, ci, m, check_stats):
        gamma, beta, bias = sd[f"{pre}{bi}.weight"], sd[f"{pre}{bi}.bias"], sd[f"{pre}{ci}.bias"]
        if training:
            z, mean, var = O.batchnorm_train(y, gamma, beta)
            if check_stats:
                node, cn = check_stats
                chk(f"bn_stats:{node}.{cn}.mean", dbg(f"{node}.{cn}.mean"), mean, 1e-4)
                chk(f"bn_stats:{node}.{cn}.invstd", dbg(f"{node}.{cn}.invstd"), torch.rsqrt(var + 1e-5), 1e-4)
                n = y.shape[0] * y.shape[2] * y.shape[3]
                rm, rv = O.updated_running_stats(sd[f"{pre}{bi}.running_mean"], sd[f"{pre}{bi}.running_var"], mean + bias, var, n)
                chk(f"bn_running:{pre}{bi}.running_mean", state[names.index(f"{pre}{bi}.running_mean")], rm, 1e-4)
                chk(f"bn_running:{pre}{bi}.running_var", state[names.index(f"{pre}{bi}.running_var")], rv, 1e-4)
                assert int(state[names.index(f"{pre}{bi}.num_batches_tracked")]) == int(sd[f"{pre}{bi}.num_batches_tracked"]) + 1
        else:
            z = O.batchnorm_eval(y + bias[None, :, None, None], gamma, beta, sd[f"{pre}{bi}.running_mean"], sd[f"{pre}{bi}.running_var"])
        h = F.relu(z)
        if m is not None:
            h = h * m[:, :, None, None]
        return bf16r(h)

    # ------------------------------------------------------------------ forward, layer by layer
    for i, (nd, pre) in enumerate(nodes):
        xin, a1, o = dbg(nd + ".in"), dbg(nd + ".a1"), dbg(nd + ".out")
        virtual = fused and nd.startswith("decoder") and f + c // 2 > 64
        if virtual:
            # inference with > 64 concat channels: the decoders read cat([x1_s, up(u3)]) as a VIRTUAL concat (the up-sampled
            # core output exists once, not once per dcat buffer): rebuild the conv input from its two sources
            up = bf16r(O.pad_to(O.upsample_bilinear2x_ac(dbg("core.up3.out")), xin.shape[2], xin.shape[3]))
            xin = torch.cat([xin[:, :f], up], dim=1)
        w1, w2 = bf16r(sd[pre + "0.weight"]), bf16r(sd[pre + "3.weight"])
        # the M per-subnetwork encoders' second conv may write an unaligned concat slice: those layers keep the unfused path
        if fused:
            chk(f"fused_layer:{nd}.c1", a1, fused_layer(xin, w1, pre, 1, 0, None))
        else:
            y1 = dbg(nd + ".c1.y")
            chk(f"fprop:{nd}.c1", y1, bf16r(O.conv3x3_reflect(xin, w1, None)))
            chk(f"bn_apply:{nd}.c1", a1, bn_relu(y1, pre, 1, 0, None, (nd, "c1")))
        ref_fused = fused_layer(a1, w2, pre, 4, 3, mask_of(i)) if fused else None
        y2 = dbg(nd + ".c2.y")
        ref_unfused = bn_relu(bf16r(O.conv3x3_reflect(a1, w2, None)), pre, 4, 3, mask_of(i), None) if fused else None
        if fused:
            # either path is legal for c2 (alignment of the destination slice decides): accept the closer one
            e_f, e_u = rel_l2(o, ref_fused), rel_l2(o, ref_unfused)
            chk(f"fused_layer:{nd}.c2", o, ref_fused if e_f <= e_u else ref_unfused)
        else:
            chk(f"fprop:{nd}.c2", y2, bf16r(O.conv3x3_reflect(a1, w2, None)))
            chk(f"bn_apply:{nd}.c2", o, bn_relu(y2, pre, 4, 3, mask_of(i), (nd, "c2")))
        # the producers write the reflect halo of every conv input
        for buf in (".in", ".a1"):
            xp = plan.debug_tensor_padded(nd + buf)
            if virtual and buf == ".in":   # only the skip slice of dcat is written
                xp = xp[:, :f]
                assert torch.equal(xp, F.pad(dbg(nd + buf)[:, :f], (1, 1, 1, 1), mode="reflect")), f"{nd}{buf} halo"
                continue
            assert torch.equal(xp, F.pad(dbg(nd + buf), (1, 1, 1, 1), mode="reflect")), f"{nd}{buf} halo"
        if "in_convs" in nd or "down" in nd and "down4" not in nd:
            assert torch.equal(dbg(nd + ".pool"), F.max_pool2d(o, 2)), f"{nd} pool"   # bit exact
    # bilinear up-sampling (align_corners) into the concat slices
    ups = [("core.up1", "core.down4", 4 * c), ("core.up2", "core.up1", 2 * c), ("core.up3", "core.up2", c)] + \
          [(f"decoder.up4s.{s}", "core.up3", f) for s in range(S)]
    for nd, src, skip_c in ups:
        if fused and nd.startswith("decoder") and f + c // 2 > 64:
            continue   # virtual concat: checked through the decoder's first conv above
        xin = dbg(nd + ".in")
        ref = bf16r(O.pad_to(O.upsample_bilinear2x_ac(dbg(src + ".out")), xin.shape[2], xin.shape[3]))
        chk(f"upsample:{nd}", xin[:, skip_c:], ref)
    # 1x1 heads (fp32 math on the stored bf16 features)
    for s in range(S):
        feat = dbg(f"decoder.up4s.{s}.out")
        ref = F.conv2d(feat, sd[f"decoder.outcs.{s}.conv.weight"], sd[f"decoder.outcs.{s}.conv.bias"])
        chk(f"head:{s}", out[:, s], ref, 1e-5)
    if not cfg["backward"]:
        chk.report()
        return

    # ------------------------------------------------------------------ backward, layer by layer
    o_leaf = out.clone().requires_grad_(True)
    w_sub = torch.softmax(torch.arange(S, dtype=torch.float32, device=dev) * 0.3, 0) * S
    (O.laplace_nll_elementwise(o_leaf[:, :, :1], o_leaf[:, :, 1:], y_true).mean(dim=(0, 2, 3, 4)) * w_sub).mean().backward()
    dout = o_leaf.grad.contiguous()
    plan.backward(dout)
    torch.cuda.synchronize()

    def bn_relu_bwd(y, Gup, pre, bi, m):
        # fp64: the checker must be more accurate than the 2e-4 bar on sums over up to 2M values per channel
        yq = y.double().requires_grad_(True)
        gq, bq = sd[f"{pre}{bi}.weight"].double().requires_grad_(True), sd[f"{pre}{bi}.bias"].double().requires_grad_(True)
        z = O.batchnorm_train(yq, gq, bq)[0]
        h = F.relu(z)
        if m is not None:
            h = h * m[:, :, None, None].double()
        (h * Gup.double()).sum().backward()
        return yq.grad.float(), gq.grad.float(), bq.grad.float()

    def conv_wgrad(xin, dy, cout, cin_):
        w = torch.zeros(cout, cin_, 3, 3, device=dev, requires_grad=True)
        O.conv3x3_reflect(xin, w, None).backward(dy)
        return w.grad

    for i, (nd, pre) in enumerate(nodes):
        xin, a1, y1, y2 = dbg(nd + ".in"), dbg(nd + ".a1"), dbg(nd + ".c1.y"), dbg(nd + ".c2.y")
        w1, w2 = bf16r(sd[pre + "0.weight"]), bf16r(sd[pre + "3.weight"])
        # second conv: upstream gradient g2 as stored by its producer
        g2 = dbg(nd + ".g2")
        dy2 = dbg(nd + ".c2.dy")
        r_dy, r_dg, r_db = bn_relu_bwd(y2, g2, pre, 4, mask_of(i))
        chk(f"bn_bwd:{nd}.c2", dy2, bf16r(r_dy))
        chk(f"bn_dgamma:{nd}.c2", G[pre + "4.weight"], r_dg, 2e-4)
        chk(f"bn_dbeta:{nd}.c2", G[pre + "4.bias"], r_db, 2e-4)
        assert float(G[pre + "3.bias"].abs().max()) == 0.0   # conv bias under train-mode BN: exactly zero gradient
        chk(f"wgrad:{nd}.c2", G[pre + "3.weight"], conv_wgrad(a1, dy2, w2.shape[0], w2.shape[1]))
        dpad2 = dbg(nd + ".c2.dpad")
        chk(f"dgrad:{nd}.c2", dpad2, bf16r(F.conv_transpose2d(dy2, w2)))
        # first conv: upstream gradient = fold_reflect(dpad2), formed in fp32 inside the fused BN-backward passes
        g1 = _fold(dpad2)
        Cm, Wn = dpad2.shape[1], dpad2.shape[3] - 2
        cp = (Cm + 7) // 8 * 8
        if not (2 * (Wn + 2) * cp * 2 <= 50 * 1024 and Wn * cp * 2 <= 16 * 1024):
            g1 = bf16r(g1)   # rows too long for the fused kernel's stage: the fold is materialised in bf16 (one more rounding)
        dy1 = dbg(nd + ".c1.dy")
        r_dy, r_dg, r_db = bn_relu_bwd(y1, g1, pre, 1, None)
        chk(f"bn_bwd:{nd}.c1", dy1, bf16r(r_dy))
        chk(f"bn_dgamma:{nd}.c1", G[pre + "1.weight"], r_dg, 2e-4)
        chk(f"bn_dbeta:{nd}.c1", G[pre + "1.bias"], r_db, 2e-4)
        chk(f"wgrad:{nd}.c1", G[pre + "0.weight"], conv_wgrad(xin, dy1, w1.shape[0], w1.shape[1]))
        if "in_convs" not in nd:
            chk(f"dgrad:{nd}.c1", dbg(nd + ".c1.dpad"), bf16r(F.conv_transpose2d(dy1, w1)))
    # heads: feature gradient + parameter gradients
    for s in range(S):
        feat = dbg(f"decoder.up4s.{s}.out")
        wh = sd[f"decoder.outcs.{s}.conv.weight"]
        chk(f"head_bwd:{s}.dfeat", dbg(f"decoder.up4s.{s}.g2"), bf16r(F.conv_transpose2d(dout[:, s], wh)))
        chk(f"head_bwd:{s}.dw", G[f"decoder.outcs.{s}.conv.weight"], torch.einsum("nkhw,nchw->kc", dout[:, s], feat)[:, :, None, None], 1e-4)
        chk(f"head_bwd:{s}.db", G[f"decoder.outcs.{s}.conv.bias"], dout[:, s].sum(dim=(0, 2, 3)), 1e-4)
    # gradient plumbing between the DoubleConvs: reflect-halo fold, concat split, bilinear adjoint, max-pool routing.
    # (two bf16 storage points on these paths: tolerance 2e-3)
    T2 = 2e-3
    t0 = None
    for s in range(S):
        part = _fold(dbg(f"decoder.up4s.{s}.c1.dpad")[:, f:])
        t0 = bf16r(part) if t0 is None else bf16r(t0 + part)
    Hh, Wh = H // 2, W // 2
    chk("glue:core.up3.g2", dbg("core.up3.g2"), bf16r(_up_adj(t0, Hh, Wh)), T2)
    chk("glue:core.up2.g2", dbg("core.up2.g2"), bf16r(_up_adj(bf16r(_fold(dbg("core.up3.c1.dpad")[:, c:])), Hh // 2, Wh // 2)), T2)
    chk("glue:core.up1.g2", dbg("core.up1.g2"), bf16r(_up_adj(bf16r(_fold(dbg("core.up2.c1.dpad")[:, 2 * c:])), Hh // 4, Wh // 4)), T2)
    chk("glue:core.down4.g2", dbg("core.down4.g2"), bf16r(_up_adj(bf16r(_fold(dbg("core.up1.c1.dpad")[:, 4 * c:])), Hh // 8, Wh // 8)), T2)
    for nd, skip_src, skip_c, pooled in (("core.down3", "core.up1", 4 * c, "core.down4"), ("core.down2", "core.up2", 2 * c, "core.down3")):
        gp = bf16r(_fold(dbg(pooled + ".c1.dpad")))
        ref = _fold(dbg(skip_src + ".c1.dpad")[:, :skip_c]) + _pool_bwd(dbg(nd + ".out"), gp)
        chk(f"glue:{nd}.g2", dbg(nd + ".g2"), bf16r(ref), T2)
    for s in range(S):
        gp = bf16r(_fold(dbg("core.down2.c1.dpad")[:, 2 * f * s: 2 * f * (s + 1)]))
        ref = _fold(dbg("core.up3.c1.dpad")[:, 2 * f * s: 2 * f * (s + 1)]) + _pool_bwd(dbg(f"encoder.down1s.{s}.out"), gp)
        chk(f"glue:encoder.down1s.{s}.g2", dbg(f"encoder.down1s.{s}.g2"), bf16r(ref), T2)
        gp = bf16r(_fold(dbg(f"encoder.down1s.{s}.c1.dpad")))
        ref = _fold(dbg(f"decoder.up4s.{s}.c1.dpad")[:, :f]) + _pool_bwd(dbg(f"encoder.in_convs.{s}.out"), gp)
        chk(f"glue:encoder.in_convs.{s}.g2", dbg(f"encoder.in_convs.{s}.g2"), bf16r(ref), T2)
    chk.report()


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_end_to_end_loss_and_graph_replay_fullshape(name):
    """Whole-network train-mode forward at the benchmarked shape against the bf16-emulating oracle run end to end on the
    GPU in fp32 (no TF32): per-subnetwork Laplace NLL rel <= 1e-3 (the robust end-to-end quantity), then
    the CUDA-graph replay of the same step (what bench.py times) reproduces the eager launches."""
    cfg = CONFIGS[name]
    S, f, cin, B, H, W = cfg["S"], cfg["f"], cfg["cin"], cfg["B"], cfg["H"], cfg["W"]
    dev = torch.device("cuda")
    torch.manual_seed(2)
    sd = {k: v.to(dev) for k, v in O.make_state_dict(cin, 2, S, f, seed=12).items()}
    names = [n for n, _, _ in O.state_dict_spec(cin, 2, S, f)]
    state = [sd[n].clone().contiguous() for n in names]
    grads = [torch.zeros_like(t) if t.dtype == torch.float32 and "running" not in n else None for n, t in zip(names, state)]
    x = torch.rand(B, S, cin, H, W, device=dev)
    y_true = torch.rand(B, S, 1, H, W, device=dev)
    plan = UNetPlan(cin, 2, S, f, B, H, W, dev)
    plan.bind(state, grads)
    with torch.no_grad():
        ref = O.mimo_unet_forward(x, sd, S, training=True, emulate_bf16=True)
        loss_ref = O.laplace_nll_elementwise(ref[:, :, :1], ref[:, :, 1:], y_true).mean(dim=(0, 2, 3, 4))
    dout = (torch.randn(B, S, 2, H, W, device=dev) * 1e-3).contiguous()
    outs, gsets = [], []
    for it in range(4):
        out = torch.empty(B, S, 2, H, W, device=dev)
        plan.forward(x, out, True)
        plan.backward(dout)
        torch.cuda.synchronize()
        outs.append(out)
        gsets.append([g.clone() for g in grads if g is not None])
    loss = O.laplace_nll_elementwise(outs[0][:, :, :1], outs[0][:, :, 1:], y_true).mean(dim=(0, 2, 3, 4))
    e_loss, e_out = rel_l2(loss, loss_ref), rel_l2(outs[0], ref)
    print(f"[{name}] end to end vs bf16 oracle: loss rel {e_loss:.3e}, output rel-L2 {e_out:.3e}")
    assert e_loss <= 1e-3
    assert e_out <= 5e-2   # ill-conditioned end to end (App. F); the per-layer gates above are the 1e-3 ones
    assert plan.graph_state == 0x1F, f"graphs expected after two eager calls, got {plan.graph_state:#x}"
    assert torch.equal(outs[0], outs[3])
    for a, b in zip(gsets[0], gsets[3]):
        assert torch.isfinite(b).all()
        assert rel_l2(b, a) <= 1e-5   # fp32 atomics in the weight gradients: summation order may differ
