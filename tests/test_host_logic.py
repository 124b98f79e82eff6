"""CPU tests of host-side logic: product module construction, loud failure without CUDA, row-view factorisation,
bucket planning and the world_size-2 gloo data-parallel exchange."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mimo_unet_b200 import MimoError
from mimo_unet_b200.functional import _common_rows_view
from mimo_unet_b200.parallel import GradientSynchronizer, allreduce_mean_
from oracle import mimo_oracle as O


def test_state_dict_layout_and_cpu_failure():
    from mimo.models.mimo_unet import MimoUnetModel
    m = MimoUnetModel(3, 2, 2, 8, 0, 0, 0, 0, 0, "laplace_nll", 0.0, 1e-3, 1, 10, 0.3)
    names = [n for n, _, _ in O.state_dict_spec(3, 2, 2, 8)]
    assert list(m.model.state_dict().keys()) == names
    assert m.hparams["trainable_params"] == sum(p.numel() for p in m.model.parameters())
    m.load_state_dict({k.replace("model.", "model._orig_mod."): v for k, v in m.state_dict().items()})  # compiled-checkpoint keys
    with pytest.raises(MimoError):
        m(torch.zeros(1, 2, 3, 32, 32))
    with pytest.raises(ValueError):
        from mimo.losses import UncertaintyLoss
        UncertaintyLoss.from_name("nope")
    assert m.compile() is None


def test_rows_view_factorisation():
    out = torch.zeros(3, 2, 2, 5, 7)
    p1, p2, y = out[:, :, :1], out[:, :, 1:], torch.zeros(3, 2, 1, 5, 7)
    ts, rows, cols, rs = _common_rows_view([p1, p2, y], p1.shape)
    assert (rows, cols, rs) == (6, 35, [70, 70, 35]) and ts[0].data_ptr() == p1.data_ptr()  # no copy of the strided views
    ts, rows, cols, rs = _common_rows_view([y, y], y.shape)
    assert (rows, cols) == (1, 210)


def test_bucket_ranges_cover_everything():
    for n in (1, 127, 128, 1000, 7383622):
        for k in (1, 4, 8):
            r = GradientSynchronizer.bucket_ranges(n, k)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:])) and len(r) <= k
            assert all(a % 128 == 0 for a, _ in r)


def _dp_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    S, f = 2, 4
    sd = O.make_state_dict(3, 2, S, f, seed=1)
    x_all, y_all = torch.rand(4, S, 3, 32, 32), torch.rand(4, S, 1, 32, 32)
    shard = slice(rank * 2, rank * 2 + 2)
    p = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.clone()) for k, v in sd.items()}
    out = O.mimo_unet_forward(x_all[shard], p, S, training=True)       # per-rank BatchNorm statistics
    loss, total = O.train_loss(out, y_all[shard], None, torch.ones(S))
    total.backward()
    learn = [v for v in p.values() if v.requires_grad]
    flat = torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).reshape(-1) for v in learn])
    sync = GradientSynchronizer(num_buckets=3)
    sync.start(flat)
    sync.wait()
    l = allreduce_mean_(loss.detach().clone())
    torch.save({"flat": flat, "loss": l}, os.path.join(tmp, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_data_parallel_exchange_gloo_world2(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["flat"], r1["flat"]) and torch.equal(r0["loss"], r1["loss"])
    # R-replica simulation of the oracle in one process (SURVEY 8e parity definition)
    torch.manual_seed(0)
    S, f = 2, 4
    sd = O.make_state_dict(3, 2, S, f, seed=1)
    x_all, y_all = torch.rand(4, S, 3, 32, 32), torch.rand(4, S, 1, 32, 32)
    flats, losses = [], []
    for r in range(2):
        p = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 and "running" not in k else v.clone()) for k, v in sd.items()}
        out = O.mimo_unet_forward(x_all[r * 2: r * 2 + 2], p, S, training=True)
        loss, total = O.train_loss(out, y_all[r * 2: r * 2 + 2], None, torch.ones(S))
        total.backward()
        flats.append(torch.cat([(v.grad if v.grad is not None else torch.zeros_like(v)).reshape(-1) for v in p.values() if v.requires_grad]))
        losses.append(loss.detach())
    assert torch.allclose(r0["flat"], (flats[0] + flats[1]) / 2, rtol=1e-5, atol=1e-8)
    assert torch.allclose(r0["loss"], (losses[0] + losses[1]) / 2, rtol=1e-6)


@pytest.mark.parametrize("cout,cin,positions,sms", [
    (336, 672, 64 * 18 * 22, 148),    # core.up1.c1 at C2
    (84, 168, 64 * 66 * 82, 148),     # core.up3.c1 at C2 (pair + single ci chunk)
    (336, 336, 64 * 10 * 12, 148),    # core.down4 at C2 (tiny map)
    (70, 100, 2 * 9 * 11, 148),       # fewer work units than SMs
    (130, 40, 3 * 8 * 12, 148),       # single ci chunk only: no two-chunk items
    (480, 960, 32 * 34 * 34, 148),    # C3 core (fbc 30)
    (65, 64, 7, 5),                   # one k-block
    (168, 84, 1428 * 64, 132),        # another SM count
])
def test_wgrad_streamk_schedule_covers_every_unit_once(cout, cin, positions, sms):
    """Host-side dump of the stream-K schedule of the weight-gradient kernel (conv_wgrad_flatk.cu): every
    (co tile, ci chunk, kh, k-block) unit is owned by exactly one CTA segment, and the CTAs' MMA counts are balanced."""
    import ctypes as C
    import numpy as np
    from mimo_unet_b200 import _lib
    lib = _lib.lib()
    grid = lib.mimo_wgrad_streamk_schedule(cout, cin, positions, sms, 0, None, 0)
    assert 1 <= grid <= sms
    n_kb = -(-positions // 128)
    co_tiles, ci_chunks = -(-cout // 128), -(-cin // 64)
    seen = np.zeros((co_tiles, ci_chunks, 3, n_kb), dtype=np.int32)
    weights = []
    buf = (C.c_int * (6 * 64))()
    for cta in range(grid):
        n = lib.mimo_wgrad_streamk_schedule(cout, cin, positions, sms, cta, buf, 64)
        assert 0 <= n <= 64
        w = 0
        for j in range(n):
            cot, cic0, nci, kh, kb0, kb1 = (buf[6 * j + k] for k in range(6))
            assert nci in (1, 2) and 0 <= kb0 < kb1 <= n_kb and 0 <= kh < 3
            seen[cot, cic0:cic0 + nci, kh, kb0:kb1] += 1
            w += nci * (kb1 - kb0)
        weights.append(w)
    assert int(seen.min()) == 1 and int(seen.max()) == 1
    total = co_tiles * ci_chunks * 3 * n_kb
    assert sum(weights) == total
    assert max(weights) - min(weights) <= 3  # equal shares up to rounding onto k-block edges


def test_bench_flop_accounting_matches_survey():
    """Numerator of bench.py's roofline: algorithmic 3x3-conv FLOPs with true channel counts: C2 forward is
    9.936 GFLOP/sample including the 1x1 heads (0.0034), dgrad skips the two image convolutions."""
    import importlib
    bench = importlib.import_module("bench")
    bench.CFG, bench.WORKLOAD = bench.WORKLOADS["C2"]
    fprop, dgrad = bench.conv_flops_per_step(64)
    heads = 2 * 2 * 21 * 2 * 128 * 160 * 64
    assert abs((fprop + heads) / 64 / 1e9 - 9.936) < 2e-3
    first = 2 * (2.0 * 128 * 160 * 21 * 3 * 9) * 64  # encoder.in_convs.{0,1} first conv: no input gradient in training
    assert abs(dgrad - (fprop - first)) / fprop < 1e-9
    assert abs((2 * fprop + dgrad + 3 * heads) / 64 / 1e9 - 29.762) < 2e-2


def test_bench_per_node_flops_follow_the_executor_node_order():
    """bench.py attributes per-launch times to layers through the executor's tag = 2 * node + conv: the FLOP table must follow
    the plan's node order (host-only plan creation, no device needed) and sum to the per-step total."""
    import ctypes as C
    import importlib
    from mimo_unet_b200 import _lib
    from mimo_unet_b200.engine import UnetConfig
    bench = importlib.import_module("bench")
    for wl in ("C2", "C3"):
        bench.CFG, bench.WORKLOAD = bench.WORKLOADS[wl]
        cfg = bench.CFG
        lib = _lib.lib()
        h = C.c_void_p()
        c = UnetConfig(cfg["in_channels"], cfg["out_channels"], cfg["num_subnetworks"], cfg["filter_base_count"], cfg["batch"],
                       cfg["height"], cfg["width"])
        _lib.check(lib.mimo_unet_plan_create(C.byref(c), C.byref(h)), "plan_create")
        try:
            n = lib.mimo_unet_num_double_convs(h)
            names = [lib.mimo_unet_node_name(h, i).decode() for i in range(n)]
        finally:
            lib.mimo_unet_plan_destroy(h)
        S = cfg["num_subnetworks"]
        expect = [f"encoder.in_convs.{s}" for s in range(S)] + [f"encoder.down1s.{s}" for s in range(S)] + \
                 ["core.down2", "core.down3", "core.down4", "core.up1", "core.up2", "core.up3"] + [f"decoder.up4s.{s}" for s in range(S)]
        assert names == expect
        node_fl = bench.conv_flops_per_node(cfg["batch"])
        assert len(node_fl) == 2 * n
        fprop, _ = bench.conv_flops_per_step(cfg["batch"])
        assert abs(sum(node_fl) - fprop) / fprop < 1e-12
        # first conv of encoder 0: Cin -> f at full resolution
        assert node_fl[0] == 2.0 * cfg["height"] * cfg["width"] * cfg["filter_base_count"] * cfg["in_channels"] * 9 * cfg["batch"]
    bench.CFG, bench.WORKLOAD = bench.WORKLOADS["C2"]


def test_runtime_grad_buffer_views_are_cached_and_rebuilt():
    """NetworkRuntime._ensure_grad_buffer: one flat fp32 buffer in state order; the per-parameter views are rebuilt only when the
    buffer is (re)created or the set of trainable tensors changes (host logic, CPU tensors)."""
    from mimo.models.mimo_components.model import MimoUNet
    from mimo_unet_b200.network import NetworkRuntime, _state_entries
    net = MimoUNet(3, 2, 2, 8)
    rt = NetworkRuntime(net)
    st = _state_entries(net)
    rt._ensure_grad_buffer(st)
    flat, views = rt.flat_grads, rt._grad_views
    n_learn = sum(1 for t in st if t.dtype == torch.float32 and t.requires_grad)
    assert flat.numel() == sum(t.numel() for t in st if t.dtype == torch.float32 and t.requires_grad)
    assert sum(v is not None for v in views) == n_learn == len(list(net.parameters()))
    off = 0
    for t, v in zip(st, views):  # contiguous slices in state order
        if v is not None:
            assert v.shape == t.shape and v.data_ptr() == flat.data_ptr() + 4 * off
            off += t.numel()
    rt._ensure_grad_buffer(st)
    assert rt.flat_grads is flat and rt._grad_views is views          # cached
    net.decoder.outcs[0].conv.bias.requires_grad_(False)              # the trainable set changed -> new buffer + views
    rt._ensure_grad_buffer(_state_entries(net))
    assert rt.flat_grads is not flat and rt._grad_views is not views
    assert rt.flat_grads.numel() == flat.numel() - 2
    rt._flat_grads = None                                            # private-buffer path of the backward
    rt._ensure_grad_buffer(_state_entries(net))
    assert rt.flat_grads is not None and rt._views_of is rt.flat_grads


def test_bench_shape_table_equals_the_oracle_table():
    """bench.py carries its own conv shape table (its GPU arm may not import oracle/): same rows as the oracle's."""
    import importlib
    from oracle import mimo_oracle as O
    bench = importlib.import_module("bench")
    for wl in ("C1", "C2", "C3"):
        bench.CFG, bench.WORKLOAD = bench.WORKLOADS[wl]
        cfg = bench.CFG
        ref = [r[:6] for r in O.conv_layer_table(cfg["in_channels"], cfg["num_subnetworks"], cfg["filter_base_count"], cfg["height"],
                                                  cfg["width"]) if r[6] == 3]
        assert bench.conv_shape_table() == ref
    bench.CFG, bench.WORKLOAD = bench.WORKLOADS["C2"]


REFERENCE_ROOT = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="reference checkout not present (GPU box)")
def test_reference_scripts_resolve_every_mimo_module_with_this_repo_first():
    """INTEGRATION.md option A: PYTHONPATH=<this repo>:<reference>. Every `mimo.*` module the reference's train/test scripts
    import must resolve -- the hot-path modules to THIS repo, the host I/O / logging glue (tasks, datasets, visualization) to
    the reference checkout through pkgutil.extend_path. Run in a subprocess so sys.path / sys.modules stay clean."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import importlib.util, os, sys
ours = {"mimo.utils", "mimo.losses", "mimo.metrics", "mimo.models.mimo_unet", "mimo.models.ensemble", "mimo.models.evidential_unet",
        "mimo.models.utils", "mimo.models.mimo_components.model", "mimo.models.mimo_components.components",
        "mimo.models.mimo_components.loss_buffer"}
theirs = {"mimo.visualization", "mimo.regularization", "mimo.datasets.nyuv2", "mimo.datasets.muad", "mimo.datasets.make3d",
          "mimo.tasks.depth.nyuv2_datamodule", "mimo.tasks.depth.callbacks", "mimo.tasks.sen12tp.sen12tp_datamodule",
          "mimo.tasks.sen12tp.callbacks"}
root, ref = sys.argv[1], sys.argv[2]
for name in sorted(ours | theirs):
    # walk the package chain by hand: find_spec imports the parents (ours: importable), not the leaf (may need lightning)
    spec = importlib.util.find_spec(name)
    assert spec is not None and spec.origin, name
    want = root if name in ours else ref
    assert os.path.abspath(spec.origin).startswith(want + os.sep), (name, spec.origin)
import mimo.losses, mimo.models.mimo_unet, mimo.models.ensemble, mimo.models.evidential_unet  # the ones scripts construct
from mimo.losses import UncertaintyLoss, EvidentialLoss
from mimo.utils import dir_path, count_trainable_parameters
print("ok")
"""
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + REFERENCE_ROOT)
    r = subprocess.run([sys.executable, "-c", code, root, REFERENCE_ROOT], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_subnetwork_stack_layout_is_8_channel_aligned():
    """The S subnetwork slices (2f channels each) at the head of the core's input / core.up3's concat buffer start at multiples of
    round_up(2f, 8) channels (host-only plan creation): no gaps when 2f is already a multiple of 8 or when there is one subnetwork."""
    import ctypes as C
    from mimo_unet_b200 import _lib
    from mimo_unet_b200.engine import UnetConfig
    lib = _lib.lib()
    for S, f, expect in ((2, 21, (42, 48, 2)), (4, 21, (42, 48, 4)), (2, 30, (60, 64, 2)), (2, 8, (0, 0, 0)), (1, 21, (0, 0, 0)),
                         (3, 13, (26, 32, 3))):
        h = C.c_void_p()
        c = UnetConfig(3, 2, S, f, 2, 32, 32)
        _lib.check(lib.mimo_unet_plan_create(C.byref(c), C.byref(h)), "plan_create")
        try:
            ln, st, n = C.c_int(), C.c_int(), C.c_int()
            _lib.check(lib.mimo_unet_stack_layout(h, C.byref(ln), C.byref(st), C.byref(n)), "stack_layout")
            assert (ln.value, st.value, n.value) == expect, (S, f)
        finally:
            lib.mimo_unet_plan_destroy(h)
