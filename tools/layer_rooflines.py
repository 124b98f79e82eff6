"""Per-layer roofline table: joins the measured per-layer, per-class times of tools/profile_layers.py (CUDA events on a B200)
with each layer's algorithmic floors -- tensor (FLOPs / sustained bf16 peak) and HBM (compulsory bytes / measured copy
bandwidth) -- and prints which floor binds and the fraction of it that was achieved. Needs no GPU.
Usage: python tools/layer_rooflines.py profiles/r01_layers_c2_final.txt [batch] > profiles/r01_layer_rooflines.txt"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def p8(c):
    return (c + 7) // 8 * 8


def main():
    path = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    peaks = {"bf16_tflops_sustained": 1372.5, "hbm_gbs": 6557.1}
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(mp):
        peaks.update(json.load(open(mp)))
    tf, bw = peaks["bf16_tflops_sustained"] * 1e12, peaks["hbm_gbs"] * 1e9
    print(f"# floors: tensor = FLOPs / {tf / 1e12:.1f} TF/s (sustained bf16), HBM = compulsory bytes / {bw / 1e9:.0f} GB/s; batch {B}")
    print("# conv bytes: read the haloed input once + write the output once (bf16, channel pitch rounded up to 8); wgrad: read dY and X once;")
    print("# bn_apply: read y + write act; bn_bwd: 2 x (read G, y) + write dy.  eff = binding floor / measured")
    hdr = f"{'layer':26s} {'cin':>4s} {'cout':>4s} {'HxW':>8s} | " + " | ".join(f"{n:>6s} {'floor':>6s} {'bnd':>3s} {'eff':>5s}" for n in ("fprop", "dgrad", "wgrad", "bnapp", "bnbwd"))
    print(hdr)
    tot_meas, tot_floor = 0.0, 0.0
    for line in open(path):
        if line.startswith("#") or line.startswith("layer") or "|" not in line:
            continue
        left, mid, _ = line.split("|")
        name, cin, cout, hw, _gf = left.split()
        cin, cout = int(cin), int(cout)
        H, W = (int(v) for v in hw.split("x"))
        fprop, dgrad, wgrad, _bnf, bnapp, bnbwd, _gather = (float(v) for v in mid.split())
        fl = 2.0 * B * H * W * cin * cout * 9
        in_b = B * (H + 2) * (W + 2) * p8(cin) * 2
        out_b = B * H * W * p8(cout) * 2
        dy_b = B * (H + 2) * (W + 2) * p8(cout) * 2
        floors = {
            "fprop": (fl / tf, (in_b + out_b) / bw),
            "dgrad": (fl / tf, (dy_b + B * (H + 2) * (W + 2) * p8(cin) * 2) / bw),
            "wgrad": (fl / tf, (dy_b + in_b) / bw),
            "bnapp": (0.0, 2 * out_b / bw),
            "bnbwd": (0.0, 5 * out_b / bw),
        }
        meas = {"fprop": fprop, "dgrad": dgrad, "wgrad": wgrad, "bnapp": bnapp, "bnbwd": bnbwd}
        cells = []
        for k in ("fprop", "dgrad", "wgrad", "bnapp", "bnbwd"):
            t_t, t_h = floors[k]
            fl_us = max(t_t, t_h) * 1e6
            m = meas[k]
            if m <= 0:
                cells.append(f"{'-':>6s} {'-':>6s} {'-':>3s} {'-':>5s}")
                continue
            tot_meas += m
            tot_floor += fl_us
            cells.append(f"{m:6.1f} {fl_us:6.1f} {'TC' if t_t >= t_h else 'HBM':>3s} {fl_us / m:5.2f}")
        print(f"{name:26s} {cin:4d} {cout:4d} {hw:>8s} | " + " | ".join(cells))
    print(f"# sum of these classes: measured {tot_meas / 1e3:.2f} ms, floors {tot_floor / 1e3:.2f} ms -> {tot_floor / tot_meas:.2f} of the roofline overall")


if __name__ == "__main__":
    main()
