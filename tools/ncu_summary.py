"""Condenses an .ncu-rep (ncu --set full) into a small text table for profiles/. Usage: python tools/ncu_summary.py rep [out]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_act_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("inst_M", "smsp__inst_executed.sum"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
]


def to_float(v, unit, name):
    v = float(v.replace(",", ""))
    u = unit.lower()
    if name.endswith("_MB") or name.endswith("_KB"):
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
        return v * scale / (1e6 if name.endswith("_MB") else 1e3)
    if name == "time_us":
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    if name == "inst_M":
        return v / 1e6
    return v


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# " + rep, "kernel | " + " | ".join(n for n, _ in WANT)]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("unnamed>::", "")
        vals = []
        for n, m in WANT:
            if m in hdr:
                i = hdr.index(m)
                try:
                    vals.append(f"{to_float(r[i], units[i], n):.2f}")
                except ValueError:
                    vals.append(r[i])
            else:
                vals.append("-")
        lines.append(name + " | " + " | ".join(vals))
    txt = "\n".join(lines)
    print(txt)
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(txt + "\n")


if __name__ == "__main__":
    main()
