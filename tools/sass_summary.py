"""SASS evidence for profiles/: per kernel of libmimo_b200's objects, the counts of the mnemonics that prove the code path
(tcgen05 MMA = UTCHMMA, TMEM alloc/ld = UTCATOM*/LDTM, TMA tile loads = UTMALDG, bulk copies = UBLKCP, mbarriers = SYNCS,
global reductions = RED/REDG) plus ptxas registers / spills. Needs no GPU.  Usage: python tools/sass_summary.py > profiles/rXX_sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mimo_unet_b200", "csrc")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "RED", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "FFMA", "MUFU"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", demangle(n))
    n = re.sub(r"^void ", "", n)
    return n.split("(")[0].replace("mimo::", "")


def main():
    regs = {}
    for log in glob.glob(os.path.join(CSRC, "*.ptxas.log")):
        cur = None
        for line in open(log):
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                regs.setdefault(cur, {})["spill"] = (int(m.group(2)), int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                regs.setdefault(cur, {})["regs"] = int(m.group(1))
    print("# cuobjdump -sass of mimo_unet_b200/csrc/*.o (sm_100a): mnemonic counts per kernel; regs / spill bytes from ptxas -v")
    print(f"{'kernel':58s} {'regs':>4s} {'spill':>9s} " + " ".join(f"{k:>7s}" for k in KEYS))
    for obj in sorted(glob.glob(os.path.join(CSRC, "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur, counts = None, collections.OrderedDict()
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and cur:
                op = m.group(1)
                for k in KEYS:
                    if op.startswith(k):
                        counts[cur][k] += 1
                        break
        print(f"## {os.path.basename(obj)}")
        for fn, c in counts.items():
            r = regs.get(fn, {})
            sp = r.get("spill", (0, 0))
            print(f"{short(fn)[:58]:58s} {r.get('regs', 0):4d} {sp[0]:4d}/{sp[1]:<4d} " + " ".join(f"{c.get(k, 0):7d}" for k in KEYS))


if __name__ == "__main__":
    main()
