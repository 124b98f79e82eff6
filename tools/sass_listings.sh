#!/bin/bash
# Raw SASS listings of the tcgen05 kernels (cuobjdump of the built objects) + a mnemonic census, committed under profiles/.
# Usage: tools/sass_listings.sh r02
R=${1:-r02}
OUT=profiles/${R}_sass
mkdir -p $OUT
cd "$(dirname "$0")/.."
for f in conv_c2 conv_flat conv_flat2 conv_wgrad_flat conv_wgrad_flatk conv_thin; do
  cuobjdump -sass mimo_unet_b200/csrc/$f.o | sed 's#/\* 0x[0-9a-f]* \*/##' | sed 's/[[:space:]]*$//' | grep -v "^$" > $OUT/$f.sass
done
# the flat kernels are instantiated four times: keep the BN = 32 instance only (the one every 21-channel layer runs)
python3 - "$OUT" <<'PY'
import re, sys, os
out = sys.argv[1]
for name, keep in (("conv_flat", "flat_kernelILi32ELi2E"), ("conv_flat2", "flat2_kernelILi32ELi2E"), ("conv_c2", "c2_kernelILb0E"),
                   ("conv_thin", "thin_kernelILi3ELi3E")):
    p = os.path.join(out, name + ".sass")
    txt = open(p).read()
    parts = re.split(r"(?=\n\s*Function : )", txt)
    kept = [parts[0]] + [q for q in parts[1:] if keep in q.split("\n")[1]]
    open(p, "w").write("".join(kept))
census = {}
for fn in sorted(os.listdir(out)):
    if not fn.endswith(".sass"):
        continue
    c = {}
    for line in open(os.path.join(out, fn)):
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for key in ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCCP", "SYNCS", "LDGSTS", "MEMBAR", "FENCE", "FFMA2"):
                if op.startswith(key):
                    k = op if key in ("UTCHMMA", "UTCBAR", "UTMALDG") else key
                    c[k] = c.get(k, 0) + 1
    census[fn] = c
with open(os.path.join(out, "census.txt"), "w") as f:
    f.write("# tcgen05 / TMA / barrier instruction census of the committed SASS listings (cuobjdump -sass, sm_100a)\n")
    for fn, c in census.items():
        f.write(fn + ": " + ", ".join(f"{k} {v}" for k, v in sorted(c.items())) + "\n")
print(open(os.path.join(out, "census.txt")).read())
PY
wc -l $OUT/*.sass
