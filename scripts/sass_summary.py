#!/usr/bin/env python
"""SASS evidence for the tensor-core kernels of librlc.so (cuobjdump needs no GPU).

    python scripts/sass_summary.py > profiles/sass_k1.txt

Per kernel: counts of the mnemonics that prove what runs where -- UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2),
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D bulk copy: there is no UTMALDG, the
operands are pre-packed in the UMMA core-matrix layout and need no tensor map), SYNCS = mbarrier operations."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "rlcontrol_b200/librlc.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UTCCP",
        "HFMA2", "FFMA2", "F2FP", "MEMBAR", "CCTL")
cnt = collections.defaultdict(collections.Counter)
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    if base in keys:
        cnt[fn][base + (".2CTA" if ".2CTA" in op else "")] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {lib}: mnemonic counts of the kernels that touch the tensor pipe / bulk copies")
for mangled, name in sorted(zip(cnt, names), key=lambda x: x[1]):
    c = cnt[mangled]
    if not any(k.startswith(("UTC", "LDTM", "STTM", "UBLKCP")) for k in c):
        continue
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    print(f"{short:44s} " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
