#!/usr/bin/env python3
"""SASS mnemonic counts per kernel of the shipped library (cuobjdump -sass): the tcgen05 / TMA / mbarrier / 256-bit access evidence.
usage: tools/sass_evidence.py [lib.so] > profiles/rNN/sass_evidence.txt"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "audiodec_b200/lib/libaudiodec_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UBLKCP", "SYNCS", "UTMALDG", "LDG.E.ENL2.256", "STG.E.ENL2.256", "MUFU.EX2", "FFMA", "ELECT"]
cur, counts = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(adec::\w+(, int)*\)|void |adec::", "", cur)
        counts[cur] = dict.fromkeys(MN, 0)
        continue
    if cur:
        for k in MN:
            if re.search(r"\b" + re.escape(k) + r"\b", line):
                counts[cur][k] += 1
print(f"SASS mnemonic counts per kernel of {lib} (cuobjdump -sass, sm_100a)")
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops,")
print("LDG/STG.E.ENL2.256 = 256-bit global accesses, MUFU.EX2 = ex2.approx (ELU); UTMALDG (tensor-map TMA) is not used: every window / weight stage is one contiguous range")
print(f"{'kernel':64s}" + "".join(f"{k.replace('.E.ENL2', ''):>11s}" for k in MN))
for k, c in counts.items():
    if "tc_conv" in k or "rvq" in k or "probe" in k or "lookup" in k or "stem" in k or "head" in k:
        print(f"{k[:64]:64s}" + "".join(f"{c[m]:11d}" for m in MN))
