#!/usr/bin/env python
"""Per-kernel SASS opcode census of the built library (cuobjdump -sass): the evidence for
'TMA-fed, packed FP32' claims (UTMALDG, SYNCS, FFMA2, LDS.128 ...).  python tools/sass_summary.py [regex]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "sdr-j-fm_b200", "libsdrjfm_b200.so")
want = re.compile(sys.argv[1] if len(sys.argv) > 1 else "frontend")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, ops = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        ops[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur:
        ops[cur][m.group(1)] += 1
dem = subprocess.run(["c++filt"], input="\n".join(ops), capture_output=True, text=True).stdout.splitlines()
for name, d in zip(ops, dem):
    if not want.search(d):
        continue
    c = ops[name]
    tot = sum(c.values())
    keys = ["UTMALDG.4D", "UTMALDG.3D", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "FFMA2", "FFMA", "FMUL", "FADD", "FADD2",
            "LDS.128", "LDS.64", "LDS", "STS.64", "STS", "LDG.E.64", "LDG.E.U16", "LDG.E", "STG.E.128", "STG.E.64", "PRMT", "I2F", "I2FP.F32.S32",
            "DMUL", "DFMA", "DADD", "MUFU.COS", "BAR.SYNC.DEFER_BLOCKING", "SHFL.UP"]
    shown = " ".join(f"{k}={c[k]}" for k in keys if c.get(k))
    print(f"{d.split('(')[0][:90]}\n    {tot} instructions: {shown}")
