"""Per-kernel counts of the Blackwell-specific SASS opcodes in libokl_b200.so (cuobjdump -sass): evidence that the hot kernels are
tcgen05 / TMA / TMEM code.  python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "okrolearn_b200", "libokl_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "UBLKCP", "HMMA", "FFMA", "LDGSTS", "LDGMC", "STGMC", "REDGMC", "RED", "ATOM"]
per = collections.OrderedDict()
cur = None
arch = None
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        per[cur]["arch:" + str(arch)] = 1
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        per[cur]["_instr"] += 1
        op = m.group(1)
        for o in ops:
            if op.startswith(o):
                per[cur][o + ("." + op.split(".")[1] if o in ("UTMALDG", "UTMASTG") and "." in op else "")] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(so)}: {len(per)} kernels; SASS opcode counts per kernel (UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store,")
print("# LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops)")
tot = collections.Counter()
for (k, c), name in zip(per.items(), demangle):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = name.split("(")[0][-90:]
    keys = [o for o in c if not o.startswith("_") and not o.startswith("arch:") and o not in ("FFMA", "RED", "ATOM", "SYNCS")]
    for o in c:
        if not o.startswith("arch:"):
            tot[o] += c[o]
    if any(o.startswith(("UTC", "UTMA", "LDTM", "LDGMC", "STGMC", "REDGMC")) for o in keys):
        print(f"{name:92s} instr {c['_instr']:6d}  " + "  ".join(f"{o}={c[o]}" for o in sorted(keys)))
print("# totals over all kernels: " + "  ".join(f"{o}={tot[o]}" for o in sorted(tot) if not o.startswith("_")))
