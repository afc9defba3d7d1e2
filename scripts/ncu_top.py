"""Top stall lines of one kernel launch from an ncu report (CUDA source lines with their SASS samples folded in):
   python scripts/ncu_top.py rep kernel_regex [launch_skip] [n_lines]"""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
nl = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for i, r in enumerate(rows):
    if 'Source' in r and any(c.startswith('Warp Stall Sampling (All') for c in r):
        hdr, start = r, i + 1
        break
if hdr is None:
    print(out[:2000]); sys.exit(1)
print(rows[0][1][:160] if len(rows[0]) > 1 else rows[0])
si = hdr.index('Source'); ai = hdr.index('Warp Stall Sampling (All Samples)')
stalls = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
# cuda,sass view: a CUDA line row is followed by its SASS rows; aggregate by the last seen CUDA row
agg = collections.OrderedDict()
cur = None
for r in rows[start:]:
    if len(r) <= ai:
        continue
    addr = r[hdr.index('Address')] if 'Address' in hdr else ''
    is_sass = addr.startswith('0x')
    if not is_sass:
        cur = r[si].strip()[:150]
        agg.setdefault(cur, [0, collections.Counter(), collections.Counter()])
        continue
    if not r[ai].isdigit() or cur is None:
        continue
    n = int(r[ai])
    agg[cur][0] += n
    for i in stalls:
        if r[i].isdigit() and int(r[i]):
            agg[cur][1][hdr[i][6:]] += int(r[i])
    if n:
        agg[cur][2][r[si].strip()[:60]] += n
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:nl]:
    print(f"{v[0]:7d} {100 * v[0] / max(tot, 1):5.1f}%  {k[:110]:110s} {v[1].most_common(3)}  sass:{v[2].most_common(2)}")
