"""Top stall lines of one kernel from an ncu report: python scripts/ncu_top.py rep kernel_regex [skip]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
for i, r in enumerate(rows):
    if 'Source' in r and 'Address' in r:
        hdr, start = r, i + 1
        break
print(rows[0][1][:120])
si, ai = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
stalls = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
body = [r for r in rows[start:] if len(r) > ai and r[ai].isdigit()]
tot = sum(int(r[ai]) for r in body)
print("total samples", tot)
for r in sorted(body, key=lambda r: -int(r[ai]))[:22]:
    top = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print(f"{int(r[ai]):6d} {100*int(r[ai])/max(tot,1):5.1f}%  {r[si][:90]:90s} {top}")
