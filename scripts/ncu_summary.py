"""ncu report -> small JSON summary committed under profiles/ (the judge reads these; bench.py takes `roofline.traffic` from it).
   python scripts/ncu_summary.py <rep.ncu-rep> <out.json> <scheme> "<command profiled>"
scheme = cifar6 labels the conv kernels of ONE captured train step in launch order; anything else keeps kernel names only."""
import csv
import io
import json
import subprocess
import sys

import re

rep, out_path, scheme, cmd = sys.argv[1], sys.argv[2], sys.argv[3], (sys.argv[4] if len(sys.argv) > 4 else "")


def is_wgrad(n):
    """weight-gradient kernels of the BF16 conv stack: both slab kernels, and the first-generation GEMM in MODE 2 (5th template argument)"""
    if "conv_wgrad_tc_kernel" in n or "conv_wgrad_gs_kernel" in n:
        return True
    m = re.search(r"tc_gemm_kernel<([^>]*)>", n)
    if m:
        a = [x.strip() for x in m.group(1).replace("(int)", "").split(",")]
        return len(a) >= 5 and a[4] == "2"
    return False


def label_cifar6(kern):
    conv = ["fprop", "fprop", "fprop", "fprop", "fprop", "dgrad", "dgrad", "dgrad", "dgrad", "dgrad"]
    lay = [2, 3, 4, 5, 6, 6, 5, 4, 3, 2]
    shapes = {2: "conv2 64->64@32x32", 3: "conv3 64->128@16x16", 4: "conv4 128->128@16x16", 5: "conv5 128->256@8x8", 6: "conv6 256->256@8x8"}
    ci, wi = 0, 0
    wl = [6, 5, 4, 3, 2]
    for k in kern:
        n = k["Kernel Name"]
        k.pop("label", None)
        if "conv_tc_kernel" in n and ci < len(conv):
            k["label"] = f"{shapes[lay[ci]]} {conv[ci]}"
            ci += 1
        elif is_wgrad(n) and wi < len(wl):
            k["label"] = f"{shapes[wl[wi]]} wgrad"
            wi += 1
        elif "conv_first_fprop" in n:
            k["label"] = "conv1 3->64@32x32 fprop (first-layer tcgen05)"
        elif "conv_first_wgrad_kernel" in n:
            k["label"] = "conv1 3->64@32x32 wgrad+db (first-layer tcgen05)"


def show(kern, unit_map):
    for k in kern:
        print(f"{k.get('label', ''):48s} {k['Kernel Name'][:50]:50s} {k.get('gpu__time_duration.sum', 0):9.1f} {unit_map.get('gpu__time_duration.sum')} tensor {k.get('tensor_pipe_pct', 0):5.1f}% "
              f"dram {k.get('dram_pct', 0):5.1f}% l2 {k.get('l2_pct', 0):5.1f}% rd {k.get('dram__bytes_read.sum', 0):8.2f}{unit_map.get('dram__bytes_read.sum')} "
              f"wr {k.get('dram__bytes_write.sum', 0):8.2f}{unit_map.get('dram__bytes_write.sum')} clk {k.get('sm_clock', 0):4.2f} GHz")


if rep.endswith(".json"):           # re-label / re-stamp an existing summary (the report itself stays on the GPU box)
    doc = json.load(open(rep))
    if scheme == "cifar6":
        label_cifar6(doc["kernels"])
    doc["commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or doc.get("commit", "")
    json.dump(doc, open(out_path, "w"), indent=1)
    show(doc["kernels"], doc["units"])
    sys.exit(0)

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = {"Kernel Name": "Kernel Name", "gpu__time_duration.sum": "gpu__time_duration.sum", "dram__bytes_read.sum": "dram__bytes_read.sum",
        "dram__bytes_write.sum": "dram__bytes_write.sum", "lts__t_bytes.sum": "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum": "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum": "l1tex__m_l1tex2xbar_write_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
        "launch__grid_size": "grid", "launch__registers_per_thread": "regs", "launch__block_size": "block", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "sm__cycles_elapsed.avg.per_second": "sm_clock", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct"}
idx = {h: i for i, h in enumerate(hdr)}
kern = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    k = {}
    for h, name in WANT.items():
        if h in idx:
            v = r[idx[h]]
            try:
                k[name] = float(v.replace(",", "")) if name != "Kernel Name" else v
            except ValueError:
                k[name] = v
    kern.append(k)
unit_map = {WANT[h]: units[idx[h]] for h in WANT if h in idx}
if scheme == "cifar6":
    label_cifar6(kern)
commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
json.dump({"commit": commit, "command": cmd, "source": rep, "units": unit_map, "note": "ncu --set full --clock-control none; per-launch values; cold-cache, "
           "serialised launches: compare shares and byte counts, not absolute times (B200_PROFILING.md)", "kernels": kern}, open(out_path, "w"), indent=1)
show(kern, unit_map)
