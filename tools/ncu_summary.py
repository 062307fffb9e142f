"""Summarise an .ncu-rep (read here, no GPU): headline metrics + hottest SASS lines with stall reasons.
    python tools/ncu_summary.py gpurun_out/attn.ncu-rep [min_pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name", "?")[:100])
    for k in KEYS:
        if k in d:
            print(f"  {k:95s} {units[hdr.index(k)]:12s} {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
stop = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] == "Address"), len(rows))   # first launch only
hdr, data = rows[start], [r for r in rows[start + 1:stop] if len(r) == len(rows[start])]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(f"-- SASS lines with >= {min_pct}% of {tot} samples")
for n, r in enumerate(data):
    s = int(r[ix["# Samples"]])
    if s >= tot * min_pct / 100:
        st = sorted(((int(r[ix[h]]), h) for h in stalls), reverse=True)[:3]
        print(f"  {n:5d} {r[ix['Source']][:70]:70s} {100 * s / tot:5.1f}%  exec {r[ix['Instructions Executed']]:>10s}  "
              + " ".join(f"{h[6:]}={v}" for v, h in st if v > 0))

# ---- stall-reason totals per SASS region (regions split at USETMAXREG / EXIT markers)
import collections
region, names = 0, {}
agg = collections.defaultdict(lambda: collections.Counter())
for n, r in enumerate(data):
    src_line = r[ix["Source"]]
    if "USETMAXREG" in src_line:
        region += 1
        names[region] = f"after line {n}: {src_line.strip()[:40]}"
    for h in stalls:
        v = int(r[ix[h]])
        if v:
            agg[region][h[6:]] += v
print("-- stall totals per region")
for reg in sorted(agg):
    tot_r = sum(agg[reg].values())
    top = ", ".join(f"{k}={100 * v / tot_r:.0f}%" for k, v in agg[reg].most_common(7))
    print(f"  region {reg} ({names.get(reg, 'prologue')}): {tot_r} samples ({100 * tot_r / tot:.1f}%): {top}")
