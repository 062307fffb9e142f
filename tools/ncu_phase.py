"""Per-phase sample/stall breakdown of the attention kernel's softmax loop from an .ncu-rep (read here, no GPU).
    python tools/ncu_phase.py gpurun_out/attn.ncu-rep [first_line last_line]
Prints every SASS line of the range with its samples and top stall reasons, so loop phases can be read off."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) == len(rows[start])]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
acc = 0
for n, r in enumerate(data):
    if n < lo or n > hi:
        continue
    s = int(r[ix["# Samples"]])
    acc += s
    st = sorted(((int(r[ix[h]]), h) for h in stalls), reverse=True)[:3]
    print(f"{n:5d} {r[ix['Source']][:64]:64s} {s:7d} {100 * s / tot:5.2f}% cum {100 * acc / tot:5.1f}% ex {r[ix['Instructions Executed']]:>10s} "
          + " ".join(f"{h[6:]}={v}" for v, h in st if v > 0))
