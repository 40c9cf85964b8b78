"""Top stall-sample SASS lines of each kernel in an .ncu-rep (source page): python tools/ncu_hot.py rep [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(raw.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
for k, b in enumerate(blocks):
    h = b["hdr"]
    iS, iN, iI = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[iN] or 0) for r in b["rows"])
    print(f"# launch {k}: {b['name']}  samples {tot}  sass lines {len(b['rows'])}")
    order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][iN] or 0))[:top]
    for i in sorted(order):
        r = b["rows"][i]
        st = sorted(((int(r[c] or 0), h[c][6:]) for c in stall_cols), reverse=True)[:2]
        print(f"  {i:5d} {int(r[iN]):6d} ({int(r[iN]) / max(tot, 1):5.1%}) x{int(r[iI] or 0):8d}  {r[iS].strip()[:70]:70s} {st}")
