"""Per-kernel totals of the warp-state samples of an `ncu --page source --csv` dump (all instructions summed):
python tools/ncu_stall_totals.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for s in secs:
    if s["name"] in seen:
        continue
    seen.add(s["name"])
    hdr = s["rows"][0]
    body = [r for r in s["rows"][1:] if len(r) == len(hdr)]
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h[6:]: sum(int(r[i] or 0) for r in body) for i, h in stall}
    allsum = sum(tot.values())
    iexe = hdr.index("Instructions Executed")
    print("=====", s["name"][:100], "samples", allsum, "warp-instructions", sum(int(r[iexe] or 0) for r in body))
    print("   ", " ".join("%s=%.1f%%" % (k, 100.0 * v / allsum) for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 100 > allsum))
