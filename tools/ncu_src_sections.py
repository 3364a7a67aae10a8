"""Per-kernel hot instructions of an `ncu --page source --csv` dump that holds several kernels:
python tools/ncu_src_sections.py src.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for s in secs:
    hdr = s["rows"][0]
    body = [r for r in s["rows"][1:] if len(r) == len(hdr)]
    ia, isrc, ismp, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[ismp] or 0) for r in body)
    print("=====", s["name"][:110], "samples", total, "instructions", len(body))
    for idx, r in enumerate(body):
        smp = int(r[ismp] or 0)
        if total and 100.0 * smp / total >= minpct:
            st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall), reverse=True)[:3]
            print("%5d %6s %5.2f%% x%-8s %-64s %s" % (idx, r[ia][-5:], 100.0 * smp / total, r[iexe], r[isrc][:64],
                                                    " ".join("%s=%d" % (n, v) for v, n in st if v)))
