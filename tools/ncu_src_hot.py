"""Hot instructions of an ncu source-page CSV: python tools/ncu_src_hot.py src.csv [top] [lo_addr hi_addr]
Prints per-instruction samples with the dominant stall reasons; with an address window prints the
window in program order (the inner loop)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = rows[2:]
ia, isrc, ismp, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
total = sum(int(r[ismp] or 0) for r in body)


def fmt(r):
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall), reverse=True)[:3]
    return "%6s %5.2f%% x%-8s %-70s %s" % (r[ia][-5:], 100.0 * int(r[ismp] or 0) / max(total, 1), r[iexe], r[isrc][:70],
                                         " ".join("%s=%d" % (n, v) for v, n in st if v))


if len(sys.argv) >= 5:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    for r in body:
        a = int(r[ia], 16)
        if lo <= a <= hi:
            print(fmt(r))
else:
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    print("total samples", total)
    agg = {}
    for r in body:
        for i, h in stall:
            agg[h[6:]] = agg.get(h[6:], 0) + int(r[i] or 0)
    print(" ".join("%s=%.1f%%" % (k, 100.0 * v / max(total, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(body, key=lambda r: -int(r[ismp] or 0))[:top]:
        print(fmt(r))
