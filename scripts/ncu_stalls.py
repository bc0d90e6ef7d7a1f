#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: stall reasons and the hottest SASS lines.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv
    python scripts/ncu_stalls.py /tmp/src.csv [top_n]
"""
import csv
import sys


def sections(path):
    rows = list(csv.reader(open(path)))
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            yield cur
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    for sec in list(sections(path)):
        hdr, data = sec["hdr"], sec["rows"]
        if not data:
            continue
        si, src = hdr.index("# Samples"), hdr.index("Source")
        stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[si]) for r in data) or 1
        print(f"== {sec['name']}: {tot} samples, {len(data)} SASS lines")
        agg = {h: sum(int(r[i]) for r in data) for i, h in stalls}
        print("   " + "  ".join(f"{h[6:]}={100 * v / tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        for r in sorted(data, key=lambda r: -int(r[si]))[:top_n]:
            st = sorted(((int(r[i]), h[6:]) for i, h in stalls), reverse=True)[:2]
            print(f"   {100 * int(r[si]) / tot:5.1f}% {r[0][-5:]} {r[src].strip()[:72]:72s} {st}")


if __name__ == "__main__":
    main()
