#!/usr/bin/env python
"""Per-source-line summary of `ncu -i rep --page source --print-source cuda,sass --csv > x.csv`, over ALL source files:
share of executed warp instructions, share of stall samples and the top stall reasons.
python tools/ncu_lines.py x.csv [top N]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fname, hdr, agg = "?", None, {}
    for r in rows:
        if r and r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        elif hdr and len(r) == len(hdr) and r[0].isdigit() and r[ie].isdigit():
            st = {hdr[i][6:]: int(r[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0}
            agg[(fname, int(r[0]))] = (int(r[ie]), int(r[isamp]) if r[isamp].isdigit() else 0, r[1].strip()[:80], st)
    tot = sum(v[0] for v in agg.values())
    tots = sum(v[1] for v in agg.values()) or 1
    print("total warp instructions", tot, "samples", tots)
    for (f, ln), (n, s, src, st) in sorted(agg.items(), key=lambda kv: -max(kv[1][0] / tot, kv[1][1] / tots))[:top]:
        topst = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print("%-18s %4d %6.2f%% inst %6.2f%% samp  %-80s %s" % (f[:18], ln, 100 * n / tot, 100 * s / tots, src, topst))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
