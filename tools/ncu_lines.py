#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export:
share of executed instructions, share of stall samples and the top stall reasons."""
import csv
import sys


def main(path, section=0, thresh=0.004):
    rows = list(csv.reader(open(path)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and "Instructions Executed" in r]
    h0 = heads[section]
    hdr = rows[h0]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    end = heads[section + 1] if section + 1 < len(heads) else len(rows)
    sec = [r for r in rows[h0 + 1:end] if len(r) == len(hdr) and r[0].isdigit()]
    tot = sum(int(r[ie]) for r in sec if r[ie].isdigit())
    tots = sum(int(r[isamp]) for r in sec if r[isamp].isdigit())
    print("total warp instructions", tot, "samples", tots)
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    for r in sec:
        if r[ie].isdigit() and (int(r[ie]) > tot * thresh or (r[isamp].isdigit() and int(r[isamp]) > tots * thresh)):
            st = {hdr[i][6:]: int(r[i]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0}
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print("%4s %6.2f%% inst %6.2f%% samp  %-72s %s" % (r[0], 100 * int(r[ie]) / tot,
                  100 * int(r[isamp]) / tots if r[isamp].isdigit() else 0, r[1].strip()[:72], top))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
