#!/usr/bin/env python
"""Per-kernel share of the LAST outer iteration in an ncu launch list (profiles/r2_launches_gamg64.csv):

    python profiles/launch_shares.py profiles/r2_launches_gamg64.csv profiles/r2_launch_shares.md

An outer iteration starts with the device copy D.prevIter <- D followed by k_bc_update; the last complete one in the
capture is taken (launch times under ncu are serialised and cold-cache: shares, not bench times)."""
import csv
import io
import re
import sys
from collections import OrderedDict


def main():
    src, out = sys.argv[1:3]
    txt = open(src).read()
    txt = txt[txt.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(txt)))
    names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "") for r in rows]
    t = [float(r["Metric Value"]) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(r["Metric Unit"], 1e-6) for r in rows]
    starts = [i for i, n in enumerate(names) if n == "k_bc_update"]
    # outer iterations: from one k_bc_update to the next (the law kernel ends it)
    if len(starts) < 2:
        raise SystemExit("fewer than two outer iterations in the capture")
    a, b = starts[-2], starts[-1]
    agg = OrderedDict()
    for i in range(a, b):
        k = agg.setdefault(names[i], [0, 0.0])
        k[0] += 1; k[1] += t[i]
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# Launch list of one outer iteration (`{src}`, launches {a}-{b - 1} of {len(rows)})\n\n")
        f.write(f"{b - a} kernel launches, {tot:.2f} ms of kernel time (serialised, cold-cache under ncu: shares, not bench times).\n\n")
        f.write("| kernel | launches | total ms | share % |\n|---|---|---|---|\n")
        for n, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {cnt} | {ms:.3f} | {100 * ms / tot:.1f} |\n")
    print(f"{b - a} launches, {tot:.2f} ms -> {out}")


if __name__ == "__main__":
    main()
