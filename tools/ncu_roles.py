"""Summarise an `ncu --page source --csv` dump of the fused kernel: poll counts of every mbarrier
wait loop (who waits for whom) and the hottest instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"hdr": None, "rows": []}
        kernels.append(cur)
        continue
    if cur is None:
        continue
    if cur["hdr"] is None:
        cur["hdr"] = r
        continue
    cur["rows"].append(r)
for ki, k in enumerate(kernels):
    h = k["hdr"]
    idx = {n: i for i, n in enumerate(h)}
    tot = sum(int(r[idx["# Samples"]]) for r in k["rows"])
    print(f"=== launch {ki}: {tot} samples, {len(k['rows'])} SASS instructions")
    for i, r in enumerate(k["rows"]):
        src = r[idx["Source"]].strip()
        if any(m in src for m in ("TRYWAIT", "UTCHMMA", "UBLKCP", "BAR.SYNC", "LDTM.x4", "USETMAXREG")):
            print(f"   {i:5d} samples {int(r[idx['# Samples']]):8d} exec {int(r[idx['Instructions Executed']]):11d}  {src[:84]}")
    top = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][idx["# Samples"]]))[:14]
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    print("   hottest:")
    for i in top:
        r = k["rows"][i]
        st = sorted(((n, int(r[idx[n]])) for n in stall_cols if int(r[idx[n]]) > 0), key=lambda kv: -kv[1])[:2]
        print(f"   {i:5d} {100 * int(r[idx['# Samples']]) / tot:5.1f}%  {r[idx['Source']].strip()[:60]:60s} {st}")
