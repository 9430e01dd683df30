"""Per-role stall summary of an `ncu --page source --csv` dump of the fused kernel (roles are delimited by the
USETMAXREG instructions), plus the hottest instructions of one role.
    python tools/ncu_regions.py source.csv <section index> [role index to list] [top N]
With --import-source the dump holds every launch twice; sections 0/2 are the SASS views of launch 0/1."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
role = int(sys.argv[3]) if len(sys.argv) > 3 else -1
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
        continue
    if cur is None:
        continue
    if cur["hdr"] is None:
        cur["hdr"] = r
        continue
    cur["rows"].append(r)
k = kernels[sec]
h = k["hdr"]
idx = {n: i for i, n in enumerate(h)}
stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
R = k["rows"]
bounds = [0] + [i for i, r in enumerate(R) if "USETMAXREG" in r[idx["Source"]]] + [len(R)]
names = ["prologue", "issuer warpgroup", "stage workers", "solver warpgroups + tail"]
tot = sum(int(r[idx["# Samples"]]) for r in R)
print(f"section {sec}: {k['name'][:70]}  {tot} samples, {len(R)} instructions")
for ri in range(len(bounds) - 1):
    a, b = bounds[ri], bounds[ri + 1]
    s = sum(int(r[idx["# Samples"]]) for r in R[a:b])
    agg = {}
    for r in R[a:b]:
        for n in stall:
            v = int(r[idx[n]])
            if v:
                agg[n[6:]] = agg.get(n[6:], 0) + v
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:6]
    print(f"  [{ri}] {names[ri] if ri < len(names) else ri}: {100 * s / tot:.1f}% of samples; " +
          ", ".join(f"{n} {100 * v / max(s, 1):.0f}%" for n, v in top))
if role >= 0:
    a, b = bounds[role], bounds[role + 1]
    top = sorted(range(a, b), key=lambda i: -int(R[i][idx["# Samples"]]))[:topn]
    for i in sorted(top):
        r = R[i]
        st = sorted(((n[6:], int(r[idx[n]])) for n in stall if int(r[idx[n]]) > 0), key=lambda kv: -kv[1])[:2]
        print(f"   {i:5d} {100 * int(r[idx['# Samples']]) / tot:5.2f}% exec {r[idx['Instructions Executed']]:>10s}  {r[idx['Source']].strip()[:64]:64s} {st}")
