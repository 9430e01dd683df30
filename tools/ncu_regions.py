"""Per-role stall summary of an `ncu --page source --csv` dump of the fused kernel.
    python tools/ncu_regions.py source.csv [top N hottest instructions per launch]
Roles are located by their instructions: the MMA issuer's code surrounds the UTCHMMA instructions, the stage workers'
code the UTMALDG instructions, the solver warpgroups start at USETMAXREG.TRY_ALLOC.  With --import-source the dump
holds every launch twice (SASS view first); only the SASS views are summarised."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
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
seen = set()
for k in kernels:
    h = k["hdr"]
    if "Source" not in h or "# Samples" not in h:
        continue
    idx = {n: i for i, n in enumerate(h)}
    R = k["rows"]
    if not any("UTCHMMA" in r[idx["Source"]] for r in R) or k["name"] in seen:
        continue
    seen.add(k["name"])
    stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[idx["# Samples"]]) for r in R)
    mma = [i for i, r in enumerate(R) if "UTCHMMA" in r[idx["Source"]]]
    tma = [i for i, r in enumerate(R) if "UTMALDG" in r[idx["Source"]]]
    alloc = [i for i, r in enumerate(R) if "TRY_ALLOC" in r[idx["Source"]]][0]
    dealloc = [i for i, r in enumerate(R) if "USETMAXREG.DEALLOC" in r[idx["Source"]]]
    first = dealloc[-1] + 1 if dealloc else 0
    # issuer and worker code are two contiguous blocks between the register hand-back and the solver entry
    if mma[0] < tma[0]:
        cut = (mma[-1] + tma[0]) // 2
        regions = [("MMA issuer warp", first, cut), ("stage worker warps", cut, alloc)]
    else:
        cut = (tma[-1] + mma[0]) // 2
        regions = [("stage worker warps", first, cut), ("MMA issuer warp", cut, alloc)]
    regions.append(("solver warpgroups + idle warps parked at the final barrier", alloc, len(R)))
    print(f"=== {k['name'][:90]}\n    {tot} samples, {len(R)} SASS instructions, {sum(int(r[idx['Instructions Executed']]) for r in R)} warp instructions executed")
    for name, a, b in regions:
        s = sum(int(r[idx["# Samples"]]) for r in R[a:b])
        ex = sum(int(r[idx["Instructions Executed"]]) for r in R[a:b])
        agg = {}
        for r in R[a:b]:
            for n in stall:
                v = int(r[idx[n]])
                if v:
                    agg[n[6:]] = agg.get(n[6:], 0) + v
        top = sorted(agg.items(), key=lambda kv: -kv[1])[:7]
        print(f"  {name}: {100 * s / tot:.1f}% of samples, {ex} warp instructions; " + ", ".join(f"{n} {100 * v / max(s, 1):.0f}%" for n, v in top))
    print("  hottest instructions:")
    for i in sorted(sorted(range(len(R)), key=lambda i: -int(R[i][idx["# Samples"]]))[:topn]):
        r = R[i]
        st = sorted(((n[6:], int(r[idx[n]])) for n in stall if int(r[idx[n]]) > 0), key=lambda kv: -kv[1])[:2]
        role = next(name for name, a, b in regions if a <= i < b) if i >= first else "prologue"
        print(f"   {100 * int(r[idx['# Samples']]) / tot:5.2f}%  exec {r[idx['Instructions Executed']]:>10s}  {r[idx['Source']].strip()[:58]:58s} {st}  [{role.split()[0]}]")
