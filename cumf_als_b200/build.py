"""Build recipe for libcumf_als_b200.so (hand-written sm_100a CUDA + C ABI).

`python -m cumf_als_b200.build` or `build()` compiles every translation unit in
cumf_als_b200/csrc with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
(no other architecture, no JIT cache: the .so is built in-tree so it travels to
the GPU box with the source snapshot) and links
    cumf_als_b200/libcumf_als_b200.so   the C ABI of include/cumf_als.h
    cumf_als_b200/cumf_als_main         the CLI (same arguments as the reference's ./main)
When the reference tree is present, its unmodified main.cpp is additionally
linked against the library (oracle/_ref/ref_main_on_b200) as a drop-in check.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = ROOT / "build" / "obj"
LIB = PKG / "libcumf_als_b200.so"
CLI = PKG / "cumf_als_main"
REFERENCE = Path(os.environ.get("CUMF_REFERENCE_DIR", "/root/reference"))

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CXXFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
            "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets"]
LIB_SOURCES = ["als_api.cu", "gram_simt.cu", "gram_tc.cu", "gram_tc2.cu", "gram_tc2_a.cu", "gram_tc2_b.cu", "gram_tc2_c.cu", "cg.cu",
               "rmse.cu", "synth.cu", "host_io.cpp"]


def _run(cmd: list[str]) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    if proc.stderr.strip() and os.environ.get("CUMF_BUILD_VERBOSE"):
        sys.stderr.write(proc.stderr)


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def _compile(src: Path, extra: list[str]) -> Path:
    obj = OBJ / (src.name + ".o")
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "cumf_als.h"]
    if _stale(obj, [src, *headers, Path(__file__)]):
        _run([NVCC, *ARCH, *CXXFLAGS, *extra, "-c", str(src), "-o", str(obj)])
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile and link everything; returns the path of the shared library."""
    if verbose:
        os.environ["CUMF_BUILD_VERBOSE"] = "1"
    if force and OBJ.exists():
        shutil.rmtree(OBJ)
    OBJ.mkdir(parents=True, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    srcs = [CSRC / s for s in LIB_SOURCES] + [CSRC / "main.cpp"]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
        objs = list(pool.map(lambda s: _compile(s, extra), srcs))
    lib_objs, main_obj = objs[:-1], objs[-1]
    if force or _stale(LIB, lib_objs):
        _run([NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, lib_objs), "-lcublas",
              "-Xlinker", "-rpath=/usr/local/cuda/lib64"])
    if force or _stale(CLI, [main_obj, LIB]):
        _run([NVCC, *ARCH, "-o", str(CLI), str(main_obj), str(LIB), "-lcublas",
              "-Xlinker", f"-rpath={PKG}", "-Xlinker", "-rpath=/usr/local/cuda/lib64"])
    _link_reference_cli()
    return LIB


def _link_reference_cli() -> None:
    """Drop-in check: the reference's unmodified main.cpp + our library (no reference kernels)."""
    ref_main = REFERENCE / "main.cpp"
    if not ref_main.exists():
        return
    out = ROOT / "oracle" / "_ref" / "ref_main_on_b200"
    out.parent.mkdir(parents=True, exist_ok=True)
    if not _stale(out, [ref_main, LIB]):
        return
    _run([NVCC, *ARCH, "-O2", "-w", "-I", str(REFERENCE), "-o", str(out), str(ref_main), str(LIB), "-lcublas",
          "-lcusparse", "-Xlinker", f"-rpath={PKG}", "-Xlinker", "-rpath=/usr/local/cuda/lib64"])


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
