"""Builds libmotionscan.so (sm_100a) in-tree with nvcc. No torch extension machinery: the product is
a plain C-ABI shared library (include/motionscan.h).

    python motion-estimated-video-trimmer_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
CSRC = HERE / "csrc"
OUT = HERE / "libmotionscan.so"
FEED_OUT = HERE / "libmscan_feed.so"  # measurement harness: decode-worker stand-in threads in front of the ABI
FEED_SRC = CSRC / "feed_harness.cpp"

SOURCES = ["ka_scan.cu", "ka_scan_cluster.cu", "kc_segments.cu", "synth.cu", "synth_host.cu", "mscan_api.cu"]
HOST_SOURCES = ["host_project.cpp"]  # plain C++ (g++): host SIMD paths selected at run time
HOST_CXXFLAGS = ["-O3", "-std=c++17", "-march=x86-64-v3", "-fPIC", "-Wall", "-Wextra", "-I/usr/local/cuda/include"]
HEADERS = [CSRC / "common.cuh", CSRC / "kernels.cuh", ROOT / "include" / "motionscan.h", ROOT / "include" / "mvgen_core.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                 # bit-exact f64 tail; the integer kernels have no float math
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-cudart", "static",           # self-contained: loads next to torch's own runtime
    "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not OUT.exists():
        return False
    t = OUT.stat().st_mtime
    deps = [CSRC / s for s in SOURCES + HOST_SOURCES] + HEADERS + [Path(__file__)]
    return all(d.stat().st_mtime <= t for d in deps)


def build_feed(force: bool = False) -> Path:
    """libmscan_feed.so: plain C++ over the public ABI (links libmotionscan.so via $ORIGIN)."""
    deps = [FEED_SRC, ROOT / "include" / "motionscan.h", OUT, Path(__file__)]
    if not force and FEED_OUT.exists() and all(d.stat().st_mtime <= FEED_OUT.stat().st_mtime for d in deps):
        return FEED_OUT
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-O3", "-std=c++17", "-march=x86-64-v3", "-fPIC", "-shared", "-pthread", "-Wall", "-o", str(FEED_OUT), str(FEED_SRC),
           f"-L{HERE}", "-lmotionscan", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return FEED_OUT


MVGEN_OUT = HERE / "libmvgen.so"  # the synthetic-stream generator alone (host side), for processes that must not
                                  # map the product library (bench.py --impl reference)


def build_mvgen(force: bool = False) -> Path:
    src = CSRC / "synth_host.cu"  # pure host code: compiled as C++ by g++
    deps = [src, ROOT / "include" / "motionscan.h", ROOT / "include" / "mvgen_core.h", Path(__file__)]
    if not force and MVGEN_OUT.exists() and all(d.stat().st_mtime <= MVGEN_OUT.stat().st_mtime for d in deps):
        return MVGEN_OUT
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    subprocess.run([cxx, "-x", "c++", "-O3", "-std=c++17", "-march=x86-64-v3", "-fPIC", "-shared", "-pthread", "-o", str(MVGEN_OUT), str(src)],
                   check=True)
    return MVGEN_OUT


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and up_to_date():
        build_feed()
        build_mvgen()
        return OUT
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = objdir / (Path(s).stem + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    for s in HOST_SOURCES:
        o = objdir / (Path(s).stem + ".o")
        cmd = [cxx, *HOST_CXXFLAGS, "-c", str(CSRC / s), "-o", str(o)]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- compilation failed on {s}\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("compilation failed")
    cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-Xcompiler", "-fPIC", "-o", str(OUT), *objs, "-lpthread", "-ldl", "-lrt"]
    subprocess.run(cmd, check=True)
    build_feed(force=True)
    build_mvgen(force=True)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
