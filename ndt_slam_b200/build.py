"""In-tree builds: the CUDA library (sm_100a), the C++ host classes, and the test-only oracle.

Everything is built with explicit nvcc / g++ command lines into the source tree so the resulting
.so files travel to the GPU box with the repo snapshot. nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "ndt_slam_b200"
CSRC = PKG / "csrc"
HOST = PKG / "host"
ORACLE = ROOT / "oracle"
REFERENCE = Path("/root/reference")

LIB_CUDA = PKG / "libndt_b200.so"
LIB_HOST = PKG / "libndt_slam_host.so"
LIB_ORACLE = ORACLE / "libndt_oracle.so"
LIB_REF = ORACLE / "_ref" / "libndt_slam_ref.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=true",            # fp64 accumulation may contract; the float32 paths use explicit __fmul_rn/__fadd_rn
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]
# grid_build.cu mirrors the reference stack's non-fused x86-64 arithmetic op for op (bit-exact grid)
NO_FMAD = {"grid_build.cu"}


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found")


def _stamp(srcs, extra="") -> str:
    h = hashlib.sha256(extra.encode())
    for s in sorted(map(str, srcs)):
        h.update(s.encode())
        h.update(Path(s).read_bytes())
    return h.hexdigest()


def _up_to_date(out: Path, stamp: str) -> bool:
    st = out.with_suffix(out.suffix + ".stamp")
    return out.exists() and st.exists() and st.read_text() == stamp


def _write_stamp(out: Path, stamp: str) -> None:
    out.with_suffix(out.suffix + ".stamp").write_text(stamp)


def _run(cmd, log: Path | None = None) -> str:
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.write_text(" ".join(map(str, cmd)) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build failed: " + " ".join(map(str, cmd)))
    return p.stdout


def build_variant(tag: str, defines: list[str]) -> Path:
    """Experimental build with extra -D flags into ndt_slam_b200/build/libndt_b200_<tag>.so (A/B tuning)."""
    objdir = PKG / "build" / tag
    objdir.mkdir(parents=True, exist_ok=True)
    out = PKG / "build" / f"libndt_b200_{tag}.so"
    objs = []
    for s in sorted(CSRC.glob("*.cu")):
        o = objdir / (s.stem + ".o")
        flags = list(NVCC_FLAGS)
        if s.name in NO_FMAD:
            flags[flags.index("--fmad=true")] = "--fmad=false"
        _run([_nvcc(), *flags, *[f"-D{d}" for d in defines], "-c", "-I", str(ROOT / "include"), "-I", str(CSRC), "-o", str(o), str(s)])
        objs.append(o)
    _run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *map(str, objs), "-lcudart"])
    return out


def build_cuda(force: bool = False) -> Path:
    """libndt_b200.so: the hand-written sm_100a kernels + the C ABI (include/ndt_b200.h)."""
    srcs = sorted(CSRC.glob("*.cu"))
    deps = srcs + sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "ndt_b200.h"]
    stamp = _stamp(deps, " ".join(NVCC_FLAGS))
    if not force and _up_to_date(LIB_CUDA, stamp):
        return LIB_CUDA
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    objs, logs = [], []
    for s in srcs:
        o = objdir / (s.stem + ".o")
        flags = list(NVCC_FLAGS)
        if s.name in NO_FMAD:
            flags[flags.index("--fmad=true")] = "--fmad=false"
        logs.append(_run([_nvcc(), *flags, "-c", "-I", str(ROOT / "include"), "-I", str(CSRC), "-o", str(o), str(s)]))
        objs.append(o)
    logs.append(_run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_CUDA),
                      *map(str, objs), "-lcudart"]))
    (PKG / "build_cuda.log").write_text("\n".join(logs))
    _write_stamp(LIB_CUDA, stamp)
    return LIB_CUDA


def build_host(force: bool = False) -> Path:
    """libndt_slam_host.so: C++ mirror of the reference classes above the C ABI + a C test harness."""
    srcs = sorted((HOST / "src").glob("*.cpp"))
    if not srcs:
        raise RuntimeError("no host sources")
    deps = srcs + sorted(p for p in (HOST / "include").rglob("*") if p.is_file()) + [ROOT / "include" / "ndt_b200.h"]
    flags = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unused-variable"]
    stamp = _stamp(deps, " ".join(flags))
    if not force and _up_to_date(LIB_HOST, stamp):
        return LIB_HOST
    build_cuda()
    cmd = ["g++", *flags, "-I", str(ROOT / "include"), "-I", str(HOST / "include"), "-I", str(HOST / "include" / "compat"),
           "-o", str(LIB_HOST), *map(str, srcs),
           "-L", str(PKG), "-lndt_b200", "-Wl,-rpath,$ORIGIN"]
    _run(cmd, log=PKG / "build_host.log")
    _write_stamp(LIB_HOST, stamp)
    return LIB_HOST


def build_oracle(force: bool = False) -> Path:
    """oracle/libndt_oracle.so: the self-contained CPU restatement (test infrastructure)."""
    srcs = [ORACLE / "ndt_oracle.cpp"]
    flags = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall"]
    stamp = _stamp(srcs + [ROOT / "include" / "ndt_b200.h"], " ".join(flags))
    if not force and _up_to_date(LIB_ORACLE, stamp):
        return LIB_ORACLE
    _run(["g++", *flags, "-o", str(LIB_ORACLE), *map(str, srcs)], log=ORACLE / "build_oracle.log")
    _write_stamp(LIB_ORACLE, stamp)
    return LIB_ORACLE


def build_ref(force: bool = False) -> Path | None:
    """oracle/_ref/libndt_slam_ref.so: the reference's own sources, compiled where they lie, against
    oracle/stubs (ROS) and oracle/minipcl (restated PCL). Only possible where /root/reference exists."""
    mk = ORACLE / "Makefile"
    if not REFERENCE.exists() or not mk.exists():
        return LIB_REF if LIB_REF.exists() else None
    _run(["make", "-C", str(ORACLE), "-j4"] + (["-B"] if force else []), log=ORACLE / "build_ref.log")
    return LIB_REF if LIB_REF.exists() else None


def build_all(force: bool = False) -> None:
    build_cuda(force)
    if (HOST / "src").exists() and list((HOST / "src").glob("*.cpp")):
        build_host(force)
    build_oracle(force)
    build_ref(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", *(p for p in (LIB_CUDA, LIB_HOST, LIB_ORACLE, LIB_REF) if p.exists()))
