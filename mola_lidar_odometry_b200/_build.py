"""In-tree builds of the native pieces (nvcc / g++), so the .so files travel with the repo snapshot.

  libmlo_b200.so   CUDA kernels + C ABI (include/mlo_b200.h), sm_100a only
  libmlo_synth.so  synthetic scene / LiDAR generator (input data, not on the measured path)

The oracle (oracle/liboracle.so) is built by `oracle/Makefile`; `build_oracle()` only drives it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB_CUDA = PKG / "libmlo_b200.so"
LIB_SYNTH = PKG / "synth" / "libmlo_synth.so"
LIB_ORACLE = ROOT / "oracle" / "liboracle.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # discrete decisions (voxel index, NN argmin, threshold) must round as written
    "-Xcompiler", "-fPIC,-O3,-ffp-contract=off,-pthread", "-shared", "-Xptxas", "-v",
]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libmlo_b200.so cannot be built")


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    srcs = sorted(CSRC.glob("*.cu")) + sorted((PKG / "host").glob("*.cpp"))
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + sorted((PKG / "host").glob("*.h")) + sorted((PKG / "host").glob("*.hpp")) + [
        ROOT / "include" / "mlo_b200.h", ROOT / "include" / "mlo_b200_host.h"]
    if force or _newer(LIB_CUDA, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(ROOT / "include"), "-I", str(CSRC), "-I", str(PKG / "host"),
               "-o", str(LIB_CUDA), *map(str, srcs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libmlo_b200.so")
        (PKG / "csrc" / "ptxas_report.txt").write_text(r.stdout + r.stderr)
    return LIB_CUDA


def build_synth(force: bool = False) -> Path:
    src = PKG / "synth" / "synth.cpp"
    if force or _newer(LIB_SYNTH, [src]):
        cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-o", str(LIB_SYNTH), str(src)]
        subprocess.run(cmd, check=True)
    return LIB_SYNTH


LIB_SYNTH_CUDA = PKG / "synth" / "libmlo_synth_cuda.so"


def build_synth_cuda(force: bool = False) -> Path:
    """synth/synth_gpu.cu: the ray caster as a CUDA kernel (bench.py's sequence workloads; input generator only)."""
    src = PKG / "synth" / "synth_gpu.cu"
    if force or _newer(LIB_SYNTH_CUDA, [src]):
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
               "-shared", "-o", str(LIB_SYNTH_CUDA), str(src), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building libmlo_synth_cuda.so")
    return LIB_SYNTH_CUDA


def build_oracle(force: bool = False) -> Path:
    if force:
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "clean"], check=True, capture_output=True)
    subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, capture_output=True)
    return LIB_ORACLE


CLI_BIN = ROOT / "apps" / "mlo-lidar-odometry-cli"


def build_cli(force: bool = False) -> Path:
    """apps/mlo-lidar-odometry-cli: the offline driver (reference: apps/mola-lidar-odometry-cli.cpp) over libmlo_b200.so."""
    src = ROOT / "apps" / "mlo-lidar-odometry-cli.cpp"
    if force or _newer(CLI_BIN, [src, ROOT / "include" / "mlo_b200_host.h", LIB_CUDA]):
        cmd = ["g++", "-O2", "-std=c++17", "-I", str(ROOT / "include"), str(src), "-o", str(CLI_BIN), "-L", str(PKG),
               "-lmlo_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../mola_lidar_odometry_b200"]
        subprocess.run(cmd, check=True)
    return CLI_BIN


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_synth(force)
    build_synth_cuda(force)
    build_cuda(force, verbose)
    build_oracle(force)
    build_cli(force)
