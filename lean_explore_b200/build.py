"""Build recipe for the native side: nvcc -> ``lean_explore_b200/liblxg.so`` (sm_100a only).

Run here (no GPU needed, nvcc cross-compiles) or through ``__graft_entry__.build()``.  The
shared library is built in-tree so that it travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "liblxg.so"
SOURCES = ["lxg_search.cu", "lxg_encoder.cu", "lxg_decoder.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "-diag-suppress", "550",
]


def _nvcc() -> str:
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cand = Path(cuda_home) / "bin" / "nvcc"
    return str(cand) if cand.exists() else "nvcc"


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    lib_mtime = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(PKG_DIR.parent / "include" / "lxg.h")
    return any(p.stat().st_mtime > lib_mtime for p in deps)


def build(force: bool = False, verbose: bool = False, defines: tuple[str, ...] = (), lib_path: Path | None = None) -> Path:
    """Compile every CUDA source of the package into one shared library.  `defines` / `lib_path`
    build an A/B variant next to the product library (e.g. ("LXG_EPI_GROUPS=2",), liblxg_g2.so;
    selected at run time with LXG_LIB_PATH)."""
    variant = lib_path is not None
    lib_path = lib_path or LIB_PATH
    if not force and not variant and not needs_build():
        return LIB_PATH
    objs = []
    for src in SOURCES:
        obj = CSRC / (Path(src).stem + (("." + lib_path.stem) if variant else "") + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
        objs.append(str(obj))
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(lib_path), *objs, "-lcudart"]
    subprocess.run(cmd, check=True)
    return lib_path


if __name__ == "__main__":
    print(build(force=True, verbose=True))
