"""In-tree build of libfpb200.so (CUDA kernels + C ABI) for sm_100a with plain nvcc.

The shared library is written next to this file so that it travels to the GPU box with the
repository snapshot; nothing is installed into site-packages.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
LIB = PKG / "libfpb200.so"
OBJ = PKG / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(ROOT / "include"),
]


# preprocess.cu must round every expression as written (bit-parity of integer decisions with the oracle)
PER_FILE_FLAGS = {"preprocess.cu": ["-fmad=false"]}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp(src: Path) -> str:
    h = hashlib.sha1()
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "fpb200.h"]):
        h.update(hdr.read_bytes())
    h.update(" ".join(NVCC_FLAGS + PER_FILE_FLAGS.get(src.name, [])).encode())
    return h.hexdigest()


def _compile(src: Path) -> Path:
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".stamp")
    want = _stamp(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, *PER_FILE_FLAGS.get(src.name, []), "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(want)
    return obj


def build(verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ and link libfpb200.so.  Incremental."""
    OBJ.mkdir(exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "--cudart", "static",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB} from {len(srcs)} sources", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(verbose=True)
