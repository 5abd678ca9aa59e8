"""In-tree build of libmtbc.so (hand-written sm_100a kernels + C ABI) with plain nvcc.

The .so lands next to this file so that it travels to the GPU box with the repo snapshot and shows up as an in-tree
native library when loaded.  No torch headers are involved: the ABI is plain C (include/mtbc.h).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libmtbc.so"

SOURCES = ["api.cu", "conv_gemm.cu", "conv_halo.cu", "convt_bwd.cu", "pack.cu", "stream_ops.cu", "stream_pipe.cu", "heads.cu", "loss.cu", "augment.cu", "residual.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (Path(cand).exists() or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [PKG_DIR.parent / "include" / "mtbc.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def lib_digest(path: Path = LIB_PATH):
    """Source digest compiled into an existing libmtbc.so (`mtbc_build_digest`), or None if it cannot be read."""
    if not path.exists():
        return None
    import ctypes
    try:
        lib = ctypes.CDLL(str(path))
        fn = lib.mtbc_build_digest
    except (OSError, AttributeError):
        return None
    fn.restype = ctypes.c_char_p
    fn.argtypes = []
    return fn().decode()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into libmtbc.so.  Skipped only when the digest COMPILED INTO the library
    on disk equals the digest of the sources on disk (no side-car stamp file that could outlive a checkout)."""
    digest = _digest()
    if not force and lib_digest() == digest:
        return LIB_PATH
    objs = []
    build_dir = PKG_DIR / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = build_dir / (src + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if src == "api.cu":
            cmd.insert(1, f'-DMTBC_BUILD_DIGEST="{digest}"')
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    log = []
    failed = False
    for src, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode != 0:
            failed = True
    (build_dir / "build.log").write_text("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed, see build/build.log")
    if verbose:
        print("\n".join(log))
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *objs]
    tmp = LIB_PATH.with_suffix(".so.tmp")
    link[link.index(str(LIB_PATH))] = str(tmp)
    subprocess.run(link, check=True)
    os.replace(tmp, LIB_PATH)   # a process that already mapped the old file keeps its inode
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
