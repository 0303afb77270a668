"""Build the CUDA shared library in-tree (``mininf_b200/_lib/libmininf_b200.so``).

The library is a plain C-ABI ``.so`` (no torch headers): one nvcc invocation over
``csrc/abi.cu`` for sm_100a. It is built in-tree so the binary travels with a repository
snapshot; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
from pathlib import Path

PACKAGE_DIR = Path(__file__).resolve().parent.parent
CSRC_DIR = PACKAGE_DIR / "csrc"
LIB_DIR = PACKAGE_DIR / "_lib"
LIB_PATH = LIB_DIR / "libmininf_b200.so"
STAMP_PATH = LIB_DIR / "libmininf_b200.stamp"
INCLUDE_DIR = PACKAGE_DIR.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for candidate in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if candidate and (os.path.sep not in candidate or os.path.exists(candidate)):
            return candidate
    raise RuntimeError("nvcc not found; set NVCC or install the CUDA toolkit")


def source_digest() -> str:
    """Digest of every source the library is compiled from."""
    digest = hashlib.sha256()
    sources = sorted(CSRC_DIR.glob("*.cu")) + sorted(CSRC_DIR.glob("*.cuh")) + \
        sorted(INCLUDE_DIR.glob("*.h"))
    for path in sources:
        digest.update(path.name.encode())
        digest.update(path.read_bytes())
    digest.update(" ".join(NVCC_FLAGS).encode())
    return digest.hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and STAMP_PATH.exists() and \
        STAMP_PATH.read_text().strip() == source_digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if it is missing or older than its sources."""
    if not force and is_current():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    command = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), str(CSRC_DIR / "abi.cu")]
    if verbose:
        command.insert(1, "-Xptxas")
        command.insert(2, "-v")
    result = subprocess.run(command, capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError(f"nvcc failed ({result.returncode}):\n{result.stdout}\n{result.stderr}")
    if verbose:
        print(result.stderr)
    STAMP_PATH.write_text(source_digest() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
