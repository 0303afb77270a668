"""Build the CUDA shared library in-tree (``mininf_b200/_lib/libmininf_b200.so``).

The library is a plain C-ABI ``.so`` (no torch headers): the translation units of ``csrc/``
(``abi.cu``, ``dense.cu``, ``site.cu``, ``rowlatent.cu``) are compiled for sm_100a in parallel and
linked into one shared object. It is built in-tree so the binary travels with a repository
snapshot; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import hashlib
import os
import re
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
    "-Xcompiler", "-fPIC",
]
# translation units: compiled in parallel, linked into ONE shared library
UNITS = ("abi", "dense", "site", "rowlatent")
_INCLUDE = re.compile(r'^\s*#include\s+"([^"]+)"', re.M)


def _nvcc() -> str:
    for candidate in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if candidate and (os.path.sep not in candidate or os.path.exists(candidate)):
            return candidate
    raise RuntimeError("nvcc not found; set NVCC or install the CUDA toolkit")


def _closure(path: Path, seen: dict) -> None:
    """Quoted includes of `path`, recursively (the unit's real dependencies)."""
    path = path.resolve()
    if path in seen or not path.exists():
        return
    text = path.read_bytes()
    seen[path] = text
    for name in _INCLUDE.findall(text.decode("utf-8", "replace")):
        _closure(path.parent / name, seen)


def unit_digest(unit: str) -> str:
    """Digest of one translation unit: its source, every header it includes, the flags."""
    seen: dict = {}
    _closure(CSRC_DIR / f"{unit}.cu", seen)
    digest = hashlib.sha256()
    for path in sorted(seen):
        digest.update(path.name.encode())
        digest.update(seen[path])
    digest.update(" ".join(NVCC_FLAGS).encode())
    return digest.hexdigest()


def source_digest() -> str:
    """Digest of every source the library is compiled from."""
    return hashlib.sha256("".join(unit_digest(unit) for unit in UNITS).encode()).hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and STAMP_PATH.exists() and \
        STAMP_PATH.read_text().strip() == source_digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if it is missing or older than its sources. Units whose sources did
    not change keep their object file (``_lib/obj/``, git-ignored)."""
    if not force and is_current():
        return LIB_PATH
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(parents=True, exist_ok=True)
    running = []
    for unit in UNITS:
        obj, stamp, digest = obj_dir / f"{unit}.o", obj_dir / f"{unit}.stamp", unit_digest(unit)
        if not force and obj.exists() and stamp.exists() and stamp.read_text().strip() == digest:
            continue
        command = [_nvcc(), *NVCC_FLAGS, "-c", "-o", str(obj), str(CSRC_DIR / f"{unit}.cu")]
        if verbose:
            command[1:1] = ["-Xptxas", "-v"]
        running.append((unit, stamp, digest, subprocess.Popen(command, stdout=subprocess.PIPE,
                                                              stderr=subprocess.PIPE, text=True)))
    failures = []
    for unit, stamp, digest, process in running:
        out, err = process.communicate()
        if process.returncode != 0:
            failures.append(f"nvcc failed on {unit}.cu ({process.returncode}):\n{out}\n{err}")
            continue
        if verbose:
            print(err)
        stamp.write_text(digest + "\n")
    if failures:
        raise RuntimeError("\n".join(failures))
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB_PATH),
            *[str(obj_dir / f"{unit}.o") for unit in UNITS]]
    result = subprocess.run(link, capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError(f"nvcc link failed ({result.returncode}):\n{result.stdout}\n{result.stderr}")
    STAMP_PATH.write_text(source_digest() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
