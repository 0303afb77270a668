"""Developer tool: build a variant of the native library with extra -D flags for dense.cu into
tools/_dbg/lib_<name>.so (load it with MNF_LIB=...). Usage: python tools/build_variant.py name -DX=1 ..."""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mininf_b200.engine import build  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
build.build()                                  # the other units are reused from _lib/obj
out_dir = Path(__file__).resolve().parent / "_dbg"
out_dir.mkdir(exist_ok=True)
obj = out_dir / f"dense_{name}.o"
subprocess.run([build._nvcc(), *build.NVCC_FLAGS, *flags, "-c", "-o", str(obj), str(build.CSRC_DIR / "dense.cu")], check=True)
lib = out_dir / f"lib_{name}.so"
others = [str(build.LIB_DIR / "obj" / f"{unit}.o") for unit in build.UNITS if unit != "dense"]
subprocess.run([build._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(lib), str(obj), *others], check=True)
print(lib)
