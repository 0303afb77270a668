"""The C-ABI shared library loads on a machine without a GPU and exports every symbol that
include/mininf_b200.h declares; the ctypes structures match the C layout."""
import ctypes
import pathlib
import re
import subprocess

import pytest

from mininf_b200.engine import abi, build

ROOT = pathlib.Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "mininf_b200.h"


@pytest.fixture(scope="module")
def library():
    build.build()
    return abi.load()


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(mnf_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(library):
    names = declared_functions()
    assert len(names) >= 10
    for name in names:
        assert hasattr(library._dll, name), f"{name} is declared in the header but not exported"
        assert name in abi.EXPORTS, f"{name} has no ctypes prototype"
    assert library.raw("mnf_abi_version")() == abi.ABI_VERSION
    match = re.search(r"#define MNF_ABI_VERSION (\d+)", HEADER.read_text())
    assert int(match.group(1)) == abi.ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    source = tmp_path / "sizes.c"
    source.write_text('''#include <stdio.h>
#include <stddef.h>
#include "mininf_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mnf_link_t), sizeof(mnf_latent_t), sizeof(mnf_site_t),
         sizeof(mnf_dense_site_t), sizeof(mnf_device_info_t), offsetof(mnf_site_t, param),
         offsetof(mnf_dense_site_t, scale), offsetof(mnf_dense_site_t, weight));
  printf("%zu %zu %zu %zu %zu\\n", sizeof(mnf_rowlatent_t), offsetof(mnf_rowlatent_t, prior_loc),
         offsetof(mnf_rowlatent_t, feat), offsetof(mnf_rowlatent_t, beta_lat), offsetof(mnf_rowlatent_t, resp_scale));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mnf_plan_desc_t), offsetof(mnf_plan_desc_t, latents),
         offsetof(mnf_plan_desc_t, flags), sizeof(mnf_row_buffers_t), sizeof(mnf_buffers_t),
         offsetof(mnf_buffers_t, workspace_bytes), offsetof(mnf_buffers_t, xrank), sizeof(mnf_adam_t),
         offsetof(mnf_adam_t, raw));
  return 0;
}''')
    binary = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(source), "-o", str(binary)], check=True)
    sizes = [int(v) for v in subprocess.run([str(binary)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(abi.Link), ctypes.sizeof(abi.Latent), ctypes.sizeof(abi.Site),
                     ctypes.sizeof(abi.DenseSite), ctypes.sizeof(abi.DeviceInfo), abi.Site.param.offset,
                     abi.DenseSite.scale.offset, abi.DenseSite.weight.offset,
                     ctypes.sizeof(abi.RowLatent), abi.RowLatent.prior_loc.offset, abi.RowLatent.feat.offset,
                     abi.RowLatent.beta_lat.offset, abi.RowLatent.resp_scale.offset,
                     ctypes.sizeof(abi.PlanDesc), abi.PlanDesc.latents.offset, abi.PlanDesc.flags.offset,
                     ctypes.sizeof(abi.RowBuffers), ctypes.sizeof(abi.Buffers), abi.Buffers.workspace_bytes.offset,
                     abi.Buffers.xrank.offset, ctypes.sizeof(abi.Adam), abi.Adam.raw.offset]


def test_argument_errors_do_not_need_a_gpu(library):
    """Null pointers are rejected before any CUDA call; the message is retrievable."""
    code = library.raw("mnf_finalize")(None, 0, 1, 1, None, None, None, 1, None, None, None, None)
    assert code == abi.E_INVALID
    assert b"mnf_finalize" in library.raw("mnf_last_error")()
    with pytest.raises(abi.NativeError, match="mnf_site_sweep"):
        library.call("mnf_site_sweep", None, 1, None, 1, 1, None, None, 0, 0, None, None)
    # the one-call step API validates its tables before touching the device
    handle = ctypes.c_void_p()
    assert library.raw("mnf_plan_create")(None, ctypes.byref(handle)) == abi.E_INVALID
    empty = abi.PlanDesc(n_particles=4, n_latent_total=2, n_latents=0)
    assert library.raw("mnf_plan_create")(ctypes.byref(empty), ctypes.byref(handle)) == abi.E_INVALID
    assert b"latent" in library.raw("mnf_last_error")()
    assert library.raw("mnf_elbo_fwd_bwd")(None, None, 0, 0, 0, None) == abi.E_INVALID
    assert library.raw("mnf_svi_step")(None, None, None, 0, 0, 0, None) == abi.E_INVALID
    assert library.raw("mnf_xrank_create")(1, 0, 10, ctypes.byref(handle), None) == abi.E_INVALID


def test_dense_kernel_query_is_host_only(library, monkeypatch):
    """mnf_dense_tf32_kernel: which tensor-core kernel a (family, p, S) shape gets - 1 dense_tc.cuh,
    2 dense_tcr.cuh, 3 only the Gram-statistics path (Normal, p <= 64, beyond 128 particles), 0 none."""
    query = library.raw("mnf_dense_tf32_kernel")
    monkeypatch.delenv("MNF_DENSE_NO_GRAM", raising=False)
    monkeypatch.delenv("MNF_DENSE_TC_KERNEL", raising=False)
    assert query(abi.NORMAL, 64, 64) == 1
    assert query(abi.BERNOULLI_LOGITS, 256, 16) == 2
    assert query(abi.BERNOULLI_LOGITS, 64, 16) == 2          # the Bernoulli epilogue prefers 16/32 particle slots
    assert query(abi.POISSON, 64, 32) == 1
    assert query(abi.NORMAL, 128, 100) == 2                  # passes of <= 32 particles
    assert query(abi.NORMAL, 64, 200) == 3 and query(abi.NORMAL, 8, 1000) == 3
    assert query(abi.NORMAL, 66, 200) == 0 and query(abi.NORMAL, 128, 200) == 0
    assert query(abi.POISSON, 64, 200) == 0 and query(abi.NORMAL, 7, 4) == 0
    monkeypatch.setenv("MNF_DENSE_NO_GRAM", "1")
    assert query(abi.NORMAL, 64, 200) == 0 and query(abi.NORMAL, 64, 64) == 1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(abi, "_LIBRARY", None)
    monkeypatch.setattr(build, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        abi.load()


def test_build_units_track_their_own_headers():
    """Every translation unit is rebuilt when (and only when) a header it includes changes."""
    assert set(build.UNITS) == {path.stem for path in build.CSRC_DIR.glob("*.cu")}
    closures = {}
    for unit in build.UNITS:
        seen = {}
        build._closure(build.CSRC_DIR / f"{unit}.cu", seen)
        closures[unit] = {path.name for path in seen}
        assert {"common.cuh", "host.h", "mininf_b200.h", f"{unit}.cu"} <= closures[unit]
    assert "site_sweep.cuh" in closures["site"] and "site_sweep.cuh" not in closures["dense"]
    assert {"dense_tc.cuh", "dense_tcr.cuh", "dense_simt.cuh"} <= closures["dense"]
    assert "rowlatent.cuh" in closures["rowlatent"] and "small.cuh" in closures["abi"]
    assert len({build.unit_digest(unit) for unit in build.UNITS}) == len(build.UNITS)


def test_the_c_consumer_compiles_and_links_against_the_header_and_the_library(tmp_path):
    """tests/c/plan_step.c (run on a B200 by tests/test_engine_gpu.py) must at least build here: a
    plain C11 translation unit over include/mininf_b200.h, linked against the in-tree library."""
    import shutil
    import subprocess
    from mininf_b200.engine import build
    if shutil.which("gcc") is None or not build.LIB_PATH.exists():
        pytest.skip("needs gcc and the built library")
    root = build.PACKAGE_DIR.parent
    binary = tmp_path / "plan_step"
    result = subprocess.run(["gcc", "-O1", "-std=c11", "-Wall", "-I", str(root / "include"), "-I", "/usr/local/cuda/include",
                             str(root / "tests" / "c" / "plan_step.c"), "-o", str(binary), str(build.LIB_PATH),
                             "-L/usr/local/cuda/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{build.LIB_DIR}",
                             "-Wl,-rpath,/usr/local/cuda/lib64"], capture_output=True, text=True)
    assert result.returncode == 0, result.stderr
    assert binary.exists()
