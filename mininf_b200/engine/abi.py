"""ctypes binding of ``include/mininf_b200.h``.

This is the only place Python touches the native library. Structures mirror the header field by
field; ``load()`` raises if the shared library has not been built - there is no Python or CPU
fallback for the ELBO path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import build as _build

ABI_VERSION = 7

# families (include/mininf_b200.h)
NORMAL, GAMMA, BETA, BERNOULLI_PROBS, BERNOULLI_LOGITS, POISSON = range(6)
T_ID, T_EXP = 0, 1
T_FROZEN = 0x80
DENSE_FP32, DENSE_TF32, DENSE_TF32_CLOSED_FORM, DENSE_F16 = 0, 1, 2, 3
SWEEP_CLOSED_FORM = 1
ST_BAD_PARAM, ST_BAD_VALUE, ST_NONFINITE, ST_XRANK_TIMEOUT, ST_RANGE = 1, 2, 4, 8, 16
STEP_ENTROPY, STEP_PRE, STEP_POST = 1, 2, 4
STEP_ALL = STEP_PRE | STEP_POST
MAX_FUSED_SITES = 4
XRANK_HANDLE_BYTES = 64

E_INVALID, E_UNSUPPORTED, E_CUDA = -1, -2, -3


class Link(C.Structure):
    _fields_ = [
        ("a_const", C.c_float), ("b_const", C.c_float),
        ("a_lat", C.c_int32), ("b_lat", C.c_int32),
        ("a_stride", C.c_int32), ("b_stride", C.c_int32),
        ("x", C.c_void_p), ("x_stride", C.c_int32), ("transform", C.c_int32),
    ]


class Latent(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("numel", C.c_int32), ("offset", C.c_int32),
        ("reserved", C.c_int32), ("p0", C.c_void_p), ("p1", C.c_void_p),
    ]


class Site(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("value_lat", C.c_int32),
        ("value", C.c_void_p), ("mask", C.c_void_p),
        ("numel", C.c_int64), ("scale", C.c_double),
        ("param", Link * 2),
    ]


class DenseSite(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("p", C.c_int32),
        ("n_rows", C.c_int64), ("ldx", C.c_int64),
        ("X", C.c_void_p), ("y", C.c_void_p), ("mask", C.c_void_p),
        ("theta_lat", C.c_int32), ("icpt_lat", C.c_int32),
        ("icpt_const", C.c_float), ("reserved", C.c_int32),
        ("scale", Link), ("weight", C.c_double),
    ]


class RowLatent(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("p", C.c_int32), ("resp_family", C.c_int32),
        ("loc", C.c_void_p), ("scale", C.c_void_p), ("grad_loc", C.c_void_p), ("grad_scale", C.c_void_p),
        ("eps", C.c_void_p), ("prior_loc", Link), ("prior_scale", Link),
        ("feat", C.c_void_p), ("feat_scale", Link), ("resp", C.c_void_p),
        ("beta_lat", C.c_int32), ("icpt_lat", C.c_int32), ("icpt_const", C.c_float),
        ("resp_transform", C.c_int32), ("resp_scale", Link),
    ]


class PlanDesc(C.Structure):
    _fields_ = [
        ("n_particles", C.c_int32), ("n_latent_total", C.c_int32), ("n_latents", C.c_int32),
        ("n_dense", C.c_int32), ("n_groups", C.c_int32), ("n_small_observed", C.c_int32),
        ("n_small_global", C.c_int32), ("n_rowlatent", C.c_int32),
        ("latents", C.POINTER(Latent)), ("dense", C.POINTER(DenseSite)), ("dense_mode", C.POINTER(C.c_int32)),
        ("group_sites", C.POINTER(Site)), ("group_sizes", C.POINTER(C.c_int32)),
        ("small_observed", C.POINTER(Site)), ("small_global", C.POINTER(Site)),
        ("rowlatent", C.POINTER(RowLatent)),
        ("flags", C.c_uint32), ("device", C.c_int32),
    ]


class RowBuffers(C.Structure):
    _fields_ = [("loc", C.c_void_p), ("scale", C.c_void_p), ("grad_loc", C.c_void_p),
                ("grad_scale", C.c_void_p), ("eps", C.c_void_p), ("loc_rw", C.c_void_p),
                ("raw_scale", C.c_void_p), ("m_loc", C.c_void_p), ("v_loc", C.c_void_p),
                ("m_scale", C.c_void_p), ("v_scale", C.c_void_p)]


class Buffers(C.Structure):
    _fields_ = [
        ("z", C.c_void_p), ("noise", C.c_void_p), ("acc", C.c_void_p), ("out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("status", C.c_void_p),
        ("noise_in", C.c_void_p), ("step_counter", C.c_void_p),
        ("rows", C.POINTER(RowBuffers)), ("xrank", C.c_void_p),
    ]


class Adam(C.Structure):
    _fields_ = [
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("raw", C.c_void_p), ("transform", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
        ("constrained", C.c_void_p), ("step", C.c_void_p),
    ]


class PredSite(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("family", C.c_int32), ("numel", C.c_int64), ("out_col", C.c_int32),
        ("transform", C.c_int32), ("param", Link * 2), ("X", C.c_void_p), ("ldx", C.c_int64),
        ("p", C.c_int32), ("theta_lat", C.c_int32), ("icpt_lat", C.c_int32), ("icpt_const", C.c_float),
    ]


PRED_DRAW, PRED_VALUE = 0, 1


class DeviceInfo(C.Structure):
    _fields_ = [
        ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
        ("max_smem_optin", C.c_int32), ("total_mem", C.c_int64),
    ]


def const_link(value: float) -> Link:
    return Link(a_const=float(value), b_const=0.0, a_lat=-1, b_lat=-1, a_stride=0, b_stride=0,
                x=None, x_stride=0, transform=T_ID)


EXPORTS = {
    "mnf_abi_version": (C.c_int, []),
    "mnf_last_error": (C.c_char_p, []),
    "mnf_device_info": (C.c_int, [C.c_int, C.POINTER(DeviceInfo)]),
    "mnf_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "mnf_rsample": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                              C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p]),
    "mnf_dense_tf32_kernel": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "mnf_dense_sweep": (C.c_int, [C.POINTER(DenseSite), C.c_int, C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "mnf_site_sweep": (C.c_int, [C.POINTER(Site), C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p]),
    "mnf_rowlatent_sweep": (C.c_int, [C.POINTER(RowLatent), C.c_void_p, C.c_int, C.c_int, C.c_uint64,
                                      C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p, C.c_void_p]),
    "mnf_small_sites": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "mnf_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mnf_masked_count": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "mnf_plan_create": (C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]),
    "mnf_plan_update": (C.c_int, [C.c_void_p, C.POINTER(PlanDesc), C.c_void_p]),
    "mnf_plan_workspace_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "mnf_plan_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "mnf_plan_destroy": (C.c_int, [C.c_void_p]),
    "mnf_elbo_fwd_bwd": (C.c_int, [C.c_void_p, C.POINTER(Buffers), C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]),
    "mnf_svi_step": (C.c_int, [C.c_void_p, C.POINTER(Buffers), C.POINTER(Adam), C.c_uint64, C.c_uint64,
                               C.c_uint32, C.c_void_p]),
    "mnf_xrank_create": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "mnf_xrank_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mnf_xrank_destroy": (C.c_int, [C.c_void_p]),
    "mnf_predictive": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64,
                                 C.c_void_p, C.c_void_p]),
}


class NativeError(RuntimeError):
    """A C-ABI call returned a negative code."""

    def __init__(self, function: str, code: int, message: str) -> None:
        super().__init__(f"{function} failed with code {code}: {message}")
        self.function = function
        self.code = code
        self.message = message


class Library:
    """The loaded shared library with typed, checked entry points."""

    def __init__(self, path: Path) -> None:
        self.path = Path(path)
        self._dll = C.CDLL(str(path))
        for name, (restype, argtypes) in EXPORTS.items():
            function = getattr(self._dll, name)  # AttributeError if a symbol is missing
            function.restype = restype
            function.argtypes = argtypes
        version = self._dll.mnf_abi_version()
        if version != ABI_VERSION:
            raise RuntimeError(f"{path} has ABI version {version}, expected {ABI_VERSION}; "
                               "rebuild with `python -m mininf_b200.engine.build --force`")

    def raw(self, name: str):
        return getattr(self._dll, name)

    def call(self, name: str, *args) -> None:
        code = getattr(self._dll, name)(*args)
        if code != 0:
            raise NativeError(name, code, self._dll.mnf_last_error().decode())

    def device_info(self, device: int = -1) -> DeviceInfo:
        info = DeviceInfo()
        self.call("mnf_device_info", device, C.byref(info))
        return info

    def workspace_bytes(self, n_particles: int, n_latent_total: int, device: int = -1) -> int:
        return int(self._dll.mnf_workspace_bytes(n_particles, n_latent_total, device))


_LIBRARY: Library | None = None


def load(build_if_missing: bool = False) -> Library:
    """Load the native library. Raises ``RuntimeError`` if it has not been built, or if it was
    built from other sources than the ones in ``csrc/`` now (a stale binary with the same ABI
    version would otherwise load silently). ``MNF_LIB=/path/to/lib.so`` loads a developer build."""
    global _LIBRARY
    if _LIBRARY is None:
        import os
        override = os.environ.get("MNF_LIB")
        if override:
            _LIBRARY = Library(Path(override))
            return _LIBRARY
        if not _build.LIB_PATH.exists():
            if build_if_missing:
                _build.build()
            else:
                raise RuntimeError(
                    f"the mininf_b200 CUDA library is missing ({_build.LIB_PATH}); build it with "
                    "`python -m mininf_b200.engine.build` - there is no CPU fallback for the ELBO "
                    "path")
        if _build.STAMP_PATH.exists() and _build.CSRC_DIR.exists() and not _build.is_current():
            raise RuntimeError(
                f"{_build.LIB_PATH} was built from different sources than mininf_b200/csrc holds now; "
                "rebuild it with `python -m mininf_b200.engine.build`")
        _LIBRARY = Library(_build.LIB_PATH)
    return _LIBRARY
