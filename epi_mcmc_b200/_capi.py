"""ctypes view of include/epimcmc.h and the loader for libepimcmc.so.

There is no CPU fallback: `load()` raises if the CUDA library has not been
built (`python -c "import __graft_entry__ as g; g.build()"`), and every compute
entry point of the library itself fails with EPI_ERR_NO_DEVICE without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libepimcmc.so")

FLAVOUR_SIMPLE, FLAVOUR_NE, FLAVOUR_AGE2Q, FLAVOUR_AGE2I, FLAVOUR_AGE2X = 0, 1, 2, 3, 4
FLAVOURS = {"simple": FLAVOUR_SIMPLE, "ne": FLAVOUR_NE, "age2q": FLAVOUR_AGE2Q, "age2i": FLAVOUR_AGE2I,
            "age2x": FLAVOUR_AGE2X}
AGE_FLAVOURS = (FLAVOUR_AGE2Q, FLAVOUR_AGE2I, FLAVOUR_AGE2X)
N_PARAMS = {FLAVOUR_SIMPLE: 8, FLAVOUR_NE: 9, FLAVOUR_AGE2Q: 45, FLAVOUR_AGE2I: 36, FLAVOUR_AGE2X: 52}
N_SERIES = {FLAVOUR_SIMPLE: 3, FLAVOUR_NE: 3, FLAVOUR_AGE2Q: 8, FLAVOUR_AGE2I: 8, FLAVOUR_AGE2X: 8}
MAX_PARAMS = 56
INVALID_DATA_OFFSET = 10000

ST_INVALID_OFFSET, ST_NA, ST_OUT_OF_BOX, ST_UNSUPPORTED = 1, 2, 8, 16
LAYOUT_AOS, LAYOUT_SOA = 0, 1
SAMPLER_RWM, SAMPLER_DEMC, SAMPLER_DRJ = 0, 1, 2
SAMPLER_FLAG_FULLDIM, SAMPLER_FLAG_REPLICATES, SAMPLER_FLAG_NO_SINGLE = 1, 2, 4

ERR_NAMES = {0: "EPI_OK", 1: "EPI_ERR_ARG", 2: "EPI_ERR_NO_DEVICE", 3: "EPI_ERR_CUDA",
             4: "EPI_ERR_NOMEM", 5: "EPI_ERR_STATE"}

_dp = C.POINTER(C.c_double)


class ModelDesc(C.Structure):
    _fields_ = [("flavour", C.c_int32), ("period", C.c_int32), ("substeps", C.c_int32),
                ("reserved", C.c_int32)]


class Data(C.Structure):
    _fields_ = [
        ("N", C.c_double), ("y_N", C.c_double), ("o_N", C.c_double),
        ("total_deaths_at_lockdown", C.c_double), ("dmort_last", C.c_double),
        ("lockdown_offset", C.c_int32), ("lockdown_transition_period", C.c_int32),
        ("d3", C.c_int32), ("d4", C.c_int32), ("d5", C.c_int32), ("d6", C.c_int32),
        ("d7", C.c_int32), ("d8", C.c_int32),
        ("d_reliable_cases", C.c_int32), ("d_reliable_hosp", C.c_int32),
        ("d_hosp_o1", C.c_int32), ("d_hosp_o2", C.c_int32),
        ("n_dhospi", C.c_int32), ("n_dhosp", C.c_int32), ("n_dmorti", C.c_int32),
        ("n_dcasei", C.c_int32), ("n_rate", C.c_int32), ("reserved", C.c_int32),
        ("dhospi", _dp), ("dmorti", _dp),
        ("y_dcasei", _dp), ("o_dcasei", _dp),
        ("y_dmorti", _dp), ("o_dmorti", _dp), ("o_wdmorti", _dp),
        ("y_ifr", _dp), ("o_ifr", _dp), ("y_hr", _dp), ("o_hr", _dp), ("g", _dp), ("gtrimp", _dp),
        ("hosp_nbinom_size", C.c_double), ("mort_nbinom_size", C.c_double),
        ("case_nbinom_size1", C.c_double), ("case_nbinom_sizey2", C.c_double),
        ("case_nbinom_sizeo2", C.c_double), ("hosp_nbinom_size1", C.c_double),
        ("hosp_nbinom_size2", C.c_double), ("hosp_last7", C.c_double),
        ("d9", C.c_int32), ("d10", C.c_int32), ("d12", C.c_int32), ("duncertain", C.c_int32),
        ("d_symp_cases", C.c_int32), ("d_all_cases", C.c_int32), ("symp_cases_factor", C.c_double),
    ]


class SamplerCfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_chains", C.c_int32), ("chain_id0", C.c_int64),
                ("seed", C.c_uint64), ("scale", C.c_double), ("de_eps", C.c_double),
                ("beta", C.c_double), ("n_rungs", C.c_int32), ("reserved", C.c_int32), ("gti_pow", C.c_double)]


_SERIES_FIELDS = ("dhospi", "dmorti", "y_dcasei", "o_dcasei", "y_dmorti", "o_dmorti", "o_wdmorti",
                  "y_ifr", "o_ifr", "y_hr", "o_hr", "g", "gtrimp")
_SCALAR_FIELDS = tuple(n for n, _ in Data._fields_ if n not in _SERIES_FIELDS and n != "reserved")


def make_data(d: dict):
    """dict (dataset JSON / ModelData.as_dict()) -> (Data struct, keepalive list of arrays)."""
    s = Data()
    keep = []
    for name in _SCALAR_FIELDS:
        if name in d and d[name] is not None:
            setattr(s, name, d[name])
    for name in _SERIES_FIELDS:
        v = d.get(name)
        if v is None:
            setattr(s, name, None)
            continue
        a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
        keep.append(a)
        setattr(s, name, a.ctypes.data_as(_dp))
    return s, keep


def as_dp(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def as_ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


_lib = None

# every symbol include/epimcmc.h declares (tests check the library exports them all)
SYMBOLS = (
    "epi_create", "epi_destroy", "epi_last_error", "epi_abi_version", "epi_n_params",
    "epi_param_name", "epi_df_params", "epi_eval", "epi_eval_submit", "epi_eval_wait", "epi_eval_device", "epi_eval_device_range", "epi_reserve", "epi_eval_detail",
    "epi_n_series", "epi_trajectories", "epi_n_series_ext", "epi_trajectories_ext", "epi_quantiles", "epi_marginalize", "epi_sampler_init", "epi_sampler_run",
    "epi_sampler_adapt", "epi_sampler_set_scale", "epi_sampler_state_device", "epi_sampler_set_population",
    "epi_sampler_get", "epi_sampler_stats", "epi_sampler_pt_stats", "epi_sampler_record", "epi_sampler_draws", "epi_sampler_draws_subset",
    "epi_sampler_draws_device", "epi_sampler_bandwidths", "epi_k2_bench",
    "epi_fp64_peak", "epi_set_timing", "epi_last_kernel_ms", "epi_launch_count", "epi_fallback_count", "epi_philox_probe",
)


def load():
    """Load libepimcmc.so (the CUDA product library).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the evaluation path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    ip = C.POINTER(C.c_int32)
    lib.epi_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(Data), C.c_int, C.POINTER(vp)]
    lib.epi_destroy.argtypes = [vp]
    lib.epi_destroy.restype = None
    lib.epi_last_error.argtypes = [vp]
    lib.epi_last_error.restype = C.c_char_p
    lib.epi_n_params.argtypes = [vp]
    lib.epi_param_name.argtypes = [vp, C.c_int]
    lib.epi_param_name.restype = C.c_char_p
    lib.epi_df_params.argtypes = [vp, _dp, _dp, _dp]
    lib.epi_eval.argtypes = [vp, _dp, i64, C.c_int, _dp, _dp, ip]
    lib.epi_eval_submit.argtypes = [vp, C.c_int, _dp, i64, C.c_int, _dp, _dp, ip]
    lib.epi_eval_wait.argtypes = [vp, C.c_int]
    lib.epi_eval_device.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.epi_eval_device_range.argtypes = [vp, vp, i64, i64, i64, i64, vp, vp, vp, vp]
    lib.epi_reserve.argtypes = [vp, i64]
    lib.epi_eval_detail.argtypes = [vp, _dp, i64, C.c_int, ip, ip, _dp]
    lib.epi_n_series.argtypes = [vp]
    lib.epi_trajectories.argtypes = [vp, _dp, i64, C.c_int, i32, _dp, ip, ip]
    lib.epi_n_series_ext.argtypes = [vp]
    lib.epi_trajectories_ext.argtypes = [vp, _dp, i64, C.c_int, i32, _dp, ip, ip]
    lib.epi_quantiles.argtypes = [vp, _dp, i64, C.c_int, i32, i32, i32, i32, _dp, i32, _dp]
    lib.epi_marginalize.argtypes = [vp, _dp, i64, C.c_int, i32, i32, i32, i32, i32, i32, i32, _dp]
    lib.epi_sampler_draws_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), ip, C.POINTER(i64)]
    lib.epi_sampler_bandwidths.argtypes = [vp, _dp]
    lib.epi_k2_bench.argtypes = [vp, i32, i64, i32, _dp, _dp, _dp]
    lib.epi_sampler_init.argtypes = [vp, C.POINTER(SamplerCfg), _dp]
    lib.epi_sampler_run.argtypes = [vp, i32]
    lib.epi_sampler_adapt.argtypes = [vp, C.c_int, C.c_double]
    lib.epi_sampler_set_scale.argtypes = [vp, C.c_double]
    lib.epi_sampler_state_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
    lib.epi_sampler_set_population.argtypes = [vp, vp, i32, i32, i64]
    lib.epi_philox_probe.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, i32, C.POINTER(C.c_uint32)]
    lib.epi_sampler_get.argtypes = [vp, _dp, _dp, _dp]
    lib.epi_sampler_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.epi_sampler_pt_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.epi_sampler_record.argtypes = [vp, i32, i32]
    lib.epi_sampler_draws.argtypes = [vp, _dp, _dp, ip]
    lib.epi_sampler_draws_subset.argtypes = [vp, i32, i32, _dp, _dp, ip]
    lib.epi_fp64_peak.argtypes = [vp, _dp]
    lib.epi_set_timing.argtypes = [vp, C.c_int]
    lib.epi_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.epi_launch_count.argtypes = [vp]
    lib.epi_launch_count.restype = i64
    lib.epi_fallback_count.argtypes = [vp]
    lib.epi_fallback_count.restype = i64
    _lib = lib
    return lib


class EpiError(RuntimeError):
    pass


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().epi_last_error(handle)
        raise EpiError(f"{ERR_NAMES.get(rc, rc)}: {msg.decode() if msg else ''}")
