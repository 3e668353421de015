"""ctypes wrapper of oracle/libepi_oracle.so — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this.  PARITY UNPINNED versus executed reference
output (R/deSolve/drjacoby are absent; see epi_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from epi_mcmc_b200 import _capi  # noqa: E402  (struct layouts of include/epimcmc.h only)

LIB = os.path.join(HERE, "libepi_oracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lib = None


def build(force: bool = False):
    src = os.path.join(HERE, "epi_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        for f in ("oracle_pgamma", "oracle_pnorm", "oracle_dnorm_log", "oracle_dlnorm_log",
                  "oracle_dnbinom_mu_log"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double] * 3
        L.oracle_age_gamma.restype = C.c_double
        L.oracle_derivs_age.argtypes = [_dp, C.c_double, _dp, _dp]
        L.oracle_derivs_agef.argtypes = [_dp, C.c_double, _dp, _dp]
        L.oracle_derivs_ager.argtypes = [_dp, C.c_double, _dp, _dp, _dp]
        L.oracle_derivs_ager.restype = None
        L.oracle_df_params.argtypes = [C.POINTER(_capi.ModelDesc), C.POINTER(_capi.Data), _dp, _dp, _dp]
        L.oracle_eval.argtypes = [C.POINTER(_capi.ModelDesc), C.POINTER(_capi.Data), _dp, C.c_int64,
                                  C.c_int, _dp, _dp, _ip, _ip, _ip, _dp]
        L.oracle_trajectory.argtypes = [C.POINTER(_capi.ModelDesc), C.POINTER(_capi.Data), _dp, C.c_int,
                                        _dp, _dp, _dp, _ip, _ip]
        L.oracle_yout_age.argtypes = [_dp, C.c_double, _dp, _dp]
        L.oracle_yout_agef.argtypes = [_dp, C.c_double, _dp, _dp]
        L.oracle_profile.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _ip, _ip]
        L.oracle_set_jitter.argtypes = [C.c_int, C.c_uint]
        L.oracle_set_jitter.restype = None
        L.oracle_set_mirror.argtypes = [C.c_int]
        L.oracle_set_mirror.restype = None
        _lib = L
    return _lib


def _desc(flavour, period, substeps):
    if isinstance(flavour, str):
        flavour = _capi.FLAVOURS[flavour]
    return _capi.ModelDesc(flavour, period, substeps, 0)


def df_params(flavour, data: dict, period=300, substeps=4):
    desc = _desc(flavour, period, substeps)
    D, keep = _capi.make_data(data)
    d = _capi.N_PARAMS[desc.flavour]
    mn, mx, init = (np.zeros(d) for _ in range(3))
    lib().oracle_df_params(C.byref(desc), C.byref(D), _capi.as_dp(mn), _capi.as_dp(mx), _capi.as_dp(init))
    return mn, mx, init


def evaluate(flavour, data: dict, theta, period, substeps=4, n_threads=1):
    """theta: (n, d) sampler-space.  Returns dict(logl, logp, status, offset, padding, terms)."""
    desc = _desc(flavour, period, substeps)
    D, keep = _capi.make_data(data)
    theta = np.ascontiguousarray(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
    n, d = theta.shape
    assert d == _capi.N_PARAMS[desc.flavour], (d, desc.flavour)
    logl, logp = np.zeros(n), np.zeros(n)
    status, offset, padding = (np.zeros(n, dtype=np.int32) for _ in range(3))
    terms = np.zeros((n, 8))
    lib().oracle_eval(C.byref(desc), C.byref(D), _capi.as_dp(theta), n, n_threads, _capi.as_dp(logl),
                      _capi.as_dp(logp), _capi.as_ip(status), _capi.as_ip(offset), _capi.as_ip(padding),
                      _capi.as_dp(terms))
    return dict(logl=logl, logp=logp, status=status, offset=offset, padding=padding, terms=terms)


def trajectory(flavour, data: dict, theta_model, period, substeps=4, calc_period=None, ext=False):
    """calculateModel(theta_model, calc_period).  Returns (series[n_series, P], states[P, neq], offset, padding);
    with ext=True the series also hold the state fields derived from the ode() matrix (age: y.E, y.I, y.R,
    o.E, o.I, o.R, Re, Rt, y.Re, o.Re; simple: E, I, R, Re, Rt)."""
    desc = _desc(flavour, period, substeps)
    D, keep = _capi.make_data(data)
    P = calc_period or period
    ns = _capi.N_SERIES[desc.flavour]
    neq = 10 if desc.flavour in _capi.AGE_FLAVOURS else 4
    th = np.ascontiguousarray(np.asarray(theta_model, dtype=np.float64))
    out, states = np.zeros((ns, P)), np.zeros((P, neq))
    n_ext = 10 if neq == 10 else 5
    extra = np.zeros((n_ext, P))
    off, pad = C.c_int32(0), C.c_int32(0)
    rc = lib().oracle_trajectory(C.byref(desc), C.byref(D), _capi.as_dp(th), P, _capi.as_dp(out),
                                 _capi.as_dp(states), _capi.as_dp(extra) if ext else None, C.byref(off), C.byref(pad))
    if rc:
        raise ValueError("latency profile too wide")
    if ext:
        out = np.vstack([out, extra])
    return out, states, off.value, pad.value


def conditioning(flavour, data, theta, period, substeps, logl, ulps=32, n_draws=3, n_threads=8):
    """How ill-conditioned the reference algorithm is at each theta: the largest change of
    logl when every reported ODE state is perturbed by up to `ulps` ulps (oracle_set_jitter).
    The parity bar of the GPU tests is max(1e-10 * |logl|, 4 x this)."""
    cond = np.zeros(len(logl))
    try:
        for k in range(n_draws):
            lib().oracle_set_jitter(ulps, 1000 + k)
            r = evaluate(flavour, data, theta, period, substeps, n_threads)
            with np.errstate(invalid="ignore"):
                dlt = np.abs(r["logl"] - logl)
            cond = np.maximum(cond, np.where(np.isfinite(dlt), dlt, 0.0))
            # the jitter turns a finite value into NA or back (the hospital ratio term 0/0 of two sums of
            # cancellation residue): nothing about this theta is determined beyond rounding
            cond = np.where(np.isnan(r["logl"]) != np.isnan(logl), np.inf, cond)
    finally:
        lib().oracle_set_jitter(0, 0)
    return cond


def set_mirror(on: bool):
    """tests only: the ODE scheme in the CUDA kernels' operation order (explicit fma) — the ODE states are then
    bit-identical to the GPU's, see epi_oracle.c"""
    lib().oracle_set_mirror(1 if on else 0)


def profile(kind, mean, sd):
    v = np.zeros(_capi.MAX_PARAMS + 32)
    kb, ke = C.c_int32(0), C.c_int32(0)
    n = lib().oracle_profile(kind, mean, sd, _capi.as_dp(v), C.byref(kb), C.byref(ke))
    if n < 0:
        raise ValueError("profile too wide")
    return v[:n].copy(), kb.value, ke.value


def derivs_age(parms45, t, y):
    p = np.ascontiguousarray(parms45, dtype=np.float64)
    yy = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros(10)
    lib().oracle_derivs_age(_capi.as_dp(p), t, _capi.as_dp(yy), _capi.as_dp(out))
    return out


def yout_age(parms, t, y):
    """Re, Rt, y.Re, o.Re of the restated RHS; 45 parms = model-agen.cpp, 37 = model-agef.cpp"""
    p = np.ascontiguousarray(parms, dtype=np.float64)
    yy = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros(4)
    (lib().oracle_yout_age if p.size == 45 else lib().oracle_yout_agef)(_capi.as_dp(p), t, _capi.as_dp(yy), _capi.as_dp(out))
    return out


def derivs_agef(parms37, t, y):
    p = np.ascontiguousarray(parms37, dtype=np.float64)
    yy = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros(10)
    lib().oracle_derivs_agef(_capi.as_dp(p), t, _capi.as_dp(yy), _capi.as_dp(out))
    return out


def derivs_ager(parms60, t, y):
    """(ydot[10], yout[9]) of the restated model-ager.cpp RHS"""
    p = np.ascontiguousarray(parms60, dtype=np.float64)
    yy = np.ascontiguousarray(y, dtype=np.float64)
    out, yo = np.zeros(10), np.zeros(9)
    lib().oracle_derivs_ager(_capi.as_dp(p), t, _capi.as_dp(yy), _capi.as_dp(out), _capi.as_dp(yo))
    return out, yo


class RefRHS:
    """The reference's own compiled RHS (oracle/_ref/model-agen.so, built by
    oracle/Makefile from /root/reference/R/models/model-agen.cpp with a stub
    R.h) driven through its deSolve ABI: initmod(odeparms) + derivs(...)."""

    def __init__(self, name="model-agen", n_parms=45):
        path = os.path.join(HERE, "_ref", name + ".so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.L = C.CDLL(path)
        self.n = n_parms
        self._parms = None
        self.ODEPARMS = C.CFUNCTYPE(None, C.POINTER(C.c_int), _dp)

    def init(self, parms):
        p = np.ascontiguousarray(parms, dtype=np.float64)
        assert p.size == self.n

        def odeparms(n_ptr, dst):
            for i in range(n_ptr[0]):
                dst[i] = p[i]

        cb = self.ODEPARMS(odeparms)
        self.L.initmod(cb)

    def derivs(self, t, y, nout=4):
        neq = C.c_int(10)
        tt = C.c_double(t)
        yy = np.ascontiguousarray(y, dtype=np.float64)
        ydot = np.zeros(10)
        yout = np.zeros(max(nout, 4))
        ip = (C.c_int * 3)(nout, 0, 0)
        self.L.derivs(C.byref(neq), C.byref(tt), _capi.as_dp(yy), _capi.as_dp(ydot), _capi.as_dp(yout), ip)
        return ydot, yout
