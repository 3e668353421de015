"""Host-side mirror of the reference's model-file API over the C ABI.

A reference model file (R/models/model-age-groups2q.R, model-simple-ode-drj-cmp.R)
defines — as R globals — `calclogl`, `calclogp`, `transformParams`,
`invTransformParams`, `calculateModel`, `calcNominalState`, `fit.paramnames`,
`keyparamnames`, `fitkeyparamnames`, `init`, `df_params`.  `Model` keeps exactly
those names (dots become underscores) and the same argument meaning, and adds
the batched entry points the GPU is for (`calclogl_batch`, `calclogpost_batch`,
`calculateModel_batch`).  Everything is computed by libepimcmc.so (CUDA); there
is no CPU path here.  R is not available in the build image, so this module is
what the tests drive; the R `.Call` shim that binds the same C entry points is in
r/ (see INTEGRATION.md).
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import _capi

DATASET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "datasets")

AGE_SERIES = ("y.S", "o.S", "y.casei", "o.casei", "y.hospi", "o.hospi", "y.deadi", "o.deadi")
SIMPLE_SERIES = ("S", "hospi", "deadi")
# the state fields derived from the other columns of the ode() matrix (epi_trajectories_ext)
AGE_SERIES_EXT = AGE_SERIES + ("y.E", "y.I", "y.R", "o.E", "o.I", "o.R", "Re", "Rt", "y.Re", "o.Re")
SIMPLE_SERIES_EXT = SIMPLE_SERIES + ("E", "I", "R", "Re", "Rt")


def load_dataset(name: str) -> dict:
    """One of the committed datasets (epi_mcmc_b200/datasets/*.json): the R globals a
    data file + settings.R would define (see tests/golden/make_fixtures.py)."""
    with open(os.path.join(DATASET_DIR, name + ".json")) as f:
        return json.load(f)


class Model:
    """One sourced model file + data file + settings.R, bound to one GPU."""

    def __init__(self, data: dict, flavour: str | None = None, period: int | None = None,
                 substeps: int = 4, device: int = 0):
        self._lib = _capi.load()
        self.data = data
        self.flavour = flavour or data["flavour"]
        self.FitTotalPeriod = int(period or data["period"])
        self.substeps = int(substeps)
        self.device = device
        desc = _capi.ModelDesc(_capi.FLAVOURS[self.flavour], self.FitTotalPeriod, self.substeps, 0)
        D, self._keep = _capi.make_data(data)
        h = C.c_void_p()
        _capi.check(self._lib.epi_create(C.byref(desc), C.byref(D), device, C.byref(h)))
        self._h = h
        self.d = self._lib.epi_n_params(h)
        self.fit_paramnames = [self._lib.epi_param_name(h, i).decode() for i in range(self.d)]
        mn, mx, init = (np.zeros(self.d) for _ in range(3))
        _capi.check(self._lib.epi_df_params(h, _capi.as_dp(mn), _capi.as_dp(mx), _capi.as_dp(init)), h)
        self.init = init
        # df_params: data.frame(name, min, max, init), model-age-groups2q.R:637-663
        self.df_params = {"name": list(self.fit_paramnames), "min": mn, "max": mx, "init": init.copy()}
        self.is_age = self.flavour in ("age2q", "age2i", "age2x")
        self.series_names_ext = AGE_SERIES_EXT if self.is_age else SIMPLE_SERIES_EXT
        if self.flavour == "age2q":  # model-age-groups2q.R:619-622
            self.keyparamnames = ["betay6", "betao6", "betayo6", "betay7", "betao7", "betayo7",
                                  "fy9", "fo9", "fyo9", "ifrred"]
            self.fitkeyparamnames = list(self.keyparamnames)
            self.series_names = AGE_SERIES
        elif self.flavour == "age2x":  # model-age-groups2x.R:699-703
            self.keyparamnames = ["betay6", "betao6", "betayo6", "betay7", "betao7", "betayo7",
                                  "betay9", "betao9", "betayo9", "ifrred"]
            self.fitkeyparamnames = list(self.keyparamnames)
            self.series_names = AGE_SERIES
        elif self.flavour == "age2i":  # model-age-groups2i.R:550-551
            self.keyparamnames = ["betay6", "betao6", "betayo6", "fy7", "fo7", "fyo7"]
            self.fitkeyparamnames = list(self.keyparamnames)
            self.series_names = AGE_SERIES
        else:  # model-simple-ode-drj-cmp.R:261-262
            self.keyparamnames = ["beta0", "betat", "R0", "Rt", "Tinf0", "Tinft"]
            self.fitkeyparamnames = ["R0", "Rt", "Tinf0", "Tinft"]
            self.series_names = SIMPLE_SERIES
        self.it = 0  # the reference's global evaluation counter (libfit.R:28-32)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.epi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- transformParams / invTransformParams -----------------------------------
    def transformParams(self, params):
        """model-age-groups2q.R:339-342 (identity); model-simple-ode-drj-cmp.R:136-142."""
        r = np.array(params, dtype=np.float64, copy=True)
        if not self.is_age:
            r[..., 4] = np.exp(r[..., 4])
        return r

    def _untransform(self, params):
        r = np.array(params, dtype=np.float64, copy=True)
        if not self.is_age:
            r[..., 4] = np.log(r[..., 4])
        return r

    def invTransformParams(self, posterior: dict) -> dict:
        """Derived columns of a posterior table (dict of arrays keyed by parameter name)."""
        p = dict(posterior)
        if self.is_age:  # model-age-groups2q.R:344-365, model-age-groups2i.R:318-351
            gamma = 1 / ((4.7 - 4) * 2)
            p["Tinf"] = 1 / gamma
            for k in ("0", "1", "2", "4") + (("6",) if self.flavour == "age2i" else ()):
                for g, pre in (("y", "betay"), ("o", "betao"), ("yo", "betayo")):
                    p[f"{g}.Rt{k}"] = np.asarray(p[pre + k]) / gamma
            if self.flavour == "age2x":  # model-age-groups2x.R:407-409
                p["y.ld2"] = np.asarray(p["betay11"]) / np.asarray(p["betay9"])
                p["o.ld2"] = np.asarray(p["betao11"]) / np.asarray(p["betao9"])
                p["yo.ld2"] = np.asarray(p["betayo11"]) / np.asarray(p["betayo9"])
            if self.flavour == "age2i":  # :342-349
                p["betay7"] = np.asarray(p["betay6"]) * np.asarray(p["fy7"])
                p["betao7"] = np.asarray(p["betao6"]) * np.asarray(p["fo7"])
                p["betayo7"] = np.asarray(p["betayo6"]) * np.asarray(p["fyo7"])
                for g, pre in (("y", "betay"), ("o", "betao"), ("yo", "betayo")):
                    p[f"{g}.Rt7"] = np.asarray(p[pre + "7"]) / gamma
        else:  # model-simple-ode-drj-cmp.R:144-154
            p["Tinc"] = 3
            p["beta0"] = np.asarray(p["R0"]) / np.asarray(p["Tinf0"])
            p["betat"] = np.asarray(p["Rt"]) / np.asarray(p["Tinft"])
            p["HR"] = np.exp(np.asarray(p["logHR"]))
        return p

    # ---- batched hot path ---------------------------------------------------------
    def calclogpost_batch(self, theta, layout="aos"):
        """theta: sampler-space proposals, (n, d) for layout 'aos' (one proposal per row —
        the memory image of an R d x n matrix) or (d, n) for 'soa'.  Returns (logl, logp,
        status): logl[i] = calclogl(transformParams(theta_i)), logp[i] = calclogp(theta_i),
        exactly what drjacoby is handed (fitMCMC-drj.R:19-21,51)."""
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
        if th.ndim == 1:
            th = th[None, :]
        n = th.shape[0] if layout == "aos" else th.shape[1]
        assert (th.shape[1] if layout == "aos" else th.shape[0]) == self.d, th.shape
        logl, logp = np.empty(n), np.empty(n)
        status = np.empty(n, dtype=np.int32)
        _capi.check(self._lib.epi_eval(self._h, _capi.as_dp(th), n,
                                       _capi.LAYOUT_AOS if layout == "aos" else _capi.LAYOUT_SOA,
                                       _capi.as_dp(logl), _capi.as_dp(logp), _capi.as_ip(status)), self._h)
        self.it += n
        return logl, logp, status

    def calclogpost_submit(self, slot: int, theta, layout="aos"):
        """Asynchronous calclogpost_batch (epi_eval_submit): queue the batch on `slot` (0 or 1) and
        return; `calclogpost_wait(slot)` hands back (logl, logp, status).  With two slots the input
        copy of the next batch overlaps the kernels of the current one."""
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
        if th.ndim == 1:
            th = th[None, :]
        n = th.shape[0] if layout == "aos" else th.shape[1]
        assert (th.shape[1] if layout == "aos" else th.shape[0]) == self.d, th.shape
        logl, logp = np.empty(n), np.empty(n)
        status = np.empty(n, dtype=np.int32)
        _capi.check(self._lib.epi_eval_submit(self._h, slot, _capi.as_dp(th), n,
                                              _capi.LAYOUT_AOS if layout == "aos" else _capi.LAYOUT_SOA,
                                              _capi.as_dp(logl), _capi.as_dp(logp), _capi.as_ip(status)), self._h)
        if not hasattr(self, "_pending"):
            self._pending = {}
        self._pending[slot] = (th, logl, logp, status)  # keep the host buffers alive until the wait
        self.it += n

    def calclogpost_wait(self, slot: int):
        _capi.check(self._lib.epi_eval_wait(self._h, slot), self._h)
        _, logl, logp, status = self._pending.pop(slot)
        return logl, logp, status

    def calclogl_batch(self, params_model):
        """calclogl over MODEL-space rows (already through transformParams)."""
        return self.calclogpost_batch(self._untransform(np.atleast_2d(params_model)))[0]

    def eval_detail(self, theta):
        th = np.ascontiguousarray(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
        n = th.shape[0]
        off, pad = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        terms = np.empty((n, 8))
        _capi.check(self._lib.epi_eval_detail(self._h, _capi.as_dp(th), n, _capi.LAYOUT_AOS, _capi.as_ip(off),
                                              _capi.as_ip(pad), _capi.as_dp(terms)), self._h)
        return off, pad, terms

    # ---- the reference's scalar API -------------------------------------------------
    def calclogl(self, params, x=None):
        """calclogl(params, x): params in MODEL space (the driver passes transformParams(theta))."""
        return float(self.calclogl_batch(np.asarray(params, dtype=np.float64)[None, :])[0])

    def calclogp(self, params, misc=None):
        """calclogp(params): the UNTRANSFORMED sampler vector (fitMCMC-drj.R:51)."""
        return float(self.calclogpost_batch(np.asarray(params, dtype=np.float64)[None, :])[1][0])

    def calculateModel_batch(self, params_model, period: int | None = None, ext: bool = False):
        """calculateModel(params, period) for each MODEL-space row.  Returns
        (series[n, n_series, period], offset[n], padding[n]); series cover days
        padding+1..padding+period in the order `self.series_names` (`self.series_names_ext` with
        ext=True: also E, I, R and the Re/Rt diagnostics)."""
        th = np.ascontiguousarray(np.atleast_2d(np.asarray(params_model, dtype=np.float64)))
        n = th.shape[0]
        P = int(period or self.FitTotalPeriod)
        ns = self._lib.epi_n_series_ext(self._h) if ext else self._lib.epi_n_series(self._h)
        out = np.empty((n, ns, P))
        off, pad = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        fn = self._lib.epi_trajectories_ext if ext else self._lib.epi_trajectories
        _capi.check(fn(self._h, _capi.as_dp(th), n, _capi.LAYOUT_AOS, P, _capi.as_dp(out),
                       _capi.as_ip(off), _capi.as_ip(pad)), self._h)
        return out, off, pad

    def calculateModel(self, params, period: int | None = None) -> dict:
        """state list of the reference (fields used downstream: libMCMC.R:151-172)."""
        series, off, pad = self.calculateModel_batch(np.asarray(params, dtype=np.float64)[None, :], period, ext=True)
        st = {name.replace(".", "_"): series[0, k] for k, name in enumerate(self.series_names_ext)}
        st["padding"] = int(pad[0])
        st["offset"] = int(off[0])
        return st

    def quantileData(self, sample, fun, startoffset: int, period: int, quantiles):
        """libMCMC.R:151-172.  sample: (n, d) posterior draws in SAMPLER space (rows of the
        posterior table; transformParams is applied like the reference does); fun: a series name
        from `self.series_names_ext` (e.g. "Re", analyzeMCMC.R:265) or a tuple of two names whose sum is wanted (the reference passes
        closures such as `function(state, params) state$y.hospi + state$o.hospi`).  Returns a
        (period, len(quantiles)) array; everything runs on the device."""
        names = (fun,) if isinstance(fun, str) else tuple(fun)
        idx = [self.series_names_ext.index(nm) for nm in names] + [-1]
        th = np.ascontiguousarray(self.transformParams(np.atleast_2d(np.asarray(sample, dtype=np.float64))))
        probs = np.ascontiguousarray(np.asarray(quantiles, dtype=np.float64))
        out = np.empty((int(period), len(probs)))
        _capi.check(self._lib.epi_quantiles(self._h, _capi.as_dp(th), th.shape[0], _capi.LAYOUT_AOS, int(period),
                                            int(startoffset), idx[0], idx[1], _capi.as_dp(probs), len(probs),
                                            _capi.as_dp(out)), self._h)
        return out

    def marginalizeData(self, sample, fun, startoffset: int, period: int, marginfun):
        """libMCMC.R:174-188: one number per posterior draw, marginfun(takeAndPad(fun(state), offset - startoffset,
        3 * period)).  fun as in quantileData; marginfun = (kind, j1, j2) with kind in "max" | "sum" | "min" | "mean"
        over v[j1..j2] (1-based, inclusive) — e.g. ("max", today, today + 60) of "Re" is predictMCMC.R:714-717.
        Returns an array of len(sample); everything runs on the device."""
        names = (fun,) if isinstance(fun, str) else tuple(fun)
        idx = [self.series_names_ext.index(nm) for nm in names] + [-1]
        kind, j1, j2 = marginfun
        th = np.ascontiguousarray(self.transformParams(np.atleast_2d(np.asarray(sample, dtype=np.float64))))
        out = np.empty(th.shape[0])
        _capi.check(self._lib.epi_marginalize(self._h, _capi.as_dp(th), th.shape[0], _capi.LAYOUT_AOS, int(period),
                                              int(startoffset), idx[0], idx[1], {"max": 0, "sum": 1, "min": 2, "mean": 3}[kind],
                                              int(j1), int(j2), _capi.as_dp(out)), self._h)
        return out

    def calcNominalState(self, state: dict) -> dict:
        """model-age-groups2q.R:367-376"""
        s = dict(state)
        if self.is_age:
            s["casei"] = s["y_casei"] + s["o_casei"]
            s["deadi"] = s["y_deadi"] + s["o_deadi"]
            s["hospi"] = s["y_hospi"] + s["o_hospi"]
            s["case"] = np.cumsum(s["casei"])
            s["died"] = np.cumsum(s["deadi"])
        return s

    # ---- measurement helpers -----------------------------------------------------------
    def fp64_peak_tflops(self) -> float:
        v = C.c_double(0)
        _capi.check(self._lib.epi_fp64_peak(self._h, C.byref(v)), self._h)
        return v.value

    def launch_count(self) -> int:
        return int(self._lib.epi_launch_count(self._h))

    def fallback_count(self) -> int:
        """chains of the last batch whose ODE passes were redone with the removed compartment integrated"""
        return int(self._lib.epi_fallback_count(self._h))

    def reserve(self, n: int):
        """epi_reserve: size the device workspace for n chains (before the first eval_device_range call)."""
        _capi.check(self._lib.epi_reserve(self._h, n), self._h)

    def eval_device_range(self, d_theta_ptr: int, ld: int, n_total: int, first: int, n: int, d_logl: int, d_logp: int,
                          d_status: int, stream: int = 0):
        """epi_eval_device_range: the chains [first, first + n) of a device-resident batch; ranges that do not overlap
        may be in flight on different streams."""
        _capi.check(self._lib.epi_eval_device_range(self._h, d_theta_ptr, ld, n_total, first, n, d_logl, d_logp,
                                                    d_status, stream), self._h)

    def eval_device(self, d_theta_ptr: int, n: int, d_logl: int, d_logp: int, d_status: int, stream: int = 0):
        """epi_eval_device on raw device pointers (theta SoA d x n)."""
        _capi.check(self._lib.epi_eval_device(self._h, d_theta_ptr, n, d_logl, d_logp, d_status, stream), self._h)

    def set_timing(self, on: bool):
        _capi.check(self._lib.epi_set_timing(self._h, int(on)), self._h)

    def last_kernel_ms(self):
        ms = (C.c_float * 4)()
        _capi.check(self._lib.epi_last_kernel_ms(self._h, ms), self._h)
        return [float(v) for v in ms]
