"""Line-by-line numpy/scipy transliteration of the reference's R model files,
written independently of oracle/epi_oracle.c (vector style, scipy nmath) and
used ONLY by tests to cross-check the oracle:

  * with integrator="rk4" it uses the same fixed-step scheme as the oracle, so
    every index/recycling/alignment rule can be compared to ~1e-12;
  * with integrator="lsoda" it uses scipy's LSODA at deSolve's defaults
    (rtol = atol = 1e-6, hmax = 1 day), i.e. the reference's own solver family,
    so the oracle's fixed-step choice can be compared to solver tolerance.

R vectors are emulated with 1-based helper accessors; NA is NaN.
Reference: R/models/model-age-groups2q.R, R/models/model-agen.cpp,
R/models/model-age-groups2i.R, R/models/model-agef.cpp,
R/models/model-simple-ode-drj-cmp.R.
"""
import numpy as np
from scipy import integrate, special, stats

InvalidDataOffset = 10000
Initial = 1


def r_round(x):
    return float(np.rint(x))  # R round(): half to even


def r_filter1(x, f):
    """stats::filter(x, f, method='convolution', sides=1)"""
    n = len(f)
    out = np.convolve(x, f)[: len(x)].astype(float)
    out[: n - 1] = np.nan
    return out


def rslice(v, a, b):
    """v[a:b] with R semantics (1-based inclusive, NA beyond the end)"""
    idx = np.arange(a, b + 1)
    out = np.full(len(idx), np.nan)
    ok = (idx >= 1) & (idx <= len(v))
    out[ok] = np.asarray(v, dtype=float)[idx[ok] - 1]
    return out


def dnbinom_log(x, mu, size):
    x, mu, size = (np.atleast_1d(np.asarray(a, dtype=float)) for a in (x, mu, size))
    n = max(len(x), len(mu), len(size))
    x, mu, size = np.resize(x, n), np.resize(mu, n), np.resize(size, n)  # R recycling
    out = stats.nbinom.logpmf(x, size, size / (size + mu))
    out[np.isnan(mu) | np.isnan(x)] = np.nan
    return out


def pmax(lo, v):
    v = np.asarray(v, dtype=float)
    return np.where(np.isnan(v), np.nan, np.maximum(lo, v))


def dnorm_log(x, mean, sd):
    return stats.norm.logpdf(x, mean, sd)


def dlnorm_log(x, meanlog, sdlog):
    return stats.lognorm.logpdf(x, s=sdlog, scale=np.exp(meanlog))


# ----------------------------------------------------------------- solvers --

def solve(rhs, Y0, times, integrator, m, breaks=(), rhs_seg=None, active=None):
    """rhs(t, y): the literal reference RHS.  For integrator='rk4' the oracle's
    breakpoint-cut scheme is used: rhs_seg(seg, t, y) evaluates the RHS with the
    schedule segment frozen to `seg`, active(t) returns the segment index."""
    times = np.asarray(times, dtype=float)
    if integrator == "lsoda":
        sol = integrate.solve_ivp(rhs, (times[0], times[-1]), Y0, method="LSODA", t_eval=times,
                                  rtol=1e-6, atol=1e-6, max_step=1.0)
        return sol.y.T
    if integrator == "lsoda_tight":
        sol = integrate.solve_ivp(rhs, (times[0], times[-1]), Y0, method="LSODA", t_eval=times,
                                  rtol=1e-12, atol=1e-9, max_step=0.25)
        return sol.y.T
    h = 1.0 / m
    y = np.array(Y0, dtype=float)
    out = [y.copy()]
    for tday in times[:-1]:
        for s in range(m):
            a = tday + s * h
            b = tday + 1.0 if s == m - 1 else tday + (s + 1) * h
            pa = a
            while True:
                inside = [c for c in breaks if pa < c < b]
                pb = min(inside) if inside else b
                hh = pb - pa
                tm = pa + 0.5 * hh
                seg = active(tm)
                k1 = rhs_seg(seg, pa, y)
                k2 = rhs_seg(seg, tm, y + 0.5 * hh * k1)
                k3 = rhs_seg(seg, tm, y + 0.5 * hh * k2)
                k4 = rhs_seg(seg, pb, y + hh * k3)
                y = y + hh / 6.0 * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
                if pb >= b:
                    break
                pa = pb
        out.append(y.copy())
    return np.array(out)


# ------------------------------------------------------------------- age2q --

class Age2q:
    G = 4.7
    Tinc1, Tinc2 = 3, 1

    def __init__(self, data, settings_period, integrator="rk4", m=4):
        self.D = data
        self.P = settings_period
        self.integrator, self.m = integrator, m
        self.a1, self.a2 = 1 / self.Tinc1, 1 / self.Tinc2
        self.gamma = 1 / ((self.G - (self.Tinc1 + self.Tinc2)) * 2)

    # model-age-groups2q.R:17-40
    @staticmethod
    def calcGammaProfile(mean, sd):
        scale = sd ** 2 / mean
        shape = mean / scale
        kbegin = max(0, int(np.ceil(mean - sd * 3)))
        kend = max(kbegin + 1, int(np.floor(mean + sd * 3)))
        ks = np.arange(kbegin, kend + 1)
        cdf = lambda q: np.where(q > 0, special.gammainc(shape, np.maximum(q, 0) / scale), 0.0)
        values = cdf(ks - 0.5) - cdf(ks + 0.5)
        values = values / np.sum(values)
        values = values[::-1]
        return dict(kbegin=-kend, kend=-kbegin, values=values)

    # model-age-groups2q.R:42-45
    @staticmethod
    def convolute(values, i1, i2, profile):
        f = r_filter1(values, profile["values"])
        return rslice(f, i1 + profile["kend"], i2 + profile["kend"])

    KIND = (1, 1, 1, 1, 0, 0, 0, 0, 1, 0)

    @staticmethod
    def active(t, T):
        for k in range(9, -1, -1):
            if t > T[k]:
                return k
        return -1

    @classmethod
    def interp_seg(cls, k, t, T, v):
        if k < 0:
            return v[0]
        if cls.KIND[k]:
            return v[k] + (t - T[k]) / (T[k + 1] - T[k]) * (v[k + 1] - v[k])
        return v[k]

    # model-agen.cpp:73-100
    @staticmethod
    def interpolate(t, T, v):
        if t > T[9]: return v[9]
        elif t > T[8]: return v[8] + (t - T[8]) / (T[9] - T[8]) * (v[9] - v[8])
        elif t > T[7]: return v[7]
        elif t > T[6]: return v[6]
        elif t > T[5]: return v[5]
        elif t > T[4]: return v[4]
        elif t > T[3]: return v[3] + (t - T[3]) / (T[4] - T[3]) * (v[4] - v[3])
        elif t > T[2]: return v[2] + (t - T[2]) / (T[3] - T[2]) * (v[3] - v[2])
        elif t > T[1]: return v[1] + (t - T[1]) / (T[2] - T[1]) * (v[2] - v[1])
        elif t > T[0]: return v[0] + (t - T[0]) / (T[1] - T[0]) * (v[1] - v[0])
        return v[0]

    def derivs(self, t, y, T, by, bo, byo, seg=None):
        Ny, No = self.D["y_N"], self.D["o_N"]
        Sy, E1y, E2y, Iy, Ry, So, E1o, E2o, Io, Ro = y
        if (min(y[:5]) < 0 or max(y[:5]) > Ny or min(y[5:]) < 0 or max(y[5:]) > No):
            return np.zeros(10)
        if seg is None:
            betay = self.interpolate(t, T, by)
            betao = self.interpolate(t, T, bo)
            betayo = self.interpolate(t, T, byo)
        else:
            betay = self.interp_seg(seg, t, T, by)
            betao = self.interp_seg(seg, t, T, bo)
            betayo = self.interp_seg(seg, t, T, byo)
        yinf = (betay * Iy) / Ny * Sy
        oinf = ((betao * Io) / No + (betayo * Iy) / Ny) * So
        return np.array([-yinf, yinf - self.a1 * E1y, self.a1 * E1y - self.a2 * E2y,
                         self.a2 * E2y - self.gamma * Iy, self.gamma * Iy,
                         -oinf, oinf - self.a1 * E1o, self.a1 * E1o - self.a2 * E2o,
                         self.a2 * E2o - self.gamma * Io, self.gamma * Io])

    def _solve(self, Y0, times, Tb, by, bo, byo):
        return solve(lambda t, y: self.derivs(t, y, Tb, by, bo, byo), Y0, times, self.integrator, self.m,
                     breaks=Tb, rhs_seg=lambda seg, t, y: self.derivs(t, y, Tb, by, bo, byo, seg=seg),
                     active=lambda t: self.active(t, Tb))

    # model-age-groups2q.R:50-337
    def calculateModel(self, params, period):
        D = self.D
        p = lambda i: params[i - 1]
        y_N, o_N = D["y_N"], D["o_N"]
        y_ifr, o_ifr, y_hr, o_hr = (np.asarray(D[k], dtype=float) for k in ("y_ifr", "o_ifr", "y_hr", "o_hr"))
        g, gtrimp = np.asarray(D["g"], dtype=float), np.asarray(D["gtrimp"], dtype=float)
        y_died_rate, o_died_rate = y_ifr[0], o_ifr[0]
        lockdown_offset, ltp = D["lockdown_offset"], D["lockdown_transition_period"]
        by = [p(1), p(4), p(7), p(7), p(26), p(29), p(33), p(37), p(37), p(40) * p(37)]
        bo = [p(2), p(5), p(8), p(27), p(27), p(30), p(34), p(38), p(38), p(41) * p(38)]
        byo = [p(3), p(6), p(9), p(28), p(28), p(31), p(35), p(39), p(39), p(42) * p(39)]
        ycase_rate, yhosp_rate, yhosp_latency, ydied_latency = p(10), p(11), p(12), p(13)
        ocase_rate, ohosp_rate, ohosp_latency, odied_latency = p(14), p(15), p(16), p(17)
        t0_morts, t0o, HLsd, DLsd, fyifr, fyhr, t3o, t4o = p(18), p(19), p(20), p(21), p(22), p(23), p(24), p(25)
        t6o, t7o, ifrred, ycase_latency, ocase_latency = p(32), p(36), p(43), p(44), p(45)
        CLsd = 5
        prof = dict(ycase=self.calcGammaProfile(ycase_latency, CLsd), ocase=self.calcGammaProfile(ocase_latency, CLsd),
                    yhosp=self.calcGammaProfile(yhosp_latency, HLsd), ohosp=self.calcGammaProfile(ohosp_latency, HLsd),
                    ydied=self.calcGammaProfile(ydied_latency, DLsd), odied=self.calcGammaProfile(odied_latency, DLsd))
        padding = max(-prof["ycase"]["kbegin"], -prof["ydied"]["kbegin"], -prof["ocase"]["kbegin"],
                      -prof["odied"]["kbegin"]) + 1
        T = padding + period
        yS = np.full(T, y_N - Initial, dtype=float)
        oS = np.full(T, o_N, dtype=float)
        Y0 = [y_N - Initial, Initial, 0, 0, 0, o_N, 0, 0, 0, 0]
        Tb = [1e10] * 10
        period1 = 60
        times = np.arange(padding + 1, padding + period1 + 1)
        out = self._solve(Y0, times, Tb, by, bo, byo)
        yS1, oS1 = yS.copy(), oS.copy()
        if len(yS1) < padding + period1:
            yS1 = np.concatenate([yS1, np.full(padding + period1 - len(yS1), np.nan)])
            oS1 = np.concatenate([oS1, np.full(padding + period1 - len(oS1), np.nan)])
        yS1[padding:padding + period1] = out[:, 0]
        oS1[padding:padding + period1] = out[:, 5]
        died = np.full(padding + period1, np.nan)
        s2 = self.convolute(yS1, padding + 1, padding + period1, prof["ydied"])
        s2o = self.convolute(oS1, padding + 1, padding + period1, prof["odied"])
        died[padding:padding + period1] = (y_N - s2) * y_died_rate + (o_N - s2o) * o_died_rate
        data_offset = InvalidDataOffset
        with np.errstate(invalid="ignore"):
            lds = np.nonzero(died > t0_morts)[0] + 1
        if len(lds) > 0:
            data_offset = lds[0] - lockdown_offset
            t2 = data_offset + lockdown_offset + ltp
            t3 = t2 + t3o
            t4 = t3 + t4o
            t5 = data_offset + D["d5"]
            t6 = t5 + t6o
            t7 = t6 + t7o
            t8 = data_offset + D["d8"]
            t9 = t8 + 7
            Tb = [data_offset + lockdown_offset + t0o, data_offset + lockdown_offset + max(t0o, 0), t2,
                  t3, t4, t5, t6, t7, t8, t9]
            times = np.arange(padding + 1, padding + period + 1)
            out = self._solve(Y0, times, Tb, by, bo, byo)
        yS[padding:] = np.resize(out[:, 0], period)  # recycled when pass 2 was skipped
        oS[padding:] = np.resize(out[:, 5], period)

        gycr = np.minimum(1, ycase_rate * y_ifr * (1 + (fyifr - 1) * g))
        gyifr = y_ifr * (1 + (fyifr - 1) * g) * (1 - ifrred * gtrimp)
        gyhr = yhosp_rate * y_hr * (1 + (fyhr - 1) * g)
        gocr = np.minimum(1, ocase_rate * o_ifr)
        goifr = o_ifr * (1 - ifrred * gtrimp)
        gohr = ohosp_rate * o_hr

        def series(S, Ng, profile, t1, rate):
            dst = np.zeros(T)
            s2 = Ng - self.convolute(S, padding + 1, padding + period, profile)
            s2i = np.concatenate([[s2[0]], np.diff(s2)])
            if data_offset != InvalidDataOffset:
                t1 = int(t1)
                t2 = min(padding + period, t1 + len(rate) - 1)
                if t1 > padding and t2 > t1:
                    dst[padding:t1] = s2i[0:t1 - padding] * rate[0]
                    dst[t1 - 1:t2] = s2i[t1 - padding - 1:t2 - padding] * rate[0:t2 - t1 + 1]
                    dst[t2 - 1:padding + period] = s2i[t2 - padding - 1:period] * rate[-1]
            else:
                dst[padding:padding + period] = s2i * rate[0]
            return dst

        st = dict(padding=padding, offset=data_offset, y_S=yS, o_S=oS)
        st["y_casei"] = series(yS, y_N, prof["ycase"], data_offset + r_round(-ydied_latency + ycase_latency), gycr)
        st["y_hospi"] = series(yS, y_N, prof["yhosp"], data_offset + r_round(-ydied_latency + yhosp_latency), gyhr)
        st["y_deadi"] = series(yS, y_N, prof["ydied"], data_offset, gyifr)
        st["o_casei"] = series(oS, o_N, prof["ocase"], data_offset + r_round(-ydied_latency + ocase_latency), gocr)
        st["o_hospi"] = series(oS, o_N, prof["ohosp"], data_offset + r_round(-ydied_latency + ohosp_latency), gohr)
        st["o_deadi"] = series(oS, o_N, prof["odied"], data_offset, goifr)
        return st

    # model-age-groups2q.R:488-602
    def calclogl(self, params):
        D = self.D
        state = self.calculateModel(params, self.P)
        if state["offset"] == InvalidDataOffset:
            state["offset"] = 1
        y_dcasei, o_dcasei = np.asarray(D["y_dcasei"], float), np.asarray(D["o_dcasei"], float)
        dhospi = np.asarray(D["dhospi"], float)
        y_dmorti, o_dmorti = np.asarray(D["y_dmorti"], float), np.asarray(D["o_dmorti"], float)
        o_wdmorti = np.asarray(D["o_wdmorti"], float)
        r, rh = D["d_reliable_cases"], D["d_reliable_hosp"]
        n_dhosp = D["n_dhosp"]

        def window(n, L):
            dstart = state["offset"]
            dend = state["offset"] + n - 1
            if dend > L:
                dend = L
                dstart = dend - n + 1
            if dstart < 1:
                dstart = 1
                dend = dstart + n - 1
            return dstart, dend

        dstart, dend = window(len(y_dcasei), len(state["y_casei"]))
        y_loglC = (np.sum(dnbinom_log(rslice(y_dcasei, 1, r - 1), pmax(0.1, rslice(state["y_casei"], dstart, dstart + r)), D["case_nbinom_size1"]))
                   + np.sum(dnbinom_log(rslice(y_dcasei, r, len(y_dcasei)), pmax(0.1, rslice(state["y_casei"], dstart + r + 1, dend)), D["case_nbinom_sizey2"])))
        o_loglC = (np.sum(dnbinom_log(rslice(o_dcasei, 1, r - 1), pmax(0.1, rslice(state["o_casei"], dstart, dstart + r)), D["case_nbinom_size1"]))
                   + np.sum(dnbinom_log(rslice(o_dcasei, r, len(o_dcasei)), pmax(0.1, rslice(state["o_casei"], dstart + r + 1, dend)), D["case_nbinom_sizeo2"])))
        dstart, dend = window(len(dhospi), len(state["y_hospi"]))
        H = state["y_hospi"] + state["o_hospi"]
        loglH = (np.sum(dnbinom_log(rslice(dhospi, 1, rh - 1), pmax(0.1, rslice(H, dstart, dstart + rh)), D["hosp_nbinom_size1"]))
                 + np.sum(dnbinom_log(rslice(dhospi, rh, n_dhosp), pmax(0.1, rslice(H, dstart + rh + 1, dend)), D["hosp_nbinom_size2"]))
                 + np.sum(dnbinom_log(rslice(dhospi, n_dhosp - 7, n_dhosp), pmax(0.1, rslice(H, dend - 7, dend)), D["hosp_nbinom_size2"] * D["hosp_last7"])))
        ytotal = np.sum(rslice(state["y_hospi"], dstart + D["d_hosp_o1"], dstart + D["d_hosp_o2"]))
        ototal = np.sum(rslice(state["o_hospi"], dstart + D["d_hosp_o1"], dstart + D["d_hosp_o2"]))
        loglHRatio = dlnorm_log(ytotal / (ytotal + ototal), np.log(56.7 / 100), np.log(1.05))
        dstart, dend = window(len(y_dmorti), len(state["y_deadi"]))
        y_loglD = np.sum(dnbinom_log(y_dmorti, pmax(0.001, rslice(state["y_deadi"], dstart, dend)), D["mort_nbinom_size"] * 2))
        o_loglD = np.sum(dnbinom_log(o_dmorti, pmax(0.001, rslice(state["o_deadi"], dstart, dend)), o_wdmorti * D["mort_nbinom_size"]))
        terms = [y_loglC, o_loglC, loglH, loglHRatio, y_loglD, o_loglD]
        return float(np.sum(terms)), terms, state

    # model-age-groups2q.R:380-486
    def calclogp(self, params):
        D, gamma = self.D, self.gamma
        p = lambda i: params[i - 1]
        lo, ltp = D["lockdown_offset"], D["lockdown_transition_period"]
        SD, lSD = np.log(2), np.log(1.5)
        betay3, betao8 = p(7), p(38)
        lp = 0.0
        lp += dnorm_log(p(19), 0, 10)
        lp += dnorm_log(lo + ltp + p(24), D["d3"], 10)
        lp += dnorm_log(lo + ltp + p(24) + p(25), D["d4"], 10)
        lp += dnorm_log(D["d5"] + p(32), D["d6"], 5)
        lp += dnorm_log(D["d5"] + p(32) + p(36), D["d7"], 4)
        for a, b in ((p(1), p(4)), (p(2), p(5)), (p(3), p(6)), (p(4), p(7)), (p(5), p(8)), (p(6), p(9)),
                     (betay3, p(26)), (p(8), p(27)), (p(9), p(28)), (p(29), p(33)), (p(30), p(34)), (p(31), p(35))):
            lp += dlnorm_log(a / b, 0, SD)
        for a, b in ((p(33), p(37)), (p(34), p(38)), (p(35), p(39))):
            lp += dlnorm_log(a / b, 0, lSD)
        lp += dnorm_log(p(38), 0.5 * gamma, 0.2 * gamma)
        lp += dnorm_log(betao8 * p(41), 0.5 * gamma, 0.2 * gamma)
        for i in (40, 41, 42):
            lp += dlnorm_log(p(i), np.log(0.85), np.log(1.07))
        lp += dnorm_log(p(20), 5, 1) + dnorm_log(p(21), 5, 1)
        lp += dnorm_log(p(44), 10, 3) + dnorm_log(p(45), 10, 3)
        lp += dnorm_log(p(12), 13, 3) + dnorm_log(p(16), 13, 3)
        lp += dnorm_log(p(13), 21, 0.5) + dnorm_log(p(17), 21, 0.5)
        lp += dlnorm_log(p(22), np.log(0.3), np.log(2)) + dlnorm_log(p(23), np.log(0.3), np.log(2))
        lp += dlnorm_log(p(43), np.log(0.35), np.log(1.3))
        lp += dlnorm_log(p(11), 0, lSD) + dlnorm_log(p(15), 0, lSD)
        return float(lp)


# ------------------------------------------------------------------- age2x --

class Age2x(Age2q):
    """R/models/model-age-groups2x.R + R/models/model-ager.cpp (the later generation), transliterated from the R file."""
    KIND = (1, 1, 1, 1, 0, 0, 0, 0, 1, 0, 1, 1, 0)
    eta = 1 / (0.75 * 356)
    CLsd = 2.5

    @staticmethod
    def active(t, T):
        for k in range(12, -1, -1):
            if t > T[k]:
                return k
        return -1

    # model-ager.cpp:88-124
    @staticmethod
    def interpolate(t, T, v):
        if t > T[12]: return v[12]
        elif t > T[11]: return v[11] + (t - T[11]) / (T[12] - T[11]) * (v[12] - v[11])
        elif t > T[10]: return v[10] + (t - T[10]) / (T[11] - T[10]) * (v[11] - v[10])
        elif t > T[9]: return v[9]
        elif t > T[8]: return v[8] + (t - T[8]) / (T[9] - T[8]) * (v[9] - v[8])
        elif t > T[7]: return v[7]
        elif t > T[6]: return v[6]
        elif t > T[5]: return v[5]
        elif t > T[4]: return v[4]
        elif t > T[3]: return v[3] + (t - T[3]) / (T[4] - T[3]) * (v[4] - v[3])
        elif t > T[2]: return v[2] + (t - T[2]) / (T[3] - T[2]) * (v[3] - v[2])
        elif t > T[1]: return v[1] + (t - T[1]) / (T[2] - T[1]) * (v[2] - v[1])
        elif t > T[0]: return v[0] + (t - T[0]) / (T[1] - T[0]) * (v[1] - v[0])
        return v[0]

    # model-ager.cpp:135-214; returns (ydot, y.i, o.i)
    def derivs_out(self, t, y, T, by, bo, byo, seg=None):
        Ny, No = self.D["y_N"], self.D["o_N"]
        Sy, E1y, E2y, Iy, Ry, So, E1o, E2o, Io, Ro = y
        if (min(y[:5]) < 0 or max(y[:5]) > Ny or min(y[5:]) < 0 or max(y[5:]) > No):
            return np.zeros(10), np.nan, np.nan
        f = (lambda v: self.interpolate(t, T, v)) if seg is None else (lambda v: self.interp_seg(seg, t, T, v))
        betay, betao, betayo = f(by), f(bo), f(byo)           # uncertain(): funcertain = 1 (:221)
        Syeff = max(0.0, Ny * 0.8 - (Ny - Sy))
        yinf = (betay * Iy) / Ny * Syeff
        oinf = ((betao * Io) / No + (betayo * Iy) / Ny) * So
        yrev, orev = self.eta * Ry, self.eta * Ro
        d = np.array([-yinf + yrev, yinf - self.a1 * E1y, self.a1 * E1y - self.a2 * E2y, self.a2 * E2y - self.gamma * Iy,
                      self.gamma * Iy - yrev,
                      -oinf + orev, oinf - self.a1 * E1o, self.a1 * E1o - self.a2 * E2o, self.a2 * E2o - self.gamma * Io,
                      self.gamma * Io - orev])
        return d, yinf, oinf

    def derivs(self, t, y, T, by, bo, byo, seg=None):
        return self.derivs_out(t, y, T, by, bo, byo, seg)[0]

    def _incidence(self, out, times, Tb, by, bo, byo):
        """the output variables y.i / o.i as deSolve reports them: one derivs() call per output time"""
        yi, oi = np.zeros(len(times)), np.zeros(len(times))
        for k, t in enumerate(times):
            _, yi[k], oi[k] = self.derivs_out(float(t), out[k], Tb, by, bo, byo)
        return yi, oi

    # model-age-groups2x.R:55-380
    def calculateModel(self, params, period):
        D = self.D
        p = lambda i: params[i - 1]
        y_N, o_N = D["y_N"], D["o_N"]
        y_ifr, o_ifr, y_hr, o_hr = (np.asarray(D[k], dtype=float) for k in ("y_ifr", "o_ifr", "y_hr", "o_hr"))
        g, gtrimp = np.asarray(D["g"], dtype=float), np.asarray(D["gtrimp"], dtype=float)
        y_died_rate, o_died_rate = y_ifr[0], o_ifr[0]
        lockdown_offset, ltp = D["lockdown_offset"], D["lockdown_transition_period"]
        first = [1, 4, 7, 25, 29, 32, 36, 40, 40, 43, 43, 50, 43]     # beta8 = beta7, beta10 = beta12 = beta9 (:56-119)
        by, bo, byo = [p(i) for i in first], [p(i + 1) for i in first], [p(i + 2) for i in first]
        ycase_rate, yhosp_rate, yhosp_latency, ydied_latency = p(10), p(11), p(12), p(13)
        ocase_rate, ohosp_rate, ohosp_latency, odied_latency = p(14), p(15), p(16), p(17)
        t0_morts, t0o, HLsd, DLsd, fyifr, fyhr = p(18), p(19), p(20), p(21), p(22), p(23)
        t3o, t4o, t6o, t7o, ifrred, ycase_latency, ocase_latency, t10o = p(24), p(28), p(35), p(39), p(46), p(47), p(48), p(49)
        CLsd = self.CLsd
        prof = dict(ycase=self.calcGammaProfile(ycase_latency, CLsd), ocase=self.calcGammaProfile(ocase_latency, CLsd),
                    yhosp=self.calcGammaProfile(yhosp_latency, HLsd), ohosp=self.calcGammaProfile(ohosp_latency, HLsd),
                    ydied=self.calcGammaProfile(ydied_latency, DLsd), odied=self.calcGammaProfile(odied_latency, DLsd))
        padding = max(-prof["ycase"]["kbegin"], -prof["ydied"]["kbegin"], -prof["ocase"]["kbegin"],
                      -prof["odied"]["kbegin"]) + 1
        T = padding + period
        Y0 = [y_N - Initial, Initial, 0, 0, 0, o_N, 0, 0, 0, 0]
        Tb = [1e10] * 13
        period1 = 60
        times = np.arange(padding + 1, padding + period1 + 1)
        out = self._solve(Y0, times, Tb, by, bo, byo)
        yi_out, oi_out = self._incidence(out, times, Tb, by, bo, byo)
        L1 = max(T, padding + period1)
        yi1, oi1 = np.zeros(L1), np.zeros(L1)
        yi1[padding:padding + period1] = yi_out
        oi1[padding:padding + period1] = oi_out
        ydeadi = self.convolute(yi1, padding + 1, padding + period1, prof["ydied"]) * y_died_rate
        odeadi = self.convolute(oi1, padding + 1, padding + period1, prof["odied"]) * o_died_rate
        died = np.full(padding + period1, np.nan)
        died[padding:] = np.cumsum(ydeadi) + np.cumsum(odeadi)
        data_offset = InvalidDataOffset
        with np.errstate(invalid="ignore"):
            lds = np.nonzero(died > t0_morts)[0] + 1
        if len(lds) > 0:
            data_offset = lds[0] - lockdown_offset
            t2 = data_offset + lockdown_offset + ltp
            t3 = t2 + t3o
            t4 = t3 + t4o
            t5 = data_offset + D["d5"]
            t6 = t5 + t6o
            t7 = t6 + t7o
            t10 = data_offset + t10o
            Tb = [data_offset + lockdown_offset + t0o, data_offset + lockdown_offset + max(t0o, 0), t2, t3, t4, t5, t6, t7,
                  data_offset + D["d8"], data_offset + D["d9"], t10, t10 + 3, data_offset + D["d12"]]
            times = np.arange(padding + 1, padding + period + 1)
            out = self._solve(Y0, times, Tb, by, bo, byo)
            yi_out, oi_out = self._incidence(out, times, Tb, by, bo, byo)
        yi, oi = np.zeros(T), np.zeros(T)
        yi[padding:] = np.resize(yi_out, period)   # recycled when pass 2 was skipped
        oi[padding:] = np.resize(oi_out, period)
        yS = np.full(T, y_N - Initial, dtype=float)
        oS = np.full(T, o_N, dtype=float)
        yS[padding:] = np.resize(out[:, 0], period)
        oS[padding:] = np.resize(out[:, 5], period)

        gycr = np.minimum(1, ycase_rate * y_ifr * (1 + (fyifr - 1) * g))
        gyifr = y_ifr * (1 + (fyifr - 1) * g) * (1 - ifrred * gtrimp)
        gyhr = yhosp_rate * y_hr * (1 + (fyhr - 1) * g)
        gocr = np.minimum(1, ocase_rate * o_ifr)
        goifr = o_ifr * (1 - ifrred * gtrimp)
        gohr = ohosp_rate * o_hr
        t_symp = data_offset + D["d_symp_cases"]
        t_all = data_offset + D["d_all_cases"] - 1

        def series(inc, profile, t1, rate, cases=False):
            dst = np.zeros(T)
            s2i = self.convolute(inc, padding + 1, padding + period, profile)
            if data_offset != InvalidDataOffset:
                t1 = int(t1)
                t2 = min(padding + period, t1 + len(rate) - 1)
                if t1 > padding and t2 > t1:
                    dst[padding:t1] = s2i[0:t1 - padding] * rate[0]
                    dst[t1 - 1:t2] = s2i[t1 - padding - 1:t2 - padding] * rate[0:t2 - t1 + 1]
                    dst[t2 - 1:padding + period] = s2i[t2 - padding - 1:period] * rate[-1]
                if cases:   # state$y.casei[t.symp:min(t.all, length)] *= symp.cases.factor (R's a:b may run downwards)
                    a, b = t_symp, min(t_all, len(dst))
                    assert 1 <= a <= len(dst) and b >= 1, "window beyond the vector: R would grow it"
                    lo_, hi_ = min(a, b), max(a, b)
                    dst[lo_ - 1:hi_] = dst[lo_ - 1:hi_] * D["symp_cases_factor"]
            else:
                dst[padding:padding + period] = s2i * rate[0]
            return dst

        st = dict(padding=padding, offset=data_offset, y_S=yS, o_S=oS, y_i=yi, o_i=oi)
        st["y_casei"] = series(yi, prof["ycase"], data_offset + r_round(-ydied_latency + ycase_latency), gycr, True)
        st["y_hospi"] = series(yi, prof["yhosp"], data_offset + r_round(-ydied_latency + yhosp_latency), gyhr)
        st["y_deadi"] = series(yi, prof["ydied"], data_offset, gyifr)
        st["o_casei"] = series(oi, prof["ocase"], data_offset + r_round(-ydied_latency + ocase_latency), gocr, True)
        st["o_hospi"] = series(oi, prof["ohosp"], data_offset + r_round(-ydied_latency + ohosp_latency), gohr)
        st["o_deadi"] = series(oi, prof["odied"], data_offset, goifr)
        return st

    # model-age-groups2x.R:547-681
    def calclogl(self, params):
        D = self.D
        state = self.calculateModel(params, self.P)
        if state["offset"] == InvalidDataOffset:
            state["offset"] = 1
        y_dcasei, o_dcasei = np.asarray(D["y_dcasei"], float), np.asarray(D["o_dcasei"], float)
        dhospi = np.asarray(D["dhospi"], float)
        y_dmorti, o_dmorti = np.asarray(D["y_dmorti"], float), np.asarray(D["o_dmorti"], float)
        o_wdmorti = np.asarray(D["o_wdmorti"], float)
        r, rh, n_dhosp, last7 = D["d_reliable_cases"], D["d_reliable_hosp"], D["n_dhosp"], D["hosp_last7"]

        def window(n, L):
            dstart = state["offset"]
            dend = state["offset"] + n - 1
            if dend > L:
                dend = L
                dstart = dend - n + 1
            if dstart < 1:
                dstart = 1
                dend = dstart + n - 1
            return dstart, dend

        nb = lambda x, mu, size, fl: np.sum(dnbinom_log(x, pmax(fl, mu), size))
        nc = len(y_dcasei)
        dstart, dend = window(nc, len(state["y_casei"]))
        y_loglC = (nb(rslice(y_dcasei, 1, r - 1), rslice(state["y_casei"], dstart, dstart + r), D["case_nbinom_size1"], 0.1)
                   + nb(rslice(y_dcasei, r, nc), rslice(state["y_casei"], dstart + r + 1, dend), D["case_nbinom_sizey2"], 0.1))
        o_loglC = (nb(rslice(o_dcasei, 1, r - 1), rslice(state["o_casei"], dstart, dstart + r), D["case_nbinom_size1"], 0.1)
                   + nb(rslice(o_dcasei, r, nc), rslice(state["o_casei"], dstart + r + 1, dend), D["case_nbinom_sizeo2"], 0.1)
                   + nb(rslice(o_dcasei, nc - 7, nc), rslice(state["o_casei"], dend - 7, dend), D["case_nbinom_sizeo2"] * last7, 0.1))
        dstart, dend = window(len(dhospi), len(state["y_hospi"]))
        H = state["y_hospi"] + state["o_hospi"]
        loglH = (nb(rslice(dhospi, 1, rh - 1), rslice(H, dstart, dstart + rh), D["hosp_nbinom_size1"], 0.1)
                 + nb(rslice(dhospi, rh, n_dhosp), rslice(H, dstart + rh + 1, dend), D["hosp_nbinom_size2"], 0.1)
                 + nb(rslice(dhospi, n_dhosp - 7, n_dhosp), rslice(H, dend - 7, dend), D["hosp_nbinom_size2"] * last7, 0.1))
        ytotal = np.sum(rslice(state["y_hospi"], dstart + D["d_hosp_o1"], dstart + D["d_hosp_o2"]))
        ototal = np.sum(rslice(state["o_hospi"], dstart + D["d_hosp_o1"], dstart + D["d_hosp_o2"]))
        loglHRatio = dlnorm_log(ytotal / (ytotal + ototal), np.log(56.7 / 100), np.log(1.05))
        nm = len(y_dmorti)
        dstart, dend = window(nm, len(state["y_deadi"]))
        y_loglD = nb(y_dmorti, rslice(state["y_deadi"], dstart, dend), D["mort_nbinom_size"] * 5, 0.001)
        o_loglD = (nb(o_dmorti, rslice(state["o_deadi"], dstart, dend), o_wdmorti * D["mort_nbinom_size"], 0.001)
                   + nb(rslice(o_dmorti, nm - 7, nm), rslice(state["o_deadi"], dend - 7, dend), D["mort_nbinom_size"] * last7, 0.001))
        terms = [y_loglC, o_loglC, loglH, loglHRatio, y_loglD, o_loglD]
        return float(np.sum(terms)), terms, state

    # model-age-groups2x.R:427-545
    def calclogp(self, params):
        D = self.D
        p = lambda i: params[i - 1]
        lo, ltp = D["lockdown_offset"], D["lockdown_transition_period"]
        SD, lSD, llSD = np.log(2), np.log(1.3), np.log(1.2)
        b = {0: 1, 1: 4, 2: 7, 3: 25, 4: 29, 5: 32, 6: 36, 7: 40, 9: 43, 11: 50}
        ratio3 = lambda i, j, m, s: sum(dlnorm_log(p(b[i] + c) / p(b[j] + c), m, s) for c in range(3))
        lp = 0.0
        lp += dnorm_log(p(19), 0, 10)
        lp += dnorm_log(lo + ltp + p(24), D["d3"], 10)
        lp += dnorm_log(lo + ltp + p(24) + p(28), D["d4"], 10)
        lp += dnorm_log(D["d5"] + p(35), D["d6"], 5)
        lp += dnorm_log(D["d5"] + p(35) + p(39), D["d7"] - lo, 2)
        lp += dnorm_log(p(49), D["d10"], 3)
        lp += ratio3(0, 1, 0, llSD)
        for i, j in ((1, 2), (2, 3), (3, 4), (5, 6), (6, 7)):
            lp += ratio3(i, j, 0, SD)
        lp += ratio3(7, 9, np.log(1.3), lSD)
        lp += ratio3(9, 11, 0, lSD)
        lp += dnorm_log(p(20), 4, 1) + dnorm_log(p(21), 5, 1)
        lp += dnorm_log(p(47), 7, 3) + dnorm_log(p(48), 7, 3)
        lp += dnorm_log(p(12), 13, 1) + dnorm_log(p(16), 13, 1)
        lp += dnorm_log(p(13), 21, 0.5) + dnorm_log(p(17), 21, 0.5)
        lp += dlnorm_log(p(22), np.log(0.6), np.log(1.5))
        lp += dlnorm_log(p(46), np.log(0.35), np.log(1.3))
        lp += dlnorm_log(p(11), 0, lSD) + dlnorm_log(p(15), 0, lSD)
        return float(lp)


# ------------------------------------------------------------------- age2i --

class Age2i(Age2q):
    """R/models/model-age-groups2i.R + R/models/model-agef.cpp (36 parameters, 8 breakpoints)."""
    KIND = (1, 1, 1, 1, 1, 1, 0, 0)

    @staticmethod
    def active(t, T):
        for k in range(7, -1, -1):
            if t > T[k]:
                return k
        return -1

    # model-agef.cpp:66-87
    @staticmethod
    def interpolate(t, T, v):
        if t > T[7]: return v[7]
        if t > T[6]: return v[6]
        elif t > T[5]: return v[5] + (t - T[5]) / (T[6] - T[5]) * (v[6] - v[5])
        elif t > T[4]: return v[4] + (t - T[4]) / (T[5] - T[4]) * (v[5] - v[4])
        elif t > T[3]: return v[3] + (t - T[3]) / (T[4] - T[3]) * (v[4] - v[3])
        elif t > T[2]: return v[2] + (t - T[2]) / (T[3] - T[2]) * (v[3] - v[2])
        elif t > T[1]: return v[1] + (t - T[1]) / (T[2] - T[1]) * (v[2] - v[1])
        elif t > T[0]: return v[0] + (t - T[0]) / (T[1] - T[0]) * (v[1] - v[0])
        return v[0]

    # model-age-groups2i.R:50-316
    def calculateModel(self, params, period):
        D = self.D
        p = lambda i: params[i - 1]
        y_N, o_N = D["y_N"], D["o_N"]
        y_ifr, o_ifr, y_hr, o_hr = (np.asarray(D[k], dtype=float) for k in ("y_ifr", "o_ifr", "y_hr", "o_hr"))
        y_died_rate, o_died_rate = y_ifr[0], o_ifr[0]
        lockdown_offset, ltp = D["lockdown_offset"], D["lockdown_transition_period"]
        by = [p(1), p(4), p(7), p(7), p(27), p(27), p(31), p(34) * p(31)]
        bo = [p(2), p(5), p(8), p(8), p(28), p(28), p(32), p(36) * p(32)]
        byo = [p(3), p(6), p(9), p(9), p(29), p(29), p(33), p(35) * p(33)]
        ycase_rate, ycase_latency, yhosp_rate, yhosp_latency, ydied_latency = p(10), p(11), p(12), p(13), p(14)
        ocase_rate, ocase_latency, ohosp_rate, ohosp_latency, odied_latency = p(15), p(16), p(17), p(18), p(19)
        t0_morts, t0o, CLsd, HLsd, DLsd, t3o, t4o, t6o = p(20), p(21), p(22), p(23), p(24), p(25), p(26), p(30)
        prof = dict(ycase=self.calcGammaProfile(ycase_latency, CLsd), ocase=self.calcGammaProfile(ocase_latency, CLsd),
                    yhosp=self.calcGammaProfile(yhosp_latency, HLsd), ohosp=self.calcGammaProfile(ohosp_latency, HLsd),
                    ydied=self.calcGammaProfile(ydied_latency, DLsd), odied=self.calcGammaProfile(odied_latency, DLsd))
        padding = max(-prof["ycase"]["kbegin"], -prof["ydied"]["kbegin"], -prof["ocase"]["kbegin"],
                      -prof["odied"]["kbegin"]) + 1
        T = padding + period
        yS = np.full(T, y_N - Initial, dtype=float)
        oS = np.full(T, o_N, dtype=float)
        Y0 = [y_N - Initial, Initial, 0, 0, 0, o_N, 0, 0, 0, 0]
        Tb = [1e10] * 8
        times = np.arange(padding + 1, padding + period + 1)
        out = self._solve(Y0, times, Tb, by, bo, byo)
        yS[padding:] = out[:, 0]
        oS[padding:] = out[:, 5]
        died = np.full(T, np.nan)
        died[padding:] = ((y_N - self.convolute(yS, padding + 1, padding + period, prof["ydied"])) * y_died_rate
                          + (o_N - self.convolute(oS, padding + 1, padding + period, prof["odied"])) * o_died_rate)
        data_offset = InvalidDataOffset
        with np.errstate(invalid="ignore"):
            lds = np.nonzero(died > t0_morts)[0] + 1
        if len(lds) > 0:
            data_offset = lds[0] - lockdown_offset
            t2 = data_offset + lockdown_offset + ltp
            t3 = t2 + t3o
            t4 = t3 + t4o
            t5 = data_offset + D["d5"]
            t6 = t5 + t6o
            t7 = data_offset + D["d7"]
            Tb = [data_offset + lockdown_offset + t0o, data_offset + lockdown_offset + max(t0o, 0), t2,
                  t3, t4, t5, t6, t7]
            out = self._solve(Y0, times, Tb, by, bo, byo)
        yS[padding:] = out[:, 0]
        oS[padding:] = out[:, 5]

        def series(S, Ng, profile, t1, valid_rate, invalid_fun):
            """valid_rate(s2i_slice, idx0, idx1) applies rate[idx0:idx1] the way the R line does"""
            dst = np.zeros(T)
            s2 = Ng - self.convolute(S, padding + 1, padding + period, profile)
            s2i = np.concatenate([[s2[0]], np.diff(s2)])
            n_rate = len(y_ifr)
            if data_offset != InvalidDataOffset:
                t1 = int(t1)
                t2 = min(padding + period, t1 + n_rate - 1)
                if t1 > padding and t2 > t1:
                    dst[padding:t1] = valid_rate(s2i[0:t1 - padding], 0, 1)
                    dst[t1 - 1:t2] = valid_rate(s2i[t1 - padding - 1:t2 - padding], 0, t2 - t1 + 1)
                    dst[t2 - 1:padding + period] = valid_rate(s2i[t2 - padding - 1:period], n_rate - 1, n_rate)
            else:
                dst[padding:padding + period] = invalid_fun(s2i)
            return dst

        st = dict(padding=padding, offset=data_offset, y_S=yS, o_S=oS)
        st["y_casei"] = series(yS, y_N, prof["ycase"], data_offset + r_round(-ydied_latency + ycase_latency),
                               lambda x, a, b: x * np.minimum(1, ycase_rate * y_ifr[a:b]),
                               lambda x: x * ycase_rate * y_died_rate)
        st["y_hospi"] = series(yS, y_N, prof["yhosp"], data_offset + r_round(-ydied_latency + yhosp_latency),
                               lambda x, a, b: x * yhosp_rate * y_hr[a:b], lambda x: x * yhosp_rate * y_hr[0])
        st["y_deadi"] = series(yS, y_N, prof["ydied"], data_offset,
                               lambda x, a, b: x * y_ifr[a:b], lambda x: x * y_died_rate)
        st["o_casei"] = series(oS, o_N, prof["ocase"], data_offset + r_round(-ydied_latency + ocase_latency),
                               lambda x, a, b: x * np.minimum(1, ocase_rate * o_ifr[a:b]),
                               lambda x: x * ocase_rate * o_died_rate)
        st["o_hospi"] = series(oS, o_N, prof["ohosp"], data_offset + r_round(-ydied_latency + ohosp_latency),
                               lambda x, a, b: x * ohosp_rate * o_hr[a:b], lambda x: x * ohosp_rate * o_hr[0])
        st["o_deadi"] = series(oS, o_N, prof["odied"], data_offset,
                               lambda x, a, b: x * o_ifr[a:b], lambda x: x * o_died_rate)
        return st

    # model-age-groups2i.R:442-536
    def calclogl(self, params):
        D = self.D
        state = self.calculateModel(params, self.P)
        if state["offset"] == InvalidDataOffset:
            state["offset"] = 1
        y_dcasei, o_dcasei = np.asarray(D["y_dcasei"], float), np.asarray(D["o_dcasei"], float)
        dhospi = np.asarray(D["dhospi"], float)
        y_dmorti, o_dmorti = np.asarray(D["y_dmorti"], float), np.asarray(D["o_dmorti"], float)
        r = D["d_reliable_cases"]
        size2 = D["case_nbinom_sizey2"]  # case_nbinom_size2 of analyses/be/age/settings.R:31
        assert size2 == D["case_nbinom_sizeo2"]

        def window(n, L):
            dstart = state["offset"]
            dend = state["offset"] + n - 1
            if dend > L:
                dend = L
                dstart = dend - n + 1
            if dstart < 1:
                dstart = 1
                dend = dstart + n - 1
            return dstart, dend

        dstart, dend = window(len(y_dcasei), len(state["y_casei"]))
        y_loglC = (np.sum(dnbinom_log(rslice(y_dcasei, 1, r - 1), pmax(0.1, rslice(state["y_casei"], dstart, dstart + r)), D["case_nbinom_size1"]))
                   + np.sum(dnbinom_log(rslice(y_dcasei, r, len(y_dcasei)), pmax(0.1, rslice(state["y_casei"], dstart + r + 1, dend)), size2)))
        o_loglC = (np.sum(dnbinom_log(rslice(o_dcasei, 1, r - 1), pmax(0.1, rslice(state["o_casei"], dstart, dstart + r)), D["case_nbinom_size1"]))
                   + np.sum(dnbinom_log(rslice(o_dcasei, r, len(o_dcasei)), pmax(0.1, rslice(state["o_casei"], dstart + r + 1, dend)), size2)))
        dstart, dend = window(len(dhospi), len(state["y_hospi"]))
        H = state["y_hospi"] + state["o_hospi"]
        loglH = np.sum(dnbinom_log(dhospi, pmax(0.001, rslice(H, dstart, dend)), D["hosp_nbinom_size"]))
        dstart, dend = window(len(y_dmorti), len(state["y_deadi"]))
        y_loglD = np.sum(dnbinom_log(y_dmorti, pmax(0.001, rslice(state["y_deadi"], dstart, dend)), D["mort_nbinom_size"]))
        o_loglD = np.sum(dnbinom_log(o_dmorti, pmax(0.001, rslice(state["o_deadi"], dstart, dend)), D["mort_nbinom_size"]))
        terms = [y_loglC, o_loglC, loglH, 0.0, y_loglD, o_loglD]
        return float(y_loglC + o_loglC + loglH + y_loglD + o_loglD), terms, state

    # model-age-groups2i.R:362-438
    def calclogp(self, params):
        D, gamma = self.D, self.gamma
        p = lambda i: params[i - 1]
        lo, ltp = D["lockdown_offset"], D["lockdown_transition_period"]
        betay7, betayo7, betao7 = p(34) * p(31), p(35) * p(33), p(36) * p(32)
        lp = 0.0
        lp += dnorm_log(p(21), 0, 10)
        lp += dnorm_log(lo + ltp + p(25), D["d3"], 10)
        lp += dnorm_log(lo + ltp + p(25) + p(26), D["d4"], 10)
        lp += dnorm_log(p(30), self.G, 3)
        for a, b in ((p(1), p(4)), (p(2), p(5)), (p(3), p(6)), (p(4), p(7)), (p(5), p(8)), (p(6), p(9))):
            lp += dnorm_log(a - b, 0, 2 * gamma)
        for a, b in ((p(7), p(27)), (p(8), p(28)), (p(9), p(29))):  # beta3 = beta2
            lp += dnorm_log(a - b, 0, 0.5 * gamma)
        lp += dnorm_log(p(22), 5, 1) + dnorm_log(p(23), 5, 1) + dnorm_log(p(24), 5, 1)
        lp += dnorm_log(p(14), 21, 4) + dnorm_log(p(19), 21, 4)
        lp += dnorm_log(p(31), 0.6, 0.2) + dnorm_log(p(32), 0.22, 0.2) + dnorm_log(p(33), 0.26, 0.2)
        lp += dnorm_log(betay7, 0.85, 0.2) + dnorm_log(betao7, 0.19, 0.05) + dnorm_log(betayo7, 0.02, 0.02)
        return float(lp)


# ------------------------------------------------------------------ simple --

class Simple:
    Tinc = 3
    died_rate = 0.007

    def __init__(self, data, settings_period, integrator="rk4", m=4, ne=False):
        self.D, self.P, self.integrator, self.m, self.ne = data, settings_period, integrator, m, ne

    # model-simple-ode-drj-cmp.R:15-34
    @staticmethod
    def calcConvolveProfile(latency, latency_sd):
        kend = min(0, int(np.ceil(latency + latency_sd * 1.5)))
        kbegin = min(kend - 1, int(np.floor(latency - latency_sd * 1.5)))
        ks = np.arange(kbegin, kend + 1)
        values = stats.norm.cdf(ks - 0.5, latency, latency_sd) - stats.norm.cdf(ks + 0.5, latency, latency_sd)
        values = values / np.sum(values)
        return dict(kbegin=kbegin, kend=kend, values=values)

    convolute = staticmethod(Age2q.convolute)

    def derivs(self, t, y, ldts, ldte, beta0, Tinf0, betat, Tinft, Nef, seg=None):
        N, a = self.D["N"], 1 / self.Tinc
        S, E, I, R = y
        if min(y) < 0 or max(y) > N:
            return np.zeros(4)
        if seg is None:
            seg = 1 if t > ldte else 0 if t > ldts else -1
        if seg == 1:
            beta, Tinf = betat, Tinft
        elif seg == 0:
            f = (t - ldts) / (ldte - ldts)
            beta, Tinf = beta0 + f * (betat - beta0), Tinf0 + f * (Tinft - Tinf0)
        else:
            beta, Tinf = beta0, Tinf0
        if self.ne:
            Ne = Nef * N
            inf = (beta * I) / Ne * max(0.0, Ne - (N - S))
        else:
            inf = (beta * I) / N * S
        return np.array([-inf, inf - a * E, a * E - I / Tinf, I / Tinf])

    def _solve(self, Y0, times, ldts, ldte, *a):
        return solve(lambda t, y: self.derivs(t, y, ldts, ldte, *a), Y0, times, self.integrator, self.m,
                     breaks=(ldts, ldte), rhs_seg=lambda seg, t, y: self.derivs(t, y, ldts, ldte, *a, seg=seg),
                     active=lambda t: 1 if t > ldte else 0 if t > ldts else -1)

    # model-simple-ode-drj-cmp.R:41-129 (params in MODEL space)
    def calculateModel(self, params, period):
        D = self.D
        N = D["N"]
        R0, Rt, Tinf0, Tinft, hosp_rate, hosp_latency, died_latency, thr = params[:8]
        Nef = params[8] if self.ne else 1.0
        beta0, betat = R0 / Tinf0, Rt / Tinft
        hp = self.calcConvolveProfile(-hosp_latency, 5)
        dp = self.calcConvolveProfile(-died_latency, 5)
        padding = max(-hp["kbegin"], -dp["kbegin"]) + 1
        T = padding + period
        S = np.full(T, N - Initial, dtype=float)
        died = np.zeros(T)
        hosp = np.zeros(T)
        Y0 = [N - Initial, Initial, 0, 0]
        times = np.arange(padding + 1, padding + period + 1)
        out = self._solve(Y0, times, 1e10, 1e10, beta0, Tinf0, betat, Tinft, Nef)
        S[padding:] = out[:, 0]
        died[padding:] = (N - self.convolute(S, padding + 1, padding + period, dp)) * self.died_rate
        data_offset = InvalidDataOffset
        lds = np.nonzero(died > thr)[0] + 1
        if len(lds) > 0:
            data_offset = lds[0] - D["lockdown_offset"]
            ldts = data_offset + D["lockdown_offset"]
            ldte = ldts + D["lockdown_transition_period"]
            out = self._solve(Y0, times, ldts, ldte, beta0, Tinf0, betat, Tinft, Nef)
        S[padding:] = out[:, 0]
        hosp[padding:] = (N - self.convolute(S, padding + 1, padding + period, hp)) * hosp_rate
        died[padding:] = (N - self.convolute(S, padding + 1, padding + period, dp)) * self.died_rate
        deadi = np.concatenate([[died[0]], np.diff(died)])
        hospi = np.concatenate([[hosp[0]], np.diff(hosp)])
        return dict(padding=padding, offset=data_offset, S=S, hospi=hospi, deadi=deadi)

    # model-simple-ode-drj-cmp.R:183-257 (params in MODEL space)
    def calclogl(self, params):
        D = self.D
        state = self.calculateModel(params, self.P)
        if state["offset"] == InvalidDataOffset:
            state["offset"] = 1
        dhospi, dmorti = np.asarray(D["dhospi"], float), np.asarray(D["dmorti"], float)
        loglLD = dnbinom_log(D["total_deaths_at_lockdown"], pmax(0.1, params[7]), D["mort_nbinom_size"])[0]

        def window(n, L):
            dstart, dend = state["offset"], state["offset"] + n - 1
            if dend > L:
                dend = L
                dstart = dend - n + 1
            if dstart < 1:
                dstart = 1
                dend = dstart + n - 1
            return dstart, dend

        dstart, dend = window(len(dhospi), len(state["hospi"]))
        loglH = np.sum(dnbinom_log(dhospi, pmax(0.1, rslice(state["hospi"], dstart, dend)), D["hosp_nbinom_size"]))
        dstart, dend = window(len(dmorti), len(state["deadi"]))
        loglD = np.sum(dnbinom_log(dmorti, pmax(0.1, rslice(state["deadi"], dstart, dend)), D["mort_nbinom_size"]))
        terms = [loglLD, loglH, loglD]
        return float(np.sum(terms)), terms, state

    # model-simple-ode-drj-cmp.R:156-179 (params = UNTRANSFORMED sampler vector)
    def calclogp(self, params):
        Tinf0, Tinft, died_latency = params[2], params[3], params[6]
        G0, Gt = self.Tinc + Tinf0 / 2, self.Tinc + Tinft / 2
        return float(dnorm_log(died_latency, 21, 4) + dnorm_log(G0, 5, 1) + dnorm_log(G0 - Gt, 0, 1.5))
