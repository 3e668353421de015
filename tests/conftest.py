import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

DATASETS = ("S120", "S365", "NE120", "A300", "A365", "I290", "X300", "DE", "SE_NE")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(ROOT, "tests", "golden", f"golden_{name}.json")) as f:
        return json.load(f)


def nan_array(lst):
    return np.array([np.nan if v is None else v for v in lst], dtype=np.float64)


def rel_err(a, b):
    """elementwise |a-b| / max(|b|, tiny); equal non-finite values (nan/nan, -inf/-inf) count as 0"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    same = (np.isnan(a) & np.isnan(b)) | ((a == b) & ~np.isfinite(a))
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    r = np.where(same, 0.0, r)
    return np.where(np.isnan(r), np.inf, r)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


def oracle_conditioning(oracle, fl, ds, theta, P, m, logl):
    return oracle.conditioning(fl, ds, theta, P, m, logl)


def logl_within_bar(got, ref, cond, tol=1e-10):
    """parity bar: |got - ref| <= max(tol * |ref|, 4 * cond); equal non-finite values pass"""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    same = (np.isnan(got) & np.isnan(ref)) | ((got == ref) & ~np.isfinite(got))
    with np.errstate(invalid="ignore"):
        ok = np.abs(got - ref) <= np.maximum(tol * np.abs(ref), 4.0 * np.asarray(cond))
    return ok | same | np.isinf(np.asarray(cond))  # cond = inf: the oracle itself flips NA <-> finite under 32-ulp jitter
