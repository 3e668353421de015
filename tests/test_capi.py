"""CPU tests of the boundary: the library loads, exports every symbol the header declares,
and refuses to compute without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "epimcmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(epi_[a-z0-9_]+)\s*\(", src)))


def test_header_and_ctypes_symbol_lists_agree():
    from epi_mcmc_b200 import _capi
    assert sorted(_capi.SYMBOLS) == _header_symbols()


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from epi_mcmc_b200 import _capi
    lib = C.CDLL(_capi.LIB_PATH)
    for name in _header_symbols():
        assert hasattr(lib, name), name
    assert lib.epi_abi_version() == 2


def test_no_cpu_fallback():
    """Without a CUDA device epi_create must fail with EPI_ERR_NO_DEVICE — the product never
    routes through the oracle or any CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from epi_mcmc_b200 import _capi
    from epi_mcmc_b200.model import Model, load_dataset
    with pytest.raises(_capi.EpiError, match="EPI_ERR_NO_DEVICE"):
        Model(load_dataset("S120"))


def test_product_does_not_reference_oracle():
    """nothing under epi_mcmc_b200/ imports, links or names the oracle"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "epi_mcmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "libepi_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_struct_layout_matches_header():
    """ctypes mirror of epi_data / epi_model_desc / epi_sampler_cfg has the C sizes"""
    import subprocess
    import tempfile
    from epi_mcmc_b200 import _capi
    src = '#include <stdio.h>\n#include "epimcmc.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(epi_model_desc), sizeof(epi_data), sizeof(epi_sampler_cfg));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(_capi.ModelDesc), C.sizeof(_capi.Data), C.sizeof(_capi.SamplerCfg)]


def test_r_shim_type_checks_against_the_header():
    """r/epimcmc_shim.c cannot be built here (no R), but it must at least type-check against include/epimcmc.h
    with a declarations-only stand-in for R's C API (tests/r_stub): every C_* wrapper calls existing entry points
    with the right argument types, and registers the arity it really has."""
    import re
    import subprocess
    from conftest import ROOT
    src = os.path.join(ROOT, "r", "epimcmc_shim.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type",
                        "-I" + os.path.join(ROOT, "tests", "r_stub"), "-I" + os.path.join(ROOT, "include"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    for name, arity in re.findall(r'\{"(C_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', text):
        m = re.search(r"SEXP " + name + r"\(([^)]*)\)", text)
        assert m and len(m.group(1).split(",")) == int(arity), name
    # every .Call in the drop-in model files names a registered routine with that many arguments
    registered = dict(re.findall(r'\{"(C_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', text))
    for fn in ("model-age-groups2q-gpu.R", "model-age-groups2i-gpu.R", "model-age-groups2x-gpu.R", "model-simple-ode-drj-cmp-gpu.R"):
        rsrc = open(os.path.join(ROOT, "r", fn)).read()
        for call in re.findall(r'\.Call\("(C_\w+)"', rsrc):
            assert call in registered, (fn, call)


def test_r_files_call_only_registered_shim_routines_with_the_right_arity():
    """r/*.R are source-only here (no R): at least every .Call("C_...", ...) they make must name a routine that
    r/epimcmc_shim.c registers, with the registered number of arguments, and every registered routine must be
    defined with that many SEXP parameters."""
    import re
    rdir = os.path.join(ROOT, "r")
    shim = open(os.path.join(rdir, "epimcmc_shim.c")).read()
    registered = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(C_\w+)",\s*\(DL_FUNC\)&\1,\s*(\d+)\}', shim)}
    assert {"C_epi_create", "C_epi_logpost", "C_epi_trajectories", "C_epi_quantiles", "C_epi_df_params",
            "C_epi_run_mcmc"} <= set(registered)
    for name, arity in registered.items():
        m = re.search(r"^SEXP %s\(([^)]*)\)" % name, shim, re.M | re.S)
        assert m, name
        assert len(re.findall(r"\bSEXP\b", m.group(1))) == arity, name

    def call_args(text, start):
        """number of top-level arguments of the call whose '(' is at text[start]"""
        depth, n, i, seen = 0, 0, start, False
        while i < len(text):
            c = text[i]
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
                if depth == 0:
                    return n + (1 if seen else 0)
            elif c == "," and depth == 1:
                n += 1
            elif c == '"':
                i = text.index('"', i + 1)
                seen = True
            elif not c.isspace() and depth >= 1:
                seen = True
            i += 1
        raise AssertionError("unbalanced call")

    n_calls = 0
    for fn in sorted(os.listdir(rdir)):
        if not fn.endswith(".R"):
            continue
        text = re.sub(r"#[^\n]*", "", open(os.path.join(rdir, fn)).read())   # strip comments
        for m in re.finditer(r'\.Call\(\s*"(C_\w+)"', text):
            name = m.group(1)
            assert name in registered, (fn, name)
            assert call_args(text, m.start() + len(".Call")) - 1 == registered[name], (fn, name)
            n_calls += 1
    assert n_calls >= 10


def test_r_model_files_pass_only_globals_the_shim_reads():
    """A misspelt name in a model file's epi_globals list would be ignored silently (the shim's lookup falls back
    to 0): every name passed must be one C_epi_create looks up, and the per-flavour essentials must be there."""
    import re
    rdir = os.path.join(ROOT, "r")
    shim = open(os.path.join(rdir, "epimcmc_shim.c")).read()
    looked_up = set(re.findall(r'\b(?:num|vec)\(g, "(\w+)"', shim))
    assert len(looked_up) > 30
    need = {"model-age-groups2q-gpu.R": {"y_N", "o_N", "y_ifr", "o_wdmorti", "d8", "hosp_last7", "g", "gtrimp"},
            "model-age-groups2i-gpu.R": {"y_N", "o_N", "y_ifr", "d7", "hosp_nbinom_size"},
            "model-age-groups2x-gpu.R": {"y_N", "o_N", "y_ifr", "o_wdmorti", "d9", "d10", "d12", "duncertain", "d_symp_cases",
                                         "d_all_cases", "symp_cases_factor", "hosp_last7", "g", "gtrimp"},
            "model-simple-ode-drj-cmp-gpu.R": {"N", "dhospi", "dmorti", "mort_nbinom_size"}}
    for fn, essentials in need.items():
        text = re.sub(r"#[^\n]*", "", open(os.path.join(rdir, fn)).read())
        start = text.index("epi_globals <- list(") + len("epi_globals <- list")
        depth, i = 0, start
        while True:
            depth += text[i] == "("
            depth -= text[i] == ")"
            i += 1
            if depth == 0:
                break
        body = text[start + 1:i - 1]
        # top-level `name = value` pairs
        keys, depth, tok = [], 0, ""
        for c in body + ",":
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
            if c == "," and depth == 0:
                keys.append(tok.split("=")[0].strip())
                tok = ""
            else:
                tok += c
        keys = {k for k in keys if k}
        assert keys <= looked_up, (fn, sorted(keys - looked_up))
        assert essentials <= keys, (fn, sorted(essentials - keys))


def test_r_driver_passes_its_warm_start_to_the_sampler():
    """the GPU driver read max.csv into `init` but started every chain from the library's compiled-in
    init (theta0 = NULL).  Static check (no R here): the starting cloud is built from `init` in R and handed to
    C_epi_run_mcmc, and the sampling phase of the shim no longer refreshes the DE-MC snapshot."""
    with open(os.path.join(ROOT, "r", "fitMCMC-gpu.R")) as f:
        src = f.read()
    call = re.search(r'\.Call\("C_epi_run_mcmc",\s*epi_handle,\s*(\w+),', src)
    assert call and call.group(1) == "theta0"
    assert re.search(r"theta0 <- matrix\(init,", src) and "df_params$min" in src and 'read.csv("max.csv")' in src
    with open(os.path.join(ROOT, "r", "epimcmc_shim.c")) as f:
        shim = f.read()
    body = shim[shim.index("SEXP C_epi_run_mcmc"):]
    burn, samp = body.split("epi_sampler_record", 1)
    assert "epi_sampler_set_population" in burn and "epi_sampler_set_population" not in samp
