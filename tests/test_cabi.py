"""CPU: the C-ABI library loads and exports every symbol include/atlasqtl_b200.h declares; host-only entry
points behave; the product refuses to run without a GPU instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "atlasqtl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aq_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from atlasqtl_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == syms


def test_host_only_entry_points():
    from atlasqtl_b200 import _lib
    lib = _lib.load()
    assert lib.aq_version() >= 100
    assert lib.aq_create(None, 0, 10, 10, 10, None, None) == -1  # AQ_EINVAL
    assert b"NULL" in lib.aq_last_error()
    assert lib.aq_destroy(None) == 0
    assert lib.aq_launch_count(None) == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from atlasqtl_b200 import _lib
    from atlasqtl_b200.device import SweepContext
    X = np.asfortranarray(np.random.default_rng(0).normal(size=(20, 8)))
    Y = np.asfortranarray(np.random.default_rng(1).normal(size=(20, 3)))
    with pytest.raises(_lib.AtlasqtlB200Error):
        SweepContext(X, Y)


def test_product_never_imports_the_oracle():
    """The oracle is the checker, never part of the product path."""
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.\.?oracle)|oracle[/.](native|vb_oracle|_ref|_build)", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "atlasqtl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_r_shim_compiles_against_the_c_abi():
    """bindings/R/atlasqtl_b200_shim.c cannot be built against R here (no R headers); it is syntax- and type-checked
    against stand-in declarations of the R API names it uses (tests/r_stub) and the real include/atlasqtl_b200.h, and
    every C-ABI call in it must name an exported symbol.  Every wrapper that hands a pointer to the library validates
    the SEXP first (unchecked sizes)."""
    import subprocess
    shim = os.path.join(ROOT, "bindings", "R", "atlasqtl_b200_shim.c")
    res = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type",
                          "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"), shim],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    text = re.sub(r"/\*.*?\*/", "", open(shim).read(), flags=re.S)
    used = set(re.findall(r"\b(aq_[a-zA-Z0-9_]+)\s*\(", text)) - {"aq_ctx", "aq_prep"}
    assert used <= set(declared_symbols()), used - set(declared_symbols())
    # the two .Call symbols of the reference (src/RcppExports.cpp:65-69) are registered with their arities
    assert re.search(r'"_atlasqtl_coreDualLoop",[^}]*15\}', text) and re.search(r'"_atlasqtl_coreDualMisLoop",[^}]*16\}', text)
    # no raw REAL()/INTEGER() of a caller-supplied argument reaches the library without a check, except the in-place
    # gam_vb of the stateless entries (whose dims define p, q and are checked to be a double matrix) and X / Y of aq_create
    for m in re.finditer(r"check\(aq_[a-z_A-Z]+\((.*?)\)\);", text, flags=re.S):
        args = m.group(1)
        raw = re.findall(r"\b(?:REAL|INTEGER)\((\w+)\)", args)
        assert set(raw) <= {"gam_vb", "X", "Y", "v", "n_obs"}, (m.group(0)[:60], raw)
