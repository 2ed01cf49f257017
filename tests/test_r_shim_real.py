"""CPU: the SHIPPED .Call shim (bindings/R/atlasqtl_b200_shim.c) compiled against the stub R runtime of tests/r_stub and
executed: routine registration (R_init_atlasqtl), arities, the argument checks that fire before the library is entered,
propagation of a library error through Rf_error(), balanced PROTECT stacks.  The end-to-end run through this shim on the
CUDA library is tests/test_gpu_r_shim_real.py."""
import re

import numpy as np
import pytest

from oracle.rlite.values import RError, RList, from_py, intv


@pytest.fixture(scope="module")
def shim():
    from r_shim_real import RealShim
    return RealShim()


def call(shim, sym, *args):
    return shim(None, [from_py(sym)] + list(args), {})


def mat(nrow, ncol, fill=0.5):
    return from_py(np.full((nrow, ncol), fill, order="F"))


def test_registration_table_matches_the_shim_and_the_reference_symbols(shim):
    import os
    from r_shim_real import SHIM_SRC
    text = open(SHIM_SRC).read()
    declared = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'\{"(_atlasqtl_\w+)",\s*\(DL_FUNC\)&\w+,\s*(\d+)\}', text))
    assert {k: v[1] for k, v in shim.routines.items()} == declared and len(declared) == 18
    # the two symbols of the reference (src/RcppExports.cpp:65-69) keep their arities
    assert shim.routines["_atlasqtl_coreDualLoop"][1] == 15 and shim.routines["_atlasqtl_coreDualMisLoop"][1] == 16
    assert shim.lib.rstub_dynamic_symbols() == 0          # R_useDynamicSymbols(dll, FALSE), as the reference
    assert os.path.exists(SHIM_SRC)


def test_unknown_symbol_and_wrong_arity_are_r_errors(shim):
    with pytest.raises(RError, match="not a registered routine"):
        call(shim, "_atlasqtl_nope", mat(2, 2))
    with pytest.raises(RError, match="expecting 3"):
        call(shim, "_atlasqtl_aq_create", mat(4, 3), mat(4, 2))


def test_argument_checks_fire_before_the_library_is_entered(shim):
    with pytest.raises(RError, match="X and Y must be double matrices"):
        call(shim, "_atlasqtl_aq_create", from_py(np.ones((4, 3), dtype=np.int64)), mat(4, 2), intv(0))
    with pytest.raises(RError, match="X and Y must be double matrices"):
        call(shim, "_atlasqtl_aq_create", from_py(np.ones(12)), mat(4, 2), intv(0))
    with pytest.raises(RError, match="same number of rows"):
        call(shim, "_atlasqtl_aq_create", mat(4, 3), mat(5, 2), intv(0))
    with pytest.raises(RError, match="X must be a double matrix"):
        call(shim, "_atlasqtl_aq_prep_x", from_py(np.ones(6)), intv(0))
    with pytest.raises(RError, match="geno must be a raw vector"):
        call(shim, "_atlasqtl_aq_prep_geno", from_py(np.ones(6)), intv(4), intv(3), from_py(1.0), intv(0))
    with pytest.raises(RError, match="already destroyed"):
        from r_shim_real import StubPtr
        call(shim, "_atlasqtl_aq_snapshot", StubPtr(shim.lib.R_MakeExternalPtr(None, None, None)))
    # stateless 15-argument entry: every matrix / vector is checked against the dims gam_vb implies
    p, q = 3, 2
    good = [mat(p, p), mat(q, p), mat(p, q), mat(p, q), mat(p, q), from_py(0.1), from_py(np.ones(q)), mat(p, q), mat(p, q),
            mat(p, q), from_py(np.ones(q)), from_py(np.ones(q)), intv(np.arange(p)), intv(np.arange(q)), from_py(1.0)]
    for k, bad, msg in ((0, mat(2, 2), r"cp_X must be 3 x 3 \(is 2 x 2\)"), (1, mat(p, q), "cp_Y_X must be 2 x 3"),
                        (2, from_py(np.ones(6)), "gam_vb must be a double matrix"), (5, from_py(np.ones(2)), "single number"),
                        (6, from_py(np.ones(3)), "log_tau_vb must have length 2"),
                        (12, from_py(np.arange(3.0)), "shuffled_ind must be an integer vector"),
                        (14, from_py("a"), None)):
        args = list(good)
        args[k] = bad
        with pytest.raises(RError, match=msg):
            call(shim, "_atlasqtl_coreDualLoop", *args)
    # 16-argument entry: cp_X_rm must be a list of q p x p matrices
    good16 = [good[0], RList([mat(p, p) for _ in range(q)]), *good[1:10], mat(p, q), *good[11:]]
    bad16 = list(good16)
    bad16[1] = RList([mat(p, p)])
    with pytest.raises(RError, match="cp_X_rm must be a list of 2 matrices"):
        call(shim, "_atlasqtl_coreDualMisLoop", *bad16)
    bad16[1] = RList([mat(p, p), mat(p, 2)])
    with pytest.raises(RError, match=r"cp_X_rm\[\[k\]\] must be 3 x 3"):
        call(shim, "_atlasqtl_coreDualMisLoop", *bad16)


def test_a_library_error_becomes_an_r_error_with_its_message(shim):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: aq_create succeeds")
    with pytest.raises(RError, match=r"atlasqtl_b200 \(-?\d+\): "):      # no CPU fallback: aq_create fails loudly
        call(shim, "_atlasqtl_aq_create", mat(8, 3), mat(8, 2), intv(0))
