"""GPU: asynchronous state hand-off (aq_snapshot / aq_snapshot_fetch) and the checkpoints written through it
(checkpoint_, reference R/utils.R:571-627)."""
import glob
import threading

import numpy as np
import pytest

from problems import make_problem, sweep_inputs

pytestmark = pytest.mark.gpu


def test_snapshot_is_frozen_while_sweeps_go_on():
    from atlasqtl_b200 import _lib
    from atlasqtl_b200.device import SweepContext
    X, Y, hyper, init = make_problem(300, 500, 700)
    p, q = X.shape[1], Y.shape[1]
    si = sweep_inputs(X, Y, init, c=0.8)
    with SweepContext(X, Y) as ctx:
        with pytest.raises(_lib.AtlasqtlB200Error):
            ctx.snapshot()                      # no state yet
        ctx.set_state(si["gam"], si["mu"])
        with pytest.raises(_lib.AtlasqtlB200Error):
            ctx.snapshot_fetch()                # no snapshot yet
        ctx.refresh_tables(si["theta"], si["zeta"], c_next=si["c"])
        before = ctx.get_state()
        ctx.snapshot()
        ctx.sweep(si["c"], si["log_sig2_inv"], si["tau"], si["log_tau"], si["sig2_beta"])
        got = ctx.snapshot_fetch()
        after = ctx.get_state()
        for key in ("gam_vb", "mu_beta_vb", "beta_vb"):
            assert np.array_equal(got[key], before[key]), key      # bitwise: the pre-sweep state
        assert np.abs(after["gam_vb"] - before["gam_vb"]).max() > 1e-6   # ... which the sweep has left behind
        # fetch on another host thread while this one keeps sweeping
        ctx.snapshot()
        box = {}
        th = threading.Thread(target=lambda: box.update(ctx.snapshot_fetch(mu=False)))
        th.start()
        for _ in range(3):
            ctx.sweep(si["c"], si["log_sig2_inv"], si["tau"], si["log_tau"], si["sig2_beta"])
        th.join()
        assert np.array_equal(box["gam_vb"], after["gam_vb"]) and np.array_equal(box["beta_vb"], after["beta_vb"])
        assert box["mu_beta_vb"] is None
        # a second snapshot replaces the first
        now = ctx.get_state()
        ctx.snapshot()
        assert np.array_equal(ctx.snapshot_fetch()["gam_vb"], now["gam_vb"])


def test_checkpoints_written_in_the_background_hold_the_state_of_their_iteration(tmp_path):
    from atlasqtl_b200 import core
    X, Y, hyper, init = make_problem(120, 90, 40)
    q = Y.shape[1]
    base = str(tmp_path) + "/"
    kw = dict(debug=True)
    full = core.atlasqtl_global_local_core_(Y, X, q, None, 1, 1e-300, 7, 0, hyper, init, checkpoint_path=base,
                                            checkpoint_rate=3, keep_checkpoints=True, **kw)
    assert full["it"] == 7
    files = sorted(glob.glob(base + "tmp_output_it_*.npz"))
    assert [f.rsplit("_", 1)[1] for f in files] == ["3.npz", "6.npz"]
    for it, f in zip((3, 6), files):
        ref = core.atlasqtl_global_local_core_(Y, X, q, None, 1, 1e-300, it, 0, hyper, init, full_output=True, **kw)
        d = np.load(f)
        assert int(d["it"]) == it
        assert np.array_equal(d["gam_vb"], ref["gam_vb"]) and np.array_equal(d["beta_vb"], ref["beta_vb"])
        assert np.array_equal(d["mu_beta_vb"], ref["mu_beta_vb"])
        assert np.array_equal(d["theta_vb"], ref["theta_vb"]) and np.array_equal(d["zeta_vb"], ref["zeta_vb"])
    # cleaned up at the end of a run unless asked otherwise (checkpoint_clean_up_, R/utils.R:612-625)
    core.atlasqtl_global_local_core_(Y, X, q, None, 1, 1e-300, 4, 0, hyper, init, checkpoint_path=base, checkpoint_rate=3, **kw)
    assert sorted(glob.glob(base + "tmp_output_it_*.npz")) == []
