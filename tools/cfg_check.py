"""Dev helper: single-sweep parity for one (n, p, q) through the C ABI vs the CPU oracle. Usage: cfg_check.py n p q [c]"""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from problems import make_problem, sweep_inputs
from oracle import native
from atlasqtl_b200.device import SweepContext
n, p, q = (int(x) for x in sys.argv[1:4])
c = float(sys.argv[4]) if len(sys.argv) > 4 else 0.8
X, Y, hyper, init = make_problem(n, p, q)
p = X.shape[1]
si = sweep_inputs(X, Y, init, c=c)
order = np.random.default_rng(5).permutation(p).astype(np.int32)
gam, mu = si["gam"].copy(order="F"), si["mu"].copy(order="F")
beta = np.asfortranarray(gam * mu)
R = native.residual(X, Y, beta)
xn = np.asfortranarray(np.sum(X ** 2, axis=0))
native.sweep_primal(X, xn, R, gam, si["log_Phi"], si["log_1_min_Phi"], si["log_sig2_inv"], si["log_tau"], beta, mu,
                    si["sig2_beta"], si["tau"], order, c=c, nthreads=8)
with SweepContext(X, Y) as ctx:
    ctx.set_order(order)
    ctx.set_state(si["gam"], si["mu"])
    R0 = ctx.get_residual()
    print("init resid err", np.abs(R0 - (Y - X @ (si["gam"] * si["mu"]))).max())
    ctx.refresh_tables(si["theta"], si["zeta"], c_next=c)
    ctx.sweep(c, si["log_sig2_inv"], si["tau"], si["log_tau"], si["sig2_beta"])
    st = ctx.get_state()
    print(f"n={n} p={p} q={q}: max|dgam|={np.abs(st['gam_vb'] - gam).max():.3e} max|dR|={np.abs(ctx.get_residual() - R).max():.3e} sweep_ms={ctx.last_sweep_ms():.3f}")
