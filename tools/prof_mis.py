"""Dev helper: time the missing-response sweep (aq_sweep_mis) on random data. Usage: prof_mis.py n p q [nsweeps] [na_frac]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from atlasqtl_b200.device import SweepContext
n, p, q = (int(x) for x in sys.argv[1:4])
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 3
frac = float(sys.argv[5]) if len(sys.argv) > 5 else 0.05
rng = np.random.default_rng(0)
X = rng.standard_normal((n, p)); X = np.asfortranarray((X - X.mean(0)) / X.std(0, ddof=1))
Y = np.asfortranarray(rng.standard_normal((n, q))); Y -= Y.mean(0)
mis = np.asfortranarray((rng.uniform(size=(n, q)) >= frac).astype(np.float64))
gam = np.asfortranarray(rng.uniform(size=(p, q)) ** 8); mu = np.asfortranarray(rng.normal(0, 0.1, (p, q)))
tau = np.ones(q)
with SweepContext(X, Y) as ctx:
    ctx.set_missing(mis)
    ctx.set_state_mis(gam, mu)
    ctx.refresh_tables(rng.normal(0, 0.1, p), rng.normal(-2, 0.3, q))
    for i in range(ns):
        ctx.sweep_mis(1.0, 0.0, 1.0, tau, np.zeros(q))
        ms = ctx.last_sweep_ms()
        print(f"mis sweep {i}: {ms:.3f} ms  {4.0*n*p*q/ms/1e9:.2f} TFLOP/s  {p*q/ms/1e6:.1f} M updates/s")
