"""Generate golden VB trajectories at the BASELINE shapes with the CPU oracle -- TEST INFRASTRUCTURE ONLY.

  python tests/golden/make_trajectory.py C4 [--threads 8]      -> tests/golden/c4_trajectory.npz
  python tests/golden/make_trajectory.py C1                     -> tests/golden/c1_trajectory.npz

Runs the restated outer loop (oracle/vb_oracle.py, R/atlasqtl_global_local_core.R:125-386) to convergence on the
seeded problem `problem(name)` and stores what the parity criteria of BASELINE.md section 5 need:

  lb_it, lb        ELBO at every evaluated iteration (R/atlasqtl_global_local_core.R:342-354)
  it, converged    iteration count at convergence (:362-375)
  sel_ppi          column-major linear indices of {gam_vb > 0.5}          (R/summarise_output.R:99-106)
  sel_fdr          column-major linear indices of {bFDR < 0.05}           (R/summarise_output.R:207-223)
  probe_idx/gam/beta   gam_vb / beta_vb at the selected pairs and at 20 000 seeded random pairs
  theta_vb, zeta_vb    final values
  in_check         checksums of the inputs, so that a test can tell "inputs drifted" from "results differ"

C1 uses sweep = "reference" (the reference's own coreLoop.cpp, p x p inputs feasible at p = 500); C4 the primal
restatement (cp_X at p = 10 000 is 0.8 GB and the dual sweep O(p^2 q) = 5e11 flop-pairs per iteration), pinned to the
reference loop by tests/test_oracle.py.  The GPU box has no /root/reference: it only reads the .npz.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

HERE = os.path.dirname(os.path.abspath(__file__))


def problem(name, seed=123):
    """Seeded inputs of BASELINE config `name` (same recipe on every machine: NumPy Generator streams)."""
    from atlasqtl_b200 import hyper_init, synthetic
    cfg = synthetic.CONFIGS[name]
    if "hotspots" in cfg:
        X, Y, pat = synthetic.simulate(cfg["n"], cfg["p"], cfg["q"], 0, 0, beta_sd=cfg["beta_sd"], seed=seed,
                                       hotspots=cfg["hotspots"])
    else:
        X, Y, pat = synthetic.simulate(cfg["n"], cfg["p"], cfg["q"], cfg["p_act"], cfg["q_act"], seed=seed)
    p = X.shape[1]
    q = Y.shape[1]
    p0 = (max(1.0, float(pat.sum(axis=0).mean())), 10.0)
    hyper = hyper_init.auto_set_hyper_(Y, p, p0)
    init = hyper_init.auto_set_init_(Y, p, p0, q, user_seed=seed)
    return X, Y, hyper, init, cfg["anneal"]


def input_checksums(X, Y, hyper, init):
    return np.array([X.sum(), np.abs(X).sum(), Y.sum(), np.abs(Y).sum(), init["gam_vb"].sum(), init["mu_beta_vb"].sum(),
                     init["theta_vb"].sum(), init["zeta_vb"].sum(), float(init["sig02_inv_vb"]), float(hyper["t02"]),
                     float(hyper["n0"][0]), float(hyper["eta"][0])])


def probe_indices(p, q, sel, seed=5):
    rng = np.random.default_rng(seed)
    extra = rng.integers(0, p * q, size=20000)
    return np.unique(np.concatenate([sel, extra])).astype(np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["C1", "C4"])
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--tol", type=float, default=0.1)
    ap.add_argument("--maxit", type=int, default=1000)
    args = ap.parse_args()
    from oracle import native, vb_oracle
    native.build()
    t0 = time.time()
    X, Y, hyper, init, anneal = problem(args.config)
    p, q = X.shape[1], Y.shape[1]
    print(f"{args.config}: n={X.shape[0]} p={p} q={q} anneal={anneal}; inputs in {time.time() - t0:.1f} s", flush=True)
    form = "reference" if args.config == "C1" and native.ref_available() else "primal"
    trace = []

    class Progress(list):
        def append(self, rec):
            super().append(rec)
            print(f"  it {rec['it']:4d} c={rec['c']:.4f} lb={rec['lb']} [{time.time() - t0:.0f} s]", flush=True)

    trace = Progress()
    out = vb_oracle.atlasqtl_global_local_core_(Y, X, q, anneal, 1, args.tol, args.maxit, hyper, init, sweep=form,
                                                trace=trace, nthreads=args.threads)
    gam = out["gam_vb"]
    sel_ppi = np.flatnonzero(gam.flatten(order="F") > 0.5).astype(np.int64)
    fdr = vb_oracle.assign_bFDR(gam)
    sel_fdr = np.flatnonzero(fdr.flatten(order="F") < 0.05).astype(np.int64)
    probe = probe_indices(p, q, np.union1d(sel_ppi, sel_fdr))
    lb_it = np.array([r["it"] for r in trace if r["lb"] is not None], dtype=np.int64)
    lb = np.array([r["lb"] for r in trace if r["lb"] is not None])
    dest = os.path.join(HERE, f"{args.config.lower()}_trajectory.npz")
    np.savez_compressed(dest, lb_it=lb_it, lb=lb, it=out["it"], converged=out["converged"], sel_ppi=sel_ppi,
                        sel_fdr=sel_fdr, probe_idx=probe, probe_gam=gam.flatten(order="F")[probe],
                        probe_beta=out["beta_vb"].flatten(order="F")[probe], theta_vb=out["theta_vb"],
                        zeta_vb=out["zeta_vb"], in_check=input_checksums(X, Y, hyper, init), tol=args.tol,
                        form=form, sum_gam=float(gam.sum()))
    print(f"wrote {dest}: it={out['it']} converged={out['converged']} |gam>0.5|={len(sel_ppi)} |bFDR<0.05|={len(sel_fdr)} "
          f"in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
