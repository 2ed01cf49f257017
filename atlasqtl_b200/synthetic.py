"""Synthetic genotype / phenotype data of the shapes BASELINE.json names.

Recipes follow the reference's own examples: X ~ Binomial(2, maf) column-standardised with the
n-1 standard deviation (R/atlasqtl.R:131-132, R/prepare_atlasqtl.R:57), effects on the first
p_act SNPs x q_act traits with a Bernoulli(0.2) pattern and N(0,1) sizes, Y = X_raw beta + N(0,1),
centred (R/atlasqtl.R:142-157, R/prepare_atlasqtl.R:83).  Seed 123 as in tests/testthat/main.R:3.
"""
import numpy as np


def standardise_genotypes(G):
    """scale(X): centre, divide by the n-1 sd; constant columns are dropped
    (R/prepare_atlasqtl.R:57-60)."""
    G = np.asarray(G, dtype=np.float64)
    mu = G.mean(axis=0)
    sd = G.std(axis=0, ddof=1)
    keep = sd > 0
    return np.asfortranarray((G[:, keep] - mu[keep]) / sd[keep]), keep


def simulate(n, p, q, p_act, q_act, maf=0.25, prob_assoc=0.2, beta_sd=1.0, seed=123, hotspots=None):
    """Returns X (n x p', standardised, Fortran), Y (n x q, centred, Fortran), pat (p' x q bool)."""
    rng = np.random.default_rng(seed)
    G = rng.binomial(2, maf, size=(n, p)).astype(np.float64)
    pat = np.zeros((p, q), dtype=bool)
    if hotspots is None:
        pat[:p_act, :q_act] = rng.random((p_act, q_act)) < prob_assoc
        for k in range(q_act):  # every active trait has at least one SNP (R/atlasqtl.R:147-150)
            if not pat[:p_act, k].any():
                pat[rng.integers(p_act), k] = True
    else:  # hotspot-dense recipe (BASELINE configs[3]): pi_h ~ Beta(1, 8), >= 50 traits each
        rows = rng.choice(p, size=hotspots, replace=False)
        for r in rows:
            k_h = max(min(50, q), rng.binomial(q, rng.beta(1.0, 8.0)))
            pat[r, rng.choice(q, size=k_h, replace=False)] = True
    beta = np.zeros((p, q))
    beta[pat] = rng.normal(0.0, beta_sd, size=int(pat.sum()))
    perm_p = rng.permutation(p)
    perm_q = rng.permutation(q)
    G, beta, pat = G[:, perm_p], beta[perm_p][:, perm_q], pat[perm_p][:, perm_q]
    Y = G @ beta + rng.normal(size=(n, q))
    X, keep = standardise_genotypes(G)
    Y = np.asfortranarray(Y - Y.mean(axis=0))
    return X, Y, pat[keep]


CONFIGS = {  # BASELINE.json configs[0..4]
    "C1": dict(n=200, p=500, q=1000, p_act=50, q_act=500, anneal=None),
    "C2": dict(n=1000, p=50000, q=20000, p_act=500, q_act=10000, anneal=(1, 2, 10)),
    "C3": dict(n=3000, p=200000, q=1500, p_act=500, q_act=750, anneal=(1, 2, 10)),
    "C4": dict(n=500, p=10000, q=5000, hotspots=20, beta_sd=0.5, anneal=(1, 2, 10)),
    "C5": dict(n=5000, p=500000, q=20000, p_act=1000, q_act=10000, anneal=(1, 2, 10)),
    # not a BASELINE config: what ONE GPU of the 8 sees of C5 (its 2500-trait slab), on a 32nd of the SNPs -- a one-GPU
    # rehearsal of the C5 code path (packed genotypes, 8-CTA sample-split clusters) for development and profiling
    "C5slab": dict(n=5000, p=16000, q=2500, p_act=100, q_act=1250, anneal=(1, 2, 10)),
}
