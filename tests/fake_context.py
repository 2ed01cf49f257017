"""TEST DOUBLE of atlasqtl_b200.device.SweepContext backed by the CPU oracle.

Lets the host-side logic of atlasqtl_b200.core (sharding over traits, the all-reduce protocol, the
re-expression of the R updates in terms of per-trait / per-SNP sums) run without a GPU, e.g. in the
world_size-2 gloo tests.  Lives under tests/ because only tests may touch oracle/.
"""
import numpy as np
from scipy import special as sp

from oracle import native

LOG_SQRT_2PI = 0.5 * np.log(2 * np.pi)


class OracleSweepContext:
    def __init__(self, X, Y, device=0, form="primal"):
        self.X = np.asfortranarray(X, dtype=np.float64)
        self.Y = np.asfortranarray(Y, dtype=np.float64)
        self.n, self.p = self.X.shape
        self.q = self.Y.shape[1]
        self.xnorm2 = np.asfortranarray(np.sum(self.X ** 2, axis=0))
        self.order = np.arange(self.p, dtype=np.int32)
        self.launches = 0

    def close(self):
        pass

    def set_order(self, shuffled_ind=None):
        self.order = (np.arange(self.p, dtype=np.int32) if shuffled_ind is None
                      else np.ascontiguousarray(shuffled_ind, dtype=np.int32))

    def _sums(self):
        beta = self.gam * self.mu
        return dict(colsum_gam=self.gam.sum(axis=0), colsum_gam_mu2=(self.gam * self.mu ** 2).sum(axis=0),
                    colsum_beta2=(beta ** 2).sum(axis=0), resid_sq=(self.R ** 2).sum(axis=0))

    def set_state(self, gam_vb, mu_beta_vb):
        self.gam = np.array(gam_vb, dtype=np.float64, order="F")
        self.mu = np.array(mu_beta_vb, dtype=np.float64, order="F")
        self.beta = np.asfortranarray(self.gam * self.mu)
        self.R = native.residual(self.X, self.Y, self.beta)
        return self._sums()

    def get_state(self, gam=True, mu=True, beta=True):
        return dict(gam_vb=self.gam.copy() if gam else None, mu_beta_vb=self.mu.copy() if mu else None,
                    beta_vb=(self.gam * self.mu) if beta else None)

    def get_residual(self):
        return self.R.copy()

    def refresh_tables(self, theta_vb, zeta_vb, c_next=1.0, want_elbo=False):
        u = np.asarray(theta_vb)[:, None] + np.asarray(zeta_vb)[None, :]
        self.log_Phi = np.asfortranarray(sp.log_ndtr(u))
        self.log_1_min_Phi = np.asfortranarray(sp.log_ndtr(-u))
        sc = 1.0 if abs(c_next - 1) < 1.5e-8 else np.sqrt(c_next)
        U = sc * u
        lp, lq = (self.log_Phi, self.log_1_min_Phi) if sc == 1.0 else (sp.log_ndtr(U), sp.log_ndtr(-U))
        m1 = np.exp(-U ** 2 / 2 - LOG_SQRT_2PI - lp)
        m1 = np.where(m1 < -U, -U, m1)
        m0 = -np.exp(-U ** 2 / 2 - LOG_SQRT_2PI - lq)
        m0 = np.where(m0 > -U, -U, m0)
        self.W, self.I0 = m1 - m0, m0
        if want_elbo:
            eps = np.finfo(np.float64).eps ** 0.75
            g = self.gam
            return float(np.sum(g * self.log_Phi + (1 - g) * self.log_1_min_Phi - g * np.log(g + eps)
                                - (1 - g) * np.log(1 - g + eps)))
        return None

    def sweep(self, c, log_sig2_inv_vb, tau_vb, log_tau_vb, sig2_beta_vb):
        native.sweep_primal(self.X, self.xnorm2, self.R, self.gam, self.log_Phi, self.log_1_min_Phi,
                            float(log_sig2_inv_vb), np.ascontiguousarray(log_tau_vb), self.beta, self.mu,
                            np.ascontiguousarray(sig2_beta_vb), np.ascontiguousarray(tau_vb), self.order, c=c)
        self.launches += 1
        out = self._sums()
        out["colsum_zpart"] = (self.gam * self.W + self.I0).sum(axis=0)
        return out

    # ---- missing responses: same contract as SweepContext.set_missing / set_state_mis / sweep_mis
    def set_missing(self, mis_pat):
        self.mis = np.asfortranarray(mis_pat, dtype=np.float64)
        self.Y = np.asfortranarray(self.Y * self.mis)
        self.xnsq = np.asfortranarray((self.X ** 2).T @ self.mis)
        return self.mis.sum(axis=0)

    def set_state_mis(self, gam_vb, mu_beta_vb):
        self.gam = np.array(gam_vb, dtype=np.float64, order="F")
        self.mu = np.array(mu_beta_vb, dtype=np.float64, order="F")
        self.beta = np.asfortranarray(self.gam * self.mu)
        self.R = np.asfortranarray(self.mis * (self.Y - self.X @ self.beta))
        out = self._sums()
        out.update(colsum_xn_gam=(self.xnsq * self.gam).sum(axis=0),
                   colsum_xn_gam_mu2=(self.xnsq * self.gam * self.mu ** 2).sum(axis=0),
                   colsum_xn_beta2=(self.xnsq * self.beta ** 2).sum(axis=0))
        return out

    def sweep_mis(self, c, log_sig2_inv_vb, sig2_inv_vb, tau_vb, log_tau_vb):
        tau = np.ascontiguousarray(tau_vb, dtype=np.float64)
        s2 = np.asfortranarray(1.0 / (c * (self.xnsq + sig2_inv_vb) * tau[None, :]))
        native.sweep_primal_mis(self.X, self.mis, self.xnsq, self.R, self.gam, self.log_Phi, self.log_1_min_Phi,
                                float(log_sig2_inv_vb), np.ascontiguousarray(log_tau_vb), self.beta, self.mu, s2, tau,
                                self.order, c=c)
        self.launches += 1
        g, m, xn = self.gam, self.mu, self.xnsq
        return dict(colsum_gam=g.sum(axis=0), colsum_gam_mu2=(g * m ** 2).sum(axis=0), colsum_sig2b_gam=(s2 * g).sum(axis=0),
                    colsum_xn_gam_mu2=(xn * g * m ** 2).sum(axis=0), colsum_xn_sig2b_gam=(xn * s2 * g).sum(axis=0),
                    colsum_xn_beta2=(xn * (g * m) ** 2).sum(axis=0), resid_sq=(self.R ** 2).sum(axis=0),
                    colsum_zpart=(g * self.W + self.I0).sum(axis=0), colsum_gam_logsig2b=(g * np.log(s2)).sum(axis=0))

    def rowsums_zpart(self):
        return (self.gam * self.W + self.I0).sum(axis=1)

    def launch_count(self):
        return self.launches

    def last_sweep_ms(self):
        return float("nan")


class SnapshotOracleSweepContext(OracleSweepContext):
    """The test double with the asynchronous hand-off of the real context (aq_snapshot / aq_snapshot_fetch): the host
    loop then writes its checkpoints on a worker thread."""

    def snapshot(self):
        self._snap = (self.gam.copy(), self.mu.copy())
        self.snapshots = getattr(self, "snapshots", 0) + 1

    def snapshot_fetch(self, gam=True, mu=True, beta=True):
        import threading
        self.fetch_threads = getattr(self, "fetch_threads", set()) | {threading.current_thread().name}
        g, m = self._snap
        return dict(gam_vb=g.copy() if gam else None, mu_beta_vb=m.copy() if mu else None,
                    beta_vb=(g * m) if beta else None)
