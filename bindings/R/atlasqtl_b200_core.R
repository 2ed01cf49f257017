# atlasqtl_b200: the VB outer loop of the `atlasqtl` R package written against the stateful C ABI
# (include/atlasqtl_b200.h through the .Call wrappers of bindings/R/atlasqtl_b200_shim.c).
#
# A maintainer of the R package replaces the body of `atlasqtl_global_local_core_`
# (reference R/atlasqtl_global_local_core.R:8-433) by a call to `atlasqtl_b200_core_` below: same arguments, same
# returned list.  No p x q or n x q object is touched by R inside the loop: the state (gam_vb, mu_beta_vb, the residual,
# the probit tables) lives on the GPU and every update that the reference writes on matrices is re-expressed through the
# per-trait / per-SNP sums the sweep returns.  atlasqtl_b200/core.py is the same loop in Python (with trait-slab
# sharding on top).  The p-, q- and scalar-sized helpers of the package are called unchanged:
#   get_annealing_ladder_ (R/utils.R:108), update_sig2_c0_vb_, update_nu_vb_, update_log_tau_vb_,
#   update_log_sig2_inv_vb_, update_annealed_lam2_inv_vb_ (R/update_vb.R), Q_approx_vec (R/utils.R:380),
#   e_tau_, e_theta_hs_, e_zeta_, e_sig2_inv_, e_sig2_inv_hs_ (R/elbo.R), checkpoint_, checkpoint_clean_up_ (R/utils.R).
#
# What replaces what (reference line numbers):
#   :40-42    Y_norm_sq, cp_X, cp_Y_X             aq_create (X tiles + Gram band + Y on the device)
#   :19-33    mis_pat, X_norm_sq, cp_X_rm         aq_set_missing (returns colSums(mis_pat))
#   :61-63, :293-295  pnorm(.., log.p = TRUE) x 2  aq_refresh_tables (also returns the p x q part of e_beta_gamma_)
#   :112-115  beta_vb, m2_beta, cp_X_Xbeta        aq_set_state / aq_set_state_mis (residual + first sums)
#   :134-150  update_*_vb_ on matrices            the same formulas on s$colsum_* / s$resid_sq
#   :162-176  coreDualLoop / coreDualMisLoop      aq_sweep / aq_sweep_mis
#   :235-237  m2_beta, Z                          s$colsum_zpart, aq_rowsums_zpart
#   :346-354  elbo_global_local_                  terms A, B, E from the sums; C, D, F, G, H unchanged
#   :414-428  output                              aq_get_state
#
# This file is executed in the test-suite by the R evaluator of oracle/rlite (there is no R in the build image) with
# `.Call` bound to the library (tests/test_r_binding.py, tests/test_gpu_r_binding.py).

# ---- small pieces of the loop -------------------------------------------------------------------------------------

# ELBO evaluation / convergence schedule of the package (thinned: look less often while far from convergence)
aq_conv_schedule_ <- function(thinned) {
  if (thinned) list(times = c(1, 5, 10, 50), batch = c(1, 10, 25, 50)) else list(times = 1, batch = 1)
}

# gam_vb and beta_vb from the device.  Two separate allocations: the library fills its arguments in place, and
# `a <- b <- matrix(..)` would make both names point at ONE buffer.
aq_fetch_state_ <- function(ctx, p, q) {
  gam <- matrix(0, p, q)
  beta <- matrix(0, p, q)
  .Call(`_atlasqtl_aq_get_state`, ctx, gam, NULL, beta)
  list(gam_vb = gam, beta_vb = beta)
}

# E[(theta_j - m0)^2] under q(theta_j), in the expanded form the package uses
aq_theta_second_moment_ <- function(theta_vb, sig2_theta_vb, m0) theta_vb^2 + sig2_theta_vb - 2 * theta_vb * m0 + m0^2


atlasqtl_b200_core_ <- function(Y, X, shr_fac_inv, anneal, df, tol, maxit, verbose, list_hyper, list_init,
                                checkpoint_path = NULL, trace_path = NULL, full_output = FALSE,
                                thinned_elbo_eval = TRUE, debug = FALSE, batch = "y", device = 0L,
                                lb_hook = NULL) {   # lb_hook(it, lb_new): called after every ELBO evaluation

  if (batch != "y") stop("Batch scheme not defined. Exit.")
  if (df != 1) stop("atlasqtl_b200 implements the df = 1 horseshoe (the only value atlasqtl() passes).")
  if (!is.null(trace_path)) stop("trace_path (plots of the hotspot variances) is not part of this path.")

  n <- nrow(Y); p <- ncol(X); q <- ncol(Y)
  say <- function(...) if (verbose != 0) cat(paste0(...))

  # ---- device context; missing responses become a 0/1 pattern and zeros
  has_na <- any(is.na(Y))
  if (has_na) {
    observed <- ifelse(is.na(Y), 0, 1)
    Y[is.na(Y)] <- 0
  }
  ctx <- .Call(`_atlasqtl_aq_create`, X, Y, as.integer(device))
  n_eff <- if (has_na) .Call(`_atlasqtl_aq_set_missing`, ctx, observed) else n     # colSums(observed) or n

  # ---- hyper-parameters and starting values (p x q starting matrices go straight to the device)
  eta <- list_hyper$eta; kappa <- list_hyper$kappa
  nu <- list_hyper$nu; rho <- list_hyper$rho
  n0 <- list_hyper$n0; t02 <- list_hyper$t02
  m0 <- list_hyper$m0; A2_inv <- list_hyper$A2_inv

  tau_vb <- list_init$tau_vb
  theta_vb <- list_init$theta_vb
  zeta_vb <- list_init$zeta_vb
  sig2_theta_vb <- list_init$sig2_theta_vb
  sig02_inv_vb <- list_init$sig02_inv_vb
  sig2_beta_vb <- list_init$sig2_beta_vb            # q-vector here; p x q (on the device only) with missing responses

  colsum_xn_m2 <- NULL
  if (has_na) {
    s <- .Call(`_atlasqtl_aq_set_state_mis`, ctx, list_init$gam_vb, list_init$mu_beta_vb)
    colsum_xn_m2 <- s$colsum_xn_gam_mu2 + sig2_beta_vb * s$colsum_xn_gam
  } else {
    s <- .Call(`_atlasqtl_aq_set_state`, ctx, list_init$gam_vb, list_init$mu_beta_vb)
  }
  colsum_m2 <- s$colsum_gam_mu2 + sig2_beta_vb * s$colsum_gam      # colSums((mu^2 + sig2_beta) * gam)
  rm(list_init)

  # ---- annealing ladder, temperature of the first iteration
  annealing <- !is.null(anneal)
  if (annealing) {
    ladder <- get_annealing_ladder_(anneal, verbose)
    c <- ladder[1]
    first_plain_it <- anneal[3]
  } else {
    c <- 1
    first_plain_it <- 1
  }
  c_s <- c          # the scale parameters are annealed with the same temperature

  sched <- aq_conv_schedule_(thinned_elbo_eval)
  sched_pos <- length(sched$batch) + 1
  batch_conv <- 1
  elbo_slack <- .Machine$double.eps^0.5

  t02_inv <- 1 / t02
  sig2_zeta_vb <- update_sig2_c0_vb_(p, t02, c = c)
  log_det_zeta <- - q * (log(t02) + log(p + t02_inv))
  nu_xi_inv_vb <- 1

  .Call(`_atlasqtl_aq_refresh_tables`, ctx, theta_vb, zeta_vb, c, FALSE)

  # sum of squares entering the rate of tau_k: |y_k - X beta_k|^2 plus the variance terms, from the sweep's sums
  tau_rate_terms <- function(s, colsum_m2, colsum_xn_m2, sig2_inv_vb) {
    if (has_na)
      s$resid_sq + sig2_inv_vb * colsum_m2 + colsum_xn_m2 - s$colsum_xn_beta2
    else
      s$resid_sq + (n - 1 + sig2_inv_vb) * colsum_m2 - (n - 1) * s$colsum_beta2
  }

  it <- 0
  converged <- FALSE
  lb_new <- -Inf

  while (!converged && it < maxit) {

    it <- it + 1
    lb_old <- lb_new
    annealed_iteration <- annealing
    if (it == 1 || it %% max(5, batch_conv) == 0) say("Iteration ", format(it), "\n")

    # ---- scalar and q-sized updates from the sums of the previous sweep (the old tau_vb enters rho_vb)
    nu_vb <- update_nu_vb_(nu, sum(s$colsum_gam), c = c)
    rho_vb <- c * (rho + sum(tau_vb * colsum_m2) / 2)
    sig2_inv_vb <- nu_vb / rho_vb

    eta_vb <- c * (eta + n_eff / 2 + s$colsum_gam / 2) - c + 1
    kappa_vb <- c * (kappa + tau_rate_terms(s, colsum_m2, colsum_xn_m2, sig2_inv_vb) / 2)
    tau_vb <- eta_vb / kappa_vb

    log_tau_vb <- update_log_tau_vb_(eta_vb, kappa_vb)
    log_sig2_inv_vb <- update_log_sig2_inv_vb_(nu_vb, rho_vb)

    # horseshoe: these read the theta_vb / sig2_theta_vb / sig02_inv_vb of the previous iteration
    L_vb <- c_s * sig02_inv_vb * shr_fac_inv * aq_theta_second_moment_(theta_vb, sig2_theta_vb, m0) / 2 / df
    rho_xi_inv_vb <- c_s * (A2_inv + sig02_inv_vb)

    # ---- the sweep over all SNP x trait pairs, on the device
    if (has_na) {
      # sig2_beta_vb(j, k) = 1 / (c (X_norm_sq(j, k) + sig2_inv_vb) tau_k) is formed on the device
      s <- .Call(`_atlasqtl_aq_sweep_mis`, ctx, c, log_sig2_inv_vb, sig2_inv_vb, tau_vb, log_tau_vb)
      colsum_m2 <- s$colsum_gam_mu2 + s$colsum_sig2b_gam
      colsum_xn_m2 <- s$colsum_xn_gam_mu2 + s$colsum_xn_sig2b_gam
    } else {
      sig2_beta_vb <- 1 / (c * (n - 1 + sig2_inv_vb) * tau_vb)
      s <- .Call(`_atlasqtl_aq_sweep`, ctx, c, log_sig2_inv_vb, tau_vb, log_tau_vb, sig2_beta_vb)
      colsum_m2 <- s$colsum_gam_mu2 + sig2_beta_vb * s$colsum_gam
    }

    # rowSums(Z), colSums(Z) with Z = (gam (imr1 - imr0) + imr0) / sqrt_c + theta_j + zeta_k
    sqrt_c <- if (isTRUE(all.equal(c, 1))) 1 else sqrt(c)
    rowsums_Z <- .Call(`_atlasqtl_aq_rowsums_zpart`, ctx, p) / sqrt_c + q * theta_vb + sum(zeta_vb)
    colsums_Z <- s$colsum_zpart / sqrt_c + sum(theta_vb) + p * zeta_vb

    # ---- p-sized updates: local scales, theta, global scale, then zeta (theta before zeta)
    if (annealing) {
      lam2_inv_vb <- update_annealed_lam2_inv_vb_(L_vb, c_s, df)
    } else {
      Q_app <- Q_approx_vec(L_vb)
      lam2_inv_vb <- 1 / (Q_app * L_vb) - 1
    }
    xi_inv_vb <- nu_xi_inv_vb / rho_xi_inv_vb

    prior_prec <- sig02_inv_vb * lam2_inv_vb * shr_fac_inv
    sig2_theta_vb <- update_sig2_c0_vb_(q, 1 / prior_prec, c = c)
    theta_vb <- c * sig2_theta_vb * (rowsums_Z + prior_prec * m0 - sum(zeta_vb))

    nu_s0_vb <- update_nu_vb_(1 / 2, p, c = c_s)
    rho_s0_vb <- c_s * (xi_inv_vb + sum(lam2_inv_vb * shr_fac_inv *
                                          aq_theta_second_moment_(theta_vb, sig2_theta_vb, m0)) / 2)
    sig02_inv_vb <- as.numeric(nu_s0_vb / rho_s0_vb)

    zeta_vb <- c * sig2_zeta_vb * (colsums_Z + t02_inv * n0 - sum(theta_vb))

    # ---- next temperature, or whether this iteration evaluates the ELBO
    c_next <- c
    want_elbo <- FALSE
    if (annealing) {
      if (it == 1 || it %% 5 == 0) say("Temperature = ", format(1 / c, digits = 4), "\n")
      c_next <- if (it < length(ladder)) ladder[it + 1] else 1
      sig2_zeta_vb <- c * sig2_zeta_vb / c_next
      if (isTRUE(all.equal(c_next, 1))) {
        annealing <- FALSE
        say("Annealing done.\n")
      }
    } else {
      want_elbo <- it <= first_plain_it + 1 || it %% batch_conv == 0 || it %% batch_conv == 1
    }

    # tables of theta_j + zeta_k for the next sweep; on demand the p x q part of e_beta_gamma_:
    # sum(gam log Phi + (1 - gam) log(1 - Phi) - gam log(gam + eps) - (1 - gam) log(1 - gam + eps))
    elbo_B_dev <- .Call(`_atlasqtl_aq_refresh_tables`, ctx, theta_vb, zeta_vb, c_next, want_elbo)

    if (want_elbo) {

      # the ELBO re-derives eta, kappa, nu, rho at c = 1 from the post-sweep sums
      eta_e <- eta + n_eff / 2 + s$colsum_gam / 2
      kappa_e <- kappa + tau_rate_terms(s, colsum_m2, colsum_xn_m2, sig2_inv_vb) / 2
      nu_e <- update_nu_vb_(nu, sum(s$colsum_gam))
      rho_e <- rho + sum(tau_vb * colsum_m2) / 2

      log_tau_e <- update_log_tau_vb_(eta_e, kappa_e)
      log_sig2_inv_e <- update_log_sig2_inv_vb_(nu_e, rho_e)
      log_sig02_inv_e <- update_log_sig2_inv_vb_(nu_s0_vb, rho_s0_vb)
      log_xi_inv_e <- update_log_sig2_inv_vb_(nu_xi_inv_vb, rho_xi_inv_vb)

      # E log p(y | .): e_y_ on sums
      term_y <- sum(n_eff * (log_tau_e - log(2 * pi)) / 2 - tau_vb * (kappa_e - colsum_m2 * sig2_inv_vb / 2 - kappa))

      # E log p(beta, gamma | .) - E log q(beta, gamma): e_beta_gamma_ on sums + the device part
      gam_log_s2b <- if (has_na) s$colsum_gam_logsig2b else s$colsum_gam * log(sig2_beta_vb)
      term_bg <- sum(s$colsum_gam * (log_sig2_inv_e / 2 + log_tau_e / 2 + 1 / 2) + gam_log_s2b / 2) -
        sum(colsum_m2 * tau_vb) * sig2_inv_vb / 2 + elbo_B_dev -
        p * q * sig2_zeta_vb / 2 - q * sum(sig2_theta_vb) / 2

      term_theta <- e_theta_hs_(lam2_inv_vb, L_vb, log_sig02_inv_e + log(shr_fac_inv), m0, theta_vb, Q_app,
                                sig02_inv_vb * shr_fac_inv, sig2_theta_vb, df)
      term_zeta <- e_zeta_(zeta_vb, n0, sig2_zeta_vb, t02_inv, log_det_zeta)
      term_tau <- e_tau_(eta, eta_e, kappa, kappa_e, log_tau_e, tau_vb)
      term_s0 <- e_sig2_inv_hs_(xi_inv_vb, nu_s0_vb, log_xi_inv_e, log_sig02_inv_e, rho_s0_vb, sig02_inv_vb)
      term_xi <- e_sig2_inv_(1 / 2, nu_xi_inv_vb, log_xi_inv_e, A2_inv, rho_xi_inv_vb, xi_inv_vb)
      term_sig <- e_sig2_inv_(nu, nu_e, log_sig2_inv_e, rho, rho_e, sig2_inv_vb)

      lb_new <- as.numeric(term_y + term_bg + term_theta + term_zeta + term_tau + term_s0 + term_xi + term_sig)
      if (!is.null(lb_hook)) lb_hook(it, lb_new)
      if (it == first_plain_it || it %% max(5, batch_conv) == 0) say("ELBO = ", format(lb_new), "\n")

      if (debug && lb_new + elbo_slack < lb_old)
        stop("ELBO not increasing monotonically. Exit. ")

      # converged when the gain is below tol; otherwise move along the thinning schedule (never backwards)
      n_above <- sum(abs(lb_new - lb_old) > sched$times * tol)
      if (n_above == 0) {
        converged <- TRUE
      } else if (n_above < sched_pos) {
        sched_pos <- n_above
        batch_conv <- sched$batch[sched_pos]
      }

    }

    if (!is.null(checkpoint_path) && !annealed_iteration && it %% 100 == 0) {   # non-annealed iterations only
      st <- aq_fetch_state_(ctx, p, q)
      checkpoint_(it, checkpoint_path, st$beta_vb, st$gam_vb, theta_vb, zeta_vb, converged, lb_new, lb_old,
                  lam2_inv_vb = lam2_inv_vb, sig02_inv_vb = sig02_inv_vb, names_x = colnames(X), names_y = colnames(Y))
      rm(st)
    }

    c <- c_s <- c_next

  }

  checkpoint_clean_up_(checkpoint_path)

  if (converged) {
    say("Converged after ", format(it), " iterations, ELBO = ", format(lb_new), "\n")
  } else if (verbose != 0) {
    warning("Maximal number of iterations reached before convergence. Exit.")
  }

  st <- aq_fetch_state_(ctx, p, q)
  .Call(`_atlasqtl_aq_destroy`, ctx)
  gam_vb <- st$gam_vb
  beta_vb <- st$beta_vb
  rm(st)

  if (full_output) {   # the sample-space state has no Gram objects to return
    return(list(beta_vb = beta_vb, eta_vb = eta_vb, gam_vb = gam_vb, kappa_vb = kappa_vb, lam2_inv_vb = lam2_inv_vb,
                nu_s0_vb = nu_s0_vb, nu_vb = nu_vb, nu_xi_inv_vb = nu_xi_inv_vb, rho_s0_vb = rho_s0_vb, rho_vb = rho_vb,
                rho_xi_inv_vb = rho_xi_inv_vb, shr_fac_inv = shr_fac_inv, sig02_inv_vb = sig02_inv_vb,
                sig2_inv_vb = sig2_inv_vb, sig2_theta_vb = sig2_theta_vb, sig2_zeta_vb = sig2_zeta_vb, tau_vb = tau_vb,
                theta_vb = theta_vb, xi_inv_vb = xi_inv_vb, zeta_vb = zeta_vb))
  }

  dimnames(gam_vb) <- dimnames(beta_vb) <- list(colnames(X), colnames(Y))
  names(theta_vb) <- colnames(X)
  names(zeta_vb) <- colnames(Y)

  list(beta_vb = beta_vb, gam_vb = gam_vb, theta_vb = theta_vb, zeta_vb = zeta_vb, n = n, p = p, q = q,
       anneal = anneal, converged = converged, it = it, maxit = maxit, tol = tol, lb_opt = lb_new,
       diff_lb = abs(lb_new - lb_old))

}


# The stateless entries keep the generated closures of the reference (R/RcppExports.R:4-10) as they are; only the
# symbols they name now resolve to the CUDA-backed shim (bindings/R/atlasqtl_b200_shim.c, zero R changes):
#   coreDualLoop(...)     -> .Call(`_atlasqtl_coreDualLoop`, ... 15 arguments ...)
#   coreDualMisLoop(...)  -> .Call(`_atlasqtl_coreDualMisLoop`, ... 16 arguments ...)
#
# Pre-processing on the device (prepare_data_, R/prepare_atlasqtl.R:57-83; optional, see INTEGRATION.md section 2e):
#   :57-72   scale(X), rm_constant_, rm_collinear_  ->  pr <- .Call(`_atlasqtl_aq_prep_x`, X, 0L)   # or _atlasqtl_aq_prep_geno
#                                                       bool_cst_x <- pr$status == 1L; bool_rmvd_x <- pr$status != 0L
#   :83      scale(Y, center = TRUE, scale = FALSE) ->  ctx <- .Call(`_atlasqtl_aq_create_prepared`, pr$prep, Y)  (replaces aq_create)
