// K4 / K5 — length-m vector kernels (single CTA of 1024 threads; m <= 8192 in the BASELINE configs, any m
// works).  Everything here is launch-latency territory; what matters is that each solver iteration needs
// only two or three of these launches and that all reductions are fixed-order.
//
//   smoother_eval  : hμ.grad / hμ.hess            phuber-smooth.jl:27-36,83-114,150-164; exponential-smooth.jl:36-50;
//                                                  log-exp-smooth.jl:47-63; ostrovskii-bach-smooth.jl:31-36,72-85
//   k_pre          : gr, Hr, ∇q = g + λ·gr, η²     prox-N-SCORE.jl:40-42,69,95-98; prox-GGN-SCORE.jl:41-43,92
//   k_tail         : damping, x+dx, scaled prox, ‖x⁺−x‖, get_reg(x⁺)
//                                                  prox-N-SCORE.jl:92-112; prox-operators.jl:8-66;
//                                                  prox-reg-utils.jl:84-119; regularizers.jl:4-39
//   k_lqn_head     : smoother, ∇q, two_loop_recursion, step size (fixed / BB), damping, prox, norms, get_reg
//                                                  prox-L-BFGS-SCORE.jl:47-68,76-146 — ONE persistent single-CTA kernel
//   k_lqn_update   : ∇q_new, γ, curvature guard, memory push, H0     prox-L-BFGS-SCORE.jl:148-162 (the other one)
//   k_ls_terms / k_ls_sum / k_ls_pick : Armijo line search with the trials on the device      utils.jl:27-35
//
// Compiled with -fmad=false: elementwise formulas round exactly like the oracle's (no silent contraction).
#pragma once
#include "common.cuh"
#include "kernels_stream.cuh"  // LossParams (line-search trials)

namespace scs {

constexpr double kJuliaEps = 2.220446049250313e-16;  // eps()

struct RegDesc {
  int kind;  // scs_reg_kind
  double lam1, lam2;
  const double* lb;  // C_set bounds (indbox), broadcast if nlb == 1
  const double* ub;
  int nlb, nub;
  const int64_t* ind;   // 3 x ngroups, column-major (1-based start, end, weight)
  int ngroups;
  const int64_t* perm;  // P.G (1-based) or nullptr
};

struct SmoothDesc {
  int kind;  // scs_smoother_kind
  double mu;
  const double* lb;  // smoother bounds after bounds_sanity_check (±Inf -> ±1e32), broadcast if n == 1
  const double* ub;
  int nlb, nub;
  const double* cdiag;  // gl: per-variable group weight (Cmat == diag(cdiag)), else nullptr
};

// ---- scalar formulas --------------------------------------------------------------------------
SCS_DEVINL double ph_val(double x, double mu) {  // pseudo_huber, phuber-smooth.jl:28-30
  const double q = mu * mu + x * x;
  return ((mu * mu - mu * sqrt(q)) + x * x) * (1.0 / sqrt(q));
}
SCS_DEVINL double ph_grad(double x, double mu) {  // huber_grad :31-33   x*(μ²+x²)^(-1/2)
  return x * (1.0 / sqrt(mu * mu + x * x));
}
SCS_DEVINL double ph_hess(double x, double mu) {  // huber_hess :34-36   μ²*(μ²+x²)^(-3/2)
  const double q = mu * mu + x * x;
  return (mu * mu) * (1.0 / (q * sqrt(q)));
}
SCS_DEVINL double osba_val(double x, double mu) {  // ostrovskii-bach-smooth.jl:28-30
  const double sq = sqrt(mu * mu + 4.0 * (x * x));
  return (((sq / 2.0 - mu / 2.0) + mu * log(((2.0 * x - sq) + mu) / x) / 2.0) - 0.6931471805599453 * mu) +
         mu * log(((sq - mu) + 2.0 * x) / x) / 2.0;
}
SCS_DEVINL double osba_grad(double x, double mu) {  // :31-33
  const double x2 = x * x, mu2 = mu * mu;
  const double sq = sqrt(mu2 + 4.0 * x2);
  const double num = (((-(mu2 * mu) + mu2 * sq) - 4.0 * x2 * mu) + 2.0 * x2 * sq) * ((mu * sq + mu2) + 4.0 * x2);
  return num / (4.0 * mu2 * (x2 * x) + 16.0 * (x2 * x2 * x));
}
SCS_DEVINL double osba_hess(double x, double mu) {  // :34-36
  const double q = mu * mu + 4.0 * (x * x);
  const double sq = sqrt(q);
  return (sq - mu) * mu / (x * x) * (1.0 / sq) / 2.0;
}

SCS_DEVINL double bnd(const double* b, int nb, int i) { return nb == 1 ? b[0] : b[i]; }

// Elementwise smoother gradient / Hessian diagonal for the kinds that need no global scalar.
SCS_DEVINL void smooth_elem(const SmoothDesc& sd, int i, double x, double& g, double& h) {
  const double mu = sd.mu;
  switch (sd.kind) {
    case 0:  // PHuber L1L2
      g = ph_grad(x, mu);
      h = ph_hess(x, mu);
      break;
    case 1: {  // PHuber IndBox, including the `-x < a` test of huber_grad_indbox
      const double a = bnd(sd.lb, sd.nlb, i), b = bnd(sd.ub, sd.nub, i);
      if (-x < a) {
        const double q = ((a * a - 2.0 * x * a) + mu * mu) + x * x;
        g = (1.0 / sqrt(q)) * (-x + a);
      } else if (x == a || x < b) {
        g = kJuliaEps;
      } else {
        const double q = ((b * b - 2.0 * b * x) + mu * mu) + x * x;
        g = (1.0 / sqrt(q)) * (b - x);
      }
      if (x <= a) {
        const double q = ((a * a - 2.0 * a * x) + mu * mu) + x * x;
        h = (mu * mu) * (1.0 / (q * sqrt(q)));
      } else if (a < x && x < b) {
        h = kJuliaEps;
      } else if (x >= b) {
        const double q = ((b * b - 2.0 * b * x) + mu * mu) + x * x;
        h = (mu * mu) * (1.0 / (q * sqrt(q)));
      } else {
        h = x + a + b;  // NaN input
      }
      break;
    }
    case 3: {  // Exponential IndBox (uses lb only)
      const double a = bnd(sd.lb, sd.nlb, i);
      const double e = exp((-x + a) / mu);
      g = -e;
      h = 1.0 / mu * e;
      break;
    }
    case 4: {  // LogExp IndBox
      const double a = bnd(sd.lb, sd.nlb, i), b = bnd(sd.ub, sd.nub, i);
      const double g1 = x <= a + mu ? ((x - a) - 2.0 * mu) / mu : (x >= b - mu ? ((x - b) + 2.0 * mu) / mu : 0.0);
      const double g2 = x < a ? mu / (a - x) : (x > b ? -mu / (b - x) : 0.0);
      g = g1 + g2;
      const double h1 = x <= a + mu ? 1.0 / mu : (x >= b - mu ? 1.0 / mu : 0.0);
      const double h2 = x < a ? mu / ((a - x) * (a - x)) : (x > b ? mu / ((b - x) * (b - x)) : 0.0);
      h = h1 + h2;
      break;
    }
    case 5:  // OsBa L1L2
      g = osba_grad(x, mu);
      h = osba_hess(x, mu);
      break;
    default:
      g = 0.0;
      h = 0.0;
  }
}

// Full smoother evaluation by the whole CTA (handles the gl kinds that need Σ Dg²).  `red` >= 32 doubles.
SCS_DEVINL void smoother_eval_cta(const SmoothDesc& sd, const double* __restrict__ x, int m,
                                  double* __restrict__ gr, double* __restrict__ hr, double* red) {
  if (sd.kind == 2 || sd.kind == 6) {
    const bool ph = sd.kind == 2;
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += kVecThreads) {
      const double dg = ph ? ph_grad(x[i], sd.mu) : osba_grad(x[i], sd.mu);
      acc += dg * dg;
    }
    const double dd = block_sum<kVecThreads>(acc, red);  // dot(Dg,Dg)
    for (int i = threadIdx.x; i < m; i += kVecThreads) {
      const double xi = x[i];
      if (ph) {
        const double c = sd.cdiag[i] * ph_val(xi, sd.mu);
        const double gc = ph_grad(c, sd.mu);
        gr[i] = gc * ph_grad(xi, sd.mu);
        hr[i] = ph_hess(c, sd.mu) * dd + gc * ph_hess(xi, sd.mu);
      } else {
        const double c = sd.cdiag[i] * osba_val(xi, sd.mu);
        const double gc = osba_grad(c, sd.mu);
        gr[i] = gc * osba_grad(xi, sd.mu);
        hr[i] = osba_hess(c, sd.mu) * dd + gc * osba_hess(xi, sd.mu);
      }
    }
  } else {
    for (int i = threadIdx.x; i < m; i += kVecThreads) {
      double g, h;
      smooth_elem(sd, i, x[i], g, h);
      gr[i] = g;
      hr[i] = h;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kVecThreads) k_smoother(SmoothDesc sd, const double* __restrict__ x, int m,
                                                          double* __restrict__ gr, double* __restrict__ hr) {
  __shared__ double red[32];
  smoother_eval_cta(sd, x, m, gr, hr, red);
}

// Scalars block (device doubles) shared by the vector kernels and copied to the host once per step.
enum {
  SC_ETASQ = 0,   // λgr'·(H⁻¹ λgr)
  SC_PRI2 = 1,    // ‖δ‖²  (δ = x⁺−x with prox, dx without)
  SC_REGNEW = 2,  // get_reg(x⁺)
  SC_NX2 = 3,     // ‖x‖²
  SC_DIFF2 = 4,   // ‖x⁺−x‖²
  SC_ERR2 = 5,    // ‖x⁺−x*‖²   (Σ (x*−x⁺)² ; the gl rel-error divides by m on the host)
  SC_LOSS = 6,    // raw loss sum S (before scale)
  SC_DG = 7,      // δ'γ   (L-BFGS)
  SC_H0 = 8,      // L-BFGS H0
  SC_PUSHED = 9,  // 1 if the memory was updated
  SC_REGX = 10,   // get_reg(x) from k_reg
  SC_GD = 11,     // ∇q'd  (line search)
  SC_GG = 12,     // γ'γ (BB) / scratch
  SC_DGBB = 13,   // δ'γ (BB)
  SC_SS = 14,     // step size the tail used (decided on the device for BB / line search)
  SC_LSFOUND = 15,  // line search: 1 once a trial step size has passed the Armijo test
  SC_COUNT = 16,    // copied to the host once per step
  SC_SCRATCH = 16,  // device-only scratch behind it: a second k_pre's η² / ‖x‖² (compute_gq), the row count all-reduce
  SC_ALLOC = 24
};

// gr, Hr at x; rhs = g + λ·gr; η² = Σ λgr_i·((1/Hr_i)·λgr_i).  g may be null (then rhs = λ·gr).
SCS_DEVINL void pre_cta(const SmoothDesc& sd, double lam, const double* __restrict__ x, const double* __restrict__ g,
                        int m, double* __restrict__ gr, double* __restrict__ hr, double* __restrict__ rhs,
                        double* __restrict__ scal, double* red) {
  smoother_eval_cta(sd, x, m, gr, hr, red);
  double acc = 0.0, nx = 0.0;
  for (int i = threadIdx.x; i < m; i += kVecThreads) {
    const double lgr = lam * gr[i];
    acc += lgr * ((1.0 / hr[i]) * lgr);
    nx += x[i] * x[i];
    if (rhs) rhs[i] = (g ? g[i] : 0.0) + lgr;
  }
  acc = block_sum<kVecThreads>(acc, red);
  nx = block_sum<kVecThreads>(nx, red);
  if (threadIdx.x == 0) {
    scal[SC_ETASQ] = acc;
    scal[SC_NX2] = nx;
  }
}
__global__ void __launch_bounds__(kVecThreads)
k_pre(SmoothDesc sd, double lam, const double* __restrict__ x, const double* __restrict__ g, int m,
      double* __restrict__ gr, double* __restrict__ hr, double* __restrict__ rhs, double* __restrict__ scal) {
  __shared__ double red[32];
  pre_cta(sd, lam, x, g, m, gr, hr, rhs, scal, red);
}

// get_reg(model, x, reg_name) by the whole CTA (regularizers.jl:4-39, prox-reg-utils.jl:101-119).
SCS_DEVINL double reg_value_cta(const RegDesc& rd, const double* __restrict__ x, int m, double* red) {
  if (rd.kind == 0 || rd.kind == 1) {
    double a = 0.0;
    for (int i = threadIdx.x; i < m; i += kVecThreads) a += rd.kind == 0 ? fabs(x[i]) : x[i] * x[i];
    return rd.lam1 * block_sum<kVecThreads>(a, red);
  } else if (rd.kind == 2) {
    double bad = 0.0;
    for (int i = threadIdx.x; i < m; i += kVecThreads)
      if (x[i] < bnd(rd.lb, rd.nlb, i) || x[i] > bnd(rd.ub, rd.nub, i)) bad = 1.0;
    bad = block_sum<kVecThreads>(bad, red);
    return bad > 0.0 ? __longlong_as_double(0x7ff0000000000000LL) : 0.0;
  } else {  // gl: λ2*Σ_g w_g*‖(Px)_g‖ + λ1*Σ|x|
    double a = 0.0;
    for (int i = threadIdx.x; i < m; i += kVecThreads) a += fabs(x[i]);
    const double l1 = block_sum<kVecThreads>(a, red);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double f = 0.0;  // per-warp partial of Σ_g w_g‖·‖ (identical on all lanes)
    for (int g = warp; g < rd.ngroups; g += kVecThreads / 32) {
      const int64_t gs = rd.ind[3 * g] - 1, ge = rd.ind[3 * g + 1];
      double s = 0.0;
      for (int64_t k = gs + lane; k < ge; k += 32) {
        const double v = rd.perm ? x[rd.perm[k] - 1] : x[k];
        s += v * v;
      }
      s = warp_sum(s);
      f += (double)rd.ind[3 * g + 2] * sqrt(s);
    }
    f = block_sum<kVecThreads>(lane == 0 ? f : 0.0, red);
    return rd.lam2 * f + rd.lam1 * l1;
  }
}

__global__ void __launch_bounds__(kVecThreads) k_reg(RegDesc rd, const double* __restrict__ x, int m,
                                                     double* __restrict__ out) {
  __shared__ double red[32];
  const double v = reg_value_cta(rd, x, m, red);
  if (threadIdx.x == 0) out[0] = v;
}

// Scaled prox applied in place to u (length m) by the whole CTA: prox-operators.jl:8-66.  hr = Hr diagonal
// (h_scale = 1/Hr is formed exactly as the reference does: t = α*λ ./ (1 ./ Hr)).
SCS_DEVINL void prox_cta(const RegDesc& rd, double ss, const double* __restrict__ hr, double* __restrict__ u,
                         int m) {
  if (rd.kind == 0) {
    for (int i = threadIdx.x; i < m; i += kVecThreads) {
      const double t = ss * rd.lam1 / (1.0 / hr[i]);
      const double v = u[i];
      u[i] = julia_sign(v) * fmax_nan(fabs(v) - t, 0.0);
    }
  } else if (rd.kind == 1) {
    for (int i = threadIdx.x; i < m; i += kVecThreads) {
      const double t = ss * rd.lam1 / (1.0 / hr[i]);
      const double v = u[i];
      u[i] = v * fmax_nan(1.0 - t / (fabs(v) * fabs(v)), 0.0);
    }
  } else if (rd.kind == 2) {
    for (int i = threadIdx.x; i < m; i += kVecThreads)
      u[i] = fmin_nan(fmax_nan(u[i], bnd(rd.lb, rd.nlb, i)), bnd(rd.ub, rd.nub, i));
  } else {  // gl: ProxL1 with t = λ1/h, then ProxL2(·, α*λ2, h) group by group (one warp per group)
    for (int i = threadIdx.x; i < m; i += kVecThreads) {
      const double t = rd.lam1 / (1.0 / hr[i]);
      const double v = u[i];
      u[i] = julia_sign(v) * fmax_nan(fabs(v) - t, 0.0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double lam = ss * rd.lam2;
    for (int g = warp; g < rd.ngroups; g += kVecThreads / 32) {
      const int64_t gs = rd.ind[3 * g] - 1, ge = rd.ind[3 * g + 1];
      double s = 0.0;
      for (int64_t k = gs + lane; k < ge; k += 32) s += u[k] * u[k];
      const double nrm = sqrt(warp_sum(s));
      const double bg = lam * (double)rd.ind[3 * g + 2];
      for (int64_t k = gs + lane; k < ge; k += 32) u[k] = u[k] * fmax_nan(1.0 - bg / ((1.0 / hr[k]) * nrm), 0.0);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kVecThreads) k_prox(RegDesc rd, double ss, const double* __restrict__ hr,
                                                      double* __restrict__ u, int m) {
  prox_cta(rd, ss, hr, u, m);
}

// Tail of every step!: α = ss/(1+Mg·η); dx = min(1,α)·(dsign·dvec); x⁺ = prox(x+dx) or x+dx.
// Also emits ‖δ‖², ‖x⁺−x‖², ‖x⁺−x*‖², get_reg(x⁺).  delta_out (optional) receives δ (L-BFGS s-vector).
// eta_sq: λgr'·H⁻¹·λgr as left by pre_cta (read by the caller after a barrier).
SCS_DEVINL void tail_cta(const RegDesc& rd, int use_prox, double ss, double Mg, double dsign, double eta_sq,
                         const double* __restrict__ x, const double* __restrict__ dvec, const double* __restrict__ hr,
                         const double* __restrict__ xstar, int m, double* __restrict__ xnew,
                         double* __restrict__ dx_out, double* __restrict__ delta_out, double* __restrict__ scal,
                         double* red) {
  const double eta = sqrt(eta_sq);
  const double alpha = ss / (1.0 + Mg * eta);
  const double safe = alpha < 1.0 ? alpha : (alpha != alpha ? alpha : 1.0);  // min(1, α), NaN-propagating
  for (int i = threadIdx.x; i < m; i += kVecThreads) {
    const double dx = safe * (dsign * dvec[i]);
    if (dx_out) dx_out[i] = dx;
    xnew[i] = x[i] + dx;
  }
  __syncthreads();
  if (use_prox) prox_cta(rd, ss, hr, xnew, m);
  double pri = 0.0, diff = 0.0, err = 0.0;
  for (int i = threadIdx.x; i < m; i += kVecThreads) {
    const double d = xnew[i] - x[i];
    diff += d * d;
    const double dl = use_prox ? d : safe * (dsign * dvec[i]);
    pri += dl * dl;
    if (delta_out) delta_out[i] = dl;
    if (xstar) {
      const double e = xstar[i] - xnew[i];
      err += e * e;
    }
  }
  pri = block_sum<kVecThreads>(pri, red);
  diff = block_sum<kVecThreads>(diff, red);
  err = block_sum<kVecThreads>(err, red);
  const double rv = reg_value_cta(rd, xnew, m, red);
  if (threadIdx.x == 0) {
    scal[SC_PRI2] = pri;
    scal[SC_DIFF2] = diff;
    scal[SC_ERR2] = err;
    scal[SC_REGNEW] = rv;
    scal[SC_SS] = ss;
  }
}
// ss_dev (optional): the step size was decided on the device (line search: scal[SC_SS]) — it overrides `ss`.
__global__ void __launch_bounds__(kVecThreads)
k_tail(RegDesc rd, int use_prox, double ss, double Mg, double dsign, const double* __restrict__ x,
       const double* __restrict__ dvec, const double* __restrict__ hr, const double* __restrict__ xstar, int m,
       double* __restrict__ xnew, double* __restrict__ dx_out, double* __restrict__ delta_out,
       double* __restrict__ scal, const double* __restrict__ ss_dev) {
  __shared__ double red[32];
  if (ss_dev) ss = ss_dev[0];
  tail_cta(rd, use_prox, ss, Mg, dsign, scal[SC_ETASQ], x, dvec, hr, xstar, m, xnew, dx_out, delta_out, scal, red);
}

// ---- L-BFGS (prox-L-BFGS-SCORE.jl) ---------------------------------------------------------------
// Memory: S, Y are cap x m row-major ring buffers; state[0] = count, state[1] = head (index of oldest).
struct LbfgsMem {
  double* S;
  double* Y;
  int64_t* state;  // {count, head}
  int cap;
};

SCS_DEVINL double dot_cta(const double* __restrict__ a, const double* __restrict__ b, int m, double* red) {
  double s = 0.0;
  for (int i = threadIdx.x; i < m; i += kVecThreads) s += a[i] * b[i];
  return block_sum<kVecThreads>(s, red);
}

// d = -∇q on the first iteration / empty memory, else two_loop_recursion(∇q)   (:47-68, :102-106)
// alpha / rho: >= 64 doubles of shared memory each.  All dots are warp-shuffle + one shared-memory stage (fixed tree).
SCS_DEVINL void lbfgs_dir_cta(const LbfgsMem& mem, int64_t iter, const double* __restrict__ gq, int m,
                              double* __restrict__ q, double* __restrict__ d, const double* __restrict__ scal,
                              double* red, double* alpha, double* rho) {
  const int cnt = (int)mem.state[0], head = (int)mem.state[1];
  if (iter == 1 || cnt == 0) {
    for (int i = threadIdx.x; i < m; i += kVecThreads) d[i] = -gq[i];
    __syncthreads();
    return;
  }
  for (int i = threadIdx.x; i < m; i += kVecThreads) q[i] = gq[i];
  __syncthreads();
  for (int t = 0; t < cnt; ++t) {  // newest -> oldest
    const int slot = (head + cnt - 1 - t) % mem.cap;
    const double* s = mem.S + (int64_t)slot * m;
    const double* y = mem.Y + (int64_t)slot * m;
    const double rho_i = 1.0 / dot_cta(y, s, m, red);
    const double alpha_i = rho_i * dot_cta(s, q, m, red);
    for (int i = threadIdx.x; i < m; i += kVecThreads) q[i] = q[i] - alpha_i * y[i];
    if (threadIdx.x == 0) {
      alpha[t] = alpha_i;
      rho[t] = rho_i;
    }
    __syncthreads();
  }
  const double H0 = scal[SC_H0];
  for (int i = threadIdx.x; i < m; i += kVecThreads) q[i] = H0 * q[i];
  __syncthreads();
  for (int t = 0; t < cnt; ++t) {  // oldest -> newest
    const int slot = (head + t) % mem.cap;
    const double* s = mem.S + (int64_t)slot * m;
    const double* y = mem.Y + (int64_t)slot * m;
    const double rho_i = rho[cnt - 1 - t], alpha_i = alpha[cnt - 1 - t];
    const double beta = rho_i * dot_cta(y, q, m, red);
    const double c = alpha_i - beta;
    for (int i = threadIdx.x; i < m; i += kVecThreads) q[i] = q[i] + s[i] * c;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < m; i += kVecThreads) d[i] = -q[i];
  __syncthreads();
}
// BB: δ = x − x_prev, γ = gq − gq_prev;  scal[SC_GG] = γ'γ, scal[SC_DGBB] = δ'γ   (utils.jl:43-48).  Returns
// (γ'γ)/(δ'γ) — used *as* the step size — to every thread.
SCS_DEVINL double bb_cta(const double* __restrict__ x, const double* __restrict__ xp, const double* __restrict__ gq,
                         const double* __restrict__ gqp, int m, double* __restrict__ scal, double* red) {
  double gg = 0.0, dg = 0.0;
  for (int i = threadIdx.x; i < m; i += kVecThreads) {
    const double de = x[i] - xp[i], ga = gq[i] - gqp[i];
    gg += ga * ga;
    dg += de * ga;
  }
  gg = block_sum<kVecThreads>(gg, red);
  dg = block_sum<kVecThreads>(dg, red);
  if (threadIdx.x == 0) {
    scal[SC_GG] = gg;
    scal[SC_DGBB] = dg;
  }
  return gg / dg;
}

// ---- the persistent ProxLQNSCORE kernels: one launch before the gradient pass, one after ------------------------------
// k_lqn_head = everything step! does at x before it needs ∇f(x⁺) (prox-L-BFGS-SCORE.jl:76-146): smoother gradient / Hessian
// diagonal, ∇q (or the one carried over from the previous step's k_lqn_update), two-loop recursion over the device-
// resident memory, the step size (fixed, or the BB estimate — decided here, no host round trip), damping, x + dx, scaled
// prox, ‖δ‖, get_reg(x⁺).  mode: 0 = ss given, 1 = BB from (x, x_prev, ∇q, ∇q_prev), 2 = direction only (a line search
// follows; the tail is k_tail with the step size it leaves in scal[SC_SS]).
__global__ void __launch_bounds__(kVecThreads)
k_lqn_head(SmoothDesc sd, RegDesc rd, LbfgsMem mem, double lam, double Mg, double ss, int mode, int use_prox,
           int64_t iter, const double* __restrict__ x, const double* __restrict__ xprev,
           const double* __restrict__ g /* A'r at x, or null: gq already holds ∇q(x) */, int m,
           double* __restrict__ gr, double* __restrict__ hr, double* __restrict__ gq,
           const double* __restrict__ gqprev, double* __restrict__ q, double* __restrict__ d,
           const double* __restrict__ xstar, double* __restrict__ xnew, double* __restrict__ dx_out,
           double* __restrict__ delta_out, double* __restrict__ scal) {
  __shared__ double red[32];
  __shared__ double alpha[64], rho[64];
  pre_cta(sd, lam, x, g, m, gr, hr, g ? gq : (double*)nullptr, scal, red);
  __syncthreads();
  lbfgs_dir_cta(mem, iter, gq, m, q, d, scal, red, alpha, rho);
  if (mode == 2) return;
  if (mode == 1) ss = bb_cta(x, xprev, gq, gqprev, m, scal, red);
  tail_cta(rd, use_prox, ss, Mg, 1.0, scal[SC_ETASQ], x, d, hr, xstar, m, xnew, dx_out, delta_out, scal, red);
}

// ∇q_new = g_new + λ·hμ.grad(x⁺);  γ = ∇q_new − ∇q;  push (δ, γ) if δ'γ > 1e-10;  H0 = γ'δ/γ'γ   (:148-162)
// gq is overwritten with ∇q_new (it is next iteration's ∇q: same x, same bits); gqprev (optional) receives the old ∇q.
SCS_DEVINL void lbfgs_update_cta(const SmoothDesc& sd, double lam, const LbfgsMem& mem, const double* __restrict__ xnew,
                                 const double* __restrict__ gnew, const double* __restrict__ delta, int m,
                                 double* __restrict__ gq, double* __restrict__ gqprev, double* __restrict__ gamma,
                                 double* __restrict__ gr_tmp, double* __restrict__ hr_tmp, double* __restrict__ scal,
                                 double* red) {
  smoother_eval_cta(sd, xnew, m, gr_tmp, hr_tmp, red);
  double dg = 0.0, gg = 0.0;
  for (int i = threadIdx.x; i < m; i += kVecThreads) {
    const double qn = gnew[i] + lam * gr_tmp[i];
    const double qo = gq[i];
    const double gm = qn - qo;
    gamma[i] = gm;
    gq[i] = qn;
    if (gqprev) gqprev[i] = qo;
    dg += delta[i] * gm;
    gg += gm * gm;
  }
  dg = block_sum<kVecThreads>(dg, red);
  gg = block_sum<kVecThreads>(gg, red);
  __syncthreads();
  const bool push = dg > 1e-10;
  if (push) {
    int cnt = (int)mem.state[0], head = (int)mem.state[1];
    int slot;
    if (cnt == mem.cap) {  // popfirst!
      slot = head;
      head = (head + 1) % mem.cap;
    } else {
      slot = (head + cnt) % mem.cap;
      cnt += 1;
    }
    double* s = mem.S + (int64_t)slot * m;
    double* y = mem.Y + (int64_t)slot * m;
    for (int i = threadIdx.x; i < m; i += kVecThreads) {
      s[i] = delta[i];
      y[i] = gamma[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mem.state[0] = cnt;
      mem.state[1] = head;
      scal[SC_H0] = dg / gg;  // dot(γ,δ)/dot(γ,γ)
    }
  }
  if (threadIdx.x == 0) {
    scal[SC_DG] = dg;
    scal[SC_PUSHED] = push ? 1.0 : 0.0;
  }
}
// k_lqn_update = what step! does after the gradient pass at x⁺ (:148-162).  On one GPU it also folds the per-cluster
// partial gradients and per-CTA loss sums of k_fused_grad itself (gpart != null; the same fixed order as k_colsum /
// k_sum_partials, so both routes give the same bits): an iteration is then k_lqn_head, k_fused_grad, k_lqn_update.
__global__ void __launch_bounds__(kVecThreads)
k_lqn_update(SmoothDesc sd, double lam, LbfgsMem mem, const double* __restrict__ xnew, double* __restrict__ gl /* [g ‖ loss] */,
             const double* __restrict__ gpart, int64_t nparts, const double* __restrict__ losspart, int64_t nloss,
             const double* __restrict__ delta, int m, double* __restrict__ gq, double* __restrict__ gqprev,
             double* __restrict__ gamma, double* __restrict__ gr_tmp, double* __restrict__ hr_tmp,
             double* __restrict__ scal) {
  __shared__ double red[32];
  if (gpart) {
    for (int j = threadIdx.x; j < m; j += kVecThreads) {  // k_colsum's order: 8 interleaved slices, then their sum
      double sl[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int64_t b = 0; b < nparts; b += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (b + q < nparts) sl[q] += gpart[(b + q) * m + j];
      }
      double t = sl[0];
#pragma unroll
      for (int q = 1; q < 8; ++q) t += sl[q];
      gl[j] = t;
    }
    double sacc = 0.0;
    for (int64_t i = threadIdx.x; i < nloss; i += kVecThreads) sacc += losspart[i];
    sacc = block_sum<kVecThreads>(sacc, red);
    if (threadIdx.x == 0) gl[m] = sacc;
    __syncthreads();
  }
  lbfgs_update_cta(sd, lam, mem, xnew, gl, delta, m, gq, gqprev, gamma, gr_tmp, hr_tmp, scal, red);
}

// ---- Armijo line search on the device (utils.jl:27-35) -----------------------------------------------------------
// The reference evaluates f(x + α d) for α = 1, 1/2, 1/4, ... — one pass over A per trial.  Here ONE pass gives
// zd = A d; with z = A x (left by the pass at x) every trial is z + α zd: a batch of kLsTrials step sizes costs one sweep
// over two row vectors.  k_ls_terms: part[b*8 + k] = Σ_{rows of block b} term(z_i + α_k (dsign zd_i)), α_k = alpha0 2^-k.
constexpr int kLsTrials = 8;
constexpr int kLsRowsPerBlock = 4096;
enum { LS_FX = 0, LS_GD = 1, LS_COUNT = 2 };
SCS_DEVINL double loss_term_only(const LossParams& lp, double z, double y) {
  if (lp.kind == 0) return log(1.0 + exp(-y * z));
  const double d = z - y;
  return d * d;
}
__global__ void __launch_bounds__(256)
k_ls_terms(const double* __restrict__ z, const double* __restrict__ zd, const double* __restrict__ y, int64_t nproc,
           int64_t row_lo, int64_t row_hi, LossParams lp, double dsign, double alpha0, const double* __restrict__ scal,
           double* __restrict__ part) {
  __shared__ double red[32];
  if (scal[SC_LSFOUND] != 0.0) return;  // an earlier batch already accepted a step size
  double acc[kLsTrials];
#pragma unroll
  for (int k = 0; k < kLsTrials; ++k) acc[k] = 0.0;
  const int64_t b0 = (int64_t)blockIdx.x * kLsRowsPerBlock;
  for (int64_t i = b0 + threadIdx.x; i < b0 + kLsRowsPerBlock && i < nproc; i += 256) {
    if (i < row_lo || i >= row_hi) continue;  // padding, or rows of another mini-batch
    const double zi = z[i], di = dsign * zd[i], yi = y[i];
    double a = alpha0;
#pragma unroll
    for (int k = 0; k < kLsTrials; ++k) {
      acc[k] += loss_term_only(lp, zi + a * di, yi);
      a *= 0.5;
    }
  }
#pragma unroll
  for (int k = 0; k < kLsTrials; ++k) {
    const double t = block_sum<256>(acc[k], red);
    if (threadIdx.x == 0) part[(int64_t)blockIdx.x * kLsTrials + k] = t;
  }
}
// sums[k] = Σ_b part[b*8 + k] in fixed order (several ranks: this is what gets all-reduced)
__global__ void __launch_bounds__(kVecThreads)
k_ls_sum(const double* __restrict__ part, int64_t nblk, const double* __restrict__ scal, double* __restrict__ sums) {
  __shared__ double red[32];
  if (scal[SC_LSFOUND] != 0.0) return;
  for (int k = 0; k < kLsTrials; ++k) {
    double s = 0.0;
    for (int64_t b = threadIdx.x; b < nblk; b += kVecThreads) s += part[b * kLsTrials + k];
    s = block_sum<kVecThreads>(s, red);
    if (threadIdx.x == 0) sums[k] = s;
  }
}
// The Armijo test over the batch, in order: accept the first α_k with  !(f(x + α_k d) > f(x) + 1e-4 α_k <∇q, d>).
// first != 0: also evaluates f(x) = fscale(loss sum at x) + get_reg(x) and <∇q, d> once (the reference recomputes the same
// values every trial) into ls[].  loss_kind / loss_p: fval = p*S (logistic) or 0.5*S/p (least squares).
__global__ void __launch_bounds__(kVecThreads)
k_ls_pick(RegDesc rd, int loss_kind, double loss_p, const double* __restrict__ x, const double* __restrict__ dvec,
          double dsign, const double* __restrict__ gq, int m, double alpha0, int first,
          const double* __restrict__ loss_sum_x, const double* __restrict__ sums, double* __restrict__ trial,
          double* __restrict__ ls, double* __restrict__ scal) {
  __shared__ double red[32];
  if (scal[SC_LSFOUND] != 0.0) return;
  if (first) {
    const double S = loss_sum_x[0];
    const double fv = loss_kind == 0 ? loss_p * S : 0.5 * S / loss_p;
    const double rv = reg_value_cta(rd, x, m, red);
    const double gd = dsign * dot_cta(gq, dvec, m, red);
    if (threadIdx.x == 0) {
      ls[LS_FX] = fv + rv;
      ls[LS_GD] = gd;
      scal[SC_REGX] = rv;
      scal[SC_GD] = dsign * gd;
    }
    __syncthreads();
  }
  const double fx = ls[LS_FX], gd = ls[LS_GD];
  double a = alpha0;
  for (int k = 0; k < kLsTrials; ++k) {
    for (int i = threadIdx.x; i < m; i += kVecThreads) trial[i] = x[i] + (dsign * a) * dvec[i];
    __syncthreads();
    const double rv = reg_value_cta(rd, trial, m, red);
    const double S = sums[k];
    const double ft = (loss_kind == 0 ? loss_p * S : 0.5 * S / loss_p) + rv;
    if (!(ft > fx + 1e-4 * a * gd)) {  // NaN accepts, like the reference's `while f(..) > ..`
      if (threadIdx.x == 0) {
        scal[SC_SS] = a;
        scal[SC_LSFOUND] = 1.0;
      }
      return;
    }
    a *= 0.5;
    __syncthreads();
  }
}

// ---- small helpers ----------------------------------------------------------------------------
// out = a + alpha*b
__global__ void k_axpy_out(const double* __restrict__ a, double alpha, const double* __restrict__ b, int m,
                           double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = a[i] + alpha * b[i];
}
// scal[slot] = a'b
__global__ void __launch_bounds__(kVecThreads) k_dot(const double* __restrict__ a, const double* __restrict__ b,
                                                     int m, double* __restrict__ scal, int slot) {
  __shared__ double red[32];
  const double s = dot_cta(a, b, m, red);
  if (threadIdx.x == 0) scal[slot] = s;
}
__global__ void __launch_bounds__(kVecThreads)
k_bb(const double* __restrict__ x, const double* __restrict__ xp, const double* __restrict__ gq,
     const double* __restrict__ gqp, int m, double* __restrict__ scal) {
  __shared__ double red[32];
  bb_cta(x, xp, gq, gqp, m, scal, red);
}

}  // namespace scs
