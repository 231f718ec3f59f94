// Scalar math shared by the device kernels and the host test hooks (same source, compiled twice).
// Every routine cites the reference code it replaces (paths under the reference's src/).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MRGP_HD __host__ __device__ __forceinline__
#else
#define MRGP_HD inline
#endif

namespace mrgp {

constexpr double kPi = 3.14159265358979323846;
constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr double kEps = 1e-45;  // Priors.py:5

// psi(x) for x > 0 (scipy.special.psi at Stats.py:36, 107, 366, 388): upward recurrence to x >= 10,
// then the asymptotic series.  |rel err| ~ 1e-16 on the arguments the model produces (1e-45 ... ~N).
MRGP_HD double digamma(double x) {
    double acc = 0.0;
    while (x < 10.0) {
        acc -= 1.0 / x;
        x += 1.0;
    }
    const double inv = 1.0 / x;
    const double i2 = inv * inv;
    const double series =
        i2 * (1.0 / 12.0 -
              i2 * (1.0 / 120.0 -
                    i2 * (1.0 / 252.0 - i2 * (1.0 / 240.0 - i2 * (1.0 / 132.0 - i2 * (691.0 / 32760.0 - i2 * (1.0 / 12.0)))))));
    return acc + (log(x) - 0.5 * inv - series);
}

// Matern spectral density S(sqrt(lambda)), KernelClass.py:80-90 exponentiated as :57-58, called from
// MRGP.py:297-303 with s = sqrt(lambda).
MRGP_HD double matern_spectral(double lambda, double nu, double l, double sf) {
    const double s = sqrt(lambda);
    const double log_arg = log(2.0 * nu) - 2.0 * log(l);
    const double arg = exp(log_arg);
    const double log_const = 0.5 * kLog2Pi + nu * log_arg;
    const double log_gamma_term = lgamma(nu + 0.5) - lgamma(nu);
    const double log_power_term = -(nu + 0.5) * log(arg + s * s);
    return exp(log(sf) + log_const + log_gamma_term + log_power_term);
}

// np.spacing(x) for x >= 0 (SanityCheck.py:40).
MRGP_HD double spacing(double x) {
    union {
        double d;
        uint64_t u;
    } v;
    v.d = x;
    v.u += 1;
    return v.d - x;
}

// LAPACK potrf on a symmetric 2x2 [[a, b], [b, c]]: success iff both pivots are > 0
// (numpy.linalg.cholesky at SanityCheck.py:59-65).
MRGP_HD bool chol2_ok(double a, double b, double c) {
    if (!(a > 0.0)) return false;
    const double l21 = b / sqrt(a);
    const double piv = c - l21 * l21;
    return piv > 0.0;
}

// Symmetric 2x2 eigen-solve: l1 >= l2 and the projector P1 = v1 v1^T (p00, p01, p11).
// Replaces numpy.linalg.eig on a symmetric input (CommonDensities.py:73-76); eigenvector signs are
// irrelevant because only v v^T is used downstream (Stats.py:375-382).
MRGP_HD void eig2(double a, double b, double c, double &l1, double &l2, double &p00, double &p01, double &p11) {
    const double m = 0.5 * (a + c);
    const double d = 0.5 * (a - c);
    const double h = sqrt(d * d + b * b);
    l1 = m + h;
    l2 = m - h;
    if (h > 0.0) {
        const double ih = 0.5 / h;
        p00 = 0.5 + d * ih;
        p11 = 0.5 - d * ih;
        p01 = b * ih;
    } else {
        p00 = 1.0;
        p01 = 0.0;
        p11 = 0.0;
    }
}

// First-order Kume-Wood saddle-point approximation of log C(kappa) and its gradient for a real Bingham
// distribution of dimension P (computeRealBinghamConstant.py:42-147).  The root of
// 1/2 sum 1/(Lam_k - t) = 1 on [0.1 - P, 0.1 - 0.5] (:72-95, brentq in the reference) is found by Newton
// from the right end of the bracket: the function is increasing and convex left of min(Lam) = 0.1 and
// non-negative at 0.1 - 0.5, so the iterates decrease monotonically to the root.
template <int P>
MRGP_HD void saddle_point(const double *kappa, double &logc, double *rho) {
    double lam[P];
    double mn = -kappa[0];
    for (int k = 1; k < P; ++k) mn = fmin(mn, -kappa[k]);
    const double adjust = 0.1 - mn;
    for (int k = 0; k < P; ++k) lam[k] = -kappa[k] + adjust;
    double t;
    if (P == 2) {
        // closed form of 1/2 (1/(u + a0) + 1/(u + a1)) = 1 with u = 0.1 - t, a_k = lam_k - 0.1 >= 0:
        // u = 1/2 [1 - (a0 + a1) + sqrt(1 + (a0 - a1)^2)], written without cancellation for a large gap
        const double a0 = lam[0] - 0.1, a1 = lam[P - 1] - 0.1;
        const double amin = fmin(a0, a1), gap = fabs(a0 - a1);
        const double u = 0.5 * (1.0 - 2.0 * amin + 1.0 / (sqrt(1.0 + gap * gap) + gap));
        t = 0.1 - u;
    } else {
        t = 0.1 - 0.5;
        for (int it = 0; it < 60; ++it) {
            double f = -1.0, fp = 0.0;
            for (int k = 0; k < P; ++k) {
                const double r = 1.0 / (lam[k] - t);
                f += 0.5 * r;
                fp += 0.5 * r * r;
            }
            const double step = f / fp;
            const double tn = t - step;
            if (!(fabs(step) > 1e-17 * fabs(t)) || tn >= t) {
                t = (tn < t) ? tn : t;
                break;
            }
            t = tn;
        }
    }
    double k2 = 0.0, k3 = 0.0, sumlog = 0.0;
    double r1[P];
    for (int k = 0; k < P; ++k) {
        r1[k] = 1.0 / (lam[k] - t);
        k2 += r1[k] * r1[k];
        k3 += r1[k] * r1[k] * r1[k];
        sumlog += log(lam[k] - t);
    }
    k2 *= 0.5;
    logc = 0.5 * (log(2.0) + (P - 1) * log(kPi) - log(k2) - sumlog) - t + adjust;
    // gradient (:125-143)
    double dk1dt = 0.0;
    for (int k = 0; k < P; ++k) dk1dt += 0.5 * r1[k] * r1[k];
    double dsumlogdt = 0.0;
    for (int k = 0; k < P; ++k) dsumlogdt -= r1[k];
    for (int k = 0; k < P; ++k) {
        const double dk1dlam = -0.5 * r1[k] * r1[k];
        const double dtdlam = -dk1dlam / dk1dt;
        const double dk2dlam = -(r1[k] * r1[k] * r1[k]) + k3 * dtdlam;
        const double dlogk2 = dk2dlam / k2;
        const double dsl = r1[k] + dsumlogdt * dtdlam;
        rho[k] = 0.5 * dlogk2 + 0.5 * dsl + dtdlam;
    }
}

struct Bingham2 {
    double b[3];      // guarded B: b00, b01, b11
    double kappa[2];  // clamped at 0 (Posteriors.py:525-526)
    double rho[2];
    double logc;
    double cov[3];    // axis_cov: c00, c01, c11
    int n_chol;       // Cholesky factorisations attempted by the guard
};

// One Bingham axis update for dy == 2: PD guard (Posteriors.py:519-523 -> SanityCheck.py:16-65), eigen-
// solve sorted descending (CommonDensities.py:73-76), saddle-point constant from the unclamped
// eigenvalues (:77), axis covariance (Stats.py:375-382).
MRGP_HD void bingham2(double a, double b, double c, Bingham2 &out) {
    int n_chol = 1;
    if (!chol2_ok(a, b, c)) {
        // nearestPD: B = (A + A^T)/2 is A itself; H = V |Lambda| V^T; A2 = (B + H)/2 = V max(Lambda,0) V^T
        const double fro = sqrt(a * a + 2.0 * b * b + c * c);
        double l1, l2, p00, p01, p11;
        eig2(a, b, c, l1, l2, p00, p01, p11);
        const double m1 = l1 > 0.0 ? l1 : 0.0, m2 = l2 > 0.0 ? l2 : 0.0;
        a = m1 * p00 + m2 * (1.0 - p00);
        b = m1 * p01 - m2 * p01;
        c = m1 * p11 + m2 * (1.0 - p11);
        ++n_chol;
        if (!chol2_ok(a, b, c)) {
            const double sp = spacing(fro);
            for (int k = 1; k <= 64; ++k) {
                eig2(a, b, c, l1, l2, p00, p01, p11);
                const double shift = -l2 * (double)(k * k) + sp;
                a += shift;
                c += shift;
                ++n_chol;
                if (chol2_ok(a, b, c)) break;
            }
        }
    }
    double l1, l2, p00, p01, p11;
    eig2(a, b, c, l1, l2, p00, p01, p11);
    const double kap[2] = {l1, l2};
    saddle_point<2>(kap, out.logc, out.rho);
    out.b[0] = a;
    out.b[1] = b;
    out.b[2] = c;
    out.kappa[0] = l1 < 0.0 ? 0.0 : l1;
    out.kappa[1] = l2 < 0.0 ? 0.0 : l2;
    out.cov[0] = out.rho[0] * p00 + out.rho[1] * (1.0 - p00);
    out.cov[1] = out.rho[0] * p01 - out.rho[1] * p01;
    out.cov[2] = out.rho[0] * p11 + out.rho[1] * (1.0 - p11);
    out.n_chol = n_chol;
}

// ------------------------------------------------------------------------------------------------
// Permutation-alignment weights (Stats.py:390-445): omega = diag(alpha) exp(lw) diag(beta) with unit row and
// column sums.  The reference finds the 2M log-scalings with MINPACK hybrd (xtol 1.5e-8); this is the
// fixed point that call approximates.  Serial statement of the algorithm that k_omega runs with one block:
//   1. shift lw by its row maxima, then by the column maxima of the result (every row and column of K
//      holds an entry 1: no overflow, no empty line), K = exp(shifted);
//   2. a few Sinkhorn sweeps (rows, columns) as a warm-up;
//   3. Newton on the column scalings v with the rows normalised exactly:
//        P_ik = K_ik v_k / sum_k' K_ik' v_k',  c = P^T 1,  (diag(c) - P^T P + 11^T/M) delta = 1 - c,
//        v_k <- v_k exp(delta_k)
//      (the Hessian of the convex dual, singular only along 1, fixed by the rank-one term); a Sinkhorn
//      column step replaces the Newton step whenever the residual did not decrease.
// Sinkhorn alone needs hundreds of sweeps on the peaked matrices of the fine layers; Newton needs < 10.
// work: K[M*M], P[M*M], S[M*M], v[M], c[M], rhs[M], dinv[M]
// ------------------------------------------------------------------------------------------------
constexpr int kOmegaWarmup = 6;
constexpr int kOmegaMaxNewton = 40;
// Step policy (the same in omega_solve_serial, k_scale and k_scale_warp).  The first step is a Sinkhorn column
// step; Sinkhorn goes on while it contracts the residual by more than kOmegaFast per step (peaked tables: rate
// ~0.1, a step costs a tenth of a Newton step); otherwise a Newton step is taken once the residual is below
// kOmegaNewtonBelow (further out the Newton matrix can lose definiteness), unless the previous Newton step failed
// to reduce the residual.
constexpr double kOmegaFast = 0.33;
constexpr double kOmegaNewtonBelow = 0.2;
enum { kOmegaNone = 0, kOmegaSinkhorn = 1, kOmegaNewton = 2 };
__host__ __device__ inline bool omega_take_newton(double err, double err_prev, int last) {
    const bool sink_fast = last == kOmegaSinkhorn && err < kOmegaFast * err_prev;
    const bool newton_ok = last != kOmegaNone && err < kOmegaNewtonBelow && !(last == kOmegaNewton && !(err < err_prev));
    return newton_ok && !sink_fast;
}
constexpr double kOmegaTol = 1e-10;   // max |column sum - 1| (rows are exact); the reference solver stops near 1e-8
constexpr int kOmegaFallbackSweeps = 2000;

inline int omega_solve_serial(const double *lw, int M, double *omega, double *K, double *P, double *S, double *v,
                              double *c, double *rhs, double *dinv) {
    for (int i = 0; i < M; ++i) {
        double mx = -INFINITY;
        for (int k = 0; k < M; ++k) mx = fmax(mx, lw[i * M + k]);
        for (int k = 0; k < M; ++k) K[i * M + k] = lw[i * M + k] - mx;
    }
    for (int k = 0; k < M; ++k) {
        double mx = -INFINITY;
        for (int i = 0; i < M; ++i) mx = fmax(mx, K[i * M + k]);
        for (int i = 0; i < M; ++i) K[i * M + k] = exp(K[i * M + k] - mx);
        v[k] = 1.0;
    }
    int iters = 0;
    int last = kOmegaNone;
    double err_prev = INFINITY;
    for (int it = 0; it < kOmegaWarmup + kOmegaMaxNewton; ++it) {
        ++iters;
        for (int i = 0; i < M; ++i) {
            double s = 0.0;
            for (int k = 0; k < M; ++k) s = fma(K[i * M + k], v[k], s);
            const double u = 1.0 / s;
            for (int k = 0; k < M; ++k) P[i * M + k] = K[i * M + k] * v[k] * u;
        }
        double err = 0.0;
        for (int k = 0; k < M; ++k) {
            double s = 0.0;
            for (int i = 0; i < M; ++i) s += P[i * M + k];
            c[k] = s;
            err = fmax(err, fabs(s - 1.0));
        }
        if (err < kOmegaTol) break;
        if (!isfinite(err)) {   // overshooting Newton step: start again from the shifts alone
            for (int k = 0; k < M; ++k) v[k] = 1.0;
            err_prev = INFINITY;
            last = kOmegaNone;
            continue;
        }
        if (!omega_take_newton(err, err_prev, last)) {
            for (int k = 0; k < M; ++k) v[k] = fmax(1e-280, fmin(1e280, v[k] / c[k]));   // Sinkhorn column step
            err_prev = err;
            last = kOmegaSinkhorn;
            continue;
        }
        err_prev = err;
        last = kOmegaNewton;
        for (int k = 0; k < M; ++k)
            for (int m = 0; m <= k; ++m) {
                double s = 0.0;
                for (int i = 0; i < M; ++i) s = fma(P[i * M + k], P[i * M + m], s);
                S[k * M + m] = ((k == m) ? c[k] : 0.0) - s + 1.0 / (double)M;
            }
        for (int k = 0; k < M; ++k) rhs[k] = 1.0 - c[k];
        for (int j = 0; j < M; ++j) {   // Cholesky, lower triangle in place (diagonal kept, 1/l_jj aside)
            const double d = sqrt(S[j * M + j]);
            dinv[j] = 1.0 / d;
            for (int i = j + 1; i < M; ++i) S[i * M + j] *= dinv[j];
            for (int r = j + 1; r < M; ++r)
                for (int q = j + 1; q <= r; ++q) S[r * M + q] -= S[r * M + j] * S[q * M + j];
        }
        for (int j = 0; j < M; ++j) {
            rhs[j] *= dinv[j];
            for (int i = j + 1; i < M; ++i) rhs[i] -= S[i * M + j] * rhs[j];
        }
        for (int j = M - 1; j >= 0; --j) {
            rhs[j] *= dinv[j];
            for (int i = 0; i < j; ++i) rhs[i] -= S[j * M + i] * rhs[j];
        }
        for (int k = 0; k < M; ++k) v[k] = fmax(1e-280, fmin(1e280, v[k] * exp(fmax(-30.0, fmin(30.0, rhs[k])))));
    }
    // last resort (not seen on model tables, reachable with synthetic near-permutation input): plain Sinkhorn
    // sweeps from the current scalings until the tolerance or the iteration cap
    for (int it = 0; it < kOmegaFallbackSweeps; ++it) {
        double err = 0.0;
        for (int i = 0; i < M; ++i) {
            double s = 0.0;
            for (int k = 0; k < M; ++k) s = fma(K[i * M + k], v[k], s);
            const double u = 1.0 / s;
            for (int k = 0; k < M; ++k) P[i * M + k] = K[i * M + k] * v[k] * u;
        }
        for (int k = 0; k < M; ++k) {
            double s = 0.0;
            for (int i = 0; i < M; ++i) s += P[i * M + k];
            c[k] = s;
            err = fmax(err, fabs(s - 1.0));
        }
        if (err < kOmegaTol || !isfinite(err)) break;
        ++iters;
        for (int k = 0; k < M; ++k) v[k] = fmax(1e-280, fmin(1e280, v[k] / c[k]));
    }
    for (int t = 0; t < M * M; ++t) omega[t] = P[t];
    return iters;
}

// Basis angle: theta = pi (x + L)/(2L) = pi (u + 1/2), u = x / (2L)  (KernelClass.py:31-35), so
// sin(theta) = cos(pi u) and cos(theta) = -sin(pi u).  Returns phi_1 = L^-1/2 sin(theta) and
// c2 = 2 cos(theta); higher orders follow phi_{i+1} = c2 phi_i - phi_{i-1} with phi_0 = 0.
// sin(pi u), cos(pi u): u = k/2 + r with k = rint(2u), |r| <= 1/4; Taylor polynomials of sin(pi r)/r (8 terms)
// and cos(pi r) (9 terms) in r^2 are exact to 1-2 ulp on that range; the quadrant k mod 4 swaps and negates.
// About 22 FP64 operations and a dozen integer ones per call (the library sincospi spends ~100 instructions,
// most of them on special values that cannot occur here); valid for |u| < 2^30.
MRGP_HD void sincospi_fast(double u, double &s, double &c) {
    const double k = rint(u + u);
    const double r = fma(k, -0.5, u);
#if defined(__CUDA_ARCH__)
    const int q = __double2int_rn(k);
#else
    const int q = (int)(long long)k;
#endif
    const double r2 = r * r;
    double ps = 7.95205400147550838e-07 * 0.0 + -2.19153534478302037e-05;
    ps = fma(ps, r2, 4.66302805767612337e-04);
    ps = fma(ps, r2, -7.37043094571434784e-03);
    ps = fma(ps, r2, 8.21458866111281910e-02);
    ps = fma(ps, r2, -5.99264529320791883e-01);
    ps = fma(ps, r2, 2.55016403987734508e+00);
    ps = fma(ps, r2, -5.16771278004996937e+00);
    ps = fma(ps, r2, 3.14159265358979312e+00);
    ps *= r;
    double pc = 4.30306958703294391e-06;
    pc = fma(pc, r2, -1.04638104924845650e-04);
    pc = fma(pc, r2, 1.92957430940392206e-03);
    pc = fma(pc, r2, -2.58068913900140508e-02);
    pc = fma(pc, r2, 2.35330630358893123e-01);
    pc = fma(pc, r2, -1.33526276885458928e+00);
    pc = fma(pc, r2, 4.05871212641676760e+00);
    pc = fma(pc, r2, -4.93480220054467900e+00);
    pc = fma(pc, r2, 1.0);
    s = (q & 1) ? pc : ps;
    c = (q & 1) ? ps : pc;
    if (q & 2) s = -s;
    if ((q + 1) & 2) c = -c;
}

MRGP_HD void basis_seed(double x, double inv2L, double rsqrtL, double &phi1, double &c2) {
    double sp, cp;
    sincospi_fast(x * inv2L, sp, cp);
    phi1 = rsqrtL * cp;
    c2 = -2.0 * sp;
}

}  // namespace mrgp
