/*
 * mrgp_port.c - CPU restatement in plain C (OpenMP) of one ciMRGP variational sweep.
 *
 * TEST INFRASTRUCTURE ONLY: this is the multi-threaded CPU baseline that bench.py times beside the GPU path
 * (`cpu_baseline`, `--impl reference`: kind "port").  The parity checker is oracle/mrgp_oracle.py (pinned to the
 * reference's own outputs, tests/golden); tests/test_port_c.py holds this file to that oracle.  Nothing in the product
 * package cimrgp_b200/ includes, links or calls it.
 *
 * Scope: MultiResolutionGaussianProcess._fit (reference src/MRGP.py:571-652), ci mode, dx = 1, dy = 2, static basis
 * intervals, region-specific noise and bias, non-informative initialisation (src/Priors.py) - BASELINE configs 3-5.
 * Every layer streams its samples twice per sweep, as the reference does (no closed-form shortcuts):
 *   pass A  (Posteriors.py:35-78)   T = Phi^T (y - fbar - b - Phi A_old),  y_tilde_i = T_i + d_i a_i  (the O(M)
 *           form of the k != i penalty of :61-78), targets inferred from the layer's own posterior for layers > 0
 *           (LatentOutputs.py:20-49);
 *   shared  (Posteriors.py:497-541, Stats.py:67-100, 375-445, CommonDensities.py:71-77,
 *           computeRealBinghamConstant.py:42-147, SanityCheck.py:16-65) Bingham axis update, scale moments, ARD,
 *           permutation weights omega (log-domain Sinkhorn to 1e-13: the fixed point the reference's fsolve approximates);
 *   pass B  (Posteriors.py:81-148, Stats.py:102-157) residual statistics with the new coefficients, bias / noise
 *           update, propagation of the latent mean / variance to the next layer (running prefix of Stats.py:126-157).
 * The basis phi_i(x) = L^-1/2 sin(pi i (x + L) / (2 L)) (KernelClass.py:21-37) is regenerated per sample by the
 * three-term recurrence from one sin / cos instead of being stored (the reference keeps J N M doubles).
 *
 * Threads: the samples of a layer are cut into chunks at region boundaries; chunks run in parallel, their partial
 * sums are combined per region in chunk order (deterministic for a chunk size).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPS0 1e-45 /* Priors.py:5 */
#define PI 3.14159265358979323846
#define CHUNK 4096
#define MAXM 64

typedef struct {
    int R;
    int64_t *off;                                  /* R + 1 */
    double *L, *S, *d;                             /* (R), (R, M), (R, M) */
    double *prec, *zeta, *ytil, *A, *A_prev, *m2, *cm2; /* (R, M[, 2]) */
    double *noise_shape, *noise_scale, *noise_mean, *noise_log_mean, *bias_prec, *bias_mean, *bias_var, *noise_scale0;
    int n_chunks;
    int64_t *c_lo, *c_hi;                          /* chunk sample ranges */
    int *c_region, *region_chunk;                  /* region of a chunk; first chunk of a region (R + 1) */
    double *part;                                  /* n_chunks x (2 M + 8) */
} Layer;

typedef struct {
    int64_t N;
    int J, M;
    const double *x, *y;                           /* borrowed: (N), (N, 2) */
    double *fbar, *fvar;                           /* (N, 2), (N): latent functions of the CURRENT layer (running prefix) */
    Layer *layer;
    double *B, *logC, *rho, *cov, *shape, *scale, *mean, *lmean, *omega; /* shared posterior / stats */
    double ard0_scale;
    double t_omega;
} Port;

static double digamma_(double x) { /* scipy.special.psi, Stats.py:36 */
    double acc = 0.0;
    while (x < 10.0) {
        acc -= 1.0 / x;
        x += 1.0;
    }
    const double inv = 1.0 / x, i2 = inv * inv;
    const double series = i2 * (1.0 / 12.0 - i2 * (1.0 / 120.0 - i2 * (1.0 / 252.0 - i2 * (1.0 / 240.0 - i2 * (1.0 / 132.0 - i2 * (691.0 / 32760.0 - i2 * (1.0 / 12.0)))))));
    return acc + (log(x) - 0.5 * inv - series);
}

static double matern_(double lam, double nu, double l, double sf) { /* KernelClass.py:80-90 */
    const double log_arg = log(2.0 * nu) - 2.0 * log(l), arg = exp(log_arg);
    return exp(log(sf) + 0.5 * log(2.0 * PI) + nu * log_arg + lgamma(nu + 0.5) - lgamma(nu) - (nu + 0.5) * log(arg + lam));
}

/* first-order saddle-point log C(kappa) and gradient, p = 2 (computeRealBinghamConstant.py:42-147; root by bisection +
 * Newton on [0.1 - p, 0.1 - 0.5] where the reference calls brentq) */
static void saddle2_(const double *kappa, double *logc, double *rho) {
    double lam[2] = {-kappa[0], -kappa[1]};
    const double adjust = 0.1 - fmin(lam[0], lam[1]);
    lam[0] += adjust;
    lam[1] += adjust;
    double lo = 0.1 - 2.0, hi = 0.1 - 0.5, t = hi;
    for (int it = 0; it < 200; ++it) {
        t = 0.5 * (lo + hi);
        const double f = 0.5 * (1.0 / (lam[0] - t) + 1.0 / (lam[1] - t)) - 1.0;
        if (f > 0.0)
            hi = t;
        else
            lo = t;
        if (hi - lo < 1e-16 * fabs(t)) break;
    }
    double r1[2], k2 = 0.0, k3 = 0.0, sumlog = 0.0;
    for (int k = 0; k < 2; ++k) {
        r1[k] = 1.0 / (lam[k] - t);
        k2 += r1[k] * r1[k];
        k3 += r1[k] * r1[k] * r1[k];
        sumlog += log(lam[k] - t);
    }
    k2 *= 0.5;
    *logc = 0.5 * (log(2.0) + log(PI) - log(k2) - sumlog) - t + adjust;
    const double dk1dt = 0.5 * (r1[0] * r1[0] + r1[1] * r1[1]), dsumlogdt = -(r1[0] + r1[1]);
    for (int k = 0; k < 2; ++k) {
        const double dk1dlam = -0.5 * r1[k] * r1[k], dtdlam = -dk1dlam / dk1dt;
        const double dk2dlam = -(r1[k] * r1[k] * r1[k]) + k3 * dtdlam;
        rho[k] = 0.5 * dk2dlam / k2 + 0.5 * (r1[k] + dsumlogdt * dtdlam) + dtdlam;
    }
}

static int chol2_(double a, double b, double c) { return a > 0.0 && (c - (b / sqrt(a)) * (b / sqrt(a))) > 0.0; }

static void eig2_(double a, double b, double c, double *l1, double *l2, double *p00, double *p01, double *p11) {
    const double m = 0.5 * (a + c), d = 0.5 * (a - c), h = sqrt(d * d + b * b);
    *l1 = m + h;
    *l2 = m - h;
    if (h > 0.0) {
        *p00 = 0.5 + 0.5 * d / h;
        *p11 = 0.5 - 0.5 * d / h;
        *p01 = 0.5 * b / h;
    } else {
        *p00 = 1.0;
        *p01 = 0.0;
        *p11 = 0.0;
    }
}

/* Posteriors.py:519-530 + Stats.py:375-382 for one 2 x 2 matrix (a, b; b, c) */
static void bingham2_(double a, double b, double c, double *Bout, double *logc, double *rho, double *cov) {
    double l1, l2, p00, p01, p11;
    if (!chol2_(a, b, c)) { /* SanityCheck.py:16-57 */
        const double fro = sqrt(a * a + 2.0 * b * b + c * c);
        eig2_(a, b, c, &l1, &l2, &p00, &p01, &p11);
        const double m1 = l1 > 0.0 ? l1 : 0.0, m2 = l2 > 0.0 ? l2 : 0.0;
        a = m1 * p00 + m2 * (1.0 - p00);
        b = m1 * p01 - m2 * p01;
        c = m1 * p11 + m2 * (1.0 - p11);
        const double sp = nextafter(fro, INFINITY) - fro;
        for (int k = 1; k <= 64 && !chol2_(a, b, c); ++k) {
            eig2_(a, b, c, &l1, &l2, &p00, &p01, &p11);
            a += -l2 * k * k + sp;
            c += -l2 * k * k + sp;
        }
    }
    eig2_(a, b, c, &l1, &l2, &p00, &p01, &p11);
    const double kap[2] = {l1, l2};
    saddle2_(kap, logc, rho);
    Bout[0] = a;
    Bout[1] = b;
    Bout[2] = c;
    cov[0] = rho[0] * p00 + rho[1] * (1.0 - p00);
    cov[1] = rho[0] * p01 - rho[1] * p01;
    cov[2] = rho[0] * p11 + rho[1] * (1.0 - p11);
}

/* Stats.py:413-420: diag(alpha) exp(lw) diag(beta) with unit row and column sums, log-domain Sinkhorn */
static void omega_(const double *lw, int M, double *omega) {
    double la[MAXM] = {0}, lb[MAXM] = {0};
    for (int it = 0; it < 100000; ++it) {
        for (int i = 0; i < M; ++i) {
            double mx = -INFINITY, s = 0.0;
            for (int k = 0; k < M; ++k) mx = fmax(mx, lw[i * M + k] + lb[k]);
            for (int k = 0; k < M; ++k) s += exp(lw[i * M + k] + lb[k] - mx);
            la[i] = -(mx + log(s));
        }
        double err = 0.0;
        for (int k = 0; k < M; ++k) {
            double mx = -INFINITY, s = 0.0;
            for (int i = 0; i < M; ++i) mx = fmax(mx, lw[i * M + k] + la[i]);
            for (int i = 0; i < M; ++i) s += exp(lw[i * M + k] + la[i] - mx);
            const double nb = -(mx + log(s));
            err = fmax(err, fabs(nb - lb[k]));
            lb[k] = nb;
        }
        if (err < 1e-13) break;
    }
    for (int i = 0; i < M; ++i) {
        double s = 0.0;
        for (int k = 0; k < M; ++k) s += (omega[i * M + k] = exp(lw[i * M + k] + la[i] + lb[k]));
        for (int k = 0; k < M; ++k) omega[i * M + k] /= s;
    }
}

static inline void seed_(double x, double L, double *f1, double *c2) {
    const double th = PI * (x + L) / (2.0 * L);
    *f1 = sin(th) / sqrt(L);
    *c2 = 2.0 * cos(th);
}

Port *port_create(const double *x, const double *y, int64_t N, int M, int J, const int64_t *const *offsets, const int32_t *n_regions,
                  double nu, double ell, double sf) {
    if (M > MAXM) return NULL;
    Port *p = (Port *)calloc(1, sizeof(Port));
    p->N = N;
    p->J = J;
    p->M = M;
    p->x = x;
    p->y = y;
    p->fbar = (double *)calloc((size_t)N * 2, sizeof(double));
    p->fvar = (double *)calloc((size_t)N, sizeof(double));
    p->layer = (Layer *)calloc(J, sizeof(Layer));
    for (int j = 0; j < J; ++j) {
        Layer *ly = &p->layer[j];
        const int R = ly->R = n_regions[j];
        ly->off = (int64_t *)malloc((R + 1) * sizeof(int64_t));
        memcpy(ly->off, offsets[j], (R + 1) * sizeof(int64_t));
#define AL(n) (double *)calloc((size_t)(n), sizeof(double))
        ly->L = AL(R); ly->S = AL(R * M); ly->d = AL(R * M); ly->prec = AL(R * M); ly->zeta = AL(R * M);
        ly->ytil = AL(R * M * 2); ly->A = AL(R * M * 2); ly->A_prev = AL(R * M * 2); ly->m2 = AL(R * M); ly->cm2 = AL(R * M);
        ly->noise_shape = AL(R); ly->noise_scale = AL(R); ly->noise_mean = AL(R); ly->noise_log_mean = AL(R);
        ly->bias_prec = AL(R); ly->bias_mean = AL(R * 2); ly->bias_var = AL(R); ly->noise_scale0 = AL(R);
        /* chunks */
        int nc = 0;
        for (int r = 0; r < R; ++r) nc += (int)((ly->off[r + 1] - ly->off[r] + CHUNK - 1) / CHUNK);
        ly->n_chunks = nc;
        ly->c_lo = (int64_t *)malloc(nc * sizeof(int64_t));
        ly->c_hi = (int64_t *)malloc(nc * sizeof(int64_t));
        ly->c_region = (int *)malloc(nc * sizeof(int));
        ly->region_chunk = (int *)malloc((R + 1) * sizeof(int));
        ly->part = AL((size_t)nc * (2 * M + 8));
        int c = 0;
        for (int r = 0; r < R; ++r) {
            ly->region_chunk[r] = c;
            for (int64_t a = ly->off[r]; a < ly->off[r + 1]; a += CHUNK, ++c) {
                ly->c_lo[c] = a;
                ly->c_hi[c] = a + CHUNK < ly->off[r + 1] ? a + CHUNK : ly->off[r + 1];
                ly->c_region[c] = r;
            }
        }
        ly->region_chunk[R] = c;
        /* K3, K1, K2: L = max |x| (BasisInterval.py:15-16, factor 1), lambda, S, d = sum phi^2 (Posteriors.py:41) */
#pragma omp parallel for schedule(dynamic)
        for (int r = 0; r < R; ++r) {
            double mx = 0.0;
            for (int64_t n = ly->off[r]; n < ly->off[r + 1]; ++n) mx = fmax(mx, fabs(x[n]));
            ly->L[r] = mx;
            for (int i = 0; i < M; ++i) {
                const double w = PI * (i + 1) / (2.0 * mx);
                ly->S[r * M + i] = matern_(w * w, nu, ell, sf);
            }
        }
#pragma omp parallel for schedule(dynamic)
        for (int cc = 0; cc < nc; ++cc) {
            double acc[MAXM] = {0};
            const double L = ly->L[ly->c_region[cc]];
            for (int64_t n = ly->c_lo[cc]; n < ly->c_hi[cc]; ++n) {
                double f, c2, fm = 0.0;
                seed_(x[n], L, &f, &c2);
                for (int i = 0; i < M; ++i) {
                    acc[i] += f * f;
                    const double fn = c2 * f - fm;
                    fm = f;
                    f = fn;
                }
            }
            memcpy(ly->part + (size_t)cc * (2 * M + 8), acc, M * sizeof(double));
        }
        for (int r = 0; r < R; ++r)
            for (int cc = ly->region_chunk[r]; cc < ly->region_chunk[r + 1]; ++cc)
                for (int i = 0; i < M; ++i) ly->d[r * M + i] += ly->part[(size_t)cc * (2 * M + 8) + i];
        /* Priors.py:74-135, Posteriors.py:17-25, Stats.py:22-49 */
        for (int r = 0; r < R; ++r) {
            ly->noise_shape[r] = EPS0;
            ly->noise_scale[r] = ly->noise_scale0[r] = (EPS0 + 1.0) * 1.0;
            ly->noise_mean[r] = EPS0 / ly->noise_scale[r];
            ly->noise_log_mean[r] = digamma_(EPS0) - log(ly->noise_scale[r]);
            ly->bias_prec[r] = EPS0;
            ly->bias_var[r] = 1.0 / EPS0;
            for (int i = 0; i < M; ++i) ly->prec[r * M + i] = 1.0 / ly->S[r * M + i];
        }
    }
    /* Priors.py:29-53, Stats.py:361-369 */
    p->B = AL(M * 3); p->logC = AL(M); p->rho = AL(M * 2); p->cov = AL(M * 3); p->shape = AL(M); p->scale = AL(M);
    p->mean = AL(M); p->lmean = AL(M); p->omega = AL(M * M);
    p->ard0_scale = EPS0 / sf;
    const double z2[2] = {0.0, 0.0};
    double logc0, rho0[2];
    saddle2_(z2, &logc0, rho0);
    for (int i = 0; i < M; ++i) {
        p->logC[i] = logc0;
        p->rho[i * 2] = rho0[0];
        p->rho[i * 2 + 1] = rho0[1];
        p->shape[i] = EPS0;
        p->scale[i] = p->ard0_scale;
        p->mean[i] = EPS0 / p->ard0_scale;
        p->lmean[i] = digamma_(EPS0) - log(p->ard0_scale);
        for (int k = 0; k < M; ++k) p->omega[i * M + k] = 1.0 / M;
    }
    return p;
}

static double now_(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

void port_sweep(Port *p) {
    const int M = p->M, J = p->J, ST = 2 * M + 8;
    const double *x = p->x, *y = p->y;
    double *pB = (double *)malloc(M * 3 * sizeof(double)), *pLogC = (double *)malloc(M * sizeof(double));
    double *pShape = (double *)malloc(M * sizeof(double)), *pScale = (double *)malloc(M * sizeof(double));
    double *lw = (double *)malloc((size_t)M * M * sizeof(double));
    memset(p->fbar, 0, (size_t)p->N * 2 * sizeof(double)); /* layer 0: Stats.py:57-62 */
    memset(p->fvar, 0, (size_t)p->N * sizeof(double));
    for (int j = 0; j < J; ++j) {
        Layer *ly = &p->layer[j];
        const int R = ly->R, infer = j > 0;
        /* previous posterior: the prior for layer 0, the posterior of layer j - 1 otherwise (MRGP.py:575 / :581) */
        for (int i = 0; i < M; ++i) {
            if (j == 0) {
                pB[i * 3] = pB[i * 3 + 1] = pB[i * 3 + 2] = 0.0;
                const double z2[2] = {0.0, 0.0};
                double r0[2];
                saddle2_(z2, &pLogC[i], r0);
                pShape[i] = EPS0;
                pScale[i] = p->ard0_scale;
            } else {
                memcpy(pB + i * 3, p->B + i * 3, 3 * sizeof(double));
                pLogC[i] = p->logC[i];
                pShape[i] = p->shape[i];
                pScale[i] = p->scale[i];
            }
        }
        /* ---- pass A ---- */
#pragma omp parallel for schedule(dynamic)
        for (int cc = 0; cc < ly->n_chunks; ++cc) {
            const int r = ly->c_region[cc];
            const double L = ly->L[r], *A = ly->A + (size_t)r * M * 2, b0 = ly->bias_mean[r * 2], b1 = ly->bias_mean[r * 2 + 1];
            double T[2 * MAXM] = {0}, phi[MAXM];
            for (int64_t n = ly->c_lo[cc]; n < ly->c_hi[cc]; ++n) {
                double f, c2, fm = 0.0, e0 = 0.0, e1 = 0.0;
                seed_(x[n], L, &f, &c2);
                for (int i = 0; i < M; ++i) {
                    phi[i] = f;
                    e0 += f * A[i * 2];
                    e1 += f * A[i * 2 + 1];
                    const double fn = c2 * f - fm;
                    fm = f;
                    f = fn;
                }
                const double fb0 = p->fbar[n * 2], fb1 = p->fbar[n * 2 + 1];
                const double t0 = infer ? e0 + (b0 + fb0) : y[n * 2], t1 = infer ? e1 + (b1 + fb1) : y[n * 2 + 1];
                const double r0 = t0 - ((fb0 + b0) + e0), r1 = t1 - ((fb1 + b1) + e1);
                for (int i = 0; i < M; ++i) {
                    T[i * 2] += phi[i] * r0;
                    T[i * 2 + 1] += phi[i] * r1;
                }
            }
            memcpy(ly->part + (size_t)cc * ST, T, 2 * M * sizeof(double));
        }
        double Bd[3 * MAXM] = {0}, W[4 * MAXM] = {0};
        for (int r = 0; r < R; ++r) {
            const double noise = ly->noise_mean[r];
            for (int i = 0; i < M; ++i) {
                double t0 = 0.0, t1 = 0.0;
                for (int cc = ly->region_chunk[r]; cc < ly->region_chunk[r + 1]; ++cc) {
                    t0 += ly->part[(size_t)cc * ST + i * 2];
                    t1 += ly->part[(size_t)cc * ST + i * 2 + 1];
                }
                const int ri = r * M + i;
                const double y0 = t0 + ly->d[ri] * ly->A[ri * 2], y1 = t1 + ly->d[ri] * ly->A[ri * 2 + 1];
                const double prec = p->mean[i] / ly->S[ri] + noise * ly->d[ri], zeta = noise / prec;
                ly->ytil[ri * 2] = y0;
                ly->ytil[ri * 2 + 1] = y1;
                ly->prec[ri] = prec;
                ly->zeta[ri] = zeta;
                const double w = 0.5 * noise * zeta;
                Bd[i * 3] += w * y0 * y0;
                Bd[i * 3 + 1] += w * y0 * y1;
                Bd[i * 3 + 2] += w * y1 * y1;
            }
        }
        /* ---- shared step ---- */
        for (int i = 0; i < M; ++i) {
            double b00 = 0.0, b01 = 0.0, b11 = 0.0, sh = 0.0, sc = 0.0;
            for (int k = 0; k < M; ++k) {
                const double w = p->omega[i * M + k];
                b00 += w * pB[k * 3];
                b01 += w * pB[k * 3 + 1];
                b11 += w * pB[k * 3 + 2];
                sh += w * pShape[k];
                sc += w * pScale[k];
            }
            bingham2_(b00 + Bd[i * 3], b01 + Bd[i * 3 + 1], b11 + Bd[i * 3 + 2], p->B + i * 3, &p->logC[i], p->rho + i * 2, p->cov + i * 3);
            W[i] = sh;
            W[MAXM + i] = sc;
        }
        for (int i = 0; i < M; ++i) { /* Stats.py:67-100 per region, ARD sum (Posteriors.py:538-541) */
            const double c00 = p->cov[i * 3], c01 = p->cov[i * 3 + 1], c11 = p->cov[i * 3 + 2];
            double msum = 0.0;
            for (int r = 0; r < R; ++r) {
                const int ri = r * M + i;
                const double y0 = ly->ytil[ri * 2], y1 = ly->ytil[ri * 2 + 1], zeta = ly->zeta[ri], prec = ly->prec[ri];
                const double cy0 = c00 * y0 + c01 * y1, cy1 = c01 * y0 + c11 * y1;
                ly->A_prev[ri * 2] = ly->A[ri * 2];
                ly->A_prev[ri * 2 + 1] = ly->A[ri * 2 + 1];
                ly->A[ri * 2] = zeta * cy0;
                ly->A[ri * 2 + 1] = zeta * cy1;
                const double ccy0 = c00 * cy0 + c01 * cy1, ccy1 = c01 * cy0 + c11 * cy1;
                ly->m2[ri] = 1.0 / prec + zeta * zeta * (y0 * cy0 + y1 * cy1);
                ly->cm2[ri] = 1.0 / prec + zeta * zeta * (y0 * (cy0 - ccy0) + y1 * (cy1 - ccy1));
                msum += ly->m2[ri] / ly->S[ri];
            }
            p->shape[i] = W[i] + 0.5 * R;
            p->scale[i] = W[MAXM + i] + 0.5 * msum;
            p->mean[i] = p->shape[i] / p->scale[i];
            p->lmean[i] = digamma_(p->shape[i]) - log(p->scale[i]);
        }
        for (int i = 0; i < M; ++i) /* Stats.py:405-412 */
            for (int k = 0; k < M; ++k) {
                const double tr = p->cov[i * 3] * pB[k * 3] + 2.0 * p->cov[i * 3 + 1] * pB[k * 3 + 1] + p->cov[i * 3 + 2] * pB[k * 3 + 2];
                lw[i * M + k] = tr - pLogC[k] + pShape[k] * log(pScale[k]) - lgamma(pShape[k]) + (pShape[k] - 1.0) * p->lmean[i] - pScale[k] * p->mean[i];
            }
        const double t0w = now_();
        omega_(lw, M, p->omega);
        p->t_omega += now_() - t0w;
        /* ---- pass B + propagation ---- */
        const Layer *nx = j + 1 < J ? &p->layer[j + 1] : NULL;
#pragma omp parallel for schedule(dynamic)
        for (int cc = 0; cc < ly->n_chunks; ++cc) {
            const int r = ly->c_region[cc];
            const double L = ly->L[r], *A = ly->A + (size_t)r * M * 2, *Ao = ly->A_prev + (size_t)r * M * 2, *cm2 = ly->cm2 + (size_t)r * M;
            const double b0 = ly->bias_mean[r * 2], b1 = ly->bias_mean[r * 2 + 1];
            double s0 = 0.0, s1 = 0.0, rr = 0.0, sfv = 0.0, sau = 0.0;
            for (int64_t n = ly->c_lo[cc]; n < ly->c_hi[cc]; ++n) {
                double f, c2, fm = 0.0, e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0, v = 0.0;
                seed_(x[n], L, &f, &c2);
                for (int i = 0; i < M; ++i) {
                    e0 += f * A[i * 2];
                    e1 += f * A[i * 2 + 1];
                    o0 += f * Ao[i * 2];
                    o1 += f * Ao[i * 2 + 1];
                    v += f * f * cm2[i];
                    const double fn = c2 * f - fm;
                    fm = f;
                    f = fn;
                }
                const double fb0 = p->fbar[n * 2], fb1 = p->fbar[n * 2 + 1], fv = p->fvar[n];
                const double t0 = infer ? o0 + (b0 + fb0) : y[n * 2], t1 = infer ? o1 + (b1 + fb1) : y[n * 2 + 1];
                const double r0 = (t0 - e0) - fb0, r1 = (t1 - e1) - fb1;
                s0 += r0;
                s1 += r1;
                rr += r0 * r0 + r1 * r1;
                sfv += fv;
                sau += v;
                if (nx) { /* Stats.py:126-157 without this layer's bias terms (added below, once they are known) */
                    p->fbar[n * 2] = fb0 + e0;
                    p->fbar[n * 2 + 1] = fb1 + e1;
                    p->fvar[n] = fv + v;
                }
            }
            double *o = ly->part + (size_t)cc * ST;
            o[0] = s0; o[1] = s1; o[2] = rr; o[3] = sfv; o[4] = sau;
        }
        for (int r = 0; r < R; ++r) { /* Posteriors.py:81-93, 132-148; Stats.py:102-124 */
            double s[5] = {0, 0, 0, 0, 0};
            for (int cc = ly->region_chunk[r]; cc < ly->region_chunk[r + 1]; ++cc)
                for (int k = 0; k < 5; ++k) s[k] += ly->part[(size_t)cc * ST + k];
            const double n = (double)(ly->off[r + 1] - ly->off[r]), bp = EPS0 + n;
            const double m0 = s[0] / bp, m1 = s[1] / bp;
            const double yvar = infer ? 1.0 / ly->noise_mean[r] : 0.0;
            ly->bias_mean[r * 2] = m0;
            ly->bias_mean[r * 2 + 1] = m1;
            ly->bias_prec[r] = bp;
            ly->bias_var[r] = 1.0 / bp;
            ly->noise_shape[r] = EPS0 + 0.5 * 2.0 * n;
            ly->noise_scale[r] = ly->noise_scale0[r] + 0.5 * (0.0 - bp * (m0 * m0 + m1 * m1) + s[2] + s[3] + s[4] + yvar);
            ly->noise_mean[r] = ly->noise_shape[r] / ly->noise_scale[r];
            ly->noise_log_mean[r] = digamma_(ly->noise_shape[r]) - log(ly->noise_scale[r]);
        }
        if (nx) { /* the layer's new bias and bias variance enter the latent functions of the next layer */
#pragma omp parallel for schedule(static)
            for (int r = 0; r < R; ++r) {
                const double b0 = ly->bias_mean[r * 2], b1 = ly->bias_mean[r * 2 + 1], bv = ly->bias_var[r];
                for (int64_t n = ly->off[r]; n < ly->off[r + 1]; ++n) {
                    p->fbar[n * 2] += b0;
                    p->fbar[n * 2 + 1] += b1;
                    p->fvar[n] += bv;
                }
            }
        }
    }
    free(pB); free(pLogC); free(pShape); free(pScale); free(lw);
}

/* field: 0 A (R, M, 2), 1 noise_mean (R), 2 bias_mean (R, 2), 3 cm2 (R, M), 4 ytil (R, M, 2), 5 noise_scale (R);
 * layer -1: 10 B (M, 3), 11 omega (M, M), 12 ard_mean (M), 13 logC (M), 14 ard_scale (M) */
int64_t port_get(const Port *p, int layer, int field, double *out) {
    const int M = p->M;
    const double *src = NULL;
    int64_t n = 0;
    if (layer >= 0) {
        const Layer *ly = &p->layer[layer];
        switch (field) {
            case 0: src = ly->A; n = (int64_t)ly->R * M * 2; break;
            case 1: src = ly->noise_mean; n = ly->R; break;
            case 2: src = ly->bias_mean; n = (int64_t)ly->R * 2; break;
            case 3: src = ly->cm2; n = (int64_t)ly->R * M; break;
            case 4: src = ly->ytil; n = (int64_t)ly->R * M * 2; break;
            case 5: src = ly->noise_scale; n = ly->R; break;
        }
    } else {
        switch (field) {
            case 10: src = p->B; n = M * 3; break;
            case 11: src = p->omega; n = (int64_t)M * M; break;
            case 12: src = p->mean; n = M; break;
            case 13: src = p->logC; n = M; break;
            case 14: src = p->scale; n = M; break;
        }
    }
    if (src && out) memcpy(out, src, n * sizeof(double));
    return n;
}

double port_omega_seconds(const Port *p) { return p->t_omega; }

int port_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void port_destroy(Port *p) {
    if (!p) return;
    for (int j = 0; j < p->J; ++j) {
        Layer *ly = &p->layer[j];
        free(ly->off); free(ly->L); free(ly->S); free(ly->d); free(ly->prec); free(ly->zeta); free(ly->ytil); free(ly->A);
        free(ly->A_prev); free(ly->m2); free(ly->cm2); free(ly->noise_shape); free(ly->noise_scale); free(ly->noise_mean);
        free(ly->noise_log_mean); free(ly->bias_prec); free(ly->bias_mean); free(ly->bias_var); free(ly->noise_scale0);
        free(ly->c_lo); free(ly->c_hi); free(ly->c_region); free(ly->region_chunk); free(ly->part);
    }
    free(p->layer); free(p->fbar); free(p->fvar); free(p->B); free(p->logC); free(p->rho); free(p->cov); free(p->shape);
    free(p->scale); free(p->mean); free(p->lmean); free(p->omega);
    free(p);
}
