/* abi_smoke.c — include/abo.h driven from a host that is neither Python nor C++: plain C99, linked against
 * libabo_cuda.so only.  Proves that the header is sufficient to bind the hot path (what the Julia `ccall` stubs of
 * julia/AboCuda.jl rely on): create -> set_params -> fit -> posterior -> acq_eval (+ top-k) -> clone -> append ->
 * nlml_batch -> GradientGP fit / posterior -> error statuses -> destroy.
 * Expected numbers: the reference's own closed-form tests (test/test_surrogates.jl:59-105, 145-170).
 *   gcc -std=c99 -O1 -I include tests/abi_smoke.c -o /tmp/abi_smoke -L abstractbayesopt.jl_b200 -labo_cuda -Wl,-rpath,$PWD/abstractbayesopt.jl_b200 -lm
 * Exit code 0 = all checks passed (prints "abi_smoke ok"), 77 = no usable CUDA device (ABO_ERR_CUDA at context creation). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "abo.h"

#define CHECK(call)                                                                                     \
    do {                                                                                                \
        int32_t rc_ = (call);                                                                           \
        if (rc_ != ABO_OK) { fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #call, rc_, abo_last_error()); return 1; } \
    } while (0)
#define NEAR(a, b, tol)                                                                                 \
    do {                                                                                                \
        if (!(fabs((a) - (b)) <= (tol))) { fprintf(stderr, "%s:%d %s = %.17g, expected %.17g\n", __FILE__, __LINE__, #a, (double)(a), (double)(b)); return 1; } \
    } while (0)

int main(void) {
    abo_ctx* ctx = NULL;
    int32_t rc = abo_ctx_create(0, &ctx);
    if (rc == ABO_ERR_CUDA) { fprintf(stderr, "no CUDA device: %s\n", abo_last_error()); return 77; }
    if (rc != ABO_OK) { fprintf(stderr, "abo_ctx_create -> %d: %s\n", rc, abo_last_error()); return 1; }
    if (abo_version() < 100) return 1;

    /* ---- StandardGP, SE kernel, l = 1, sigma^2 = 1, noise 0.1; xs = [0, .5, 1], ys = [0, .25, 1] (G1) */
    abo_gp* gp = NULL;
    CHECK(abo_gp_create(ctx, ABO_KERNEL_SE, 1, 1, &gp));
    CHECK(abo_gp_set_params(gp, 1.0, 1.0, 0.1, NULL));
    const double X[3] = {0.0, 0.5, 1.0}, y[3] = {0.0, 0.25, 1.0};
    int64_t info = -1, n = 0;
    CHECK(abo_gp_fit(gp, X, y, 3, &info));
    CHECK(abo_gp_n(gp, &n));
    if (info != 0 || n != 3) return 1;
    const double xq[1] = {0.25};
    double mean = 0, var = 0;
    CHECK(abo_gp_posterior(gp, xq, 1, 1, &mean, &var));
    NEAR(mean, 0.1771247751991296, 1e-10);
    NEAR(var, 0.050320225208722924, 1e-10);

    /* ---- nlml at (log 1, log 1) (G2) with its analytic gradient checked by central differences */
    double theta[6] = {0.0, 0.0, 1e-5, 0.0, -1e-5, 0.0}, val[3], grad[6];
    int32_t ninfo[3];
    CHECK(abo_nlml_batch(gp, X, y, 3, theta, 3, val, grad, ninfo));
    NEAR(val[0], 2.6769327097262567, 1e-10);
    NEAR(grad[0], (val[1] - val[2]) / 2e-5, 1e-6);

    /* ---- acquisitions on ys = [2, 1, .5] (G3): EI, PI, UCB + the stable top-k of a small candidate set */
    const double y3[3] = {2.0, 1.0, 0.5};
    CHECK(abo_gp_fit(gp, X, y3, 3, &info));
    const double cand[5] = {0.25, 0.9, 0.25, 0.1, 0.6};
    double scores[5], top_val[3], p_ei[2] = {0.01, 0.5}, p_ucb[1] = {2.0};
    int64_t top_idx[3];
    CHECK(abo_acq_eval(gp, ABO_ACQ_EI, p_ei, cand, 5, scores, 3, top_idx, top_val));
    NEAR(scores[0], 3.11345832411526e-07, 1e-9 * 3.2e-7);
    if (scores[0] != scores[2]) return 1;                                   /* same point, same bits */
    if (!(top_val[0] >= top_val[1] && top_val[1] >= top_val[2]) || top_val[0] != scores[top_idx[0]]) return 1;
    CHECK(abo_acq_eval(gp, ABO_ACQ_PI, p_ei, cand, 5, scores, 0, NULL, NULL));
    NEAR(scores[0], 6.608138679027337e-06, 1e-9 * 6.7e-6);
    CHECK(abo_acq_eval(gp, ABO_ACQ_UCB, p_ucb, cand, 5, scores, 3, top_idx, top_val));
    NEAR(scores[0], -1.0186125700256665, 1e-10);
    if (scores[0] == scores[2] && !(top_idx[0] != 2 || top_idx[1] != 0)) return 1;   /* ties keep ascending index order */
    double g_scores[5], g_grad[5];
    CHECK(abo_acq_eval_grad(gp, ABO_ACQ_UCB, p_ucb, cand, 5, g_scores, g_grad, NULL, NULL));
    NEAR(g_scores[0], scores[0], 1e-12);

    /* ---- Base.copy + O(n^2) append: the clone is untouched, the appended posterior equals a re-fit */
    abo_gp *snap = NULL, *refit = NULL;
    CHECK(abo_gp_clone(gp, &snap));
    const double xn[1] = {0.75}, yn[1] = {0.7};
    CHECK(abo_gp_append(gp, xn, yn, &info));
    CHECK(abo_gp_n(gp, &n));
    if (n != 4) return 1;
    CHECK(abo_gp_n(snap, &n));
    if (n != 3) return 1;
    const double X4[4] = {0.0, 0.5, 1.0, 0.75}, y4[4] = {2.0, 1.0, 0.5, 0.7};
    CHECK(abo_gp_create(ctx, ABO_KERNEL_SE, 1, 1, &refit));
    CHECK(abo_gp_set_params(refit, 1.0, 1.0, 0.1, NULL));
    CHECK(abo_gp_fit(refit, X4, y4, 4, &info));
    double m_a, v_a, m_r, v_r, m_s, v_s;
    CHECK(abo_gp_posterior(gp, xq, 1, 1, &m_a, &v_a));
    CHECK(abo_gp_posterior(refit, xq, 1, 1, &m_r, &v_r));
    CHECK(abo_gp_posterior(snap, xq, 1, 1, &m_s, &v_s));
    NEAR(m_a, m_r, 1e-12); NEAR(v_a, v_r, 1e-12);
    NEAR(m_s, 1.467255970550952, 1e-10);                                    /* G3 mean: the snapshot still holds 3 points */

    /* ---- failure protocol across the ABI (test/test_bayesian_opt.jl:749-786): noise 0 and a near-duplicate point */
    abo_gp* bad = NULL;
    CHECK(abo_gp_create(ctx, ABO_KERNEL_SE, 2, 1, &bad));
    CHECK(abo_gp_set_params(bad, 1.0, 1.0, 0.0, NULL));
    const double Xb[6] = {-1.0, -1.0, 5.0, -5.0, -1.0 + 1e-12, -1.0 + 1e-12}, yb[3] = {2.0, 50.0, 2.0};
    rc = abo_gp_fit(bad, Xb, yb, 3, &info);
    if (rc != ABO_ERR_NOT_POSDEF || info != 3) { fprintf(stderr, "expected NOT_POSDEF at pivot 3, got %d / %lld\n", rc, (long long)info); return 1; }
    rc = abo_gp_posterior(bad, Xb, 1, 1, &mean, &var);
    if (rc != ABO_ERR_NOT_FITTED) return 1;

    /* ---- GradientGP (p = d + 1), out-major observations; posterior value + all outputs */
    abo_gp* gg = NULL;
    CHECK(abo_gp_create(ctx, ABO_KERNEL_SE, 2, 3, &gg));
    CHECK(abo_gp_set_params(gg, 1.0, 1.0, 0.1, NULL));
    const double Xg[6] = {0.0, 0.0, 0.5, 0.5, 1.0, 1.0};
    const double yg[9] = {1.0, 0.5, 0.0, /* d/dx1 */ 0.1, 0.0, -0.1, /* d/dx2 */ 0.1, 0.0, -0.1};
    CHECK(abo_gp_fit(gg, Xg, yg, 3, &info));
    const double xg[2] = {0.25, 0.25};
    double gm[3], gv[3], cov[9];
    CHECK(abo_gp_posterior(gg, xg, 1, 3, gm, gv));
    CHECK(abo_gp_posterior_cov(gg, xg, 1, 3, cov));
    NEAR(cov[0], gv[0], 1e-12); NEAR(cov[4], gv[1], 1e-12); NEAR(cov[8], gv[2], 1e-12);
    NEAR(cov[1], cov[3], 1e-14);
    const int32_t ids[2] = {ABO_ACQ_UCB, ABO_ACQ_GRADNORM_UCB};
    const double wts[2] = {0.5, 0.5}, pars[4] = {2.0, 0.0, 1.5, 0.0};
    double ms_[1];
    CHECK(abo_acq_eval_multi(gg, 2, ids, wts, pars, xg, 1, ms_, 0, NULL, NULL));
    {   /* the same two members from the posterior pieces read back above */
        double m1 = gm[1], m2 = gm[2], s11 = cov[4], s12 = cov[5], s22 = cov[8];
        double mu_sq = m1 * m1 + m2 * m2 + s11 + s22;
        double var_sq = 4 * (m1 * (s11 * m1 + s12 * m2) + m2 * (s12 * m1 + s22 * m2)) + 2 * (s11 * s11 + 2 * s12 * s12 + s22 * s22);
        double gn = -mu_sq + 1.5 * sqrt(var_sq > 1e-12 ? var_sq : 1e-12);
        double ucb = -gm[0] + 2.0 * sqrt(gv[0] > 0 ? gv[0] : 0);
        NEAR(ms_[0], 0.5 * ucb + 0.5 * gn, 1e-12);
    }

    int64_t launches = 0;
    CHECK(abo_ctx_launch_count(ctx, &launches));
    if (launches < 10) return 1;
    CHECK(abo_gp_destroy(gp)); CHECK(abo_gp_destroy(snap)); CHECK(abo_gp_destroy(refit)); CHECK(abo_gp_destroy(bad)); CHECK(abo_gp_destroy(gg));
    CHECK(abo_ctx_destroy(ctx));
    printf("abi_smoke ok (%lld kernel launches)\n", (long long)launches);
    return 0;
}
