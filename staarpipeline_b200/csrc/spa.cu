// spa.cu — K6: saddle-point approximation p-values for imbalanced binary traits.
//
// Replaces Individual_Score_Test_SPA and its 13 helpers (reference src/Individual_Score_Test_SPA.cpp:38-533).
// One CTA per variant; the control flow of the reference (Newton-Raphson, its max_iter quirk, golden-section
// bracket + bisection fallback, every "_alt if not finite" retry, the ==1.0 / NaN fallbacks and the final clamp)
// runs uniformly in all threads, and every O(n) CGF sum is a block-parallel pass with a fixed reduction tree.
// What changes versus the reference is only work the reference repeats: K1 and K2 at the same point share one
// pass (one exp per sample), and G_tilde = G - XXWX_inv*(XW*G) (:61) is formed per column into an L2/HBM
// scratch slot instead of as a second n x p matrix.  FP64-ALU bound (exp/div per sample per pass).
#include "common.cuh"

namespace staar {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxQ = 64;

struct SpaCol {
    const double *mu;   // muhat, length n
    const double *g;    // G_tilde column, length n
    int64_t n;
    double *red;        // shared double[32]
    long long passes;
};

__device__ __forceinline__ bool is_bad(double v) { return !isfinite(v); }   // (R_finite(v)==0)||check_is_na(v)
__device__ __forceinline__ bool same_sign(double a, double b) { return (a >= 0) == (b >= 0); }   // :531-533

// K1 (:453-467) and K2 (:488-505) at the same x in one pass
__device__ void pass_K1K2(SpaCol &c, double x, double q, double &k1, double &k2)
{
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = threadIdx.x; i < c.n; i += kThreads) {
        const double mu = c.mu[i], g = c.g[i];
        const double e = exp(-x * g);
        const double den = mu + (1 - mu) * e;
        s1 += -mu * g + mu * g / den;
        s2 += (mu * (1 - mu) * (g * g) * e) / (den * den);
    }
    k1 = block_sum(s1, c.red) - q;
    k2 = block_sum(s2, c.red);
    c.passes++;
}
__device__ double pass_K1_alt(SpaCol &c, double x, double q)   // :471-485
{
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < c.n; i += kThreads) {
        const double mu = c.mu[i], g = c.g[i];
        const double e = exp(x * g);
        s += -mu * g + mu * g * e / (mu * e + (1 - mu));
    }
    c.passes++;
    return block_sum(s, c.red) - q;
}
__device__ double pass_K2(SpaCol &c, double x)                  // :488-505
{
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < c.n; i += kThreads) {
        const double mu = c.mu[i], g = c.g[i];
        const double e = exp(-x * g);
        const double den = mu + (1 - mu) * e;
        s += (mu * (1 - mu) * (g * g) * e) / (den * den);
    }
    c.passes++;
    return block_sum(s, c.red);
}
__device__ double pass_K2_alt(SpaCol &c, double x)              // :508-525
{
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < c.n; i += kThreads) {
        const double mu = c.mu[i], g = c.g[i];
        const double den = mu * exp(x * g) + (1 - mu);
        s += (mu * (1 - mu) * (g * g)) / (den * den);
    }
    c.passes++;
    return block_sum(s, c.red);
}
__device__ double pass_K(SpaCol &c, double x)                   // :421-433
{
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < c.n; i += kThreads) {
        const double mu = c.mu[i], g = c.g[i];
        s += -x * mu * g + log(1 - mu + mu * exp(x * g));
    }
    c.passes++;
    return block_sum(s, c.red);
}
__device__ double pass_K_alt(SpaCol &c, double x)               // :437-450
{
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < c.n; i += kThreads) {
        const double mu = c.mu[i], g = c.g[i];
        s += log((1 - mu) * exp(-x * g) + mu);
    }
    c.passes++;
    return block_sum(s, c.red);
}
// K(x) and K2(x) of the tail formula in one pass (:129,:140)
__device__ void pass_K_K2(SpaCol &c, double x, double &k, double &k2)
{
    double s = 0.0, s2 = 0.0;
    for (int64_t i = threadIdx.x; i < c.n; i += kThreads) {
        const double mu = c.mu[i], g = c.g[i];
        s += -x * mu * g + log(1 - mu + mu * exp(x * g));
        const double e = exp(-x * g);
        const double den = mu + (1 - mu) * e;
        s2 += (mu * (1 - mu) * (g * g) * e) / (den * den);
    }
    k = block_sum(s, c.red);
    k2 = block_sum(s2, c.red);
    c.passes++;
}
__device__ double K1_fb(SpaCol &c, double x, double q)          // K1, "_alt if not finite"
{
    double k1, k2;
    pass_K1K2(c, x, q, k1, k2);
    if (is_bad(k1)) k1 = pass_K1_alt(c, x, q);
    return k1;
}

// NR_Binary_SPA (:158-233)
__device__ double NR_spa(SpaCol &c, double q, double tol, int max_iter, long long &nr_iter)
{
    double xi = 0.0, xi_update = 0.0, k1, k2;
    pass_K1K2(c, xi, q, k1, k2);
    if (fabs(k1) > tol) xi_update = xi - k1 / k2;                  // :167-170 (no _alt retry here)
    int no_iter = 0;
    while (true) {                                                 // :177
        if (isnan(xi_update) || !isfinite(xi_update)) break;
        if (!(fabs(xi_update - xi) > tol)) break;
        pass_K1K2(c, xi_update, q, k1, k2);
        if (!(fabs(k1) > tol)) break;
        if (!(no_iter < max_iter)) break;
        no_iter = no_iter + 1;
        xi = xi_update;
        double numerator = k1;                                     // K1(xi), same point as the loop test
        if (is_bad(numerator)) numerator = pass_K1_alt(c, xi, q);
        double denominator = k2;
        if (is_bad(denominator)) denominator = pass_K2_alt(c, xi);
        xi_update = xi - numerator / denominator;
    }
    nr_iter += no_iter;
    if (is_bad(xi_update)) xi_update = xi;                         // :200-203
    if (no_iter == max_iter) {                                     // :211-230, literal (xhat still 0.0 in w)
        double xhat = 0.0;
        double w = sqrt(2 * (xhat * q - pass_K(c, xhat)));
        xhat = xi_update;
        if (xhat < 0) w = -w;
        double ki = xhat * sqrt(pass_K2(c, xhat));
        if (is_bad(ki)) ki = xhat * sqrt(pass_K2_alt(c, xhat));
        if (fabs(w + log(ki / w) / w) < 38) xi_update = 0.0;
    }
    return xi_update;
}

// the tail formula shared by Saddle_Binary_SPA (:129-152) and Saddle_Binary_SPA_Bisection (:260-283)
__device__ double tail_from_xhat(SpaCol &c, double xhat, double q, bool lower)
{
    double K, K2;
    pass_K_K2(c, xhat, K, K2);
    double w = sqrt(2 * (xhat * q - K));
    if (is_bad(w)) w = sqrt(2 * (xhat * q - pass_K_alt(c, xhat)));
    if (xhat < 0) w = -w;
    double ki = xhat * sqrt(K2);
    if (is_bad(ki)) ki = xhat * sqrt(pass_K2_alt(c, xhat));
    if (fabs(xhat) < 1e-04) return 1.0;
    return pnorm_std(w + log(ki / w) / w, lower);
}

// goldenSectionSearchForSignChange (:360-417)
__device__ void golden_search(SpaCol &c, double a, double b, double q, double tol, int max_iter, double &x0, double &x1)
{
    int iter = 0;
    x0 = a; x1 = b;
    const double phi = (1 + sqrt(5.0)) / 2;
    double K1x0 = K1_fb(c, x0, q);
    double K1x1 = K1_fb(c, x1, q);
    while ((iter < max_iter) && (fabs(x1 - x0) > tol) && same_sign(K1x0, K1x1)) {
        x1 = b - (b - a) / phi;
        x0 = a + (b - a) / phi;
        K1x0 = K1_fb(c, x0, q);
        K1x1 = K1_fb(c, x1, q);
        if (K1x1 < K1x0) b = x0; else a = x1;
        iter++;
    }
}

// Bisection_Binary_SPA (:290-358)
__device__ double bisection(SpaCol &c, double q, double xmin, double xmax, double tol)
{
    double xupper = xmax, xlower = xmin, x0 = 0.0, K1x0 = 1.0;
    double K1left = K1_fb(c, xlower, q);
    double K1right = K1_fb(c, xupper, q);
    if (!is_bad(K1right) && !is_bad(K1left)) {
        if (!same_sign(K1left, K1right)) {
            while ((fabs(xupper - xlower) > tol) && (fabs(K1x0) > tol)) {
                x0 = (xupper + xlower) / 2.0;
                K1x0 = K1_fb(c, x0, q);
                if (K1x0 == 0) break;
                if (same_sign(K1left, K1x0)) xlower = x0; else xupper = x0;
                K1left = K1_fb(c, xlower, q);              // re-evaluated every iteration (:342-352)
                K1right = K1_fb(c, xupper, q);
            }
        }
    }
    return x0;
}

__device__ double saddle_bisect(SpaCol &c, double q, double tol, int max_iter, bool lower)   // :235-286
{
    double a, b;
    golden_search(c, -100.0, 100.0, q, tol, max_iter, a, b);
    double xmin = a, xmax = b;
    if (!(a < b)) { xmin = b; xmax = a; }
    const double xhat = bisection(c, q, xmin, xmax, tol);
    return tail_from_xhat(c, xhat, q, lower);
}

// grid-stride over variants; CTA `blockIdx.x` owns scratch slot blockIdx.x (n doubles) for G_tilde.
__global__ void __launch_bounds__(kThreads) spa_kernel(const double *__restrict__ G, int64_t n, int64_t p,
                                                       const double *__restrict__ XW,          // q x n
                                                       const double *__restrict__ XXWX_inv,    // n x q
                                                       int q, const double *__restrict__ residuals,
                                                       const double *__restrict__ muhat, double tol, int max_iter,
                                                       double *__restrict__ gt_scratch, double *__restrict__ pvalue,
                                                       long long *__restrict__ trace)
{
    __shared__ double red[32];
    __shared__ double sb[kMaxQ];
    double *gt = gt_scratch + (size_t)blockIdx.x * n;
    for (int64_t v = blockIdx.x; v < p; v += gridDim.x) {
        const double *g = G + v * n;
        // b = XW * g (:61 inner product), sum0 = residuals' g (:64, raw G)
        double s0 = 0.0;
        for (int64_t j = threadIdx.x; j < n; j += kThreads) s0 += residuals[j] * g[j];
        const double sum0 = block_sum(s0, red);
        for (int c0 = 0; c0 < q; c0++) {
            double s = 0.0;
            for (int64_t j = threadIdx.x; j < n; j += kThreads) {
                const double gj = g[j];
                if (gj != 0.0) s += XW[c0 + j * q] * gj;
            }
            s = block_sum(s, red);
            if (threadIdx.x == 0) sb[c0] = s;
        }
        __syncthreads();
        for (int64_t j = threadIdx.x; j < n; j += kThreads) {
            double s = 0.0;
            for (int c0 = 0; c0 < q; c0++) s += XXWX_inv[j + (int64_t)c0 * n] * sb[c0];
            gt[j] = g[j] - s;
        }
        __syncthreads();

        SpaCol col{muhat, gt, n, red, 0};
        long long nr_iter = 0, bits = 0;
        const double aq = fabs(sum0);
        double res = 1.0;
        double xhat = NR_spa(col, aq, tol, max_iter, nr_iter);                 // :78
        double respart1 = tail_from_xhat(col, xhat, aq, false);
        if (isnan(respart1) || respart1 == 1.0) {                              // :80-83
            bits |= 1;
            respart1 = saddle_bisect(col, aq, tol, max_iter, false);
        }
        if (isnan(respart1) || respart1 == 1.0) {                              // :85-88
            bits |= 2;
            res = 1.0;
        } else {
            xhat = NR_spa(col, -aq, tol, max_iter, nr_iter);                   // :90
            double respart2 = tail_from_xhat(col, xhat, -aq, true);
            if (isnan(respart2) || respart2 == 1.0) {
                bits |= 4;
                respart2 = saddle_bisect(col, -aq, tol, max_iter, true);
            }
            if (isnan(respart2) || respart2 == 1.0) { bits |= 8; res = 1.0; }
            else res = respart1 + respart2;
        }
        if (res > 1) { bits |= 16; res = 1.0; }                                // :107-110
        if (threadIdx.x == 0) {
            pvalue[v] = res;
            if (trace) {
                trace[4 * v + 0] = col.passes; trace[4 * v + 1] = nr_iter; trace[4 * v + 2] = 0;
                trace[4 * v + 3] = bits | (nr_iter << 8);
            }
        }
        __syncthreads();
    }
}

}  // namespace

int spa_scratch_slots(staar_ctx *ctx, int64_t p) { return (int)std::min<int64_t>(p, (int64_t)ctx->sm_count * 2); }

int launch_spa(staar_ctx *ctx, const double *dG, int64_t n, int64_t p, const double *dXW, const double *dXXWX_inv,
               int q, const double *d_res, const double *d_mu, double tol, int max_iter, double *d_gt,
               double *d_pvalue, long long *d_trace)
{
    if (p == 0) return STAAR_OK;
    if (q > kMaxQ) { set_error("SPA: q = %d exceeds the supported maximum of %d", q, kMaxQ); return STAAR_ERR_ARG; }
    spa_kernel<<<spa_scratch_slots(ctx, p), kThreads, 0, ctx->stream>>>(dG, n, p, dXW, dXXWX_inv, q, d_res, d_mu, tol,
                                                                        max_iter, d_gt, d_pvalue, d_trace);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
