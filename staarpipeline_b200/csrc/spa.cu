// spa.cu — K6: saddle-point approximation p-values for imbalanced binary traits.
//
// Replaces Individual_Score_Test_SPA and its 13 helpers (reference src/Individual_Score_Test_SPA.cpp:38-533).
// One thread-block CLUSTER per variant: the control flow of the reference (Newton-Raphson, its max_iter quirk,
// golden-section bracket + bisection fallback, every "_alt if not finite" retry, the ==1.0 / NaN fallbacks and the
// final clamp) runs uniformly in every thread of every CTA of the cluster, and each O(n) CGF sum is a pass in which
// CTA r of the cluster covers samples [r*len, (r+1)*len); partial sums are exchanged through distributed shared
// memory and added in rank order, so all CTAs continue with bit-identical values.  Why a cluster: the number of
// passes per variant ranges from ~15 to several thousand (a Newton-Raphson run that exhausts max_iter = 1000 and
// falls back to the bracket search), a pass is FP64-pipe bound on the SM it runs on (exp + divisions per sample), so
// the long variants set the makespan unless each one is spread over several SMs.  Variants are handed out through an
// atomic counter.  What changes versus the reference is only work the reference repeats: K1 and K2 at the same point
// share one pass (one exp per sample), and G_tilde = G - XXWX_inv*(XW*G) (:61) is formed per column into an L2-resident
// scratch slot instead of as a second n x p matrix.  The column comes either from the FP64 matrix of the frozen
// signature or straight from resident 2-bit columns (flip / imputation as matrix_flip_mean|minor, packed.cuh).
#include <cooperative_groups.h>
#include "common.cuh"
#include "packed.cuh"

namespace staar {

namespace {

namespace cg = cooperative_groups;

constexpr int kThreads = 512;
constexpr int kMaxQ = 64;
constexpr int kMaxCluster = 16;     // 16 needs the non-portable cluster size attribute (B200: supported)

struct SpaCol {
    const double *mu;   // muhat, length n
    const double *g;    // G_tilde column, length n
    int64_t lo, hi;     // this CTA's samples
    double *red;        // shared double[3][32]: warp partials of up to three sums
    double *slots;      // shared double[2][3][kMaxCluster]: [parity][value][rank], written by every CTA of the cluster
    unsigned csize, crank;
    int parity;
    long long passes;
};

// Up to three cluster-wide sums with a fixed tree: warp shuffles -> CTA partial -> one DSMEM store per peer -> cluster
// barrier -> rank-ordered add.  Every thread of every CTA returns the same bits.  The slot buffers alternate so that a
// CTA that runs ahead writes the buffer its peers are not reading (a buffer is reused only after the next barrier).
constexpr int kMaxSums = 3;
template <int K> __device__ __forceinline__ void cluster_sums(SpaCol &c, double (&v)[K])
{
    static_assert(K <= kMaxSums, "slot layout");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) c.red[32 * k + w] = v[k];
    }
    __syncthreads();
    cg::cluster_group cluster = cg::this_cluster();
    double *slot = c.slots + c.parity * kMaxSums * kMaxCluster;
    if (w == 0) {
        double t[K];
#pragma unroll
        for (int k = 0; k < K; k++) t[k] = warp_sum(lane < kThreads / 32 ? c.red[32 * k + lane] : 0.0);
        if (lane < (int)c.csize) {
            double *peer = cluster.map_shared_rank(slot, lane);
#pragma unroll
            for (int k = 0; k < K; k++) peer[k * kMaxCluster + c.crank] = t[k];
        }
    }
    cluster.sync();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double t = 0.0;
        for (unsigned r = 0; r < c.csize; r++) t += slot[k * kMaxCluster + r];
        v[k] = t;
    }
    c.parity ^= 1;
}
__device__ __forceinline__ void cluster_sum2(SpaCol &c, double a, double b, double &A, double &B)
{
    double v[2] = {a, b};
    cluster_sums(c, v);
    A = v[0]; B = v[1];
}
__device__ __forceinline__ double cluster_sum(SpaCol &c, double a)
{
    double v[1] = {a};
    cluster_sums(c, v);
    return v[0];
}

__device__ __forceinline__ bool is_bad(double v) { return !isfinite(v); }   // (R_finite(v)==0)||check_is_na(v)
__device__ __forceinline__ bool same_sign(double a, double b) { return (a >= 0) == (b >= 0); }   // :531-533

// One pass over this CTA's samples, kUnroll samples per thread in flight: the loads of the next batch are issued
// before the arithmetic of the current one (the loop is otherwise latency-bound: ~120 dependent instructions per sample
// behind two L2 loads at 4 warps per scheduler).
constexpr int kUnroll = 4;
template <class Body> __device__ __forceinline__ void sweep(const SpaCol &c, Body body)
{
    double m[kUnroll], g[kUnroll];
    int64_t i = c.lo + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kUnroll; k++) {
        const int64_t j = i + (int64_t)k * kThreads;
        const bool ok = j < c.hi;
        m[k] = ok ? c.mu[j] : 0.5;
        g[k] = ok ? c.g[j] : 0.0;
    }
    while (i < c.hi) {
        double mc[kUnroll], gc[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; k++) { mc[k] = m[k]; gc[k] = g[k]; }
        const int64_t i0 = i;
        i += (int64_t)kUnroll * kThreads;
#pragma unroll
        for (int k = 0; k < kUnroll; k++) {
            const int64_t j = i + (int64_t)k * kThreads;
            const bool ok = j < c.hi;
            m[k] = ok ? c.mu[j] : 0.5;
            g[k] = ok ? c.g[j] : 0.0;
        }
        if (i0 + (int64_t)(kUnroll - 1) * kThreads < c.hi) {
#pragma unroll
            for (int k = 0; k < kUnroll; k++) body(mc[k], gc[k]);
        } else {
#pragma unroll
            for (int k = 0; k < kUnroll; k++)
                if (i0 + (int64_t)k * kThreads < c.hi) body(mc[k], gc[k]);
        }
    }
}

// K1 (:453-467) and K2 (:488-505) at the same x in one pass
__device__ void pass_K1K2(SpaCol &c, double x, double q, double &k1, double &k2)
{
    double s1 = 0.0, s2 = 0.0;
    sweep(c, [&](double mu, double g) {
        const double e = exp(-x * g);
        const double r = 1.0 / (mu + (1 - mu) * e);            // one reciprocal serves both quotients
        s1 += -mu * g + mu * g * r;
        s2 += (mu * (1 - mu) * (g * g) * e) * r * r;
    });
    cluster_sum2(c, s1, s2, k1, k2);
    k1 -= q;
    c.passes++;
}
// The Newton-Raphson pass: K1, K2 and, from the same exponential, K2_alt (:508-525).  With e = exp(-x g) and
// r = 1/(mu + (1-mu) e):  mu exp(x g) + (1-mu) = (mu + (1-mu) e)/e, so the summand of K2_alt,
// mu(1-mu)g^2 / (mu exp(x g) + 1-mu)^2, equals mu(1-mu)g^2 (e r)^2 — no second exponential, no second pass.  Where e
// overflowed (the case that makes K2 NaN and sends the reference to K2_alt) e r is taken at its limit 1/(1-mu), which
// is what the reference's expression evaluates to there (exp(x g) = 0).  Matters because a run that leaves the range of
// exp() keeps stepping with K2_alt until max_iter: one pass per step instead of two.
__device__ void pass_K1K2_nr(SpaCol &c, double x, double q, double &k1, double &k2, double &k2_alt)
{
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    sweep(c, [&](double mu, double g) {
        const double e = exp(-x * g);
        const double r = 1.0 / (mu + (1 - mu) * e);
        const double a = mu * (1 - mu) * (g * g);
        s1 += -mu * g + mu * g * r;
        s2 += a * e * r * r;
        const double t = isinf(e) ? 1.0 / (1 - mu) : e * r;
        s3 += a * t * t;
    });
    double v[3] = {s1, s2, s3};
    cluster_sums(c, v);
    k1 = v[0] - q; k2 = v[1]; k2_alt = v[2];
    c.passes++;
}
__device__ double pass_K1_alt(SpaCol &c, double x, double q)   // :471-485
{
    double s = 0.0;
    sweep(c, [&](double mu, double g) {
        const double e = exp(x * g);
        s += -mu * g + mu * g * e / (mu * e + (1 - mu));
    });
    c.passes++;
    return cluster_sum(c, s) - q;
}
__device__ double pass_K2(SpaCol &c, double x)                  // :488-505
{
    double s = 0.0;
    sweep(c, [&](double mu, double g) {
        const double e = exp(-x * g);
        const double den = mu + (1 - mu) * e;
        s += (mu * (1 - mu) * (g * g) * e) / (den * den);
    });
    c.passes++;
    return cluster_sum(c, s);
}
__device__ double pass_K2_alt(SpaCol &c, double x)              // :508-525
{
    double s = 0.0;
    sweep(c, [&](double mu, double g) {
        const double den = mu * exp(x * g) + (1 - mu);
        s += (mu * (1 - mu) * (g * g)) / (den * den);
    });
    c.passes++;
    return cluster_sum(c, s);
}
__device__ double pass_K(SpaCol &c, double x)                   // :421-433
{
    double s = 0.0;
    sweep(c, [&](double mu, double g) { s += -x * mu * g + log(1 - mu + mu * exp(x * g)); });
    c.passes++;
    return cluster_sum(c, s);
}
__device__ double pass_K_alt(SpaCol &c, double x)               // :437-450
{
    double s = 0.0;
    sweep(c, [&](double mu, double g) { s += log((1 - mu) * exp(-x * g) + mu); });
    c.passes++;
    return cluster_sum(c, s);
}
// K(x) and K2(x) of the tail formula in one pass (:129,:140)
__device__ void pass_K_K2(SpaCol &c, double x, double &k, double &k2)
{
    double s = 0.0, s2 = 0.0;
    sweep(c, [&](double mu, double g) {
        s += -x * mu * g + log(1 - mu + mu * exp(x * g));
        const double e = exp(-x * g);
        const double r = 1.0 / (mu + (1 - mu) * e);
        s2 += (mu * (1 - mu) * (g * g) * e) * r * r;
    });
    cluster_sum2(c, s, s2, k, k2);
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
    double xi = 0.0, xi_update = 0.0, k1, k2, k2_alt;
    pass_K1K2(c, xi, q, k1, k2);
    if (fabs(k1) > tol) xi_update = xi - k1 / k2;                  // :167-170 (no _alt retry here)
    int no_iter = 0;
    while (true) {                                                 // :177
        if (isnan(xi_update) || !isfinite(xi_update)) break;
        if (!(fabs(xi_update - xi) > tol)) break;
        pass_K1K2_nr(c, xi_update, q, k1, k2, k2_alt);
        if (!(fabs(k1) > tol)) break;
        if (!(no_iter < max_iter)) break;
        no_iter = no_iter + 1;
        xi = xi_update;
        double numerator = k1;                                     // K1(xi), same point as the loop test
        if (is_bad(numerator)) numerator = pass_K1_alt(c, xi, q);
        double denominator = k2;
        if (is_bad(denominator)) denominator = k2_alt;             // K2_alt(xi), from the same pass
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

// Where a variant's column comes from: the FP64 matrix of the frozen signature, or resident 2-bit columns.
struct SpaInput {
    const double *G;                // n x p column-major (dense mode) or nullptr
    const uint8_t *packed;          // packed mode: resident block, `stride` bytes per column, null-model sample order
    size_t stride;
    const long long *variant;       // packed mode: column number of entry v
    const int32_t *inv_perm;        // packed mode: original sample -> position (nullptr: identity)
    int impute;
};

// Persistent clusters; cluster k owns scratch slot k (n doubles) for G_tilde; variants come from the counter `work`.
__global__ void __launch_bounds__(kThreads) spa_kernel(SpaInput in, int64_t n, int64_t p,
                                                       const double *__restrict__ XW,          // q x n
                                                       const double *__restrict__ XXWX_inv,    // n x q
                                                       int q, const double *__restrict__ residuals,
                                                       const double *__restrict__ muhat, double tol, int max_iter,
                                                       double *__restrict__ gt_scratch, double *__restrict__ pvalue,
                                                       long long *__restrict__ trace, unsigned long long *work)
{
    __shared__ double red[32 * kMaxSums];
    __shared__ double slots[2 * kMaxSums * kMaxCluster];
    __shared__ double sb[kMaxQ + 1];                          // b = XW g; slot kMaxQ = residuals'g
    __shared__ double wacc[kThreads / 32][kMaxQ + 1];         // per-warp partials of the same
    __shared__ double xch[kMaxCluster][kMaxQ + 1];            // per-CTA partials, written by every CTA of the cluster
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned csize = cluster.num_blocks(), crank = cluster.block_rank();
    const int64_t len = ((n + csize - 1) / csize + 3) & ~(int64_t)3;
    const int64_t lo = min((long long)n, (long long)crank * len), hi = min((long long)n, (long long)(lo + len));
    double *gt = gt_scratch + (size_t)(blockIdx.x / csize) * n;
    SpaCol col{muhat, gt, lo, hi, red, slots, csize, crank, 0, 0};
    while (true) {
        // next variant: rank 0 draws, the cluster sum broadcasts (exact: integers below 2^53)
        double draw = 0.0;
        if (crank == 0 && threadIdx.x == 0) draw = (double)atomicAdd(work, 1ull);
        const int64_t v = (int64_t)cluster_sum(col, draw);
        if (v >= p) break;
        col.passes = 0;
        // flip / imputation value of a resident column from its code counts (matrix_flip_mean|minor on integers)
        const double *gd = in.G ? in.G + v * n : nullptr;
        const uint8_t *pc = in.G ? nullptr : in.packed + (size_t)in.variant[v] * in.stride;
        VariantFlip vf{0.0, 0.0, 0.0, 0};
        if (!in.G) {
            int c1 = 0, c2 = 0, cm = 0;
            const uint4 *v4 = reinterpret_cast<const uint4 *>(pc);
            for (int64_t k = (int64_t)crank * kThreads + threadIdx.x; k < (int64_t)(in.stride >> 4); k += (int64_t)csize * kThreads) {
                const uint4 wd = __ldg(v4 + k);
                count_word(wd.x, c1, c2, cm); count_word(wd.y, c1, c2, cm); count_word(wd.z, c1, c2, cm); count_word(wd.w, c1, c2, cm);
            }
            double cnt[3] = {(double)c1, (double)c2, (double)cm};
            cluster_sums(col, cnt);
            vf = flip_from_counts(PackedCounts{(int)cnt[0], (int)cnt[1], (int)cnt[2]}, n, in.impute);
        }
        // One sweep over this CTA's samples, a warp per 32 consecutive samples: the column in FP64 (caller's sample order)
        // goes to the scratch slot, sum0 = residuals'g (:64, raw G), and b = XW g (:61) is accumulated over the carriers
        // only — lane c holds covariate c (and c + 32), so a carrier costs one coalesced read of its q-vector of XW.
        double s0 = 0.0, b0 = 0.0, b1 = 0.0;
        for (int64_t j0 = lo + 32 * warp; j0 < hi; j0 += kThreads) {
            const int64_t j = j0 + lane;
            double gj = 0.0;
            if (j < hi) {
                if (gd) gj = gd[j];
                else {
                    const int64_t pos = in.inv_perm ? in.inv_perm[j] : j;
                    uint32_t code = ((uint32_t)__ldg(pc + (pos >> 2)) >> (2 * (pos & 3))) & 3u;
                    if (vf.flip && !(code & 1u)) code ^= 2u;
                    gj = code == 3u ? vf.mu : (double)code;
                }
                gt[j] = gj;
                s0 += residuals[j] * gj;
            }
            unsigned carriers = __ballot_sync(0xffffffffu, gj != 0.0);
            while (carriers) {
                const int src = __ffs(carriers) - 1;
                carriers &= carriers - 1;
                const double gs = __shfl_sync(0xffffffffu, gj, src);
                const double *xw = XW + (j0 + src) * q;
                if (lane < q) b0 += xw[lane] * gs;
                if (lane + 32 < q) b1 += xw[lane + 32] * gs;
            }
        }
        s0 = warp_sum(s0);
        wacc[warp][lane] = b0; wacc[warp][lane + 32] = b1;
        if (lane == 0) wacc[warp][kMaxQ] = s0;
        __syncthreads();
        if (threadIdx.x <= kMaxQ) {                     // warps in order, then CTAs in rank order: fixed summation tree
            double t = 0.0;
            for (int ww = 0; ww < kThreads / 32; ww++) t += wacc[ww][threadIdx.x];
            for (unsigned r = 0; r < csize; r++) cluster.map_shared_rank(&xch[0][0], r)[crank * (kMaxQ + 1) + threadIdx.x] = t;
        }
        cluster.sync();
        if (threadIdx.x <= kMaxQ) {
            double t = 0.0;
            for (unsigned r = 0; r < csize; r++) t += xch[r][threadIdx.x];
            sb[threadIdx.x] = t;
        }
        __syncthreads();
        const double sum0 = sb[kMaxQ];
        // G_tilde = g - XXWX_inv b (:61), two samples per thread in flight
        for (int64_t j = lo + threadIdx.x; j < hi; j += 2 * kThreads) {
            const int64_t j2 = j + kThreads;
            const bool ok2 = j2 < hi;
            double sa = 0.0, sc = 0.0;
#pragma unroll 4
            for (int c0 = 0; c0 < q; c0++) {
                const double bc = sb[c0];
                sa += XXWX_inv[j + (int64_t)c0 * n] * bc;
                if (ok2) sc += XXWX_inv[j2 + (int64_t)c0 * n] * bc;
            }
            gt[j] -= sa;
            if (ok2) gt[j2] -= sc;
        }
        __syncthreads();

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
        if (crank == 0 && threadIdx.x == 0) {
            pvalue[v] = res;
            if (trace) {
                trace[4 * v + 0] = col.passes; trace[4 * v + 1] = nr_iter; trace[4 * v + 2] = 0;
                trace[4 * v + 3] = bits | (nr_iter << 8);
            }
        }
    }
    cluster.sync();      // no CTA exits while a peer may still address its shared memory
}

int resident_clusters(staar_ctx *ctx, int width);

// samples per variant decide the cluster width (a function of n and of the device only, so a variant's result does not
// depend on which other variants share the call).  For long columns the width is the one among 8, 7, 6 that lets
// resident clusters cover most SMs: a cluster lives inside one GPC, and on this B200 eight-wide clusters reach only 120
// of the 148 SMs (15 clusters) while six-wide ones reach 132-144 (measured: 0.408 s vs 0.384 s for 940 variants at n = 4e5).
int cluster_width(staar_ctx *ctx, int64_t n)
{
    static const int cap = getenv("STAAR_B200_SPA_CLUSTER") ? atoi(getenv("STAAR_B200_SPA_CLUSTER")) : 0;
    if (cap > 0) return std::min(cap, kMaxCluster);
    if (n < (1 << 17)) return n >= (1 << 15) ? 4 : n >= (1 << 13) ? 2 : 1;
    if (ctx->spa_best_width == 0) {               // per context (= per device): occupancy depends on the GPC layout
        int cover = -1;
        for (int w = 8; w >= 6; w--) {
            const int c = resident_clusters(ctx, w) * w;
            if (c > cover) { cover = c; ctx->spa_best_width = w; }
        }
    }
    return ctx->spa_best_width;
}

int resident_clusters(staar_ctx *ctx, int width)
{
    int *cached = ctx->spa_resident;              // per context, see cluster_width
    if (cached[width] > 0) return cached[width];
    if (width > 8 && cudaFuncSetAttribute(spa_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        set_error("SPA: cluster width %d is not available on this device", width);
        return -1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ctx->sm_count * 4 / width * width));
    cfg.blockDim = dim3(kThreads);
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)width; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int k = 0;
    if (cudaOccupancyMaxActiveClusters(&k, spa_kernel, &cfg) != cudaSuccess || k <= 0) {
        cudaGetLastError();
        k = std::max(1, ctx->sm_count / width);
    }
    cached[width] = k;
    return k;
}

}  // namespace

int spa_scratch_slots(staar_ctx *ctx, int64_t n, int64_t p)
{
    return (int)std::min<int64_t>(std::max<int64_t>(p, 1), resident_clusters(ctx, cluster_width(ctx, n)));
}

static int launch_spa_input(staar_ctx *ctx, const SpaInput &in, int64_t n, int64_t p, const double *dXW,
                            const double *dXXWX_inv, int q, const double *d_res, const double *d_mu, double tol,
                            int max_iter, double *d_gt, double *d_pvalue, long long *d_trace)
{
    if (p == 0) return STAAR_OK;
    if (q > kMaxQ) { set_error("SPA: q = %d exceeds the supported maximum of %d", q, kMaxQ); return STAAR_ERR_ARG; }
    const int width = cluster_width(ctx, n), clusters = spa_scratch_slots(ctx, n, p);
    if (clusters < 1) return STAAR_ERR_CUDA;
    STAAR_TRY(ctx->ctr.reserve(sizeof(unsigned long long)));
    unsigned long long *work = ctx->ctr.as<unsigned long long>();
    STAAR_CUDA(cudaMemsetAsync(work, 0, sizeof(unsigned long long), ctx->stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * width));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)width; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    STAAR_CUDA(cudaLaunchKernelEx(&cfg, spa_kernel, in, n, p, dXW, dXXWX_inv, q, d_res, d_mu, tol, max_iter, d_gt, d_pvalue,
                                  d_trace, work));
    ctx->launches++;
    return STAAR_OK;
}

int launch_spa(staar_ctx *ctx, const double *dG, int64_t n, int64_t p, const double *dXW, const double *dXXWX_inv,
               int q, const double *d_res, const double *d_mu, double tol, int max_iter, double *d_gt,
               double *d_pvalue, long long *d_trace)
{
    SpaInput in{dG, nullptr, 0, nullptr, nullptr, 0};
    return launch_spa_input(ctx, in, n, p, dXW, dXXWX_inv, q, d_res, d_mu, tol, max_iter, d_gt, d_pvalue, d_trace);
}

int launch_spa_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, const long long *d_variant,
                      const int32_t *d_inv_perm, int impute, int64_t n, int64_t p, const double *dXW,
                      const double *dXXWX_inv, int q, const double *d_res, const double *d_mu, double tol, int max_iter,
                      double *d_gt, double *d_pvalue, long long *d_trace)
{
    SpaInput in{nullptr, d_packed, stride, d_variant, d_inv_perm, impute};
    return launch_spa_input(ctx, in, n, p, dXW, dXXWX_inv, q, d_res, d_mu, tol, max_iter, d_gt, d_pvalue, d_trace);
}

}  // namespace staar
