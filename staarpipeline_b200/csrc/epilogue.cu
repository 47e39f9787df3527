// epilogue.cu — K5/K7: per-variant statistics from the contracted operands.
//
// One thread per variant.  Inputs are the rows of L = G'[residuals | Sigma_iX | X_adj | PX_adj] (K2) and the
// quadratic forms g_a' S g_b (K3/K4).  Reproduces, entry by entry, what the reference reads out of its p x p Cov:
//   single trait : Cov_ii = quad - t' cov t                       src/Individual_Score_Test.cpp:43,45-64
//   conditional  : ... - a' M b - b' M a + a' M (PA'A) M a         src/Individual_Score_Test_cond.cpp:49,51-72
//   k traits     : C = Cov(id,id), det(C)==0 guard, U' inv(C) U    src/Individual_Score_Test_multi.cpp:45-65
// with chi-square log-tails that stay finite far beyond 1e-308 (common.cuh: log_gamma_q).
#include "common.cuh"

namespace staar {

namespace {

constexpr int kMaxTraits = 8;

struct EpiParams {
    int64_t p_total, pp;
    int k;                 // 0: single-trait outputs; >= 1: k-trait block test
    const double *L; int ldL;
    int q; const double *cov;
    const double *quad;    // pp * kk * kk
    int m; const double *M; const double *MQM; int colA0, colPA0;
    double *Score, *Score_se, *pl, *Est, *Est_se;
};

__device__ __forceinline__ double bilinear(const double *x, const double *C, const double *y, int q)
{
    double s = 0.0;
    for (int b = 0; b < q; b++) {
        double xb = 0.0;
        for (int a = 0; a < q; a++) xb += x[a] * C[a + b * q];
        s += xb * y[b];
    }
    return s;
}

// LU with partial pivoting on a k x k column-major matrix in local memory; returns the determinant
// (0 on an exact-zero pivot) and, if inv != nullptr, the inverse.  Same algorithm as the oracle's stand-in
// for arma::det / arma::inv.
__device__ double lu_det_inv(double *A, int k, double *inv)
{
    int piv[kMaxTraits];
    int sign = 1;
    for (int c = 0; c < k; c++) {
        int pr = c; double best = fabs(A[c + c * k]);
        for (int r = c + 1; r < k; r++) {
            double v = fabs(A[r + c * k]);
            if (v > best) { best = v; pr = r; }
        }
        piv[c] = pr;
        if (pr != c) {
            for (int cc = 0; cc < k; cc++) { double t = A[c + cc * k]; A[c + cc * k] = A[pr + cc * k]; A[pr + cc * k] = t; }
            sign = -sign;
        }
        const double d = A[c + c * k];
        if (d == 0.0) return 0.0;
        for (int r = c + 1; r < k; r++) {
            const double f = A[r + c * k] / d;
            A[r + c * k] = f;
            for (int cc = c + 1; cc < k; cc++) A[r + cc * k] -= f * A[c + cc * k];
        }
    }
    double det = (double)sign;
    for (int c = 0; c < k; c++) det *= A[c + c * k];
    if (inv) {
        for (int c = 0; c < k; c++) {
            double *x = inv + c * k;
            for (int r = 0; r < k; r++) x[r] = (r == c) ? 1.0 : 0.0;
            for (int r = 0; r < k; r++) { int pr = piv[r]; if (pr != r) { double t = x[r]; x[r] = x[pr]; x[pr] = t; } }
            for (int r = 0; r < k; r++) for (int cc = 0; cc < r; cc++) x[r] -= A[r + cc * k] * x[cc];
            for (int r = k - 1; r >= 0; r--) {
                for (int cc = r + 1; cc < k; cc++) x[r] -= A[r + cc * k] * x[cc];
                x[r] /= A[r + r * k];
            }
        }
    }
    return det;
}

__global__ void __launch_bounds__(128) epilogue_kernel(EpiParams P)
{
    const int kk = P.k > 0 ? P.k : 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.pp; i += (int64_t)gridDim.x * blockDim.x) {
        double C[kMaxTraits * kMaxTraits], U[kMaxTraits];
        for (int a = 0; a < kk; a++) {
            const double *La = P.L + ((int64_t)a * P.pp + i) * P.ldL;
            U[a] = La[0];
            P.Score[(int64_t)a * P.pp + i] = U[a];
            for (int b = 0; b < kk; b++) {
                const double *Lb = P.L + ((int64_t)b * P.pp + i) * P.ldL;
                double c = P.quad[(i * kk + a) * kk + b];
                if (P.q > 0) c -= bilinear(La + 1, P.cov, Lb + 1, P.q);
                if (P.m > 0) {
                    c -= bilinear(La + P.colA0, P.M, Lb + P.colPA0, P.m);
                    c -= bilinear(La + P.colPA0, P.M, Lb + P.colA0, P.m);
                    c += bilinear(La + P.colA0, P.MQM, Lb + P.colA0, P.m);
                }
                C[a + b * kk] = c;
            }
        }
        if (P.k == 0) {
            single_trait_stats(U[0], C[0], P.Score_se[i], P.pl[i], P.Est[i], P.Est_se[i]);
        } else {
            double W[kMaxTraits * kMaxTraits], Ci[kMaxTraits * kMaxTraits];
            for (int e = 0; e < kk * kk; e++) W[e] = C[e];
            const double det = lu_det_inv(W, kk, Ci);
            if (det == 0.0) {
                P.pl[i] = 0.0;
            } else {
                const double stat = bilinear(U, Ci, U, kk);
                P.pl[i] = -log_gamma_q(0.5 * (double)kk, 0.5 * stat);
            }
        }
    }
}

}  // namespace

int launch_epilogue_general(staar_ctx *ctx, int64_t p_total, int k, const double *dL, int ldL, int q,
                            const double *d_cov, const double *d_quad, int m, const double *dM, const double *dMQM,
                            int colA0, int colPA0, double *dScore, double *dScore_se, double *dpl, double *dEst,
                            double *dEst_se)
{
    if (p_total == 0) return STAAR_OK;
    if (k > kMaxTraits) { set_error("n_pheno = %d exceeds the supported maximum of %d", k, kMaxTraits); return STAAR_ERR_ARG; }
    EpiParams P;
    P.p_total = p_total; P.k = k; P.pp = k > 0 ? p_total / k : p_total;
    P.L = dL; P.ldL = ldL; P.q = q; P.cov = d_cov; P.quad = d_quad;
    P.m = m; P.M = dM; P.MQM = dMQM; P.colA0 = colA0; P.colPA0 = colPA0;
    P.Score = dScore; P.Score_se = dScore_se; P.pl = dpl; P.Est = dEst; P.Est_se = dEst_se;
    int blocks = (int)std::min<int64_t>((P.pp + 127) / 128, (int64_t)ctx->sm_count * 8);
    epilogue_kernel<<<blocks, 128, 0, ctx->stream>>>(P);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
