// abi.cu — the extern "C" entry points declared in include/staar_b200.h.
//
// Host-side orchestration only: operand residency (content-hashed null-model cache), H2D/D2H of the per-call
// genotype block and results, and the kernel sequence of each reference function.  No arithmetic of the score
// tests happens on the host except the m x m (X_adj'X_adj)^-1 algebra of the conditional test, which the
// reference recomputes per call on n x m operands (src/Individual_Score_Test_cond.cpp:48-49) and which here
// acts on m x m matrices already reduced on the device.
#include <chrono>
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <thread>

namespace staar {
uint64_t hash_bytes(const void *data, size_t bytes, uint64_t seed);
int nullmodel_sparse_upload(staar_ctx *ctx, int64_t n, int64_t q, const double *Sv, const uint64_t *Sr,
                            const uint64_t *Sc, const double *SX, const double *cov, const double *res, uint64_t fp,
                            uint64_t res_fp, bool want_sliced);
int nullmodel_dense_upload(staar_ctx *ctx, int64_t n, const double *P, const double *res, uint64_t fp);

namespace {

struct Guard {   // every entry point runs on the context's device
    explicit Guard(staar_ctx *c) { cudaSetDevice(c->device); }
};

#define CHECK_CTX(ctx)                                           \
    if (!(ctx)) { staar::set_error("null context"); return STAAR_ERR_ARG; } \
    staar::Guard guard__(ctx)

// Large transfers from / to PAGEABLE host memory (the Rcpp boundary hands over ordinary R vectors: 4 GB of G per chunk,
// and matrix_flip_* returns as much): a plain cudaMemcpy stages them through one driver thread at ~10 GB/s.  Here host
// threads copy 32 MB slices into / out of a ring of pinned buffers while the DMA engine works on the previous slices,
// which runs at the PCIe rate.  Pinned memory and small transfers go straight to cudaMemcpyAsync.
constexpr size_t kStageBytes = 32u << 20;
constexpr size_t kStageMin = 64u << 20;

static bool is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

static int stage_init(staar_ctx *ctx)
{
    if (ctx->stage[0]) return STAAR_OK;
    for (int b = 0; b < staar_ctx::kStageBufs; b++) {
        STAAR_CUDA(cudaMallocHost(&ctx->stage[b], kStageBytes));
        STAAR_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[b], cudaEventDisableTiming));
    }
    return STAAR_OK;
}

static void parallel_memcpy(void *dst, const void *src, size_t bytes)
{
    // copy threads: the ranks of a box share its cores (torchrun sets LOCAL_WORLD_SIZE); STAAR_B200_STAGE_THREADS overrides
    static const unsigned nthreads = [] {
        if (const char *e = getenv("STAAR_B200_STAGE_THREADS")) return (unsigned)std::max(1, atoi(e));
        const unsigned ranks = getenv("LOCAL_WORLD_SIZE") ? (unsigned)std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1u;
        return std::min(16u, std::max(1u, std::thread::hardware_concurrency() / ranks));
    }();
    const size_t piece = (bytes / nthreads + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthreads; t++) {
        const size_t off = t * piece;
        if (off >= bytes) break;
        th.emplace_back([=] { memcpy(static_cast<char *>(dst) + off, static_cast<const char *>(src) + off, std::min(piece, bytes - off)); });
    }
    memcpy(dst, src, std::min(piece, bytes));
    for (auto &x : th) x.join();
}

int h2d(staar_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return STAAR_OK;
    if (bytes < kStageMin || !is_pageable(src)) {
        STAAR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return STAAR_OK;
    }
    STAAR_TRY(stage_init(ctx));
    int b = 0;
    for (size_t off = 0; off < bytes; off += kStageBytes, b = (b + 1) % staar_ctx::kStageBufs) {
        const size_t len = std::min(kStageBytes, bytes - off);
        STAAR_CUDA(cudaEventSynchronize(ctx->stage_ev[b]));                   // the DMA that last read this buffer is done
        parallel_memcpy(ctx->stage[b], static_cast<const char *>(src) + off, len);
        STAAR_CUDA(cudaMemcpyAsync(static_cast<char *>(dst) + off, ctx->stage[b], len, cudaMemcpyHostToDevice, ctx->stream));
        STAAR_CUDA(cudaEventRecord(ctx->stage_ev[b], ctx->stream));
    }
    return STAAR_OK;
}

int d2h(staar_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0 || !dst) return STAAR_OK;
    if (bytes < kStageMin || !is_pageable(dst)) {
        STAAR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return STAAR_OK;
    }
    STAAR_TRY(stage_init(ctx));
    // slices are copied out one ring position behind the DMA; on return everything is in dst
    const int nb = staar_ctx::kStageBufs;
    const size_t nslice = (bytes + kStageBytes - 1) / kStageBytes;
    for (size_t s = 0; s < nslice + (size_t)(nb - 1); s++) {
        if (s >= (size_t)(nb - 1)) {                                           // drain slice s - (nb - 1)
            const size_t k = s - (nb - 1), off = k * kStageBytes, len = std::min(kStageBytes, bytes - off);
            STAAR_CUDA(cudaEventSynchronize(ctx->stage_ev[k % nb]));
            parallel_memcpy(static_cast<char *>(dst) + off, ctx->stage[k % nb], len);
        }
        if (s < nslice) {
            const size_t off = s * kStageBytes, len = std::min(kStageBytes, bytes - off);
            STAAR_CUDA(cudaMemcpyAsync(ctx->stage[s % nb], static_cast<const char *>(src) + off, len, cudaMemcpyDeviceToHost, ctx->stream));
            STAAR_CUDA(cudaEventRecord(ctx->stage_ev[s % nb], ctx->stream));
        }
    }
    return STAAR_OK;
}
int sync(staar_ctx *ctx)
{
    STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
    return STAAR_OK;
}

// ---- null-model residency -------------------------------------------------------------------------------
int ensure_sparse(staar_ctx *ctx, int64_t n, int64_t q, const double *Sv, const uint64_t *Sr, const uint64_t *Sc,
                  const double *SX, const double *cov, const double *res)
{
    if (!Sv || !Sr || !Sc || !SX || !cov || !res) { set_error("null-model operand pointer is NULL"); return STAAR_ERR_ARG; }
    const uint64_t nnz = Sc[n];
    uint64_t fp = hash_bytes(Sv, nnz * 8, 11) ^ hash_bytes(Sr, nnz * 8, 12) ^ hash_bytes(Sc, (n + 1) * 8, 13) ^
                  hash_bytes(SX, (size_t)n * q * 8, 14) ^ hash_bytes(cov, (size_t)q * q * 8, 15) ^
                  ((uint64_t)n * 0x9E3779B97F4A7C15ull) ^ (uint64_t)q;
    const uint64_t rfp = hash_bytes(res, (size_t)n * 8, 16);
    NullSparse &ns = ctx->ns;
    if (ns.set && ns.fingerprint == fp && ns.n == n && ns.q == q) {
        if (ns.res_fingerprint != rfp) {          // same Sigma_i / Sigma_iX / cov, new residual vector
            STAAR_TRY(ctx->tmp.reserve(sizeof(double) * n));
            STAAR_TRY(h2d(ctx, ctx->tmp.p, res, sizeof(double) * n));
            STAAR_TRY(launch_set_Y_col0(ctx, n, ns.ldY, ctx->tmp.as<double>(), ns.Y));
            ns.res_fingerprint = rfp;
            ns.sliced = false;
        }
        return STAAR_OK;
    }
    return nullmodel_sparse_upload(ctx, n, q, Sv, Sr, Sc, SX, cov, res, fp, rfp, false);
}

int ensure_dense(staar_ctx *ctx, int64_t n, const double *P)
{
    if (!P) { set_error("P is NULL"); return STAAR_ERR_ARG; }
    // every byte of P is hashed (threaded, ~0.4 s for 3.2 GB — small next to the upload it saves): a P that differs in
    // a few entries only must not be served from the resident copy
    const uint64_t fp = hash_bytes(P, (size_t)n * n * sizeof(double), 21) ^ (uint64_t)n;
    if (ctx->nd.set && ctx->nd.fingerprint == fp && ctx->nd.n == n) return STAAR_OK;
    return nullmodel_dense_upload(ctx, n, P, nullptr, fp);      // new allocation: resident residuals are gone (nd.res_set = false)
}

// ---- genotype block on the device ---------------------------------------------------------------------------
int upload_dense_G(staar_ctx *ctx, const double *G, int64_t n, int64_t p)
{
    STAAR_TRY(ctx->G.reserve(sizeof(double) * (size_t)n * std::max<int64_t>(p, 1)));
    return h2d(ctx, ctx->G.p, G, sizeof(double) * (size_t)n * p);
}

int upload_csc_G(staar_ctx *ctx, const double *Gv, const uint64_t *Gr, const uint64_t *Gc, int64_t n, int64_t p)
{
    if (!Gv || !Gr || !Gc) { set_error("sparse G pointer is NULL"); return STAAR_ERR_ARG; }
    const uint64_t nnz = Gc[p];
    // the scatter kernel writes G[c n + row]: a bad index would be an out-of-bounds device write
    for (int64_t c = 0; c < p; c++)
        if (Gc[c] > Gc[c + 1]) { set_error("sparse G: column pointers are not non-decreasing"); return STAAR_ERR_ARG; }
    for (uint64_t e = 0; e < nnz; e++)
        if (Gr[e] >= (uint64_t)n) { set_error("sparse G: row index out of range"); return STAAR_ERR_ARG; }
    STAAR_TRY(ctx->G.reserve(sizeof(double) * (size_t)n * std::max<int64_t>(p, 1)));
    const size_t bytes = nnz * 16 + (size_t)(p + 1) * 8;
    STAAR_TRY(ctx->pk.reserve(bytes + 64));
    double *dval = ctx->pk.as<double>();
    int64_t *drow = reinterpret_cast<int64_t *>(dval + nnz);
    int64_t *dcp = drow + nnz;
    STAAR_TRY(h2d(ctx, dval, Gv, nnz * 8));
    STAAR_TRY(h2d(ctx, drow, Gr, nnz * 8));            // uint64 indices < 2^63: same bits as int64
    STAAR_TRY(h2d(ctx, dcp, Gc, (size_t)(p + 1) * 8));
    return launch_csc_to_dense(ctx, dval, drow, dcp, n, p, ctx->G.as<double>());
}

// ---- host LU for the m x m algebra of the conditional test ----------------------------------------------
int host_inv(std::vector<double> &A, int k, std::vector<double> &inv)
{
    std::vector<int> piv(k);
    for (int c = 0; c < k; c++) {
        int pr = c; double best = fabs(A[c + c * k]);
        for (int r = c + 1; r < k; r++) if (fabs(A[r + c * k]) > best) { best = fabs(A[r + c * k]); pr = r; }
        piv[c] = pr;
        if (pr != c) for (int cc = 0; cc < k; cc++) std::swap(A[c + cc * k], A[pr + cc * k]);
        const double d = A[c + c * k];
        if (d == 0.0) return 1;
        for (int r = c + 1; r < k; r++) {
            const double f = A[r + c * k] / d; A[r + c * k] = f;
            for (int cc = c + 1; cc < k; cc++) A[r + cc * k] -= f * A[c + cc * k];
        }
    }
    inv.assign((size_t)k * k, 0.0);
    for (int c = 0; c < k; c++) {
        double *x = inv.data() + (size_t)c * k;
        x[c] = 1.0;
        for (int r = 0; r < k; r++) if (piv[r] != r) std::swap(x[r], x[piv[r]]);
        for (int r = 0; r < k; r++) for (int cc = 0; cc < r; cc++) x[r] -= A[r + cc * k] * x[cc];
        for (int r = k - 1; r >= 0; r--) {
            for (int cc = r + 1; cc < k; cc++) x[r] -= A[r + cc * k] * x[cc];
            x[r] /= A[r + r * k];
        }
    }
    return 0;
}

struct CondOperands { int m = 0; double *dM = nullptr, *dMQM = nullptr; const double *dY = nullptr; int ldY = 0, colA0 = 0, colPA0 = 0; };

// PX_adj, (X_adj'X_adj)^-1 and M (PX_adj'X_adj) M on the device / m x m host algebra; Ycond = [res|SX|A|PA]
// X_adj == nullptr: the adjustment matrix is already in place on the device (first n*m doubles of ctx->cond, see
// staar_packed_cond_refit) and d_res holds the refit residual vector; nothing is cached across calls in that mode.
int prepare_cond(staar_ctx *ctx, const double *X_adj, int64_t m, const double *cov_host, CondOperands &co,
                 const double *d_res = nullptr)
{
    NullSparse &ns = ctx->ns;
    const int64_t n = ns.n, q = ns.q;
    const int ldc = (int)((1 + q + 2 * m + 7) / 8 * 8), ld2 = (int)((1 + 2 * m + 7) / 8 * 8);
    // aux layout: A (n*m) | PA (n*m) | Ycond (n*ldc) | Y2 (n*ld2) | T1 (q*m) | M (m*m) | MQM (m*m)
    const size_t need = sizeof(double) * ((size_t)2 * n * m + (size_t)n * ldc + (size_t)n * ld2 + q * m + 2 * m * m);
    // same X_adj, same null model, same residual vector as the previous call: everything below is still on the device
    const uint64_t key = X_adj ? hash_bytes(X_adj, sizeof(double) * (size_t)n * m, 31) ^ ns.fingerprint ^ (ns.res_fingerprint * 3) ^ (uint64_t)m : 0;
    const void *before = ctx->cond.p;
    STAAR_TRY(ctx->cond.reserve(need));
    if (ctx->cond.p != before) {
        if (!X_adj) { set_error("internal: conditional scratch moved under a device-resident X_adj"); return STAAR_ERR_STATE; }
        ctx->cond_key = 0;                                            // buffer grew: contents are gone
    }
    double *dA = ctx->cond.as<double>(), *dPA = dA + n * m, *dYc = dPA + n * m, *dY2 = dYc + (size_t)n * ldc;
    double *dT1 = dY2 + (size_t)n * ld2, *dM = dT1 + q * m, *dMQM = dM + m * m;
    if (ctx->cond_key == key && ctx->cond_m == (int)m && ctx->cond_ld == ldc && key != 0) {
        co.m = (int)m; co.dM = dM; co.dMQM = dMQM; co.dY = dYc; co.ldY = ldc; co.colA0 = (int)(1 + q); co.colPA0 = (int)(1 + q + m);
        return STAAR_OK;
    }
    ctx->cond_key = 0;
    if (X_adj) STAAR_TRY(h2d(ctx, dA, X_adj, sizeof(double) * n * m));
    // trans(Sigma_iX) * X_adj  (q x m): contraction of the m columns of A against Y = [res | SX]
    STAAR_TRY(ctx->L.reserve(sizeof(double) * (size_t)m * std::max(ns.ldY, ld2)));
    STAAR_TRY(launch_contract_dense(ctx, dA, n, m, ns.Y, ns.ldY, (int)(1 + q), ctx->L.as<double>(), ns.ldY));
    std::vector<double> LA((size_t)m * ns.ldY);
    STAAR_TRY(d2h(ctx, LA.data(), ctx->L.p, sizeof(double) * LA.size()));
    STAAR_TRY(sync(ctx));
    std::vector<double> T1((size_t)q * m);
    for (int64_t c = 0; c < m; c++)
        for (int64_t a = 0; a < q; a++) {
            double s = 0;
            for (int64_t b = 0; b < q; b++) s += cov_host[a + b * q] * LA[(size_t)c * ns.ldY + 1 + b];
            T1[a + c * q] = s;
        }
    STAAR_TRY(h2d(ctx, dT1, T1.data(), sizeof(double) * q * m));
    STAAR_TRY(launch_px_adj(ctx, n, ns.row_ptr, ns.col, ns.val, dA, (int)m, ns.SX, (int)q, dT1, dPA));   // :48
    // X_adj'X_adj and PX_adj'X_adj via one more contraction against [A | PA]
    STAAR_TRY(launch_build_Y(ctx, n, ld2, nullptr, nullptr, 0, dA, dPA, (int)m, dY2));
    STAAR_TRY(launch_contract_dense(ctx, dA, n, m, dY2, ld2, (int)(1 + 2 * m), ctx->L.as<double>(), ld2));
    std::vector<double> L2((size_t)m * ld2);
    STAAR_TRY(d2h(ctx, L2.data(), ctx->L.p, sizeof(double) * L2.size()));
    STAAR_TRY(sync(ctx));
    std::vector<double> AtA((size_t)m * m), PAtA((size_t)m * m), M, tmp((size_t)m * m), MQM((size_t)m * m);
    for (int64_t a = 0; a < m; a++)
        for (int64_t b = 0; b < m; b++) {
            AtA[a + b * m] = L2[(size_t)a * ld2 + 1 + b];          // sum_j A[j,a] A[j,b]
            PAtA[a + b * m] = L2[(size_t)b * ld2 + 1 + m + a];     // sum_j PA[j,a] A[j,b]
        }
    if (host_inv(AtA, (int)m, M) != 0) {
        set_error("inv(): matrix is singular (trans(X_adj)*X_adj)");
        return STAAR_ERR_SINGULAR;
    }
    for (int64_t a = 0; a < m; a++)
        for (int64_t b = 0; b < m; b++) {
            double s = 0;
            for (int64_t c = 0; c < m; c++) s += M[a + c * m] * PAtA[c + b * m];
            tmp[a + b * m] = s;
        }
    for (int64_t a = 0; a < m; a++)
        for (int64_t b = 0; b < m; b++) {
            double s = 0;
            for (int64_t c = 0; c < m; c++) s += tmp[a + c * m] * M[c + b * m];
            MQM[a + b * m] = s;
        }
    STAAR_TRY(h2d(ctx, dM, M.data(), sizeof(double) * m * m));
    STAAR_TRY(h2d(ctx, dMQM, MQM.data(), sizeof(double) * m * m));
    // Ycond: column 0 is the refit residual vector of THIS call (already in ns.Y column 0 via ensure_sparse)
    if (!d_res) {
        STAAR_TRY(ctx->tmp.reserve(sizeof(double) * n));
        STAAR_CUDA(cudaMemcpy2DAsync(ctx->tmp.p, sizeof(double), ns.Y, sizeof(double) * ns.ldY, sizeof(double), n,
                                     cudaMemcpyDeviceToDevice, ctx->stream));
        d_res = ctx->tmp.as<double>();
    }
    STAAR_TRY(launch_build_Y(ctx, n, ldc, d_res, ns.SX, (int)q, dA, dPA, (int)m, dYc));
    STAAR_TRY(sync(ctx));       // host vectors above go out of scope
    ctx->cond_key = key; ctx->cond_m = (int)m; ctx->cond_ld = ldc;
    co.m = (int)m; co.dM = dM; co.dMQM = dMQM; co.dY = dYc; co.ldY = ldc; co.colA0 = (int)(1 + q); co.colPA0 = (int)(1 + q + m);
    return STAAR_OK;
}

// ---- the sparse-Sigma family on a device-resident dense G ----------------------------------------------
// k == 0: single trait (5 outputs of length p); k >= 1: block test (Score length p, pvalue_log length p/k).
int run_sparse_family(staar_ctx *ctx, int64_t n, int64_t p, int k, const CondOperands *co,
                      double *Score, double *Score_se, double *pl, double *Est, double *Est_se)
{
    NullSparse &ns = ctx->ns;
    const int kk = k > 0 ? k : 1;
    if (k > 0 && p % k != 0) { set_error("number of columns of G (%lld) is not a multiple of n_pheno (%d)", (long long)p, k); return STAAR_ERR_ARG; }
    const int64_t pp = p / kk, npairs = pp * kk * kk;
    const double *dY = co ? co->dY : ns.Y;
    const int ldY = co ? co->ldY : ns.ldY;
    const int ncols = (int)(1 + ns.q + (co ? 2 * co->m : 0));
    const double *dG = ctx->G.as<double>();
    STAAR_TRY(ctx->L.reserve(sizeof(double) * (size_t)std::max<int64_t>(p, 1) * ldY));
    STAAR_TRY(launch_contract_dense(ctx, dG, n, p, dY, ldY, ncols, ctx->L.as<double>(), ldY));
    STAAR_TRY(ctx->tmp.reserve(sizeof(int32_t) * 2 * (size_t)npairs + sizeof(double) * (size_t)npairs + 64));
    double *dquad = ctx->tmp.as<double>();
    int32_t *dcolA = reinterpret_cast<int32_t *>(dquad + npairs), *dcolB = dcolA + npairs;
    STAAR_TRY(launch_fill_pairs(ctx, pp, kk, dcolA, dcolB));
    STAAR_TRY(launch_quad_pairs_csr(ctx, dG, n, ns.row_ptr, ns.col, ns.val, dcolA, dcolB, npairs, dquad));
    STAAR_TRY(ctx->out.reserve(sizeof(double) * 5 * (size_t)std::max<int64_t>(p, 1)));
    double *o = ctx->out.as<double>();
    STAAR_TRY(launch_epilogue_general(ctx, p, k, ctx->L.as<double>(), ldY, (int)ns.q, ns.cov, dquad, co ? co->m : 0,
                                      co ? co->dM : nullptr, co ? co->dMQM : nullptr, co ? co->colA0 : 0,
                                      co ? co->colPA0 : 0, o, o + p, o + 2 * p, o + 3 * p, o + 4 * p));
    STAAR_TRY(d2h(ctx, Score, o, sizeof(double) * p));
    if (k == 0) {
        STAAR_TRY(d2h(ctx, Score_se, o + p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, pl, o + 2 * p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, Est, o + 3 * p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, Est_se, o + 4 * p, sizeof(double) * p));
    } else {
        STAAR_TRY(d2h(ctx, pl, o + 2 * p, sizeof(double) * pp));
    }
    return sync(ctx);
}

// ---- the dense-P family ----------------------------------------------------------------------------------
int run_dense_family(staar_ctx *ctx, int64_t n, int64_t p, int k, const double *residuals,
                     double *Score, double *Score_se, double *pl, double *Est, double *Est_se)
{
    const int kk = k > 0 ? k : 1;
    if (k > 0 && p % k != 0) { set_error("number of columns of G (%lld) is not a multiple of n_pheno (%d)", (long long)p, k); return STAAR_ERR_ARG; }
    const int64_t pp = p / kk, npairs = pp * kk * kk;
    const double *dG = ctx->G.as<double>();
    // Uscore = trans(residuals)*G: contraction against the single column [res]
    const int ldY = 8;
    STAAR_TRY(ctx->aux.reserve(sizeof(double) * ((size_t)n * ldY + (size_t)n)));
    double *dYr = ctx->aux.as<double>(), *dres = dYr + (size_t)n * ldY;
    STAAR_TRY(h2d(ctx, dres, residuals, sizeof(double) * n));
    STAAR_TRY(launch_build_Y(ctx, n, ldY, dres, nullptr, 0, nullptr, nullptr, 0, dYr));
    STAAR_TRY(ctx->L.reserve(sizeof(double) * (size_t)std::max<int64_t>(p, 1) * ldY));
    STAAR_TRY(launch_contract_dense(ctx, dG, n, p, dYr, ldY, 1, ctx->L.as<double>(), ldY));
    // W = P*G, Cov entries = g_a' W[:,b]   (Individual_Score_Test_denseGRM.cpp:37)
    STAAR_TRY(ctx->tmp2.reserve(sizeof(double) * (size_t)n * std::max<int64_t>(p, 1)));
    // note: launch_contract_dense may have used tmp2 for split-K partials; they are consumed by now (stream order)
    STAAR_TRY(launch_dense_P_times_G(ctx, ctx->nd.P, dG, n, p, ctx->tmp2.as<double>()));
    STAAR_TRY(ctx->tmp.reserve(sizeof(int32_t) * 2 * (size_t)npairs + sizeof(double) * (size_t)npairs + 64));
    double *dquad = ctx->tmp.as<double>();
    int32_t *dcolA = reinterpret_cast<int32_t *>(dquad + npairs), *dcolB = dcolA + npairs;
    STAAR_TRY(launch_fill_pairs(ctx, pp, kk, dcolA, dcolB));
    STAAR_TRY(launch_dot_pairs(ctx, dG, ctx->tmp2.as<double>(), n, dcolA, dcolB, npairs, dquad));
    STAAR_TRY(ctx->out.reserve(sizeof(double) * 5 * (size_t)std::max<int64_t>(p, 1)));
    double *o = ctx->out.as<double>();
    STAAR_TRY(launch_epilogue_general(ctx, p, k, ctx->L.as<double>(), ldY, 0, nullptr, dquad, 0, nullptr, nullptr, 0, 0,
                                      o, o + p, o + 2 * p, o + 3 * p, o + 4 * p));
    STAAR_TRY(d2h(ctx, Score, o, sizeof(double) * p));
    if (k == 0) {
        STAAR_TRY(d2h(ctx, Score_se, o + p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, pl, o + 2 * p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, Est, o + 3 * p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, Est_se, o + 4 * p, sizeof(double) * p));
    } else {
        STAAR_TRY(d2h(ctx, pl, o + 2 * p, sizeof(double) * pp));
    }
    return sync(ctx);
}

// Sparse genotype columns against the dense P: carriers only (quadform.cu).  Taken when no column has more than n/8
// stored entries — the R driver sends MAF < 0.05 here, i.e. <= 0.0975 n carriers; denser input goes through P*G.
bool csc_is_sparse_enough(const uint64_t *Gr, const uint64_t *Gc, int64_t n, int64_t p)
{
    if (!Gr || !Gc) return false;
    uint64_t worst = 0;
    for (int64_t j = 0; j < p; j++) {
        worst = std::max<uint64_t>(worst, Gc[j + 1] - Gc[j]);
        for (uint64_t e = Gc[j] + 1; e < Gc[j + 1]; e++)
            if (Gr[e] <= Gr[e - 1]) return false;          // unsorted / duplicate rows (not an arma::sp_mat): dense route
    }
    return worst <= (uint64_t)(n / 8);
}

// upload of a sparse genotype block as it is (values | int64 rows | int64 column pointers) into ctx->pk
int upload_csc_carriers(staar_ctx *ctx, const double *Gv, const uint64_t *Gr, const uint64_t *Gc, int64_t n, int64_t p,
                        double **dval, int64_t **drow, int64_t **dcp)
{
    if (!Gv || !Gr || !Gc) { set_error("sparse G pointer is NULL"); return STAAR_ERR_ARG; }
    const uint64_t nnz = Gc[p];
    for (uint64_t e = 0; e < nnz; e++)
        if (Gr[e] >= (uint64_t)n) { set_error("sparse G: row index out of range"); return STAAR_ERR_ARG; }
    STAAR_TRY(ctx->pk.reserve(nnz * 16 + (size_t)(p + 1) * 8 + 64));
    *dval = ctx->pk.as<double>();
    *drow = reinterpret_cast<int64_t *>(*dval + nnz);
    *dcp = *drow + nnz;
    STAAR_TRY(h2d(ctx, *dval, Gv, nnz * 8));
    STAAR_TRY(h2d(ctx, *drow, Gr, nnz * 8));            // uint64 indices < 2^63: same bits as int64
    STAAR_TRY(h2d(ctx, *dcp, Gc, (size_t)(p + 1) * 8));
    return STAAR_OK;
}

// Individual_Score_Test_sp / _sp_multi on sparse columns: contraction and quadratic forms over the stored entries
int run_sparse_family_carriers(staar_ctx *ctx, const double *Gv, const uint64_t *Gr, const uint64_t *Gc, int64_t n, int64_t p,
                               int k, double *Score, double *Score_se, double *pl, double *Est, double *Est_se)
{
    NullSparse &ns = ctx->ns;
    const int kk = k > 0 ? k : 1;
    if (k > 0 && p % k != 0) { set_error("number of columns of G (%lld) is not a multiple of n_pheno (%d)", (long long)p, k); return STAAR_ERR_ARG; }
    const int64_t pp = p / kk, npairs = pp * kk * kk;
    double *dval; int64_t *drow, *dcp;
    STAAR_TRY(upload_csc_carriers(ctx, Gv, Gr, Gc, n, p, &dval, &drow, &dcp));
    const int ldY = ns.ldY, ncols = (int)(1 + ns.q);
    STAAR_TRY(ctx->L.reserve(sizeof(double) * (size_t)std::max<int64_t>(p, 1) * ldY));
    STAAR_CUDA(cudaMemsetAsync(ctx->L.p, 0, sizeof(double) * (size_t)std::max<int64_t>(p, 1) * ldY, ctx->stream));
    STAAR_TRY(launch_contract_carriers(ctx, dval, drow, dcp, p, ns.Y, ldY, ncols, ctx->L.as<double>(), ldY));
    STAAR_TRY(ctx->tmp.reserve(sizeof(int32_t) * 2 * (size_t)npairs + sizeof(double) * (size_t)npairs + 64));
    double *dquad = ctx->tmp.as<double>();
    int32_t *dcolA = reinterpret_cast<int32_t *>(dquad + npairs), *dcolB = dcolA + npairs;
    STAAR_TRY(launch_fill_pairs(ctx, pp, kk, dcolA, dcolB));
    STAAR_TRY(launch_quad_pairs_carriers_csr(ctx, dval, drow, dcp, ns.row_ptr, ns.col, ns.val, dcolA, dcolB, npairs, dquad));
    STAAR_TRY(ctx->out.reserve(sizeof(double) * 5 * (size_t)std::max<int64_t>(p, 1)));
    double *o = ctx->out.as<double>();
    STAAR_TRY(launch_epilogue_general(ctx, p, k, ctx->L.as<double>(), ldY, (int)ns.q, ns.cov, dquad, 0, nullptr, nullptr, 0, 0,
                                      o, o + p, o + 2 * p, o + 3 * p, o + 4 * p));
    STAAR_TRY(d2h(ctx, Score, o, sizeof(double) * p));
    if (k == 0) {
        STAAR_TRY(d2h(ctx, Score_se, o + p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, pl, o + 2 * p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, Est, o + 3 * p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, Est_se, o + 4 * p, sizeof(double) * p));
    } else {
        STAAR_TRY(d2h(ctx, pl, o + 2 * p, sizeof(double) * pp));
    }
    return sync(ctx);
}

int run_dense_family_carriers(staar_ctx *ctx, const double *Gv, const uint64_t *Gr, const uint64_t *Gc, int64_t n, int64_t p,
                              int k, const double *residuals, double *Score, double *Score_se, double *pl, double *Est,
                              double *Est_se)
{
    const int kk = k > 0 ? k : 1;
    if (k > 0 && p % k != 0) { set_error("number of columns of G (%lld) is not a multiple of n_pheno (%d)", (long long)p, k); return STAAR_ERR_ARG; }
    const int64_t pp = p / kk, npairs = pp * kk * kk;
    double *dval; int64_t *drow, *dcp;
    STAAR_TRY(upload_csc_carriers(ctx, Gv, Gr, Gc, n, p, &dval, &drow, &dcp));
    const int ldY = 8;
    STAAR_TRY(ctx->aux.reserve(sizeof(double) * (size_t)n));
    double *dres = ctx->aux.as<double>();
    STAAR_TRY(h2d(ctx, dres, residuals, sizeof(double) * n));
    STAAR_TRY(ctx->L.reserve(sizeof(double) * (size_t)std::max<int64_t>(p, 1) * ldY));
    STAAR_CUDA(cudaMemsetAsync(ctx->L.p, 0, sizeof(double) * (size_t)std::max<int64_t>(p, 1) * ldY, ctx->stream));
    STAAR_TRY(launch_score_carriers(ctx, dval, drow, dcp, p, dres, ctx->L.as<double>(), ldY));
    STAAR_TRY(ctx->tmp.reserve(sizeof(int32_t) * 2 * (size_t)npairs + sizeof(double) * (size_t)npairs + 64));
    double *dquad = ctx->tmp.as<double>();
    int32_t *dcolA = reinterpret_cast<int32_t *>(dquad + npairs), *dcolB = dcolA + npairs;
    STAAR_TRY(launch_fill_pairs(ctx, pp, kk, dcolA, dcolB));
    STAAR_TRY(launch_quad_pairs_carriers_P(ctx, dval, drow, dcp, n, ctx->nd.P, dcolA, dcolB, npairs, dquad));
    STAAR_TRY(ctx->out.reserve(sizeof(double) * 5 * (size_t)std::max<int64_t>(p, 1)));
    double *o = ctx->out.as<double>();
    STAAR_TRY(launch_epilogue_general(ctx, p, k, ctx->L.as<double>(), ldY, 0, nullptr, dquad, 0, nullptr, nullptr, 0, 0,
                                      o, o + p, o + 2 * p, o + 3 * p, o + 4 * p));
    STAAR_TRY(d2h(ctx, Score, o, sizeof(double) * p));
    if (k == 0) {
        STAAR_TRY(d2h(ctx, Score_se, o + p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, pl, o + 2 * p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, Est, o + 3 * p, sizeof(double) * p));
        STAAR_TRY(d2h(ctx, Est_se, o + 4 * p, sizeof(double) * p));
    } else {
        STAAR_TRY(d2h(ctx, pl, o + 2 * p, sizeof(double) * pp));
    }
    return sync(ctx);
}

int check_dims(int64_t n, int64_t p)
{
    if (n <= 0 || p < 0) { set_error("bad dimensions n=%lld p=%lld", (long long)n, (long long)p); return STAAR_ERR_ARG; }
    return STAAR_OK;
}

// matrix_flip_* of a large block in pageable host memory: the matrix goes to the device, through the flip kernel and
// back in column blocks of one staging buffer each, the upload of block b+1 and the download of block b running at
// the same time (PCIe is full duplex; two pinned rings, two streams), host threads filling and draining the rings.
static int flip_entry_pipelined(staar_ctx *ctx, const double *G, int64_t n, int64_t p, bool minor, double *Geno_out,
                                double *dAF, double *dMAF)
{
    STAAR_TRY(stage_init(ctx));
    const int nbuf = staar_ctx::kStageBufs;
    if (!ctx->stage_out[0])
        for (int b = 0; b < nbuf; b++) {
            STAAR_CUDA(cudaMallocHost(&ctx->stage_out[b], kStageBytes));
            STAAR_CUDA(cudaEventCreateWithFlags(&ctx->stage_out_ev[b], cudaEventDisableTiming));
            STAAR_CUDA(cudaEventCreateWithFlags(&ctx->stage_k_ev[b], cudaEventDisableTiming));
        }
    if (!ctx->stream2) STAAR_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    const int64_t cols = std::max<int64_t>(1, (int64_t)(kStageBytes / (sizeof(double) * (size_t)n)));
    const int64_t nblk = (p + cols - 1) / cols;
    const int lag = 2;
    double *dG = ctx->G.as<double>();
    for (int64_t b = 0; b < nblk + lag; b++) {
        if (b < nblk) {
            const int s = (int)(b % nbuf);
            const int64_t c0 = b * cols, nc = std::min(cols, p - c0);
            const size_t off = (size_t)c0 * n, len = sizeof(double) * (size_t)nc * n;
            STAAR_CUDA(cudaEventSynchronize(ctx->stage_ev[s]));
            parallel_memcpy(ctx->stage[s], G + off, len);
            STAAR_CUDA(cudaMemcpyAsync(dG + off, ctx->stage[s], len, cudaMemcpyHostToDevice, ctx->stream));
            STAAR_CUDA(cudaEventRecord(ctx->stage_ev[s], ctx->stream));
            STAAR_TRY(launch_flip_dense(ctx, dG + off, n, nc, minor, true, dAF + c0, dMAF + c0));
            STAAR_CUDA(cudaEventRecord(ctx->stage_k_ev[s], ctx->stream));
            STAAR_CUDA(cudaStreamWaitEvent(ctx->stream2, ctx->stage_k_ev[s], 0));
            STAAR_CUDA(cudaMemcpyAsync(ctx->stage_out[s], dG + off, len, cudaMemcpyDeviceToHost, ctx->stream2));
            STAAR_CUDA(cudaEventRecord(ctx->stage_out_ev[s], ctx->stream2));
        }
        if (b >= lag) {
            const int64_t k = b - lag, c0 = k * cols, nc = std::min(cols, p - c0);
            STAAR_CUDA(cudaEventSynchronize(ctx->stage_out_ev[k % nbuf]));
            parallel_memcpy(Geno_out + (size_t)c0 * n, ctx->stage_out[k % nbuf], sizeof(double) * (size_t)nc * n);
        }
    }
    return STAAR_OK;
}

int flip_entry(staar_ctx *ctx, const double *G, int64_t n, int64_t p, bool minor, double *Geno_out, double *AF, double *MAF)
{
    STAAR_TRY(check_dims(n, p));
    if (!G || !AF || !MAF) { set_error("matrix_flip: NULL pointer"); return STAAR_ERR_ARG; }
    STAAR_TRY(ctx->out.reserve(sizeof(double) * 2 * (size_t)std::max<int64_t>(p, 1)));
    double *o = ctx->out.as<double>();
    const size_t bytes = sizeof(double) * (size_t)n * p;
    if (Geno_out && bytes >= kStageMin && sizeof(double) * (size_t)n <= kStageBytes && is_pageable(G) && is_pageable(Geno_out)) {
        STAAR_TRY(ctx->G.reserve(bytes));
        STAAR_TRY(flip_entry_pipelined(ctx, G, n, p, minor, Geno_out, o, o + p));
    } else {
        STAAR_TRY(upload_dense_G(ctx, G, n, p));
        STAAR_TRY(launch_flip_dense(ctx, ctx->G.as<double>(), n, p, minor, Geno_out != nullptr, o, o + p));
        STAAR_TRY(d2h(ctx, Geno_out, ctx->G.p, bytes));
    }
    STAAR_TRY(d2h(ctx, AF, o, sizeof(double) * p));
    STAAR_TRY(d2h(ctx, MAF, o + p, sizeof(double) * p));
    return sync(ctx);
}

}  // namespace

// content-hashed residency of the sparse null model, for multi.cu
int ensure_sparse_shared(staar_ctx *ctx, int64_t n, int64_t q, const double *Sv, const uint64_t *Sr, const uint64_t *Sc,
                         const double *SX, const double *cov, const double *res)
{
    Guard g(ctx);
    return ensure_sparse(ctx, n, q, Sv, Sr, Sc, SX, cov, res);
}
}  // namespace staar

using namespace staar;

extern "C" {

int staar_matrix_flip_mean(staar_ctx *ctx, const double *G, int64_t n, int64_t p, double *Geno_out, double *AF, double *MAF)
{
    CHECK_CTX(ctx);
    return flip_entry(ctx, G, n, p, false, Geno_out, AF, MAF);
}
int staar_matrix_flip_minor(staar_ctx *ctx, const double *G, int64_t n, int64_t p, double *Geno_out, double *AF, double *MAF)
{
    CHECK_CTX(ctx);
    return flip_entry(ctx, G, n, p, true, Geno_out, AF, MAF);
}

int staar_individual_score_test(staar_ctx *ctx, const double *G, int64_t n, int64_t p, const double *Sv,
                                const uint64_t *Sr, const uint64_t *Sc, const double *SX, int64_t q, const double *cov,
                                const double *res, double *Score, double *Score_se, double *pl, double *Est, double *Est_se)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    STAAR_TRY(ensure_sparse(ctx, n, q, Sv, Sr, Sc, SX, cov, res));
    STAAR_TRY(upload_dense_G(ctx, G, n, p));
    return run_sparse_family(ctx, n, p, 0, nullptr, Score, Score_se, pl, Est, Est_se);
}

int staar_individual_score_test_sp(staar_ctx *ctx, const double *Gv, const uint64_t *Gr, const uint64_t *Gc, int64_t n,
                                   int64_t p, const double *Sv, const uint64_t *Sr, const uint64_t *Sc, const double *SX,
                                   int64_t q, const double *cov, const double *res, double *Score, double *Score_se,
                                   double *pl, double *Est, double *Est_se)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    STAAR_TRY(ensure_sparse(ctx, n, q, Sv, Sr, Sc, SX, cov, res));
    if (csc_is_sparse_enough(Gr, Gc, n, p))
        return run_sparse_family_carriers(ctx, Gv, Gr, Gc, n, p, 0, Score, Score_se, pl, Est, Est_se);
    STAAR_TRY(upload_csc_G(ctx, Gv, Gr, Gc, n, p));
    return run_sparse_family(ctx, n, p, 0, nullptr, Score, Score_se, pl, Est, Est_se);
}

int staar_individual_score_test_denseGRM(staar_ctx *ctx, const double *G, int64_t n, int64_t p, const double *P,
                                         const double *res, double *Score, double *Score_se, double *pl, double *Est,
                                         double *Est_se)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    STAAR_TRY(ensure_dense(ctx, n, P));
    STAAR_TRY(upload_dense_G(ctx, G, n, p));
    return run_dense_family(ctx, n, p, 0, res, Score, Score_se, pl, Est, Est_se);
}

int staar_individual_score_test_sp_denseGRM(staar_ctx *ctx, const double *Gv, const uint64_t *Gr, const uint64_t *Gc,
                                            int64_t n, int64_t p, const double *P, const double *res, double *Score,
                                            double *Score_se, double *pl, double *Est, double *Est_se)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    STAAR_TRY(ensure_dense(ctx, n, P));
    if (csc_is_sparse_enough(Gr, Gc, n, p))
        return run_dense_family_carriers(ctx, Gv, Gr, Gc, n, p, 0, res, Score, Score_se, pl, Est, Est_se);
    STAAR_TRY(upload_csc_G(ctx, Gv, Gr, Gc, n, p));
    return run_dense_family(ctx, n, p, 0, res, Score, Score_se, pl, Est, Est_se);
}

int staar_individual_score_test_multi(staar_ctx *ctx, const double *G, int64_t n, int64_t p, const double *Sv,
                                      const uint64_t *Sr, const uint64_t *Sc, const double *SX, int64_t q,
                                      const double *cov, const double *res, int n_pheno, double *Score, double *pl)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (n_pheno < 1) { set_error("n_pheno must be >= 1"); return STAAR_ERR_ARG; }
    STAAR_TRY(ensure_sparse(ctx, n, q, Sv, Sr, Sc, SX, cov, res));
    STAAR_TRY(upload_dense_G(ctx, G, n, p));
    return run_sparse_family(ctx, n, p, n_pheno, nullptr, Score, nullptr, pl, nullptr, nullptr);
}

int staar_individual_score_test_sp_multi(staar_ctx *ctx, const double *Gv, const uint64_t *Gr, const uint64_t *Gc,
                                         int64_t n, int64_t p, const double *Sv, const uint64_t *Sr, const uint64_t *Sc,
                                         const double *SX, int64_t q, const double *cov, const double *res, int n_pheno,
                                         double *Score, double *pl)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (n_pheno < 1) { set_error("n_pheno must be >= 1"); return STAAR_ERR_ARG; }
    STAAR_TRY(ensure_sparse(ctx, n, q, Sv, Sr, Sc, SX, cov, res));
    if (csc_is_sparse_enough(Gr, Gc, n, p))
        return run_sparse_family_carriers(ctx, Gv, Gr, Gc, n, p, n_pheno, Score, nullptr, pl, nullptr, nullptr);
    STAAR_TRY(upload_csc_G(ctx, Gv, Gr, Gc, n, p));
    return run_sparse_family(ctx, n, p, n_pheno, nullptr, Score, nullptr, pl, nullptr, nullptr);
}

int staar_individual_score_test_denseGRM_multi(staar_ctx *ctx, const double *G, int64_t n, int64_t p, const double *P,
                                               const double *res, int n_pheno, double *Score, double *pl)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (n_pheno < 1) { set_error("n_pheno must be >= 1"); return STAAR_ERR_ARG; }
    STAAR_TRY(ensure_dense(ctx, n, P));
    STAAR_TRY(upload_dense_G(ctx, G, n, p));
    return run_dense_family(ctx, n, p, n_pheno, res, Score, nullptr, pl, nullptr, nullptr);
}

int staar_individual_score_test_sp_denseGRM_multi(staar_ctx *ctx, const double *Gv, const uint64_t *Gr,
                                                  const uint64_t *Gc, int64_t n, int64_t p, const double *P,
                                                  const double *res, int n_pheno, double *Score, double *pl)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (n_pheno < 1) { set_error("n_pheno must be >= 1"); return STAAR_ERR_ARG; }
    STAAR_TRY(ensure_dense(ctx, n, P));
    if (csc_is_sparse_enough(Gr, Gc, n, p))
        return run_dense_family_carriers(ctx, Gv, Gr, Gc, n, p, n_pheno, res, Score, nullptr, pl, nullptr, nullptr);
    STAAR_TRY(upload_csc_G(ctx, Gv, Gr, Gc, n, p));
    return run_dense_family(ctx, n, p, n_pheno, res, Score, nullptr, pl, nullptr, nullptr);
}

int staar_individual_score_test_cond(staar_ctx *ctx, const double *G, int64_t n, int64_t p, const double *Sv,
                                     const uint64_t *Sr, const uint64_t *Sc, const double *SX, int64_t q,
                                     const double *cov, const double *X_adj, int64_t m, const double *res,
                                     double *Score, double *Score_se, double *pl, double *Est, double *Est_se)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (!X_adj || m < 1) { set_error("X_adj must have at least one column"); return STAAR_ERR_ARG; }
    STAAR_TRY(ensure_sparse(ctx, n, q, Sv, Sr, Sc, SX, cov, res));
    CondOperands co;
    STAAR_TRY(prepare_cond(ctx, X_adj, m, cov, co));
    STAAR_TRY(upload_dense_G(ctx, G, n, p));
    return run_sparse_family(ctx, n, p, 0, &co, Score, Score_se, pl, Est, Est_se);
}

int staar_individual_score_test_cond_multi(staar_ctx *ctx, const double *G, int64_t n, int64_t p, const double *Sv,
                                           const uint64_t *Sr, const uint64_t *Sc, const double *SX, int64_t q,
                                           const double *cov, const double *X_adj, int64_t m, const double *res,
                                           int n_pheno, double *Score, double *pl)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (!X_adj || m < 1) { set_error("X_adj must have at least one column"); return STAAR_ERR_ARG; }
    if (n_pheno < 1) { set_error("n_pheno must be >= 1"); return STAAR_ERR_ARG; }
    STAAR_TRY(ensure_sparse(ctx, n, q, Sv, Sr, Sc, SX, cov, res));
    CondOperands co;
    STAAR_TRY(prepare_cond(ctx, X_adj, m, cov, co));
    STAAR_TRY(upload_dense_G(ctx, G, n, p));
    return run_sparse_family(ctx, n, p, n_pheno, &co, Score, nullptr, pl, nullptr, nullptr);
}

int staar_individual_score_test_SPA(staar_ctx *ctx, const double *G, int64_t n, int64_t p, const double *XW,
                                    const double *XXWX_inv, int64_t q, const double *res, const double *muhat,
                                    double tol, int max_iter, double *pvalue, int64_t *trace)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (!G || !XW || !XXWX_inv || !res || !muhat || !pvalue) { set_error("SPA: NULL pointer"); return STAAR_ERR_ARG; }
    STAAR_TRY(upload_dense_G(ctx, G, n, p));
    const int slots = spa_scratch_slots(ctx, n, p);
    if (slots < 1) return STAAR_ERR_CUDA;
    // aux: XW (q*n) | XXWX_inv (n*q) | res (n) | mu (n) | gt scratch (slots*n)
    STAAR_TRY(ctx->aux.reserve(sizeof(double) * ((size_t)2 * n * q + 2 * (size_t)n + (size_t)slots * n)));
    double *dXW = ctx->aux.as<double>(), *dXX = dXW + n * q, *dres = dXX + n * q, *dmu = dres + n, *dgt = dmu + n;
    STAAR_TRY(h2d(ctx, dXW, XW, sizeof(double) * n * q));
    STAAR_TRY(h2d(ctx, dXX, XXWX_inv, sizeof(double) * n * q));
    STAAR_TRY(h2d(ctx, dres, res, sizeof(double) * n));
    STAAR_TRY(h2d(ctx, dmu, muhat, sizeof(double) * n));
    STAAR_TRY(ctx->out.reserve(sizeof(double) * (size_t)std::max<int64_t>(p, 1) + sizeof(long long) * 4 * (size_t)std::max<int64_t>(p, 1)));
    double *dpv = ctx->out.as<double>();
    long long *dtr = reinterpret_cast<long long *>(dpv + std::max<int64_t>(p, 1));
    STAAR_TRY(launch_spa(ctx, ctx->G.as<double>(), n, p, dXW, dXX, (int)q, dres, dmu, tol, max_iter, dgt, dpv,
                         trace ? dtr : nullptr));
    STAAR_TRY(d2h(ctx, pvalue, dpv, sizeof(double) * p));
    if (trace) STAAR_TRY(d2h(ctx, trace, dtr, sizeof(long long) * 4 * p));
    return sync(ctx);
}

// ============================================================================================ packed fast path
size_t staar_packed_stride(int64_t n) { return (size_t)(((n + 3) / 4 + 15) / 16 * 16); }

int staar_nullmodel_set_sparse(staar_ctx *ctx, int64_t n, int64_t q, const double *Sv, const uint64_t *Sr,
                               const uint64_t *Sc, const double *SX, const double *cov, const double *res)
{
    CHECK_CTX(ctx);
    STAAR_TRY(ensure_sparse(ctx, n, q, Sv, Sr, Sc, SX, cov, res));
    if (!ctx->ns.sliced) STAAR_TRY(build_sliced_operand(ctx, nullptr));
    return STAAR_OK;
}

int staar_nullmodel_set_sparse_device(staar_ctx *ctx, int64_t n, int64_t q, const int64_t *d_row_ptr,
                                      const int32_t *d_col, const double *d_val, const double *d_SX,
                                      const double *d_cov, const double *d_res)
{
    CHECK_CTX(ctx);
    // Operands arrive in device memory the caller owns (filled by an NCCL broadcast).  The set-up work (diagonal,
    // kinship flags, row-major Y, INT8 slicing) is a once-per-run host pass, so stage them through the host path.
    if (n <= 0 || q < 0) { set_error("bad dimensions"); return STAAR_ERR_ARG; }
    std::vector<int64_t> rp(n + 1);
    STAAR_CUDA(cudaMemcpy(rp.data(), d_row_ptr, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost));
    const int64_t nnz = rp[n];
    std::vector<int32_t> col(nnz); std::vector<double> val(nnz), SX((size_t)n * q), cov((size_t)q * q), res(n);
    STAAR_CUDA(cudaMemcpy(col.data(), d_col, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost));
    STAAR_CUDA(cudaMemcpy(val.data(), d_val, sizeof(double) * nnz, cudaMemcpyDeviceToHost));
    STAAR_CUDA(cudaMemcpy(SX.data(), d_SX, sizeof(double) * n * q, cudaMemcpyDeviceToHost));
    STAAR_CUDA(cudaMemcpy(cov.data(), d_cov, sizeof(double) * q * q, cudaMemcpyDeviceToHost));
    STAAR_CUDA(cudaMemcpy(res.data(), d_res, sizeof(double) * n, cudaMemcpyDeviceToHost));
    // CSR of a symmetric matrix is its CSC
    std::vector<uint64_t> Sr(nnz), Sc(n + 1);
    for (int64_t e = 0; e < nnz; e++) Sr[e] = (uint64_t)col[e];
    for (int64_t r = 0; r <= n; r++) Sc[r] = (uint64_t)rp[r];
    return staar_nullmodel_set_sparse(ctx, n, q, val.data(), Sr.data(), Sc.data(), SX.data(), cov.data(), res.data());
}

int staar_packed_flip_stats_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p,
                                   int impute, double *d_AF, double *d_MAF, int32_t *d_flip, int32_t *d_nmiss)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (stride % 16 != 0 || stride < (size_t)((n + 3) / 4)) { set_error("packed stride must be a multiple of 16 and >= ceil(n/4)"); return STAAR_ERR_ARG; }
    STAAR_TRY(launch_flip_stats_packed(ctx, d_packed, stride, n, p, impute, d_AF, d_MAF, d_flip, d_nmiss));
    return sync(ctx);
}

int staar_packed_score_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                              double *d_AF, double *d_MAF, double *d_Score, double *d_Score_se, double *d_pl,
                              double *d_Est, double *d_Est_se)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    NullSparse &ns = ctx->ns;
    if (!ns.set) { set_error("packed score: call staar_nullmodel_set_sparse first"); return STAAR_ERR_STATE; }
    if (ns.n != n) { set_error("packed score: n = %lld does not match the null model (n = %lld)", (long long)n, (long long)ns.n); return STAAR_ERR_ARG; }
    if (stride % 16 != 0 || stride < (size_t)((n + 3) / 4)) { set_error("packed stride must be a multiple of 16 and >= ceil(n/4)"); return STAAR_ERR_ARG; }
    if (!ns.sliced) STAAR_TRY(build_sliced_operand(ctx, nullptr));
    if (!ns.sliced) { set_error("packed score supports at most %d fixed-effect columns (q = %lld)", kColsPerSlice - 3, (long long)ns.q); return STAAR_ERR_ARG; }
    if (p == 0) return STAAR_OK;
    // scratch: flip (p i32) | mu (p) | quad (p) | Sm (16p) | hi (16p i64) | lo (16p i64)
    const size_t pi = (size_t)((p + 1) / 2 * 2);
    STAAR_TRY(ctx->tmp.reserve(sizeof(int32_t) * pi + sizeof(double) * (2 * (size_t)p + 16 * (size_t)p) + sizeof(long long) * 32 * (size_t)p));
    int32_t *dflip = ctx->tmp.as<int32_t>();
    double *dmu = reinterpret_cast<double *>(dflip + pi), *dquad = dmu + p, *dSm = dquad + p;
    long long *dhi = reinterpret_cast<long long *>(dSm + 16 * p), *dlo = dhi + 16 * p;
    const bool prof = ctx->profiling;
    ctx->last_fused = false;
    if (ctx->packed_path == 0 && !ctx->use_mma_sync && fused_path_available(ctx)) {
        // single pass: fused contraction + scanners -> warp-per-variant finaliser -> (rare) redo of the variants whose
        // orientation guess was wrong or whose missing list overflowed, through the block-per-variant scan.
        // ev[0..3] bracket: fused kernel | finaliser | redo.
        const int msub = fused_missing_sub(n);
        STAAR_TRY(ctx->fscan.reserve(sizeof(FusedScanOut) * (size_t)p));
        STAAR_TRY(ctx->mlist.reserve(sizeof(uint2) * 16 * (size_t)msub * (size_t)p));
        STAAR_TRY(ctx->msubcnt.reserve(sizeof(uint16_t) * 16 * (size_t)p));
        uint16_t *dmc = ctx->msubcnt.as<uint16_t>();
        STAAR_TRY(ctx->redo.reserve(sizeof(long long) * (size_t)p + 2 * sizeof(unsigned long long)));
        FusedScanOut *dscan = ctx->fscan.as<FusedScanOut>();
        uint2 *dml = ctx->mlist.as<uint2>();
        unsigned long long *dcount = ctx->redo.as<unsigned long long>();
        long long *dredo = reinterpret_cast<long long *>(dcount + 2);
        STAAR_CUDA(cudaMemsetAsync(dcount, 0, 2 * sizeof(unsigned long long), ctx->stream));
        if (prof) STAAR_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
        STAAR_TRY(launch_fused_tc(ctx, d_packed, stride, n, p, dhi, dlo, dscan, dml, msub, dmc));
        if (prof) STAAR_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
        STAAR_TRY(launch_fused_finalize(ctx, n, p, impute, dhi, dlo, dscan, dml, msub, dmc, d_AF, d_MAF, d_Score, d_Score_se, d_pl,
                                        d_Est, d_Est_se, dredo, dcount));
        if (prof) STAAR_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
        if (stride + 8192 <= 227 * 1024) {
            STAAR_TRY(launch_scan_packed(ctx, d_packed, stride, n, p, impute, dlo, d_AF, d_MAF, dflip, dmu, dquad, dSm, dredo, dcount));
            STAAR_TRY(launch_finalize_packed(ctx, p, dflip, dmu, dquad, dSm, dhi, dlo, d_Score, d_Score_se, d_pl, d_Est, d_Est_se,
                                             dredo, dcount));
        } else {
            // columns too long for the block-per-variant kernel's shared memory: nothing may need it
            unsigned long long pending = 0;
            STAAR_CUDA(cudaMemcpyAsync(&pending, dcount, sizeof pending, cudaMemcpyDeviceToHost, ctx->stream));
            STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
            if (pending) {
                set_error("packed score: %llu variants need the block-per-variant scan, but a column of %zu bytes does not fit "
                          "in shared memory (n > ~9e5): use the FP64 entry points", pending, stride);
                return STAAR_ERR_ARG;
            }
        }
        if (prof) STAAR_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
        ctx->last_fused = true;
        return STAAR_OK;
    }
    // two passes: contraction (raw codes; also yields the exact sum of codes per variant) -> scan (flip decision from
    // that sum and its own missing count, sparse terms) -> finaliser.  ev[0..3] bracket the three kernels.
    if (prof) STAAR_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    // The tcgen05 contraction also lists the missing samples of every variant (its expander threads see every word and
    // have slack), so the scan knows the flip up front and runs without a detection pass; variants whose lists were
    // truncated (missing rate far above the list capacity) go through the scan's own detection path afterwards.
    // (short columns of unrelated samples are the one case where the listing costs more than the detection it saves:
    // n = 1e4, no kinship: 5.79 vs 5.42 ms per 2 M variants)
    bool lists = ctx->use_missing_lists && !ctx->use_mma_sync && !ctx->use_tc_pair && (n >= 16384 || ctx->ns.kin_words > 0);
    const int mcap = (int)std::min<int64_t>(2048, std::max<int64_t>(96, ((n / 1000 + 64) + 31) / 32 * 32)), pcap = 2 * mcap;
    if (lists) {
        // list scratch: 16 mcap + 4 pcap + 16 bytes per variant (3.9 KB at n = 1e5).  If the device cannot hold it next to
        // the caller's columns, the call still succeeds: the scan detects the missing samples itself.
        if (ctx->mlist.reserve(sizeof(uint2) * 2 * (size_t)mcap * (size_t)p) != STAAR_OK ||
            ctx->plist.reserve(sizeof(uint32_t) * (size_t)pcap * (size_t)p) != STAAR_OK ||
            ctx->msubcnt.reserve(sizeof(uint32_t) * 2 * (size_t)p) != STAAR_OK || ctx->pcnt.reserve(sizeof(uint32_t) * (size_t)p) != STAAR_OK ||
            ctx->redo.reserve(sizeof(long long) * (size_t)p + 2 * sizeof(unsigned long long)) != STAAR_OK) {
            (void)cudaGetLastError();
            ctx->mlist.release(); ctx->plist.release();
            lists = false;
        }
    }
    if (lists) {
        uint2 *dml = ctx->mlist.as<uint2>();
        uint32_t *dmc = ctx->msubcnt.as<uint32_t>();
        unsigned long long *dcount = ctx->redo.as<unsigned long long>();
        long long *dredo = reinterpret_cast<long long *>(dcount + 2);
        STAAR_TRY(launch_contract_tc(ctx, d_packed, stride, n, p, dhi, dlo, dml, dmc, mcap));
        if (prof) STAAR_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
        STAAR_TRY(launch_expand_lists(ctx, dml, dmc, mcap, p, ctx->plist.as<uint32_t>(), ctx->pcnt.as<uint32_t>(), pcap, dredo, dcount));
        STAAR_TRY(launch_scan_packed(ctx, d_packed, stride, n, p, impute, dlo, d_AF, d_MAF, dflip, dmu, dquad, dSm, nullptr, nullptr,
                                     ctx->plist.as<uint32_t>(), ctx->pcnt.as<uint32_t>(), pcap));
        STAAR_TRY(launch_scan_packed(ctx, d_packed, stride, n, p, impute, dlo, d_AF, d_MAF, dflip, dmu, dquad, dSm, dredo, dcount));
    } else {
    if (ctx->use_mma_sync) STAAR_TRY(launch_contract_packed(ctx, d_packed, stride, n, p, dhi, dlo));
    else STAAR_TRY(launch_contract_tc(ctx, d_packed, stride, n, p, dhi, dlo));
    if (prof) STAAR_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    STAAR_TRY(launch_scan_packed(ctx, d_packed, stride, n, p, impute, dlo, d_AF, d_MAF, dflip, dmu, dquad, dSm));
    }
    if (prof) STAAR_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    STAAR_TRY(launch_finalize_packed(ctx, p, dflip, dmu, dquad, dSm, dhi, dlo, d_Score, d_Score_se, d_pl, d_Est, d_Est_se));
    if (prof) STAAR_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    return STAAR_OK;      // asynchronous on ctx->stream: the caller times with events / staar_ctx_synchronize
}

// Test hook: the intermediate arrays of the LAST staar_packed_score_device call (any pointer may be NULL).
// hi/lo: p x 16 exact integer contraction results (X = hi * 2^24 + lo in units of col_scale[c]; slot 15 = hom-minor
// sum against diag(Sigma_i)); Sm: p x 16 column sums over missing samples; quad: kinship off-diagonal part.
int staar_packed_debug_dump(staar_ctx *ctx, int64_t p, long long *hi, long long *lo, double *Sm, double *quad,
                            double *mu, int32_t *flip, double *col_scale)
{
    CHECK_CTX(ctx);
    const size_t pi = (size_t)((p + 1) / 2 * 2);
    int32_t *dflip = ctx->tmp.as<int32_t>();
    double *dmu = reinterpret_cast<double *>(dflip + pi), *dquad = dmu + p, *dSm = dquad + p;
    long long *dhi = reinterpret_cast<long long *>(dSm + 16 * p), *dlo = dhi + 16 * p;
    STAAR_TRY(sync(ctx));
    STAAR_TRY(d2h(ctx, hi, dhi, sizeof(long long) * 16 * p));
    STAAR_TRY(d2h(ctx, lo, dlo, sizeof(long long) * 16 * p));
    STAAR_TRY(d2h(ctx, Sm, dSm, sizeof(double) * 16 * p));
    STAAR_TRY(d2h(ctx, quad, dquad, sizeof(double) * p));
    STAAR_TRY(d2h(ctx, mu, dmu, sizeof(double) * p));
    STAAR_TRY(d2h(ctx, flip, dflip, sizeof(int32_t) * p));
    STAAR_TRY(d2h(ctx, col_scale, ctx->ns.col_scale, sizeof(double) * kColsPerSlice));
    return sync(ctx);
}

int staar_ctx_set_packed_path(staar_ctx *ctx, int path)
{
    CHECK_CTX(ctx);
    if (path != STAAR_PACKED_PATH_FUSED && path != STAAR_PACKED_PATH_SPLIT) { set_error("unknown packed path %d", path); return STAAR_ERR_ARG; }
    ctx->packed_path = path;
    return STAAR_OK;
}

int staar_ctx_last_packed_path_was_fused(staar_ctx *ctx) { return ctx && ctx->last_fused ? 1 : 0; }

int staar_packed_debug_fused(staar_ctx *ctx, int64_t p, double *acc, int32_t *cnt, int32_t *cm, int32_t *guess,
                             uint64_t *redo_count)
{
    CHECK_CTX(ctx);
    if (!ctx->last_fused) { set_error("the last packed score call did not take the fused route"); return STAAR_ERR_STATE; }
    STAAR_TRY(sync(ctx));
    std::vector<FusedScanOut> h((size_t)p);
    STAAR_TRY(d2h(ctx, h.data(), ctx->fscan.p, sizeof(FusedScanOut) * (size_t)p));
    unsigned long long rc = 0;
    STAAR_TRY(d2h(ctx, &rc, ctx->redo.p, sizeof rc));
    STAAR_TRY(sync(ctx));
    for (int64_t i = 0; i < p; i++) {
        if (acc) acc[i] = h[i].acc;
        if (cnt) cnt[i] = h[i].cnt;
        if (cm) cm[i] = h[i].cm;
        if (guess) guess[i] = h[i].guess;
    }
    if (redo_count) *redo_count = rc;
    return STAAR_OK;
}

int staar_packed_debug_timers(staar_ctx *ctx, int64_t ctas, uint32_t *out)
{
    CHECK_CTX(ctx);
    STAAR_TRY(sync(ctx));
    STAAR_TRY(d2h(ctx, out, ctx->tmp2.p, sizeof(uint32_t) * 256 * (size_t)ctas));
    return sync(ctx);
}

int staar_ctx_set_profiling(staar_ctx *ctx, int on)
{
    CHECK_CTX(ctx);
    if (on && !ctx->ev[0]) for (int i = 0; i < 4; i++) STAAR_CUDA(cudaEventCreate(&ctx->ev[i]));
    ctx->profiling = on != 0;
    return STAAR_OK;
}

int staar_ctx_last_stage_ms(staar_ctx *ctx, float ms[3])
{
    CHECK_CTX(ctx);
    if (!ctx->ev[0]) { set_error("profiling was never enabled on this context"); return STAAR_ERR_STATE; }
    STAAR_CUDA(cudaEventSynchronize(ctx->ev[3]));
    if (ctx->last_fused) {
        // fused path: ms[] = redo (block-per-variant scan of the listed variants), fused kernel, finaliser
        STAAR_CUDA(cudaEventElapsedTime(&ms[0], ctx->ev[2], ctx->ev[3]));
        STAAR_CUDA(cudaEventElapsedTime(&ms[1], ctx->ev[0], ctx->ev[1]));
        STAAR_CUDA(cudaEventElapsedTime(&ms[2], ctx->ev[1], ctx->ev[2]));
        return STAAR_OK;
    }
    // kernels run contraction, scan, finaliser; ms[] is reported as scan, contraction, finaliser
    STAAR_CUDA(cudaEventElapsedTime(&ms[0], ctx->ev[1], ctx->ev[2]));
    STAAR_CUDA(cudaEventElapsedTime(&ms[1], ctx->ev[0], ctx->ev[1]));
    STAAR_CUDA(cudaEventElapsedTime(&ms[2], ctx->ev[2], ctx->ev[3]));
    return STAAR_OK;
}

namespace staar {
namespace {
// caller-order packed columns already in ctx->pk: re-order into the resident layout, score, copy the 7 arrays out
int score_from_pk(staar_ctx *ctx, size_t stride, int64_t n, int64_t p, int impute, double *AF, double *MAF, double *Score,
                  double *Score_se, double *pl, double *Est, double *Est_se)
{
    NullSparse &ns = ctx->ns;
    const uint8_t *d_cols = ctx->pk.as<uint8_t>();
    if (ns.permuted) {      // host columns are in the caller's sample order: re-order into the resident layout
        STAAR_TRY(ctx->pk2.reserve(stride * (size_t)std::max<int64_t>(p, 1)));
        STAAR_TRY(launch_permute_packed(ctx, d_cols, ctx->pk2.as<uint8_t>(), stride, n, p, ns.perm));
        d_cols = ctx->pk2.as<uint8_t>();
    }
    STAAR_TRY(ctx->out.reserve(sizeof(double) * 7 * (size_t)std::max<int64_t>(p, 1)));
    double *o = ctx->out.as<double>();
    STAAR_TRY(staar_packed_score_device(ctx, d_cols, stride, n, p, impute, o, o + p, o + 2 * p, o + 3 * p, o + 4 * p,
                                        o + 5 * p, o + 6 * p));
    double *dst[7] = {AF, MAF, Score, Score_se, pl, Est, Est_se};
    for (int a = 0; a < 7; a++) STAAR_TRY(d2h(ctx, dst[a], o + (size_t)a * p, sizeof(double) * p));
    return sync(ctx);
}
}  // namespace
}  // namespace staar

int staar_packed_score_host(staar_ctx *ctx, const uint8_t *packed, size_t stride, int64_t n, int64_t p, int impute,
                            double *AF, double *MAF, double *Score, double *Score_se, double *pl, double *Est,
                            double *Est_se)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    NullSparse &ns = ctx->ns;
    if (!ns.set) { set_error("packed score: call staar_nullmodel_set_sparse first"); return STAAR_ERR_STATE; }
    STAAR_TRY(ctx->pk.reserve(stride * (size_t)std::max<int64_t>(p, 1)));
    STAAR_TRY(h2d(ctx, ctx->pk.p, packed, stride * (size_t)p));
    return score_from_pk(ctx, stride, n, p, impute, AF, MAF, Score, Score_se, pl, Est, Est_se);
}

int staar_packed_analysis_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                                 double mac_cutoff, double spa_p_cutoff, int64_t *d_index, double *d_AF, double *d_MAF,
                                 double *d_Score, double *d_Score_se, double *d_pl, double *d_Est, double *d_Est_se,
                                 int64_t *d_spa_index, int64_t *n_tested, int64_t *n_spa)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (!n_tested) { set_error("packed_analysis: n_tested is NULL"); return STAAR_ERR_ARG; }
    // all variants first (scratch), then the driver's filter: MAF > (mac_cutoff - 0.5)/n/2  (R/Individual_Analysis.R:242,324)
    STAAR_TRY(ctx->out.reserve(sizeof(double) * 7 * (size_t)std::max<int64_t>(p, 1) + 4 * sizeof(long long)));
    double *o = ctx->out.as<double>();
    long long *d_tot = reinterpret_cast<long long *>(o + 7 * (size_t)std::max<int64_t>(p, 1));
    STAAR_TRY(staar_packed_score_device(ctx, d_packed, stride, n, p, impute, o, o + p, o + 2 * p, o + 3 * p, o + 4 * p,
                                        o + 5 * p, o + 6 * p));
    const double *in7[7] = {o, o + p, o + 2 * p, o + 3 * p, o + 4 * p, o + 5 * p, o + 6 * p};
    double *out7[7] = {d_AF, d_MAF, d_Score, d_Score_se, d_pl, d_Est, d_Est_se};
    long long *sel = reinterpret_cast<long long *>(d_spa_index);
    if (!sel) {                                  // no SPA stage requested: selection goes to scratch and is ignored
        STAAR_TRY(ctx->tmp2.reserve(sizeof(long long) * (size_t)std::max<int64_t>(p, 1)));
        sel = ctx->tmp2.as<long long>();
    }
    STAAR_TRY(launch_compact_results(ctx, p, in7, (mac_cutoff - 0.5) / (double)n / 2.0, d_spa_index ? spa_p_cutoff : 0.0,
                                     reinterpret_cast<long long *>(d_index), out7, sel, d_tot));
    long long tot[2] = {0, 0};
    STAAR_TRY(d2h(ctx, tot, d_tot, sizeof tot));
    STAAR_TRY(sync(ctx));
    *n_tested = tot[0];
    if (n_spa) *n_spa = tot[1];
    return STAAR_OK;
}

int staar_nullmodel_set_spa(staar_ctx *ctx, int64_t n, int64_t q, const double *XW, const double *XXWX_inv,
                            const double *residuals, const double *muhat)
{
    CHECK_CTX(ctx);
    if (n <= 0 || q <= 0 || !XW || !XXWX_inv || !residuals || !muhat) { set_error("nullmodel_set_spa: bad arguments"); return STAAR_ERR_ARG; }
    NullSpa &sp = ctx->nspa;
    sp.release();
    STAAR_CUDA(cudaMalloc(&sp.XW, sizeof(double) * n * q));
    STAAR_CUDA(cudaMalloc(&sp.XXWX_inv, sizeof(double) * n * q));
    STAAR_CUDA(cudaMalloc(&sp.residuals, sizeof(double) * n));
    STAAR_CUDA(cudaMalloc(&sp.muhat, sizeof(double) * n));
    STAAR_TRY(h2d(ctx, sp.XW, XW, sizeof(double) * n * q));
    STAAR_TRY(h2d(ctx, sp.XXWX_inv, XXWX_inv, sizeof(double) * n * q));
    STAAR_TRY(h2d(ctx, sp.residuals, residuals, sizeof(double) * n));
    STAAR_TRY(h2d(ctx, sp.muhat, muhat, sizeof(double) * n));
    STAAR_TRY(sync(ctx));
    sp.n = n; sp.q = q; sp.set = true;
    return STAAR_OK;
}

int staar_packed_spa_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, const int64_t *d_variant,
                            int64_t n_sel, int impute, double tol, int max_iter, double *d_pvalue, int64_t *d_trace)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, n_sel));
    const NullSpa &sp = ctx->nspa;
    if (!sp.set || sp.n != n) { set_error("packed SPA: call staar_nullmodel_set_spa for n = %lld first", (long long)n); return STAAR_ERR_STATE; }
    if (stride % 16 != 0 || stride < (size_t)((n + 3) / 4)) { set_error("packed stride must be a multiple of 16 and >= ceil(n/4)"); return STAAR_ERR_ARG; }
    if (n_sel == 0) return STAAR_OK;
    // resident columns are in the sparse null model's sample order (identity when none is set or it is unrelated)
    const int32_t *d_inv = (ctx->ns.set && ctx->ns.n == n && ctx->ns.permuted) ? ctx->ns.inv_perm : nullptr;
    const int slots = spa_scratch_slots(ctx, n, n_sel);
    if (slots < 1) return STAAR_ERR_CUDA;
    STAAR_TRY(ctx->aux.reserve(sizeof(double) * (size_t)slots * n));
    STAAR_TRY(launch_spa_packed(ctx, d_packed, stride, reinterpret_cast<const long long *>(d_variant), d_inv, impute, n, n_sel,
                                sp.XW, sp.XXWX_inv, (int)sp.q, sp.residuals, sp.muhat, tol, max_iter, ctx->aux.as<double>(),
                                d_pvalue, reinterpret_cast<long long *>(d_trace)));
    return sync(ctx);
}

int staar_nullmodel_set_dense(staar_ctx *ctx, int64_t n, const double *P, const double *residuals)
{
    CHECK_CTX(ctx);
    if (n <= 0 || !P || !residuals) { set_error("nullmodel_set_dense: bad arguments"); return STAAR_ERR_ARG; }
    STAAR_TRY(ensure_dense(ctx, n, P));
    STAAR_TRY(h2d(ctx, ctx->nd.residuals, residuals, sizeof(double) * n));
    STAAR_TRY(sync(ctx));
    ctx->nd.res_set = true;
    return STAAR_OK;
}

int staar_packed_score_dense_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                                    double *d_AF, double *d_MAF, double *d_Score, double *d_Score_se, double *d_pl,
                                    double *d_Est, double *d_Est_se)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    // res_set: a frozen-signature denseGRM call with another P replaces the resident P and leaves no residuals behind
    if (!ctx->nd.set || !ctx->nd.res_set || ctx->nd.n != n) { set_error("packed dense-GRM score: call staar_nullmodel_set_dense for n = %lld first", (long long)n); return STAAR_ERR_STATE; }
    if (stride % 16 != 0 || stride < (size_t)((n + 3) / 4)) { set_error("packed stride must be a multiple of 16 and >= ceil(n/4)"); return STAAR_ERR_ARG; }
    if (p == 0) return STAAR_OK;
    STAAR_TRY(packed_score_dense(ctx, d_packed, stride, n, p, impute, d_AF, d_MAF, d_Score, d_Score_se, d_pl, d_Est, d_Est_se));
    return sync(ctx);
}

int staar_packed_score_multi_device(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int n_pheno,
                                    int impute, double *d_AF, double *d_MAF, double *d_Score, double *d_pl)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    if (n_pheno < 1) { set_error("n_pheno must be >= 1"); return STAAR_ERR_ARG; }
    if (!ctx->ns.set || ctx->ns.n != n * n_pheno) {
        set_error("packed multi-trait score: call staar_nullmodel_set_sparse with the %lld x %lld null model first",
                  (long long)(n * n_pheno), (long long)(n * n_pheno));
        return STAAR_ERR_STATE;
    }
    if (stride % 16 != 0 || stride < (size_t)((n + 3) / 4)) { set_error("packed stride must be a multiple of 16 and >= ceil(n/4)"); return STAAR_ERR_ARG; }
    if (p == 0) return STAAR_OK;
    STAAR_TRY(packed_score_multi(ctx, d_packed, stride, n, p, n_pheno, impute, d_AF, d_MAF, d_Score, d_pl));
    return sync(ctx);
}

int staar_packed_cond(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, const int64_t *d_variant,
                      int64_t n_sel, int impute, const double *X_adj, int64_t m, double *Score, double *Score_se, double *pl,
                      double *Est, double *Est_se)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, n_sel));
    NullSparse &ns = ctx->ns;
    if (!ns.set || ns.n != n) { set_error("packed conditional test: call staar_nullmodel_set_sparse (with the refit residuals) for n = %lld first", (long long)n); return STAAR_ERR_STATE; }
    if (!X_adj || m < 1) { set_error("packed conditional test: X_adj is empty"); return STAAR_ERR_ARG; }
    if (stride % 16 != 0 || stride < (size_t)((n + 3) / 4)) { set_error("packed stride must be a multiple of 16 and >= ceil(n/4)"); return STAAR_ERR_ARG; }
    if (n_sel == 0) return STAAR_OK;
    std::vector<double> cov_host((size_t)ns.q * ns.q);
    STAAR_TRY(d2h(ctx, cov_host.data(), ns.cov, sizeof(double) * cov_host.size()));
    STAAR_TRY(sync(ctx));
    CondOperands co;
    STAAR_TRY(prepare_cond(ctx, X_adj, m, cov_host.data(), co));
    STAAR_TRY(ctx->G.reserve(sizeof(double) * (size_t)n * n_sel));
    STAAR_TRY(launch_decode_selected(ctx, d_packed, stride, n, reinterpret_cast<const long long *>(d_variant), n_sel, impute,
                                     ns.permuted ? ns.perm : nullptr, ctx->G.as<double>(), nullptr, nullptr));
    return run_sparse_family(ctx, n, n_sel, 0, &co, Score, Score_se, pl, Est, Est_se);
}

// Conditional analysis of one group of tested variants that share their known loci, refit included
// (R/Individual_Analysis_cond.R:183-212: per tested variant, lm(scaled.residuals ~ genotype_adj + X - 1) — or
// ~ genotype_adj with an intercept — gives the refit residuals and X_adj = model.matrix, then a p = 1 call of
// Individual_Score_Test_cond).  Everything runs on resident columns: the known variants are decoded (flipped / imputed
// as matrix_flip_* does) into X_adj on the device, the least-squares refit is solved from the column-scaled Gram matrix
// (device contraction, m x m Cholesky on the host) with two rounds of iterative refinement — the accuracy of a QR
// solve — and the refit residuals never leave the device.
int staar_packed_cond_refit(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, const int64_t *d_tested,
                            int64_t n_tested, const int64_t *d_known, int64_t n_known, const double *X, int64_t qx, int method,
                            int impute, double *Score, double *Score_se, double *pl, double *Est, double *Est_se,
                            double *coef_out)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, n_tested));
    NullSparse &ns = ctx->ns;
    if (!ns.set || ns.n != n) { set_error("conditional refit: call staar_nullmodel_set_sparse (with the null model's scaled residuals) for n = %lld first", (long long)n); return STAAR_ERR_STATE; }
    if (n_known < 1 || !d_known) { set_error("conditional refit: no known variants"); return STAAR_ERR_ARG; }
    if (method != 0 && method != 1) { set_error("conditional refit: method must be 0 (optimal) or 1 (naive)"); return STAAR_ERR_ARG; }
    if (method == 0 && (!X || qx < 1)) { set_error("conditional refit: method 'optimal' needs the covariate matrix X"); return STAAR_ERR_ARG; }
    if (stride % 16 != 0 || stride < (size_t)((n + 3) / 4)) { set_error("packed stride must be a multiple of 16 and >= ceil(n/4)"); return STAAR_ERR_ARG; }
    if (n_tested == 0) return STAAR_OK;
    const int64_t q = ns.q;
    const int64_t m = method == 0 ? n_known + qx : 1 + n_known;            // model.matrix: [G_known | X] or [1 | G_known]
    // the scratch layout of prepare_cond: X_adj first
    const int ldc = (int)((1 + q + 2 * m + 7) / 8 * 8), ld2 = (int)((1 + 2 * m + 7) / 8 * 8);
    STAAR_TRY(ctx->cond.reserve(sizeof(double) * ((size_t)2 * n * m + (size_t)n * ldc + (size_t)n * ld2 + q * m + 2 * m * m)));
    ctx->cond_key = 0;
    double *dA = ctx->cond.as<double>();
    double *dG_known = method == 0 ? dA : dA + n;
    STAAR_TRY(launch_decode_selected(ctx, d_packed, stride, n, reinterpret_cast<const long long *>(d_known), n_known, impute,
                                     ns.permuted ? ns.perm : nullptr, dG_known, nullptr, nullptr));
    if (method == 0) {
        STAAR_TRY(h2d(ctx, dA + (size_t)n * n_known, X, sizeof(double) * (size_t)n * qx));
    } else {
        std::vector<double> ones((size_t)n, 1.0);
        STAAR_TRY(h2d(ctx, dA, ones.data(), sizeof(double) * (size_t)n));
        STAAR_TRY(sync(ctx));
    }
    // ---- least squares: Gram matrix and X_adj' r by one contraction against [r | X_adj]
    // (pk2, not tmp2: the dense contraction keeps its split-K partial sums in tmp2)
    STAAR_TRY(ctx->pk2.reserve(sizeof(double) * ((size_t)3 * std::max<int64_t>(n, 64) + (size_t)n * ld2 + 64)));
    double *d_r = ctx->pk2.as<double>(), *d_r1 = d_r + n, *d_coef = d_r1 + n, *dY2 = d_r + 3 * (size_t)std::max<int64_t>(n, 64);     // r | r1 | coef (<= 64) | [r | X_adj]
    if (m > 64) { set_error("conditional refit: at most 64 adjustment columns"); return STAAR_ERR_ARG; }
    STAAR_CUDA(cudaMemcpy2DAsync(d_r, sizeof(double), ns.Y, sizeof(double) * ns.ldY, sizeof(double), n, cudaMemcpyDeviceToDevice, ctx->stream));
    STAAR_TRY(launch_build_Y(ctx, n, ld2, d_r, nullptr, 0, dA, nullptr, (int)m, dY2));
    STAAR_TRY(ctx->L.reserve(sizeof(double) * (size_t)m * std::max<int>(ld2, 8)));
    STAAR_TRY(launch_contract_dense(ctx, dA, n, m, dY2, ld2, (int)(1 + m), ctx->L.as<double>(), ld2));
    std::vector<double> L2((size_t)m * ld2);
    STAAR_TRY(d2h(ctx, L2.data(), ctx->L.p, sizeof(double) * L2.size()));
    STAAR_TRY(sync(ctx));
    std::vector<double> scale((size_t)m), R((size_t)m * m), b((size_t)m), coef((size_t)m, 0.0), delta((size_t)m);
    for (int64_t a = 0; a < m; a++) {
        const double d = L2[(size_t)a * ld2 + 1 + a];
        if (!(d > 0.0)) { set_error("conditional refit: an adjustment column is identically zero (singular design)"); return STAAR_ERR_SINGULAR; }
        scale[a] = 1.0 / sqrt(d);
    }
    for (int64_t a = 0; a < m; a++)                                       // Cholesky of the scaled Gram matrix (lower, in R)
        for (int64_t c = 0; c <= a; c++) {
            double sacc = L2[(size_t)a * ld2 + 1 + c] * scale[a] * scale[c];
            for (int64_t k2 = 0; k2 < c; k2++) sacc -= R[a + k2 * m] * R[c + k2 * m];
            if (a == c) {
                if (!(sacc > 1e-13)) { set_error("conditional refit: the design [known variants | covariates] is rank deficient"); return STAAR_ERR_SINGULAR; }
                R[a + a * m] = sqrt(sacc);
            } else {
                R[a + c * m] = sacc / R[c + c * m];
            }
        }
    auto solve = [&](const std::vector<double> &rhs, std::vector<double> &x) {     // (D G D) y = D rhs, x = D y
        std::vector<double> y((size_t)m);
        for (int64_t a = 0; a < m; a++) {
            double sacc = rhs[a] * scale[a];
            for (int64_t k2 = 0; k2 < a; k2++) sacc -= R[a + k2 * m] * y[k2];
            y[a] = sacc / R[a + a * m];
        }
        for (int64_t a = m - 1; a >= 0; a--) {
            double sacc = y[a];
            for (int64_t k2 = a + 1; k2 < m; k2++) sacc -= R[k2 + a * m] * y[k2];
            y[a] = sacc / R[a + a * m];
        }
        for (int64_t a = 0; a < m; a++) x[a] = y[a] * scale[a];
    };
    for (int64_t a = 0; a < m; a++) b[a] = L2[(size_t)a * ld2];
    for (int round = 0; round < 3; round++) {                             // solve, then refine twice on the true residual
        solve(b, delta);
        for (int64_t a = 0; a < m; a++) coef[a] += delta[a];
        STAAR_TRY(h2d(ctx, d_coef, coef.data(), sizeof(double) * m));
        STAAR_TRY(launch_residual_update(ctx, n, dA, (int)m, d_coef, d_r, d_r1));     // r1 = r - X_adj coef
        if (round == 2) break;
        STAAR_TRY(launch_build_Y(ctx, n, 8, d_r1, nullptr, 0, nullptr, nullptr, 0, dY2));
        STAAR_TRY(launch_contract_dense(ctx, dA, n, m, dY2, 8, 1, ctx->L.as<double>(), 8));
        std::vector<double> L1((size_t)m * 8);
        STAAR_TRY(d2h(ctx, L1.data(), ctx->L.p, sizeof(double) * L1.size()));
        STAAR_TRY(sync(ctx));
        for (int64_t a = 0; a < m; a++) b[a] = L1[(size_t)a * 8];
    }
    STAAR_TRY(sync(ctx));
    if (coef_out) memcpy(coef_out, coef.data(), sizeof(double) * m);
    // ---- the conditional test with X_adj and the refit residuals in place
    std::vector<double> cov_host((size_t)q * q);
    STAAR_TRY(d2h(ctx, cov_host.data(), ns.cov, sizeof(double) * cov_host.size()));
    STAAR_TRY(sync(ctx));
    CondOperands co;
    STAAR_TRY(prepare_cond(ctx, nullptr, m, cov_host.data(), co, d_r1));
    STAAR_TRY(ctx->G.reserve(sizeof(double) * (size_t)n * n_tested));
    STAAR_TRY(launch_decode_selected(ctx, d_packed, stride, n, reinterpret_cast<const long long *>(d_tested), n_tested, impute,
                                     ns.permuted ? ns.perm : nullptr, ctx->G.as<double>(), nullptr, nullptr));
    return run_sparse_family(ctx, n, n_tested, 0, &co, Score, Score_se, pl, Est, Est_se);
}

int staar_nullmodel_sample_order(staar_ctx *ctx, int64_t n, int32_t *order)
{
    CHECK_CTX(ctx);
    NullSparse &ns = ctx->ns;
    if (!ns.set || ns.n != n) { set_error("sample_order: null model not set for n = %lld", (long long)n); return STAAR_ERR_STATE; }
    if (!order) { set_error("sample_order: NULL pointer"); return STAAR_ERR_ARG; }
    for (int64_t j = 0; j < n; j++) order[j] = ns.permuted ? ns.h_perm[j] : (int32_t)j;
    return STAAR_OK;
}

int staar_packed_reorder_device(staar_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, size_t stride, int64_t n, int64_t p)
{
    CHECK_CTX(ctx);
    STAAR_TRY(check_dims(n, p));
    NullSparse &ns = ctx->ns;
    if (!ns.set || ns.n != n) { set_error("packed_reorder: null model not set for n = %lld", (long long)n); return STAAR_ERR_STATE; }
    if (stride % 16 != 0 || stride < (size_t)((n + 3) / 4)) { set_error("packed stride must be a multiple of 16 and >= ceil(n/4)"); return STAAR_ERR_ARG; }
    if (d_in == d_out) { set_error("packed_reorder: in-place re-ordering is not supported"); return STAAR_ERR_ARG; }
    if (!ns.permuted) {
        STAAR_CUDA(cudaMemcpyAsync(d_out, d_in, stride * (size_t)p, cudaMemcpyDeviceToDevice, ctx->stream));
        return sync(ctx);
    }
    STAAR_TRY(launch_permute_packed(ctx, d_in, d_out, stride, n, p, ns.perm));
    return sync(ctx);
}

// Share of a pinned FP64 chunk's variants that the GPU packs itself over PCIe.  Starts at 0.3 (measured optimum on the
// 16-vCPU B200 box: 0.25-0.35; 0.45 makes the PCIe side the bottleneck) and is re-balanced per context after every
// chunk from the two measured packing rates, so a box with slower host memory shifts work to the GPU and vice versa.
// STAAR_B200_GPU_PACK_SHARE fixes it (0 disables GPU packing).
double staar_chunk_gpu_pack_share(void)
{
    static const double share = getenv("STAAR_B200_GPU_PACK_SHARE") ? atof(getenv("STAAR_B200_GPU_PACK_SHARE")) : 0.3;
    return share < 0.0 ? 0.0 : (share > 1.0 ? 1.0 : share);
}

double staar_ctx_gpu_pack_share(staar_ctx *ctx)
{
    if (!ctx) return staar_chunk_gpu_pack_share();
    return ctx->pack_share < 0.0 ? staar_chunk_gpu_pack_share() : ctx->pack_share;
}

// One chunk of the single-variant driver (R/Individual_Analysis.R:186-196,264): a `$dosage` block in host memory
// -> flip/impute + Individual_Score_Test statistics, without materialising Geno on the host.  0/1/2/NA dosages are
// packed to 2 bits on the host (threaded, AVX-512 when the CPU has it) so 1/32 of the FP64 bytes cross PCIe; real-valued
// FP64 blocks take the FP64 kernels.  elem: 8 = FP64 (NaN / <= -1 missing), 4 = int32 (R integer matrix, NA_integer_
// or any value <= -1 missing), 1 = uint8 (SeqArray raw dosage, 0xFF missing).
namespace staar {
namespace {
// Upper bound of the share of a pinned FP64 block the GPU packs itself (reading the host matrix over its own PCIe
// link).  With all cores behind one rank the host packers are the faster side and 0.6 keeps the split stable; when the
// ranks of a box share the cores (fewer than 8 host threads per rank) nearly everything may go to the GPU — every GPU
// has its own link, the host threads do not multiply.
static double pack_share_cap()
{
    static const int ranks = getenv("LOCAL_WORLD_SIZE") ? std::max(1, atoi(getenv("LOCAL_WORLD_SIZE"))) : 1;
    const int per_rank = std::max(1, (int)std::thread::hardware_concurrency() / ranks);
    return per_rank < 8 ? 0.95 : 0.6;
}

int chunk_impl(staar_ctx *ctx, const void *G, int elem, int64_t n, int64_t p, int impute, int threads, double *AF,
               double *MAF, double *Score, double *Score_se, double *pl, double *Est, double *Est_se)
{
    STAAR_TRY(check_dims(n, p));
    NullSparse &ns = ctx->ns;
    if (!ns.set || ns.n != n) { set_error("individual_analysis_chunk: null model not set for n = %lld", (long long)n); return STAAR_ERR_STATE; }
    if (p == 0) return STAAR_OK;
    const size_t stride = staar_packed_stride(n);
    bool packed_ok = false;
    if (ns.q + 2 <= kColsPerSlice - 1) {
        STAAR_TRY(pinned_reserve(ctx, stride * (size_t)p));
        STAAR_TRY(ctx->pk.reserve(stride * (size_t)p + 16));
        uint8_t *dst = static_cast<uint8_t *>(ctx->pinned);
        // FP64 blocks in pinned (device-visible) host memory: the GPU packs the LAST share of the variants itself,
        // reading the host matrix over PCIe, while the host threads pack the rest -- the two bandwidths add up
        int64_t p_gpu = 0;
        int *d_bad = reinterpret_cast<int *>(ctx->pk.as<uint8_t>() + stride * (size_t)p);
        const double *G_dev = nullptr;
        if (elem == 8) {
            if (ctx->pack_share < 0.0) ctx->pack_share = staar_chunk_gpu_pack_share();
            const double share = ctx->pack_share;
            cudaPointerAttributes at;
            if (share > 0.0 && cudaPointerGetAttributes(&at, G) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
                G_dev = static_cast<const double *>(at.devicePointer);
                p_gpu = std::min<int64_t>(p, (int64_t)(share * (double)p));
            } else {
                cudaGetLastError();               // pageable memory: not an error, the host packs everything
            }
        }
        const int64_t p_cpu = p - p_gpu;
        if (p_gpu > 0) {
            if (!ctx->ev_pack[0]) { STAAR_CUDA(cudaEventCreate(&ctx->ev_pack[0])); STAAR_CUDA(cudaEventCreate(&ctx->ev_pack[1])); }
            STAAR_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
            STAAR_CUDA(cudaEventRecord(ctx->ev_pack[0], ctx->stream));
            STAAR_TRY(launch_pack_f64(ctx, G_dev + (size_t)p_cpu * n, n, p_gpu, ctx->pk.as<uint8_t>() + stride * (size_t)p_cpu, stride, d_bad));
            STAAR_CUDA(cudaEventRecord(ctx->ev_pack[1], ctx->stream));
        }
        const auto t_host0 = std::chrono::steady_clock::now();
        int rc = elem == 8   ? staar_pack_dosage_host(static_cast<const double *>(G), n, p_cpu, dst, stride, threads)
                 : elem == 4 ? staar_pack_dosage_host_i32(static_cast<const int32_t *>(G), n, p, dst, stride, threads)
                             : staar_pack_dosage_host_u8(static_cast<const uint8_t *>(G), n, p, dst, stride, threads);
        const double t_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
        packed_ok = rc == STAAR_OK;
        if (!packed_ok && elem != 8) return rc;        // integer dosages other than 0/1/2/NA are an error
        if (packed_ok) {
            STAAR_TRY(h2d(ctx, ctx->pk.p, dst, stride * (size_t)p_cpu));
            if (p_gpu > 0) {
                int bad = 0;
                STAAR_TRY(d2h(ctx, &bad, d_bad, sizeof(int)));
                STAAR_TRY(sync(ctx));
                packed_ok = bad == 0;
                // re-balance for the next chunk: both sides ran concurrently, so the rates include their contention
                static const bool fixed = getenv("STAAR_B200_GPU_PACK_SHARE") != nullptr;
                float t_gpu = 0.f;
                if (!fixed && p_cpu > 0 && cudaEventElapsedTime(&t_gpu, ctx->ev_pack[0], ctx->ev_pack[1]) == cudaSuccess &&
                    t_gpu > 0.f && t_host > 0.0) {
                    const double r_cpu = (double)p_cpu / t_host, r_gpu = (double)p_gpu / (double)t_gpu;
                    const double target = r_gpu / (r_cpu + r_gpu);
                    ctx->pack_share = std::min(pack_share_cap(), std::max(0.05, 0.75 * ctx->pack_share + 0.25 * target));
                }
            }
        } else {
            STAAR_TRY(sync(ctx));
        }
    }
    if (packed_ok) return score_from_pk(ctx, stride, n, p, impute, AF, MAF, Score, Score_se, pl, Est, Est_se);
    if (elem != 8) { set_error("individual_analysis_chunk: integer dosages need q <= %d covariates", kColsPerSlice - 3); return STAAR_ERR_ARG; }
    // real-valued genotypes: FP64 flip kernel in place on the device copy, then the FP64 score kernels
    STAAR_TRY(upload_dense_G(ctx, static_cast<const double *>(G), n, p));
    STAAR_TRY(ctx->out.reserve(sizeof(double) * 5 * (size_t)p));
    double *o = ctx->out.as<double>();
    STAAR_TRY(launch_flip_dense(ctx, ctx->G.as<double>(), n, p, impute == STAAR_IMPUTE_MINOR, true, o, o + p));
    STAAR_TRY(d2h(ctx, AF, o, sizeof(double) * p));
    STAAR_TRY(d2h(ctx, MAF, o + p, sizeof(double) * p));
    STAAR_TRY(sync(ctx));
    return run_sparse_family(ctx, n, p, 0, nullptr, Score, Score_se, pl, Est, Est_se);
}

// ---- pipelined form: submit() packs and enqueues, wait() collects.  While the GPU copies and scores chunk i the host
// threads already pack chunk i+1 (the synchronous call leaves the host idle for the H2D + kernels + D2H tail).
int chunk_submit(staar_ctx *ctx, const void *G, int elem, int64_t n, int64_t p, int impute, int threads, int s)
{
    STAAR_TRY(check_dims(n, p));
    if (s < 0 || s > 1) { set_error("chunk_submit: slot must be 0 or 1"); return STAAR_ERR_ARG; }
    NullSparse &ns = ctx->ns;
    if (!ns.set || ns.n != n) { set_error("chunk_submit: null model not set for n = %lld", (long long)n); return STAAR_ERR_STATE; }
    if (ns.q + 2 > kColsPerSlice - 1) { set_error("chunk_submit: the packed path needs q <= %d covariates", kColsPerSlice - 3); return STAAR_ERR_ARG; }
    if (!G || p == 0) { set_error("chunk_submit: empty block"); return STAAR_ERR_ARG; }
    staar_ctx::ChunkSlot &sl = ctx->slot[s];
    if (sl.busy) { set_error("chunk_submit: slot %d still holds a chunk (call staar_chunk_wait first)", s); return STAAR_ERR_STATE; }
    if (!sl.done) {
        STAAR_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
        STAAR_CUDA(cudaEventCreate(&sl.ev0)); STAAR_CUDA(cudaEventCreate(&sl.ev1));
    }
    if (!ctx->stream2) STAAR_CUDA(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    const size_t stride = staar_packed_stride(n);
    if (stride * (size_t)p > sl.pinned_cap) {
        if (sl.pinned) cudaFreeHost(sl.pinned);
        sl.pinned = nullptr; sl.pinned_cap = 0;
        STAAR_CUDA(cudaMallocHost(&sl.pinned, stride * (size_t)p));
        sl.pinned_cap = stride * (size_t)p;
    }
    STAAR_TRY(ctx->pk.reserve(stride * (size_t)p + 16));
    STAAR_TRY(sl.res.reserve(sizeof(double) * 7 * (size_t)p + 16));
    double *o = sl.res.as<double>();
    int *d_bad = reinterpret_cast<int *>(o + 7 * (size_t)p);
    int64_t p_gpu = 0;
    const double *G_dev = nullptr;
    if (elem == 8) {
        if (ctx->pack_share < 0.0) ctx->pack_share = staar_chunk_gpu_pack_share();
        cudaPointerAttributes at;
        if (ctx->pack_share > 0.0 && cudaPointerGetAttributes(&at, G) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            G_dev = static_cast<const double *>(at.devicePointer);
            p_gpu = std::min<int64_t>(p, (int64_t)(ctx->pack_share * (double)p));
        } else {
            cudaGetLastError();
        }
    }
    const int64_t p_cpu = p - p_gpu;
    STAAR_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    if (p_gpu > 0) {
        STAAR_CUDA(cudaEventRecord(sl.ev0, ctx->stream));
        STAAR_TRY(launch_pack_f64(ctx, G_dev + (size_t)p_cpu * n, n, p_gpu, ctx->pk.as<uint8_t>() + stride * (size_t)p_cpu, stride, d_bad));
        STAAR_CUDA(cudaEventRecord(sl.ev1, ctx->stream));
    }
    uint8_t *dst = static_cast<uint8_t *>(sl.pinned);
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = elem == 8   ? staar_pack_dosage_host(static_cast<const double *>(G), n, p_cpu, dst, stride, threads)
                   : elem == 4 ? staar_pack_dosage_host_i32(static_cast<const int32_t *>(G), n, p, dst, stride, threads)
                               : staar_pack_dosage_host_u8(static_cast<const uint8_t *>(G), n, p, dst, stride, threads);
    sl.t_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rc != STAAR_OK && elem != 8) { STAAR_TRY(sync(ctx)); return rc; }
    sl.host_ok = rc == STAAR_OK;
    sl.G = G; sl.elem = elem; sl.impute = impute; sl.threads = threads; sl.n = n; sl.p = p; sl.p_cpu = p_cpu; sl.p_gpu = p_gpu;
    if (sl.host_ok) {
        STAAR_TRY(h2d(ctx, ctx->pk.p, dst, stride * (size_t)p_cpu));
        const uint8_t *d_cols = ctx->pk.as<uint8_t>();
        if (ns.permuted) {
            STAAR_TRY(ctx->pk2.reserve(stride * (size_t)p));
            STAAR_TRY(launch_permute_packed(ctx, d_cols, ctx->pk2.as<uint8_t>(), stride, n, p, ns.perm));
            d_cols = ctx->pk2.as<uint8_t>();
        }
        STAAR_TRY(staar_packed_score_device(ctx, d_cols, stride, n, p, impute, o, o + p, o + 2 * p, o + 3 * p, o + 4 * p,
                                            o + 5 * p, o + 6 * p));
    }
    STAAR_CUDA(cudaEventRecord(sl.done, ctx->stream));
    sl.busy = true;
    return STAAR_OK;
}

int chunk_wait(staar_ctx *ctx, int s, double *AF, double *MAF, double *Score, double *Score_se, double *pl, double *Est,
               double *Est_se)
{
    if (s < 0 || s > 1) { set_error("chunk_wait: slot must be 0 or 1"); return STAAR_ERR_ARG; }
    staar_ctx::ChunkSlot &sl = ctx->slot[s];
    if (!sl.busy) { set_error("chunk_wait: slot %d holds no chunk", s); return STAAR_ERR_STATE; }
    sl.busy = false;
    const int64_t p = sl.p;
    double *o = sl.res.as<double>();
    int bad = 0;
    if (sl.host_ok) {
        // read back on the second stream: the first may already hold the next chunk's work behind this one
        STAAR_CUDA(cudaStreamWaitEvent(ctx->stream2, sl.done, 0));
        double *dstv[7] = {AF, MAF, Score, Score_se, pl, Est, Est_se};
        for (int a = 0; a < 7; a++)
            if (dstv[a]) STAAR_CUDA(cudaMemcpyAsync(dstv[a], o + (size_t)a * p, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream2));
        STAAR_CUDA(cudaMemcpyAsync(&bad, o + 7 * (size_t)p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream2));
        STAAR_CUDA(cudaStreamSynchronize(ctx->stream2));
        static const bool fixed = getenv("STAAR_B200_GPU_PACK_SHARE") != nullptr;
        float t_gpu = 0.f;
        if (!fixed && sl.p_gpu > 0 && sl.p_cpu > 0 && cudaEventElapsedTime(&t_gpu, sl.ev0, sl.ev1) == cudaSuccess && t_gpu > 0.f &&
            sl.t_host > 0.0) {
            const double r_cpu = (double)sl.p_cpu / sl.t_host, r_gpu = (double)sl.p_gpu / (double)t_gpu;
            ctx->pack_share = std::min(pack_share_cap(), std::max(0.05, 0.75 * ctx->pack_share + 0.25 * r_gpu / (r_cpu + r_gpu)));
        }
        if (bad == 0) return STAAR_OK;
    } else {
        STAAR_CUDA(cudaEventSynchronize(sl.done));
    }
    // real-valued genotypes somewhere in the block: the synchronous entry point takes the FP64 kernels
    if (ctx->slot[s ^ 1].busy) STAAR_CUDA(cudaEventSynchronize(ctx->slot[s ^ 1].done));
    return chunk_impl(ctx, sl.G, sl.elem, sl.n, sl.p, sl.impute, sl.threads, AF, MAF, Score, Score_se, pl, Est, Est_se);
}
}  // namespace
}  // namespace staar

int staar_individual_analysis_chunk(staar_ctx *ctx, const double *G, int64_t n, int64_t p, int impute, int threads,
                                    double *AF, double *MAF, double *Score, double *Score_se, double *pl, double *Est,
                                    double *Est_se)
{
    CHECK_CTX(ctx);
    return chunk_impl(ctx, G, 8, n, p, impute, threads, AF, MAF, Score, Score_se, pl, Est, Est_se);
}
int staar_individual_analysis_chunk_i32(staar_ctx *ctx, const int32_t *G, int64_t n, int64_t p, int impute, int threads,
                                        double *AF, double *MAF, double *Score, double *Score_se, double *pl,
                                        double *Est, double *Est_se)
{
    CHECK_CTX(ctx);
    return chunk_impl(ctx, G, 4, n, p, impute, threads, AF, MAF, Score, Score_se, pl, Est, Est_se);
}
int staar_individual_analysis_chunk_u8(staar_ctx *ctx, const uint8_t *G, int64_t n, int64_t p, int impute, int threads,
                                       double *AF, double *MAF, double *Score, double *Score_se, double *pl,
                                       double *Est, double *Est_se)
{
    CHECK_CTX(ctx);
    return chunk_impl(ctx, G, 1, n, p, impute, threads, AF, MAF, Score, Score_se, pl, Est, Est_se);
}

int staar_chunk_submit(staar_ctx *ctx, const double *G, int64_t n, int64_t p, int impute, int threads, int slot)
{
    CHECK_CTX(ctx);
    return chunk_submit(ctx, G, 8, n, p, impute, threads, slot);
}
int staar_chunk_submit_i32(staar_ctx *ctx, const int32_t *G, int64_t n, int64_t p, int impute, int threads, int slot)
{
    CHECK_CTX(ctx);
    return chunk_submit(ctx, G, 4, n, p, impute, threads, slot);
}
int staar_chunk_submit_u8(staar_ctx *ctx, const uint8_t *G, int64_t n, int64_t p, int impute, int threads, int slot)
{
    CHECK_CTX(ctx);
    return chunk_submit(ctx, G, 1, n, p, impute, threads, slot);
}
int staar_chunk_wait(staar_ctx *ctx, int slot, double *AF, double *MAF, double *Score, double *Score_se, double *pl,
                     double *Est, double *Est_se)
{
    CHECK_CTX(ctx);
    return chunk_wait(ctx, slot, AF, MAF, Score, Score_se, pl, Est, Est_se);
}

int staar_synth_packed_device(staar_ctx *ctx, uint8_t *d_packed, size_t stride, int64_t n, int64_t first_variant,
                              int64_t p, uint64_t seed, const uint64_t *d_thr_hom, const uint64_t *d_thr_het,
                              const uint64_t *d_thr_miss, const uint8_t *d_store_alt, const int32_t *d_sample_order)
{
    CHECK_CTX(ctx);
    STAAR_TRY(launch_synth_packed(ctx, d_packed, stride, n, first_variant, p, seed, d_thr_hom, d_thr_het, d_thr_miss,
                                  d_store_alt, d_sample_order));
    return sync(ctx);
}

}  // extern "C"
