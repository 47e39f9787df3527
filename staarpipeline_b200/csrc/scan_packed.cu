// scan_packed.cu — K1 + K3 for packed dosages, and the per-variant finalisation (K5) of the packed path.
//
// scan_packed_kernel (one warp per variant, two sweeps over the 2-bit column):
//   sweep 1: popcount the codes -> AF, MAF, flip decision, imputation value — bit-exact with
//            matrix_flip_mean / matrix_flip_minor (src/matrix_flip_mean.cpp:23-70, src/matrix_flip_minor.cpp:23-95);
//   sweep 2 (the column is L1/L2-hot): flip the codes in the bit domain and visit only the NON-ZERO genotypes
//            (carriers and imputed entries, a few % of a rare variant's column) to accumulate the sparse quadratic
//            form g' Sigma_i g (the (trans(G)*Sigma_i)*G factor of src/Individual_Score_Test.cpp:43, diagonal entry
//            only): hom-minor corrections and the kinship off-diagonal rows of related carriers — the linear part
//            sum_j s_jj g_j rides in the tensor-core contraction — plus the column sums over missing entries that
//            the integer contraction (contract_packed.cu) leaves out.
// finalize_packed_kernel (one thread per variant): integer contraction -> FP64, imputation correction mu * Sm,
//   Cov_ii = quad - t' cov t, and the statistics loop of src/Individual_Score_Test.cpp:45-64.
//
// Integer/byte work, HBM/L2-bound: 128-bit coalesced loads, popc/ffs bit tricks, warp-shuffle reductions with a
// fixed tree (no atomics => 1/2/4/8-GPU runs are bit-identical).
#include "common.cuh"
#include "packed.cuh"

namespace staar {

namespace {

struct ScanParams {
    const uint8_t *packed; size_t stride; int64_t n, p; int impute;
    const double *diag;                    // Sigma_i[j,j]
    const uint8_t *relmask;                // 2-bit mask, 11 = sample has kinship off-diagonals
    const NullSparse::RelRecord *rel;      // per-sample inline off-diagonal entries (64 B)
    const NullSparse::SampleInfo *sinfo;   // CSR fallback for rows with more than 5 off-diagonals
    const NullSparse::OffEntry *off_ent;
    const double *Y; int ldY, ncol;
    double *AF, *MAF; int32_t *flip; double *mu, *quad, *Sm;
};

constexpr int kScanWarps = 8;
constexpr int kListCap = 1024;          // genotypes scanned per warp iteration (32 lanes x 8 bytes x 4)

__device__ __forceinline__ double code_value(uint32_t code, double mu)
{
    return code == 3u ? mu : (double)code;
}

// decoded (post-flip) genotype of sample k in a packed column
__device__ __forceinline__ double lookup_value(const uint8_t *colp, int64_t k, bool flip, double mu)
{
    uint32_t code = (__ldg(colp + (k >> 2)) >> (2 * (k & 3))) & 3u;
    if (flip && code != 3u && code != 1u) code ^= 2u;
    return code_value(code, mu);
}

// One warp per variant, two sweeps over the 2-bit column.
//  sweep 1: popcounts -> AF / MAF / flip / imputation value.
//  sweep 2: only the genotypes the integer contraction cannot account for are visited:
//            code 2 after the flip (hom-minor: g^2 = g + 2, contributes 2 s_jj to the quadratic form),
//            code 3 (missing: column sums of Y for the imputation correction),
//            carriers whose sample has kinship off-diagonals (relmask): sum_k s_jk g_j g_k.
//           They are compacted (bit tricks only) into a per-warp shared-memory list, then processed lane-parallel
//           so the latency-bound gathers have 32 independent loads in flight per instruction.
// The diagonal part sum_j s_jj g_j of g' Sigma_i g rides in the tensor-core contraction as an extra column of Y.
template <bool HAS_OFF>
__global__ void __launch_bounds__(kScanWarps * 32) scan_packed_kernel(ScanParams P)
{
    __shared__ uint32_t s_list[kScanWarps][kListCap];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nvec8 = (int)(P.stride >> 3);
    uint32_t *list = s_list[wib];
    const uint2 *rmask = reinterpret_cast<const uint2 *>(P.relmask);
    for (int64_t i = warp; i < P.p; i += nwarps) {
        const uint8_t *colp = P.packed + (size_t)i * P.stride;
        const PackedCounts pc = count_packed_column(colp, P.stride, lane);
        const VariantFlip vf = flip_from_counts(pc, P.n, P.impute);
        const bool flip = vf.flip != 0;
        const double mu = vf.mu;
        double quad = 0.0, sm = 0.0;           // sm: lane c (< ncol) accumulates sum over missing j of Y[j][c]
        const uint2 *vecs = reinterpret_cast<const uint2 *>(colp);
        for (int k0 = 0; k0 < nvec8; k0 += 32) {
            const int k = k0 + lane;
            uint2 q2 = make_uint2(0u, 0u), r2 = make_uint2(0u, 0u);
            if (k < nvec8) {
                q2 = __ldg(vecs + k);
                if (HAS_OFF) r2 = __ldg(rmask + k);
            }
            uint32_t w[2] = {q2.x, q2.y}, sel[2];
            const uint32_t rm[2] = {r2.x, r2.y};
            const int64_t base = (int64_t)k * 32;                   // first sample of this lane's 8 bytes
            int cnt = 0;
#pragma unroll
            for (int wi = 0; wi < 2; wi++) {
                if (flip) {
                    w[wi] = flip_word(w[wi]);
                    const int64_t valid = P.n - (base + wi * 16);  // padding codes would flip to 2
                    if (valid < 16) w[wi] &= (valid <= 0) ? 0u : ((1u << (2 * valid)) - 1u);
                }
                const uint32_t hi = (w[wi] >> 1) & 0x55555555u;                      // codes 2 and 3
                const uint32_t nz = (w[wi] | (w[wi] >> 1)) & 0x55555555u;            // any carrier
                sel[wi] = hi | (nz & rm[wi]);                                         // one bit per selected sample
                cnt += __popc(sel[wi]);
            }
            if (!__any_sync(0xffffffffu, cnt != 0)) continue;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            int pos = incl - cnt;
#pragma unroll
            for (int wi = 0; wi < 2; wi++) {
                uint32_t ss = sel[wi];
                while (ss) {
                    const int b = __ffs(ss) - 1;                     // even bit position = 2 * field
                    ss &= ss - 1;
                    const uint32_t code = (w[wi] >> b) & 3u;
                    list[pos++] = ((uint32_t)(base + wi * 16 + (b >> 1)) << 2) | code;
                }
            }
            __syncwarp();
            for (int e0 = 0; e0 < total; e0 += 32) {
                const int e = e0 + lane;
                const bool live = e < total;
                const uint32_t ent = live ? list[e] : 0u;
                const uint32_t code = ent & 3u;
                const int64_t j = ent >> 2;
                if (live) {
                    const double v = code_value(code, mu);
                    if (code == 2u) quad += 2.0 * P.diag[j];                       // g^2 - g for g = 2
                    if (HAS_OFF && v != 0.0) {
                        const bool related = (__ldg(P.relmask + (j >> 2)) >> (2 * (j & 3))) & 1u;
                        if (related) {
                            const NullSparse::RelRecord rr = P.rel[j];
                            double s = 0.0;
                            if (rr.count <= 5) {
#pragma unroll
                                for (int t = 0; t < 5; t++)
                                    if (t < rr.count) s += rr.val[t] * lookup_value(colp, rr.col[t], flip, mu);
                            } else {
                                const NullSparse::SampleInfo si = P.sinfo[j];
                                const NullSparse::OffEntry *oe = P.off_ent + si.off_start;
                                for (uint32_t t = 0; t < si.off_count; t++) s += oe[t].val * lookup_value(colp, oe[t].col, flip, mu);
                            }
                            quad += v * s;
                        }
                    }
                }
                // missing entries: all lanes cooperate on the Y row (lane c adds column c)
                uint32_t mm = __ballot_sync(0xffffffffu, live && code == 3u);
                while (mm) {
                    const int src = __ffs(mm) - 1;
                    mm &= mm - 1;
                    const int64_t jm = __shfl_sync(0xffffffffu, ent, src) >> 2;
                    if (lane < P.ncol) sm += P.Y[jm * P.ldY + lane];
                }
            }
            __syncwarp();
        }
        quad = warp_sum(quad);
        if (lane == 0) {
            P.AF[i] = vf.af; P.MAF[i] = vf.maf; P.flip[i] = vf.flip; P.mu[i] = mu; P.quad[i] = quad;
        }
        if (lane < kColsPerSlice) P.Sm[i * kColsPerSlice + lane] = (lane < P.ncol) ? sm : 0.0;
    }
}

struct FinParams {
    int64_t p; int q;
    const double *mu, *quad, *Sm;
    const long long *hi, *lo;
    const double *scale, *cov;
    double *Score, *Score_se, *pl, *Est, *Est_se;
};

__global__ void __launch_bounds__(128) finalize_packed_kernel(FinParams P)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.p; i += (int64_t)gridDim.x * blockDim.x) {
        const double mu = P.mu[i];
        double L[kColsPerSlice];
        for (int c = 0; c <= P.q; c++) {
            // the integer contraction saw the (already flipped) codes with missing as 0; add mu * sum_missing(y)
            const long long hi = P.hi[i * 16 + c], lo = P.lo[i * 16 + c];
            L[c] = ((double)hi * 16777216.0 + (double)lo) * P.scale[c] + mu * P.Sm[i * kColsPerSlice + c];
        }
        const double U = L[0];
        double tct = 0.0;
        for (int b = 0; b < P.q; b++) {
            double xb = 0.0;
            for (int a = 0; a < P.q; a++) xb += L[1 + a] * P.cov[a + b * P.q];
            tct += xb * L[1 + b];
        }
        // g' Sigma_i g = sum_j s_jj g_j (column q+1 of the contraction, non-missing samples)
        //                + [2 s_jj over hom-minor carriers + kinship off-diagonals]  (scan_packed_kernel)
        //                + mu^2 * sum_missing s_jj
        const long long shi = P.hi[i * 16 + P.q + 1], slo = P.lo[i * 16 + P.q + 1];
        const double dlin = ((double)shi * 16777216.0 + (double)slo) * P.scale[P.q + 1];
        const double quad = dlin + P.quad[i] + mu * mu * P.Sm[i * kColsPerSlice + P.q + 1];
        const double C = quad - tct;
        P.Score[i] = U;
        single_trait_stats(U, C, P.Score_se[i], P.pl[i], P.Est[i], P.Est_se[i]);
    }
}

}  // namespace

int launch_scan_packed(staar_ctx *ctx, const uint8_t *d_packed, size_t stride, int64_t n, int64_t p, int impute,
                       double *dAF, double *dMAF, int32_t *dflip, double *d_mu, double *d_quad, double *d_Sm)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    if (n >= (1ll << 30)) { set_error("packed scan: n too large for 30-bit sample indices"); return STAAR_ERR_ARG; }
    ScanParams P;
    P.packed = d_packed; P.stride = stride; P.n = n; P.p = p; P.impute = impute;
    P.diag = ns.diag; P.relmask = ns.relmask; P.rel = ns.rel; P.sinfo = ns.sinfo; P.off_ent = ns.off_ent;
    P.Y = ns.Y; P.ldY = ns.ldY; P.ncol = (int)ns.q + 2;
    P.AF = dAF; P.MAF = dMAF; P.flip = dflip; P.mu = d_mu; P.quad = d_quad; P.Sm = d_Sm;
    int64_t blocks = std::min<int64_t>((p + kScanWarps - 1) / kScanWarps, (int64_t)ctx->sm_count * 8);
    if (ns.nnz_offdiag > 0) scan_packed_kernel<true><<<(int)blocks, kScanWarps * 32, 0, ctx->stream>>>(P);
    else scan_packed_kernel<false><<<(int)blocks, kScanWarps * 32, 0, ctx->stream>>>(P);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

int launch_finalize_packed(staar_ctx *ctx, int64_t p, const double *d_mu, const double *d_quad,
                           const double *d_Sm, const long long *d_hi, const long long *d_lo, double *dScore,
                           double *dScore_se, double *dpl, double *dEst, double *dEst_se)
{
    if (p == 0) return STAAR_OK;
    const NullSparse &ns = ctx->ns;
    FinParams P;
    P.p = p; P.q = (int)ns.q; P.mu = d_mu; P.quad = d_quad; P.Sm = d_Sm;
    P.hi = d_hi; P.lo = d_lo; P.scale = ns.col_scale; P.cov = ns.cov;
    P.Score = dScore; P.Score_se = dScore_se; P.pl = dpl; P.Est = dEst; P.Est_se = dEst_se;
    int blocks = (int)std::min<int64_t>((p + 127) / 128, (int64_t)ctx->sm_count * 8);
    finalize_packed_kernel<<<blocks, 128, 0, ctx->stream>>>(P);
    ctx->launches++;
    STAAR_CUDA(cudaGetLastError());
    return STAAR_OK;
}

}  // namespace staar
