// ctx.cu — context, error state, device scratch and the resident null-model operands.
//
// The Rcpp boundary passes Sigma_i, Sigma_iX, cov and residuals (or a 3.2 GB P) on EVERY chunk call
// (reference src/RcppExports.cpp:19-31; R/Individual_Analysis.R:264).  They are uploaded once and kept in HBM,
// keyed on a content hash, so that only genotypes cross PCIe per call (SURVEY.md §7 "null-model residency").
#include "common.cuh"
#include <cstring>
#include <algorithm>

namespace staar {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *last_error() { return g_err; }

int DevBuf::reserve(size_t bytes)
{
    if (bytes <= cap) return STAAR_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        p = nullptr;
        set_error("cudaMalloc(%zu bytes): %s", want, cudaGetErrorString(e));
        return STAAR_ERR_CUDA;
    }
    cap = want;
    return STAAR_OK;
}
void DevBuf::release()
{
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
}

int pinned_reserve(staar_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->pinned_cap) return STAAR_OK;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr; ctx->pinned_cap = 0;
    STAAR_CUDA(cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_cap = bytes;
    return STAAR_OK;
}

void NullSparse::release()
{
    if (owned) {
        cudaFree(row_ptr); cudaFree(col); cudaFree(val); cudaFree(SX); cudaFree(cov);
    }
    cudaFree(diag); cudaFree(sinfo); cudaFree(off_ent); cudaFree(relmask); cudaFree(rel); cudaFree(up_ent); cudaFree(Y); cudaFree(Yfrag); cudaFree(col_scale);
    *this = NullSparse();
}
void NullDense::release()
{
    cudaFree(P); cudaFree(residuals);
    *this = NullDense();
}

// 4-lane multiplicative hash over 8-byte words (content key of host operands; ~8 GB/s)
uint64_t hash_bytes(const void *data, size_t bytes, uint64_t seed)
{
    const uint64_t *w = static_cast<const uint64_t *>(data);
    const size_t nw = bytes / 8;
    uint64_t h[4] = {seed ^ 0x9E3779B97F4A7C15ull, seed ^ 0xC2B2AE3D27D4EB4Full, seed ^ 0x165667B19E3779F9ull,
                     seed ^ 0x27D4EB2F165667C5ull};
    size_t i = 0;
    for (; i + 4 <= nw; i += 4)
        for (int l = 0; l < 4; l++) { h[l] = (h[l] ^ w[i + l]) * 0x100000001B3ull; h[l] ^= h[l] >> 29; }
    for (; i < nw; i++) { h[0] = (h[0] ^ w[i]) * 0x100000001B3ull; h[0] ^= h[0] >> 29; }
    const uint8_t *b = static_cast<const uint8_t *>(data) + nw * 8;
    for (size_t k = 0; k < bytes % 8; k++) h[1] = (h[1] ^ b[k]) * 0x100000001B3ull;
    uint64_t r = h[0];
    for (int l = 1; l < 4; l++) r = (r ^ h[l]) * 0x9E3779B97F4A7C15ull + (r >> 31);
    return r ^ bytes;
}

// strided sample of a large dense matrix (P is 3.2 GB at n = 20k): every `step`-th word plus head and tail
uint64_t hash_sampled(const double *data, size_t count, uint64_t seed)
{
    uint64_t h = seed ^ count;
    const size_t step = std::max<size_t>(1, count / 65536);
    for (size_t i = 0; i < count; i += step) {
        uint64_t w; memcpy(&w, data + i, 8);
        h = (h ^ w) * 0x100000001B3ull; h ^= h >> 29;
    }
    const size_t edge = std::min<size_t>(count, 4096);
    return h ^ hash_bytes(data, edge * 8, 1) ^ hash_bytes(data + (count - edge), edge * 8, 2);
}

// ------------------------------------------------------------------------------------------------
// sparse null model: CSC(uint64) host -> CSR(int64/int32) device, Y = [res | Sigma_iX] row-major, diag, flags
// ------------------------------------------------------------------------------------------------
int nullmodel_sparse_upload(staar_ctx *ctx, int64_t n, int64_t q, const double *Sv, const uint64_t *Sr,
                            const uint64_t *Sc, const double *SX, const double *cov, const double *res,
                            uint64_t fp, uint64_t res_fp, bool want_sliced)
{
    NullSparse &ns = ctx->ns;
    ns.release();
    if (n <= 0 || q < 0) { set_error("null model: bad dimensions n=%lld q=%lld", (long long)n, (long long)q); return STAAR_ERR_ARG; }
    if (n >= (1ll << 31)) { set_error("null model: n too large for int32 column indices"); return STAAR_ERR_ARG; }
    const int64_t nnz = (int64_t)Sc[n];
    // CSC -> CSR (transpose; exact for any matrix, and Sigma_i is symmetric anyway)
    std::vector<int64_t> row_ptr(n + 1, 0);
    for (int64_t e = 0; e < nnz; e++) {
        if (Sr[e] >= (uint64_t)n) { set_error("Sigma_i: row index out of range"); return STAAR_ERR_ARG; }
        row_ptr[Sr[e] + 1]++;
    }
    for (int64_t r = 0; r < n; r++) row_ptr[r + 1] += row_ptr[r];
    std::vector<int32_t> col(nnz);
    std::vector<double> val(nnz);
    std::vector<int64_t> pos(row_ptr.begin(), row_ptr.end() - 1);
    std::vector<double> diag(n, 0.0);
    int64_t nnz_off = 0;
    for (int64_t c = 0; c < n; c++)
        for (uint64_t e = Sc[c]; e < Sc[c + 1]; e++) {
            const int64_t r = (int64_t)Sr[e];
            const int64_t d = pos[r]++;
            col[d] = (int32_t)c; val[d] = Sv[e];
            if (r == c) diag[r] += Sv[e];
            else if (Sv[e] != 0.0) nnz_off++;
        }
    std::vector<NullSparse::SampleInfo> sinfo(n);
    std::vector<NullSparse::OffEntry> off_ent(std::max<int64_t>(nnz_off, 1));
    {
        int64_t o = 0;
        for (int64_t r = 0; r < n; r++) {
            sinfo[r].diag = diag[r]; sinfo[r].off_start = (uint32_t)o;
            for (int64_t e = row_ptr[r]; e < row_ptr[r + 1]; e++)
                if (col[e] != r && val[e] != 0.0) { off_ent[o].val = val[e]; off_ent[o].col = col[e]; off_ent[o].pad = 0; o++; }
            sinfo[r].off_count = (uint32_t)(o - sinfo[r].off_start);
        }
    }
    const int ldY = (int)(((2 + q) + 7) / 8 * 8);
    std::vector<double> Y((size_t)n * ldY, 0.0);
    for (int64_t j = 0; j < n; j++) {
        Y[(size_t)j * ldY] = res[j];
        for (int64_t c = 0; c < q; c++) Y[(size_t)j * ldY + 1 + c] = SX[j + c * n];
        Y[(size_t)j * ldY + 1 + q] = diag[j];
    }
    ns.n = n; ns.q = q; ns.nnz = nnz; ns.nnz_offdiag = nnz_off; ns.ldY = ldY; ns.owned = true;
    STAAR_CUDA(cudaMalloc(&ns.row_ptr, sizeof(int64_t) * (n + 1)));
    STAAR_CUDA(cudaMalloc(&ns.col, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
    STAAR_CUDA(cudaMalloc(&ns.val, sizeof(double) * std::max<int64_t>(nnz, 1)));
    STAAR_CUDA(cudaMalloc(&ns.diag, sizeof(double) * n));
    // relatives mask + inline records for the packed carrier scan (upper triangle)
    const size_t pstride = staar_packed_stride(n);
    std::vector<uint8_t> relmask(pstride, 0);
    std::vector<NullSparse::CarrierRec> rel;
    std::vector<NullSparse::OffEntry> up_ent;
    if (nnz_off > 0) {
        rel.resize(n);
        for (int64_t r = 0; r < n; r++) {
            NullSparse::CarrierRec &rr = rel[r];
            memset(&rr, 0, sizeof rr);
            rr.start = (uint32_t)up_ent.size();
            for (uint32_t t = 0; t < sinfo[r].off_count; t++) {
                const NullSparse::OffEntry &oe = off_ent[sinfo[r].off_start + t];
                if (oe.col <= r) continue;
                if (rr.count < 2) { rr.val[rr.count] = oe.val; rr.col[rr.count] = oe.col; }
                up_ent.push_back(oe);
                rr.count++;
            }
            if (rr.count > 0) relmask[r >> 2] |= (uint8_t)(3u << (2 * (r & 3)));
        }
        if (up_ent.empty()) up_ent.resize(1);
        STAAR_CUDA(cudaMalloc(&ns.rel, sizeof(NullSparse::CarrierRec) * n));
        STAAR_CUDA(cudaMemcpyAsync(ns.rel, rel.data(), sizeof(NullSparse::CarrierRec) * n, cudaMemcpyHostToDevice, ctx->stream));
        STAAR_CUDA(cudaMalloc(&ns.up_ent, sizeof(NullSparse::OffEntry) * up_ent.size()));
        STAAR_CUDA(cudaMemcpyAsync(ns.up_ent, up_ent.data(), sizeof(NullSparse::OffEntry) * up_ent.size(), cudaMemcpyHostToDevice, ctx->stream));
        STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    STAAR_CUDA(cudaMalloc(&ns.relmask, pstride));
    STAAR_CUDA(cudaMemcpyAsync(ns.relmask, relmask.data(), pstride, cudaMemcpyHostToDevice, ctx->stream));
    STAAR_CUDA(cudaMalloc(&ns.sinfo, sizeof(NullSparse::SampleInfo) * n));
    STAAR_CUDA(cudaMalloc(&ns.off_ent, sizeof(NullSparse::OffEntry) * off_ent.size()));
    STAAR_CUDA(cudaMalloc(&ns.SX, sizeof(double) * std::max<int64_t>(n * q, 1)));
    STAAR_CUDA(cudaMalloc(&ns.cov, sizeof(double) * std::max<int64_t>(q * q, 1)));
    STAAR_CUDA(cudaMalloc(&ns.Y, sizeof(double) * (size_t)n * ldY));
    cudaStream_t st = ctx->stream;
    STAAR_CUDA(cudaMemcpyAsync(ns.row_ptr, row_ptr.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(ns.col, col.data(), sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(ns.val, val.data(), sizeof(double) * nnz, cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(ns.diag, diag.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(ns.sinfo, sinfo.data(), sizeof(NullSparse::SampleInfo) * n, cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(ns.off_ent, off_ent.data(), sizeof(NullSparse::OffEntry) * off_ent.size(), cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(ns.SX, SX, sizeof(double) * n * q, cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(ns.cov, cov, sizeof(double) * q * q, cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(ns.Y, Y.data(), sizeof(double) * (size_t)n * ldY, cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaStreamSynchronize(st));
    ns.fingerprint = fp; ns.res_fingerprint = res_fp;
    ns.set = true;
    if (want_sliced) STAAR_TRY(build_sliced_operand(ctx, Y.data()));
    return STAAR_OK;
}

int nullmodel_dense_upload(staar_ctx *ctx, int64_t n, const double *P, const double *res, uint64_t fp)
{
    NullDense &nd = ctx->nd;
    nd.release();
    STAAR_CUDA(cudaMalloc(&nd.P, sizeof(double) * (size_t)n * n));
    STAAR_CUDA(cudaMalloc(&nd.residuals, sizeof(double) * n));
    STAAR_CUDA(cudaMemcpyAsync(nd.P, P, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, ctx->stream));
    STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
    nd.n = n; nd.fingerprint = fp; nd.set = true;
    (void)res;
    return STAAR_OK;
}

}  // namespace staar

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int staar_abi_version(void) { return 1; }
const char *staar_last_error(void) { return staar::last_error(); }

int staar_ctx_create(int device, staar_ctx **out)
{
    if (!out) { staar::set_error("staar_ctx_create: out is NULL"); return STAAR_ERR_ARG; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        staar::set_error("no CUDA device available (%s): this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return STAAR_ERR_CUDA;
    }
    if (device < 0 || device >= count) { staar::set_error("device %d out of range (0..%d)", device, count - 1); return STAAR_ERR_ARG; }
    STAAR_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    STAAR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        staar::set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return STAAR_ERR_CUDA;
    }
    staar_ctx *ctx = new staar_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete ctx; staar::set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return STAAR_ERR_CUDA; }
    *out = ctx;
    return STAAR_OK;
}

void staar_ctx_destroy(staar_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->ns.release(); ctx->nd.release();
    ctx->G.release(); ctx->L.release(); ctx->aux.release(); ctx->out.release(); ctx->tmp.release();
    ctx->tmp2.release(); ctx->pk.release(); ctx->ctr.release();
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void *staar_ctx_stream(staar_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int staar_ctx_synchronize(staar_ctx *ctx)
{
    if (!ctx) { staar::set_error("null context"); return STAAR_ERR_ARG; }
    STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
    return STAAR_OK;
}

uint64_t staar_ctx_launch_count(staar_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
