// ctx.cu — context, error state, device scratch and the resident null-model operands.
//
// The Rcpp boundary passes Sigma_i, Sigma_iX, cov and residuals (or a 3.2 GB P) on EVERY chunk call
// (reference src/RcppExports.cpp:19-31; R/Individual_Analysis.R:264).  They are uploaded once and kept in HBM,
// keyed on a content hash, so that only genotypes cross PCIe per call (SURVEY.md §7 "null-model residency").
#include <thread>
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
    if (permuted) cudaFree(Yp);
    cudaFree(diag); cudaFree(diag_p); cudaFree(sinfo_p); cudaFree(off_ent_p); cudaFree(relmask); cudaFree(rel); cudaFree(up_ent); cudaFree(Y);
    cudaFree(Yfrag); cudaFree(Ytc); cudaFree(Ypair); cudaFree(col_scale); cudaFree(col_total); cudaFree(perm); cudaFree(inv_perm); cudaFree(kmask); cudaFree(wdesc); cudaFree(xterms); cudaFree(lut); cudaFree(lutf);
    *this = NullSparse();
}
void NullSpa::release()
{
    cudaFree(XW); cudaFree(XXWX_inv); cudaFree(residuals); cudaFree(muhat);
    *this = NullSpa();
}
void NullDense::release()
{
    cudaFree(P); cudaFree(residuals);
    *this = NullDense();
}

// 4-lane multiplicative hash over 8-byte words (content key of host operands; ~8 GB/s)
static uint64_t hash_bytes_serial(const void *data, size_t bytes, uint64_t seed)
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

// Fingerprint of a host operand (the Rcpp signatures pass the null model on every call; the device copy is reused when
// the fingerprint matches).  Every byte is hashed; operands of several MB are split into fixed 4 MB pieces hashed by
// host threads and combined in order, so the value does not depend on the thread count.
uint64_t hash_bytes(const void *data, size_t bytes, uint64_t seed)
{
    constexpr size_t kPiece = 4u << 20;
    if (bytes < 2 * kPiece) return hash_bytes_serial(data, bytes, seed);
    const size_t pieces = (bytes + kPiece - 1) / kPiece;
    std::vector<uint64_t> part(pieces);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nthreads = (unsigned)std::min<size_t>(std::min(hw, 8u), pieces);
    auto work = [&](unsigned tid) {
        for (size_t k = tid; k < pieces; k += nthreads) {
            const size_t off = k * kPiece, len = std::min(kPiece, bytes - off);
            part[k] = hash_bytes_serial(static_cast<const uint8_t *>(data) + off, len, seed + 0x51ull * (k + 1));
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    return hash_bytes_serial(part.data(), pieces * 8, seed) ^ bytes;
}

// ------------------------------------------------------------------------------------------------
// sparse null model: CSC(uint64) host -> CSR(int64/int32) device, Y = [res | Sigma_iX | diag] row-major,
// and the packed path's "null-model sample order" with its per-family kinship tables (common.cuh).
// ------------------------------------------------------------------------------------------------
namespace {

template <typename T> int upload_vec(staar_ctx *ctx, T **dst, const std::vector<T> &src)
{
    const size_t count = std::max<size_t>(src.size(), 1);
    STAAR_CUDA(cudaMalloc(reinterpret_cast<void **>(dst), sizeof(T) * count));
    ctx->ns.reg(dst, sizeof(T) * count);
    if (!src.empty())
        STAAR_CUDA(cudaMemcpyAsync(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, ctx->stream));
    return STAAR_OK;
}

struct Family { int32_t first_member; int32_t size; };   // members: fam_members[first_member .. +size), natural indices

}  // namespace

int nullmodel_sparse_upload(staar_ctx *ctx, int64_t n, int64_t q, const double *Sv, const uint64_t *Sr,
                            const uint64_t *Sc, const double *SX, const double *cov, const double *res,
                            uint64_t fp, uint64_t res_fp, bool want_sliced)
{
    NullSparse &ns = ctx->ns;
    ns.release();
    if (n <= 0 || q < 0) { set_error("null model: bad dimensions n=%lld q=%lld", (long long)n, (long long)q); return STAAR_ERR_ARG; }
    if (n >= (1ll << 30)) { set_error("null model: n too large for 30-bit sample indices"); return STAAR_ERR_ARG; }
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
    std::vector<int64_t> wpos(row_ptr.begin(), row_ptr.end() - 1);
    std::vector<double> diag(n, 0.0);
    int64_t nnz_off = 0;
    // union-find over the off-diagonal graph: families = connected components
    std::vector<int32_t> parent(n);
    for (int64_t j = 0; j < n; j++) parent[j] = (int32_t)j;
    auto find = [&](int32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (int64_t c = 0; c < n; c++)
        for (uint64_t e = Sc[c]; e < Sc[c + 1]; e++) {
            const int64_t r = (int64_t)Sr[e];
            const int64_t d = wpos[r]++;
            col[d] = (int32_t)c; val[d] = Sv[e];
            if (r == c) diag[r] += Sv[e];
            else if (Sv[e] != 0.0) {
                nnz_off++;
                const int32_t a = find((int32_t)r), b = find((int32_t)c);
                if (a != b) parent[std::max(a, b)] = std::min(a, b);
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

    // ---- families and the null-model sample order --------------------------------------------------------
    const size_t pstride = staar_packed_stride(n);
    const bool tables_ok = nnz_off > 0 && pstride + 4096 <= 227 * 1024;     // column must fit in shared memory
    std::vector<int32_t> fam_members;            // natural indices, grouped by family (multi-member families only)
    std::vector<Family> fams;
    std::vector<int32_t> singles;
    if (nnz_off > 0) {
        std::vector<int32_t> root(n), count(n, 0), start(n, -1);
        for (int64_t j = 0; j < n; j++) { root[j] = find((int32_t)j); count[root[j]]++; }
        int32_t off = 0;
        for (int64_t j = 0; j < n; j++)
            if (root[j] == j && count[j] > 1) { start[j] = off; fams.push_back(Family{off, count[j]}); off += count[j]; }
        fam_members.resize(off);
        std::vector<int32_t> fill(n, 0);
        for (int64_t j = 0; j < n; j++) {
            const int32_t r = root[j];
            if (count[r] > 1) fam_members[start[r] + fill[r]++] = (int32_t)j;   // ascending natural index within a family
            else singles.push_back((int32_t)j);
        }
    }
    // ---- layout: 16-sample words of four 4-sample groups.  A family of <= 4 members lives inside one group (its
    // members first, unrelated samples fill the rest; two pairs share a group); a family of 5..16 members takes
    // consecutive groups of one word in chunks of 4.  Words holding such a family come first ("cross" words: they
    // also need member x chunk terms); every within-group pair is covered by that group's look-up table.
    struct Item { std::vector<int32_t> fam; int groups; };      // groups of 4 slots this item occupies in a word
    std::vector<Item> items;
    std::vector<int32_t> generic_fams;
    size_t singles_needed = 0;
    if (tables_ok) {
        std::vector<int32_t> pairs;
        for (size_t f = 0; f < fams.size(); f++) {
            const int sz = fams[f].size;
            if (sz > 16) generic_fams.push_back((int32_t)f);
            else if (sz == 2) pairs.push_back((int32_t)f);
            else items.push_back(Item{{(int32_t)f}, (sz + 3) / 4});
        }
        for (size_t i = 0; i + 1 < pairs.size(); i += 2) items.push_back(Item{{pairs[i], pairs[i + 1]}, 1});
        if (pairs.size() & 1) items.push_back(Item{{pairs.back()}, 1});
    } else {
        for (size_t f = 0; f < fams.size(); f++) generic_fams.push_back((int32_t)f);
    }
    // bins = words of 4 groups, best-fit decreasing by group count; multi-group items first
    struct Bin { std::vector<int32_t> item; int used = 0; bool cross = false; };
    std::vector<Bin> bins;
    {
        std::vector<int32_t> order(items.size());
        for (size_t i = 0; i < items.size(); i++) order[i] = (int32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return items[a].groups > items[b].groups; });
        std::vector<std::vector<int32_t>> open(5);                  // open[r] = bins with r free groups
        for (int32_t it : order) {
            const int g = items[it].groups;
            int r = g;
            while (r <= 4 && open[r].empty()) r++;
            int32_t b;
            if (r <= 4) { b = open[r].back(); open[r].pop_back(); }
            else { b = (int32_t)bins.size(); bins.emplace_back(); r = 4; }
            bins[b].item.push_back(it); bins[b].used += g; bins[b].cross |= g > 1;
            if (r - g > 0) open[r - g].push_back(b);
        }
    }
    auto bin_members = [&](const Bin &b) { int m = 0; for (int32_t it : b.item) for (int32_t f : items[it].fam) m += fams[f].size; return m; };
    // words that cannot be topped up with unrelated samples go generic (cohorts where nearly everyone is related)
    std::vector<int32_t> bin_order(bins.size());
    for (size_t b = 0; b < bins.size(); b++) bin_order[b] = (int32_t)b;
    std::stable_sort(bin_order.begin(), bin_order.end(), [&](int32_t a, int32_t b) {
        if (bins[a].cross != bins[b].cross) return bins[a].cross;              // cross words first
        return bin_members(bins[a]) > bin_members(bins[b]);
    });
    std::vector<int32_t> table_bins;
    {
        size_t reserved = 0;
        for (int32_t b : bin_order) {
            const size_t gap = (size_t)(16 - bin_members(bins[b]));
            if (gap <= singles.size() - reserved) { table_bins.push_back(b); reserved += gap; }
            else for (int32_t it : bins[b].item) for (int32_t f : items[it].fam) generic_fams.push_back(f);
        }
        singles_needed = reserved;
    }
    (void)singles_needed;
    std::vector<int32_t> perm;                    // position -> natural index
    perm.reserve(n);
    size_t single_next = 0;
    struct Placed { int32_t fam; int64_t pos0; };
    std::vector<Placed> table_placed, generic_placed;
    const bool permute = tables_ok && !table_bins.empty();
    ns.x_words = 0;
    if (permute) {
        for (int32_t b : table_bins) {
            const int64_t word0 = (int64_t)perm.size();
            for (int32_t it : bins[b].item) {
                // an item starts on a group boundary; families of one item are laid out back to back
                for (int32_t f : items[it].fam) {
                    table_placed.push_back(Placed{f, (int64_t)perm.size()});
                    for (int m = 0; m < fams[f].size; m++) perm.push_back(fam_members[fams[f].first_member + m]);
                }
                while (((int64_t)perm.size() - word0) % 4 != 0) perm.push_back(singles[single_next++]);
            }
            while ((int64_t)perm.size() - word0 < 16) perm.push_back(singles[single_next++]);
            if (bins[b].cross) ns.x_words = (int64_t)perm.size() / 16;
        }
        const int64_t tab_samples = (int64_t)perm.size();
        for (int32_t f : generic_fams) {
            generic_placed.push_back(Placed{f, (int64_t)perm.size()});
            for (int m = 0; m < fams[f].size; m++) perm.push_back(fam_members[fams[f].first_member + m]);
        }
        ns.tab_words = tab_samples / 16;                                   // table words come first: [0, tab_words)
        ns.kin_words = ((int64_t)perm.size() + 15) / 16;
        while (single_next < singles.size()) perm.push_back(singles[single_next++]);
    } else {
        // identity order; every family on the generic path (positions == natural indices)
        for (size_t f = 0; f < fams.size(); f++) generic_placed.push_back(Placed{(int32_t)f, -1});
        ns.tab_words = 0;
        int64_t last = -1;
        for (int32_t j : fam_members) last = std::max<int64_t>(last, j);
        ns.kin_words = nnz_off > 0 ? last / 16 + 1 : 0;
    }
    if (permute && (int64_t)perm.size() != n) { set_error("internal: sample order is not a permutation"); return STAAR_ERR_ARG; }
    std::vector<int32_t> inv;
    if (permute) { inv.resize(n); for (int64_t pz = 0; pz < n; pz++) inv[perm[pz]] = (int32_t)pz; }
    auto pos_of = [&](int32_t natural) { return permute ? inv[natural] : natural; };
    auto nat_of = [&](int64_t pz) { return permute ? perm[pz] : (int32_t)pz; };
    ns.permuted = permute;

    // ---- per-position kinship rows (both directions): corrections for missing related samples ------------
    std::vector<NullSparse::SampleInfo> sinfo(n);
    std::vector<NullSparse::OffEntry> off_ent;
    off_ent.reserve(std::max<int64_t>(nnz_off, 1));
    for (int64_t pz = 0; pz < n; pz++) {
        const int32_t r = nat_of(pz);
        sinfo[pz].diag = diag[r]; sinfo[pz].off_start = (uint32_t)off_ent.size();
        for (int64_t e = row_ptr[r]; e < row_ptr[r + 1]; e++)
            if (col[e] != r && val[e] != 0.0) off_ent.push_back(NullSparse::OffEntry{val[e], pos_of(col[e]), 0});
        sinfo[pz].off_count = (uint32_t)(off_ent.size() - sinfo[pz].off_start);
    }
    // ---- table words: one look-up table per 4-sample group (all related pairs inside the group), and for words
    //      holding a family of more than 4 members the member x later-chunk terms ----------------------------
    std::vector<uint32_t> kmask((size_t)(ns.tab_words + 7) / 4 * 4, 0u);   // per table word: both bits of every related slot
    std::vector<NullSparse::WordDesc> wdesc((size_t)std::max<int64_t>(ns.x_words, 1));
    memset(wdesc.data(), 0, sizeof(NullSparse::WordDesc) * wdesc.size());
    std::vector<NullSparse::KinTerm> xterms;
    // group tables at a fixed stride: entry idx of group g (= 4 * word + group-in-word) lives at lut[(g << 8) + idx], so a
    // look-up needs no descriptor; the member x chunk tables of the cross words follow
    std::vector<double> lut((size_t)std::max<int64_t>(ns.tab_words, 1) * 4 * 256, 0.0), lutf;
    {
        auto codeval = [](int idx, int m) { const int c = (idx >> (2 * m)) & 3; return c == 3 ? 0.0 : (double)c; };
        // kinship entry between two positions (0 when unrelated)
        auto s_between = [&](int64_t a, int64_t b2) {
            const NullSparse::SampleInfo &si = sinfo[a];
            double v = 0.0;
            for (uint32_t t = 0; t < si.off_count; t++) if (off_ent[si.off_start + t].col == b2) v += off_ent[si.off_start + t].val;
            return v;
        };
        for (int64_t g = 0; g < ns.tab_words * 4; g++) {
            const int64_t p0 = g * 4;
            double S[4][4];
            int hi = -1;                                             // highest slot holding a related sample
            uint32_t mask = 0u;
            bool any_pair = false;
            for (int a = 0; a < 4; a++) {
                if (p0 + a < n && sinfo[p0 + a].off_count > 0) { hi = a; mask |= 3u << (2 * a); }
                for (int b2 = 0; b2 < 4; b2++) {
                    S[a][b2] = (a != b2 && p0 + a < n && p0 + b2 < n) ? s_between(p0 + a, p0 + b2) : 0.0;
                    any_pair |= S[a][b2] != 0.0;
                }
            }
            if (hi < 0) continue;                                    // only unrelated samples in this group
            kmask[g >> 2] |= mask << (8 * (g & 3));
            if (!any_pair) continue;                                 // e.g. the lone last member of a 5-member family
            const int entries = 1 << (2 * (hi + 1));
            for (int idx = 0; idx < entries; idx++) {
                double v = 0.0;
                for (int m = 0; m <= hi; m++)
                    for (int m2 = m + 1; m2 <= hi; m2++) v += 2.0 * S[m][m2] * codeval(idx, m) * codeval(idx, m2);
                lut[((size_t)g << 8) + idx] = v;
            }
        }
        if (lut.size() >= (1ull << 32) - (1u << 20)) { set_error("kinship tables exceed 2^32 entries"); return STAAR_ERR_ARG; }
        // cross terms of a family of 5..16 members: for every pair of chunks the LARGER chunk indexes a table
        // (<= 256 entries) per member of the smaller one:  sum_j code_j * T_j[codes of the other chunk]
        std::vector<std::vector<NullSparse::KinTerm>> word_terms((size_t)ns.x_words);
        for (const Placed &pl : table_placed) {
            const int f = fams[pl.fam].size;
            if (f <= 4) continue;
            const int64_t word = pl.pos0 / 16;
            const int o0 = (int)(pl.pos0 % 16);                      // multiple of 4: chunks coincide with groups
            if (word >= ns.x_words || o0 % 4 != 0) { set_error("internal: large family not on a cross word"); return STAAR_ERR_ARG; }
            const int nch = (f + 3) / 4;
            for (int A = 0; A < nch; A++)
                for (int B = A + 1; B < nch; B++) {
                    const int sa = std::min(4, f - 4 * A), sb = std::min(4, f - 4 * B);
                    const int I = sa >= sb ? A : B, M = sa >= sb ? B : A;          // index chunk, multiplier chunk
                    const int si_ = std::min(4, f - 4 * I), smul = std::min(4, f - 4 * M);
                    for (int j = 4 * M; j < 4 * M + smul; j++) {
                        double sj[4] = {0, 0, 0, 0};
                        bool any = false;
                        for (int m = 0; m < si_; m++) { sj[m] = s_between(pl.pos0 + j, pl.pos0 + 4 * I + m); any |= sj[m] != 0.0; }
                        if (!any) continue;
                        NullSparse::KinTerm t{(uint32_t)lut.size(), (uint8_t)(2 * (o0 + 4 * I)), (uint8_t)((1 << (2 * si_)) - 1),
                                              (uint8_t)(2 * (o0 + j)), 0};
                        for (int idx = 0; idx < (1 << (2 * si_)); idx++) {
                            double v = 0.0;
                            for (int m = 0; m < si_; m++) v += 2.0 * sj[m] * codeval(idx, m);
                            lut.push_back(v);
                        }
                        while (lut.size() % 16) lut.push_back(0.0);
                        word_terms[word].push_back(t);
                    }
                }
        }
        for (int64_t w = 0; w < ns.x_words; w++) {
            wdesc[w].start = (uint32_t)xterms.size();
            wdesc[w].count = (uint32_t)word_terms[w].size();
            xterms.insert(xterms.end(), word_terms[w].begin(), word_terms[w].end());
        }
    }
    ns.lut_entries = (int64_t)lut.size();
    // ---- generic families: upper-triangle records in positions -------------------------------------------
    std::vector<uint8_t> relmask(pstride, 0);
    std::vector<NullSparse::CarrierRec> rel;
    std::vector<NullSparse::OffEntry> up_ent;
    ns.n_generic = 0;
    if (!generic_placed.empty()) {
        rel.resize(n);
        memset(rel.data(), 0, sizeof(NullSparse::CarrierRec) * rel.size());
        for (const Placed &pl : generic_placed) {
            for (int m = 0; m < fams[pl.fam].size; m++) {
                const int64_t pz = pos_of(fam_members[fams[pl.fam].first_member + m]);
                NullSparse::CarrierRec &rr = rel[pz];
                rr.start = (uint32_t)up_ent.size();
                const NullSparse::SampleInfo &si = sinfo[pz];
                for (uint32_t t = 0; t < si.off_count; t++) {
                    const NullSparse::OffEntry &oe = off_ent[si.off_start + t];
                    if (oe.col <= pz) continue;
                    if (rr.count < 2) { rr.val[rr.count] = oe.val; rr.col[rr.count] = oe.col; }
                    up_ent.push_back(oe);
                    rr.count++;
                }
                if (rr.count > 0) relmask[pz >> 2] |= (uint8_t)(3u << (2 * (pz & 3)));
                ns.n_generic++;
            }
        }
    }
    // ---- Y rows in null-model order ----------------------------------------------------------------------
    if (permute) ns.h_perm = perm;

    cudaStream_t st = ctx->stream;
    STAAR_TRY(upload_vec(ctx, &ns.row_ptr, row_ptr));
    STAAR_TRY(upload_vec(ctx, &ns.col, col));
    STAAR_TRY(upload_vec(ctx, &ns.val, val));
    STAAR_TRY(upload_vec(ctx, &ns.diag, diag));
    STAAR_TRY(upload_vec(ctx, &ns.sinfo_p, sinfo));
    {
        std::vector<double> diag_p((size_t)n + 64, 0.0);
        for (int64_t pz = 0; pz < n; pz++) diag_p[pz] = sinfo[pz].diag;
        STAAR_TRY(upload_vec(ctx, &ns.diag_p, diag_p));
        // Tables of the fused path (fused_tc.cu): same layout and offsets as `lut`, but a group's table is indexed by the
        // codes of ALL four samples of the group (unrelated fillers included) and also carries the hom-minor part of
        // g' diag(Sigma_i) g, 2 s_jj [g_j == 2]: one look-up per group serves both terms.  2 * Sigma_i[j,j] by position
        // follows the tables (words outside the table region gather it per carrier).
        lutf = lut;
        for (int64_t g = 0; g < ns.tab_words * 4; g++) {
            const int64_t p0 = g * 4;
            const double *old = &lut[(size_t)g << 8];
            uint32_t relmask4 = 0u;                                  // index bits of the related slots
            for (int a = 0; a < 4; a++) if (sinfo[p0 + a].off_count > 0) relmask4 |= 3u << (2 * a);
            for (int idx = 0; idx < 256; idx++) {
                int ridx = 0;                                        // missing (3) never occurs in an index; map it to 0
                double v = 0.0;
                for (int a = 0; a < 4; a++) {
                    const int c = (idx >> (2 * a)) & 3;
                    if (c != 3) ridx |= c << (2 * a);
                    if (c == 2) v += 2.0 * sinfo[p0 + a].diag;
                }
                lutf[((size_t)g << 8) + idx] = old[ridx & relmask4] + v;
            }
        }
        ns.diag2_off = (int64_t)lutf.size();
        for (size_t pz = 0; pz < diag_p.size(); pz++) lutf.push_back(2.0 * diag_p[pz]);
        // Unrelated samples of a linear mixed model all carry the same diagonal entry: find the longest suffix of
        // positions with one value (the scan multiplies a count by it instead of gathering, scan_packed.cu:hom_vec).
        int64_t first = n;
        while (first > 0 && memcmp(&diag_p[first - 1], &diag_p[n - 1], sizeof(double)) == 0) first--;
        // never inside the table words: their hom-minor carriers are served by the (folded) group tables of the fused path
        const int64_t vec = std::max<int64_t>((first + 63) / 64, (ns.tab_words * 16 + 63) / 64);
        ns.diag_uniform_vec = (n > 0 && vec * 64 < n) ? vec : -1;
        ns.diag_uniform_val = n > 0 ? diag_p[n - 1] : 0.0;
    }
    STAAR_TRY(upload_vec(ctx, &ns.off_ent_p, off_ent));
    STAAR_TRY(upload_vec(ctx, &ns.kmask, kmask));
    STAAR_TRY(upload_vec(ctx, &ns.wdesc, wdesc));
    STAAR_TRY(upload_vec(ctx, &ns.xterms, xterms));
    STAAR_TRY(upload_vec(ctx, &ns.lut, lut));
    STAAR_TRY(upload_vec(ctx, &ns.lutf, lutf));
    STAAR_TRY(upload_vec(ctx, &ns.relmask, relmask));
    if (!rel.empty()) {
        STAAR_TRY(upload_vec(ctx, &ns.rel, rel));
        STAAR_TRY(upload_vec(ctx, &ns.up_ent, up_ent));
    }
    if (permute) {
        STAAR_TRY(upload_vec(ctx, &ns.perm, perm));
        std::vector<int32_t> inv((size_t)n);
        for (int64_t pz = 0; pz < n; pz++) inv[perm[pz]] = (int32_t)pz;
        STAAR_TRY(upload_vec(ctx, &ns.inv_perm, inv));
    }
    STAAR_TRY(upload_vec(ctx, &ns.Y, Y));
    STAAR_CUDA(cudaMalloc(&ns.SX, sizeof(double) * std::max<int64_t>(n * q, 1)));
    STAAR_CUDA(cudaMalloc(&ns.cov, sizeof(double) * std::max<int64_t>(q * q, 1)));
    ns.reg(&ns.SX, sizeof(double) * std::max<int64_t>(n * q, 1));
    ns.reg(&ns.cov, sizeof(double) * std::max<int64_t>(q * q, 1));
    STAAR_CUDA(cudaMemcpyAsync(ns.SX, SX, sizeof(double) * n * q, cudaMemcpyHostToDevice, st));
    STAAR_CUDA(cudaMemcpyAsync(ns.cov, cov, sizeof(double) * q * q, cudaMemcpyHostToDevice, st));
    if (permute) { STAAR_CUDA(cudaMalloc(&ns.Yp, sizeof(double) * (size_t)n * ldY)); ns.reg(&ns.Yp, sizeof(double) * (size_t)n * ldY); }
    else ns.Yp = ns.Y;
    STAAR_CUDA(cudaStreamSynchronize(st));
    ns.fingerprint = fp; ns.res_fingerprint = res_fp;
    ns.set = true;
    STAAR_TRY(refresh_permuted_Y(ctx, Y.data()));
    if (want_sliced) STAAR_TRY(build_sliced_operand(ctx, Y.data()));
    return STAAR_OK;
}

// Yp <- rows of the (natural-order) host operand in null-model order; no-op when the order is the identity
int refresh_permuted_Y(staar_ctx *ctx, const double *hY)
{
    NullSparse &ns = ctx->ns;
    if (!ns.permuted) return STAAR_OK;
    const int64_t n = ns.n;
    const int ldY = ns.ldY;
    std::vector<double> Yp((size_t)n * ldY);
    for (int64_t pz = 0; pz < n; pz++) memcpy(&Yp[(size_t)pz * ldY], hY + (size_t)ns.h_perm[pz] * ldY, sizeof(double) * ldY);
    STAAR_CUDA(cudaMemcpyAsync(ns.Yp, Yp.data(), sizeof(double) * Yp.size(), cudaMemcpyHostToDevice, ctx->stream));
    STAAR_CUDA(cudaStreamSynchronize(ctx->stream));
    return STAAR_OK;
}

// Replica of src's resident sparse null model on dst's device (same process): every registered device array is copied
// device to device; nothing of the host-side build (families, sample order, tables, digit slicing) is repeated.
int nullmodel_sparse_clone(staar_ctx *dst, const staar_ctx *src)
{
    STAAR_CUDA(cudaSetDevice(dst->device));
    dst->ns.release();
    const NullSparse &s = src->ns;
    if (!s.set) { set_error("nullmodel clone: the source context holds no sparse null model"); return STAAR_ERR_STATE; }
    NullSparse &d = dst->ns;
    d = s;                                               // scalars, host vectors; pointers are replaced below
    d.row_ptr = nullptr; d.col = nullptr; d.val = nullptr; d.diag = nullptr; d.perm = nullptr; d.inv_perm = nullptr;
    d.Yp = nullptr; d.kmask = nullptr; d.wdesc = nullptr; d.xterms = nullptr; d.lut = nullptr; d.lutf = nullptr;
    d.sinfo_p = nullptr; d.diag_p = nullptr; d.off_ent_p = nullptr; d.relmask = nullptr; d.rel = nullptr; d.up_ent = nullptr;
    d.SX = nullptr; d.Y = nullptr; d.cov = nullptr; d.Yfrag = nullptr; d.Ytc = nullptr; d.Ypair = nullptr; d.col_scale = nullptr; d.col_total = nullptr;
    d.set = false;
    for (const auto &a : s.arrays) {
        void **slot = reinterpret_cast<void **>(reinterpret_cast<char *>(&d) + a.first);
        void *const from = *reinterpret_cast<void *const *>(reinterpret_cast<const char *>(&s) + a.first);
        if (!from) continue;
        STAAR_CUDA(cudaMalloc(slot, a.second));
        STAAR_CUDA(cudaMemcpyPeerAsync(*slot, dst->device, from, src->device, a.second, dst->stream));
    }
    if (!s.permuted) d.Yp = d.Y;
    STAAR_CUDA(cudaStreamSynchronize(dst->stream));
    d.owned = true;
    d.set = true;
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
    nd.n = n; nd.fingerprint = fp; nd.set = true; nd.res_set = false;
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
    {
        const char *sel = getenv("STAAR_B200_CONTRACT");
        ctx->use_mma_sync = sel && strcmp(sel, "mma") == 0;
        ctx->use_tc_pair = sel && strcmp(sel, "pair") == 0;
        const char *ml = getenv("STAAR_B200_MISSING_LISTS");
        ctx->use_missing_lists = !(ml && atoi(ml) == 0);
        const char *path = getenv("STAAR_B200_PACKED_PATH");
        // default: the two-kernel route — the single-pass kernel is parity-equal but, on B200, bound by the ALU pipe of
        // its scanner warps and not faster; STAAR_B200_PACKED_PATH=fused selects it
        ctx->packed_path = path && strcmp(path, "fused") == 0 ? 0 : 1;
        const char *tiles = getenv("STAAR_B200_FUSED_TILES");
        ctx->fused_tiles = tiles && atoi(tiles) == 1 ? 1 : 2;
    }
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
    ctx->ns.release(); ctx->nd.release(); ctx->nspa.release();
    ctx->G.release(); ctx->L.release(); ctx->aux.release(); ctx->out.release(); ctx->tmp.release();
    ctx->tmp2.release(); ctx->pk.release(); ctx->pk2.release(); ctx->ctr.release(); ctx->cond.release();
    ctx->fscan.release(); ctx->mlist.release(); ctx->msubcnt.release(); ctx->redo.release();
    ctx->plist.release(); ctx->pcnt.release();
    for (auto &e : ctx->ev_pack) if (e) cudaEventDestroy(e);
    for (auto &sl : ctx->slot) {
        if (sl.pinned) cudaFreeHost(sl.pinned);
        sl.res.release();
        if (sl.done) { cudaEventDestroy(sl.done); cudaEventDestroy(sl.ev0); cudaEventDestroy(sl.ev1); }
    }
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    for (int b = 0; b < staar_ctx::kStageBufs; b++) {
        if (ctx->stage[b]) cudaFreeHost(ctx->stage[b]);
        if (ctx->stage_ev[b]) cudaEventDestroy(ctx->stage_ev[b]);
        if (ctx->stage_out[b]) cudaFreeHost(ctx->stage_out[b]);
        if (ctx->stage_out_ev[b]) cudaEventDestroy(ctx->stage_out_ev[b]);
        if (ctx->stage_k_ev[b]) cudaEventDestroy(ctx->stage_k_ev[b]);
    }
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
