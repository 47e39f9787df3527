// multi.cu — several GPUs of one box behind one handle of the C ABI: one staar_ctx per device,
// variants of a call sharded contiguously across the devices and served by one host thread per device, the sparse null
// model built once (host pass on the first device) and replicated device to device.  There is no data-path
// collective: every variant is independent given the null model (R/Individual_Analysis.R loops over variant chunks).
// The Rcpp shim opens such a handle when STAAR_B200_DEVICES lists more than one device (rshim/staar_b200_shim.cpp).
#include <string>
#include <thread>
#include <vector>
#include "common.cuh"

struct staar_multi {
    std::vector<staar_ctx *> ctx;
};

namespace staar {
namespace {

// run f(i, a, b) for shard i = columns [a, b) on one host thread per device; first error wins
template <typename F> int for_shards(staar_multi *m, int64_t p, F f)
{
    const int nd = (int)m->ctx.size();
    std::vector<int> rc(nd, STAAR_OK);
    std::vector<std::string> msg(nd);
    std::vector<std::thread> th;
    for (int i = 0; i < nd; i++) {
        const int64_t a = p * i / nd, b = p * (i + 1) / nd;
        th.emplace_back([&, i, a, b]() {
            if (b > a || i == 0) rc[i] = f(i, a, b);
            if (rc[i] != STAAR_OK) msg[i] = staar_last_error();      // the error text lives in this worker thread
        });
    }
    for (auto &t : th) t.join();
    for (int i = 0; i < nd; i++)
        if (rc[i] != STAAR_OK) { set_error("device %d: %s", m->ctx[i]->device, msg[i].c_str()); return rc[i]; }
    return STAAR_OK;
}

// the null model the frozen signatures pass on every call: hashed and built once on the first device, cloned to the rest
int ensure_sparse_all(staar_multi *m, int64_t n, int64_t q, const double *Sv, const uint64_t *Sr, const uint64_t *Sc,
                      const double *SX, const double *cov, const double *res)
{
    STAAR_TRY(ensure_sparse_shared(m->ctx[0], n, q, Sv, Sr, Sc, SX, cov, res));
    const NullSparse &s = m->ctx[0]->ns;
    for (size_t i = 1; i < m->ctx.size(); i++) {
        const NullSparse &d = m->ctx[i]->ns;
        if (d.set && d.fingerprint == s.fingerprint && d.res_fingerprint == s.res_fingerprint && d.sliced == s.sliced) continue;
        STAAR_TRY(nullmodel_sparse_clone(m->ctx[i], m->ctx[0]));
    }
    return STAAR_OK;
}

}  // namespace
}  // namespace staar

using namespace staar;

extern "C" {

int staar_multi_create(const int *devices, int n_devices, staar_multi **out)
{
    if (!out || !devices || n_devices < 1) { set_error("staar_multi_create: bad arguments"); return STAAR_ERR_ARG; }
    *out = nullptr;
    staar_multi *m = new staar_multi();
    for (int i = 0; i < n_devices; i++) {
        staar_ctx *c = nullptr;
        const int rc = staar_ctx_create(devices[i], &c);
        if (rc != STAAR_OK) {
            for (staar_ctx *x : m->ctx) staar_ctx_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    *out = m;
    return STAAR_OK;
}

void staar_multi_destroy(staar_multi *m)
{
    if (!m) return;
    for (staar_ctx *c : m->ctx) staar_ctx_destroy(c);
    delete m;
}

int staar_multi_size(const staar_multi *m) { return m ? (int)m->ctx.size() : 0; }

staar_ctx *staar_multi_ctx(staar_multi *m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

int staar_multi_nullmodel_set_sparse(staar_multi *m, int64_t n, int64_t q, const double *Sv, const uint64_t *Sr,
                                     const uint64_t *Sc, const double *SX, const double *cov, const double *res)
{
    if (!m) { set_error("null handle"); return STAAR_ERR_ARG; }
    STAAR_TRY(staar_nullmodel_set_sparse(m->ctx[0], n, q, Sv, Sr, Sc, SX, cov, res));
    return for_shards(m, (int64_t)m->ctx.size(), [&](int i, int64_t, int64_t) {
        return i == 0 ? STAAR_OK : nullmodel_sparse_clone(m->ctx[i], m->ctx[0]);
    });
}

int staar_multi_individual_analysis_chunk(staar_multi *m, const double *G, int64_t n, int64_t p, int impute, int threads,
                                          double *AF, double *MAF, double *Score, double *Score_se, double *pvalue_log,
                                          double *Est, double *Est_se)
{
    if (!m) { set_error("null handle"); return STAAR_ERR_ARG; }
    const int nd = (int)m->ctx.size();
    unsigned hw = std::thread::hardware_concurrency();
    const int per = threads > 0 ? std::max(1, threads / nd) : std::max(1, (int)(hw ? hw : 8) / nd);
    return for_shards(m, p, [&](int i, int64_t a, int64_t b) {
        return staar_individual_analysis_chunk(m->ctx[i], G + (size_t)a * n, n, b - a, impute, per, AF + a, MAF + a, Score + a,
                                               Score_se + a, pvalue_log + a, Est + a, Est_se + a);
    });
}

int staar_multi_matrix_flip_mean(staar_multi *m, const double *G, int64_t n, int64_t p, double *Geno_out, double *AF, double *MAF)
{
    if (!m) { set_error("null handle"); return STAAR_ERR_ARG; }
    return for_shards(m, p, [&](int i, int64_t a, int64_t b) {
        return staar_matrix_flip_mean(m->ctx[i], G + (size_t)a * n, n, b - a, Geno_out ? Geno_out + (size_t)a * n : nullptr, AF + a, MAF + a);
    });
}

int staar_multi_matrix_flip_minor(staar_multi *m, const double *G, int64_t n, int64_t p, double *Geno_out, double *AF, double *MAF)
{
    if (!m) { set_error("null handle"); return STAAR_ERR_ARG; }
    return for_shards(m, p, [&](int i, int64_t a, int64_t b) {
        return staar_matrix_flip_minor(m->ctx[i], G + (size_t)a * n, n, b - a, Geno_out ? Geno_out + (size_t)a * n : nullptr, AF + a, MAF + a);
    });
}

int staar_multi_individual_score_test(staar_multi *m, const double *G, int64_t n, int64_t p, const double *Sv,
                                      const uint64_t *Sr, const uint64_t *Sc, const double *SX, int64_t q, const double *cov,
                                      const double *res, double *Score, double *Score_se, double *pvalue_log, double *Est,
                                      double *Est_se)
{
    if (!m) { set_error("null handle"); return STAAR_ERR_ARG; }
    STAAR_TRY(ensure_sparse_all(m, n, q, Sv, Sr, Sc, SX, cov, res));       // one host build, replicas by device copies
    return for_shards(m, p, [&](int i, int64_t a, int64_t b) {
        return staar_individual_score_test(m->ctx[i], G + (size_t)a * n, n, b - a, Sv, Sr, Sc, SX, q, cov, res, Score + a,
                                           Score_se + a, pvalue_log + a, Est + a, Est_se + a);
    });
}

}  // extern "C"
