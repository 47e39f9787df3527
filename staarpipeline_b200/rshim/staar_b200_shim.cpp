// staar_b200_shim.cpp — the reference-side binding: drop-in replacements for the 13 native functions of
// STAARpipeline, with the SAME C++ signatures the generated glue expects (src/RcppExports.cpp:18,33,50,66,83,
// 96,110,126,141,154,168,184,218), forwarding to the C ABI (include/staar_b200.h).
//
// How it is used: in a checkout of the reference, delete src/matrix_flip_*.cpp and
// src/Individual_Score_Test*.cpp, add this file, and add `-lstaar_b200` to PKG_LIBS in src/Makevars.
// src/RcppExports.cpp, R/RcppExports.R and every R caller (Individual_Analysis(), Sliding_Window(),
// Gene_Centric_*()) stay byte-for-byte unchanged.
//
// R / Rcpp / RcppArmadillo are not installed in the build image, so this file is compile-checked against the
// stand-in headers in oracle/ref_stub/ (tests/test_rshim_compiles.py); it uses only memptr(), n_rows, n_cols,
// size(), values / row_indices / col_ptrs, List::create(Named()=...), which are real RcppArmadillo API.
// [[Rcpp::depends(RcppArmadillo)]]
// [[Rcpp::interfaces(r, cpp)]]
#include <RcppArmadillo.h>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include "staar_b200.h"

using namespace Rcpp;

namespace {

// one context per R session (R is single-threaded; one .Call at a time — SURVEY.md §8b "Threading")
staar_ctx *ctx()
{
    static staar_ctx *c = nullptr;
    if (!c) {
        const char *dev = std::getenv("STAAR_B200_DEVICE");
        if (staar_ctx_create(dev ? std::atoi(dev) : 0, &c) != STAAR_OK)
            throw std::runtime_error(std::string("staar_b200: ") + staar_last_error());
    }
    return c;
}

// STAAR_B200_DEVICES="0,1,2,3": several GPUs of the box behind one handle (staar_b200.h section C); the columns of
// matrix_flip_* and of the dense-G Individual_Score_Test are then sharded over the devices, each moving its share
// over its own PCIe link.  NULL when the variable is unset or names a single device.
staar_multi *multi()
{
    static staar_multi *m = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char *list = std::getenv("STAAR_B200_DEVICES");
        std::vector<int> dev;
        for (const char *c = list; c && *c;) {
            char *end = nullptr;
            const long v = std::strtol(c, &end, 10);
            if (end == c) break;
            dev.push_back((int)v);
            c = (*end == ',') ? end + 1 : end;
        }
        if (dev.size() > 1 && staar_multi_create(dev.data(), (int)dev.size(), &m) != STAAR_OK)
            throw std::runtime_error(std::string("staar_b200: ") + staar_last_error());
    }
    return m;
}

// non-zero status -> C++ exception -> R error through BEGIN_RCPP/END_RCPP (src/RcppExports.cpp:20,30)
void check(int rc)
{
    if (rc != STAAR_OK) throw std::runtime_error(staar_last_error());
}

// the C ABI takes CSC indices as uint64 (ARMA_64BIT_WORD, reference src/Makevars:2): the reinterpret_casts below are only
// valid for 8-byte arma::uword
static_assert(sizeof(arma::uword) == 8, "build with -DARMA_64BIT_WORD=1 (reference src/Makevars)");
template <typename SpMat> const uint64_t *rows(const SpMat &m) { return reinterpret_cast<const uint64_t *>(&m.row_indices[0]); }
template <typename SpMat> const uint64_t *cols(const SpMat &m) { return reinterpret_cast<const uint64_t *>(&m.col_ptrs[0]); }
template <typename SpMat> const double *vals(const SpMat &m) { return &m.values[0]; }

List five(arma::vec &Score, arma::vec &Score_se, arma::vec &pvalue_log, arma::vec &Est, arma::vec &Est_se)
{
    return List::create(Named("Score") = Score, Named("Score_se") = Score_se, Named("pvalue_log") = pvalue_log,
                        Named("Est") = Est, Named("Est_se") = Est_se);
}

struct Out5 {
    arma::vec Score, Score_se, pvalue_log, Est, Est_se;
    explicit Out5(arma::uword p) { Score.zeros(p); Score_se.zeros(p); pvalue_log.zeros(p); Est.zeros(p); Est_se.zeros(p); }
};

}  // namespace

// [[Rcpp::export]]
List matrix_flip_mean(arma::mat G)
{
    arma::vec AF, MAF;
    AF.zeros(G.n_cols); MAF.zeros(G.n_cols);
    if (staar_multi *m = multi()) check(staar_multi_matrix_flip_mean(m, G.memptr(), G.n_rows, G.n_cols, G.memptr(), AF.memptr(), MAF.memptr()));
    else check(staar_matrix_flip_mean(ctx(), G.memptr(), G.n_rows, G.n_cols, G.memptr(), AF.memptr(), MAF.memptr()));
    return List::create(Named("Geno") = G, Named("AF") = AF, Named("MAF") = MAF);
}

// [[Rcpp::export]]
List matrix_flip_minor(arma::mat G)
{
    arma::vec AF, MAF;
    AF.zeros(G.n_cols); MAF.zeros(G.n_cols);
    if (staar_multi *m = multi()) check(staar_multi_matrix_flip_minor(m, G.memptr(), G.n_rows, G.n_cols, G.memptr(), AF.memptr(), MAF.memptr()));
    else check(staar_matrix_flip_minor(ctx(), G.memptr(), G.n_rows, G.n_cols, G.memptr(), AF.memptr(), MAF.memptr()));
    return List::create(Named("Geno") = G, Named("AF") = AF, Named("MAF") = MAF);
}

// [[Rcpp::export]]
List Individual_Score_Test(arma::mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov, arma::vec residuals)
{
    Out5 o(G.n_cols);
    if (staar_multi *m = multi())
        check(staar_multi_individual_score_test(m, G.memptr(), G.n_rows, G.n_cols, vals(Sigma_i), rows(Sigma_i), cols(Sigma_i),
                                                Sigma_iX.memptr(), Sigma_iX.n_cols, cov.memptr(), residuals.memptr(),
                                                o.Score.memptr(), o.Score_se.memptr(), o.pvalue_log.memptr(), o.Est.memptr(),
                                                o.Est_se.memptr()));
    else
    check(staar_individual_score_test(ctx(), G.memptr(), G.n_rows, G.n_cols, vals(Sigma_i), rows(Sigma_i), cols(Sigma_i),
                                      Sigma_iX.memptr(), Sigma_iX.n_cols, cov.memptr(), residuals.memptr(),
                                      o.Score.memptr(), o.Score_se.memptr(), o.pvalue_log.memptr(), o.Est.memptr(),
                                      o.Est_se.memptr()));
    return five(o.Score, o.Score_se, o.pvalue_log, o.Est, o.Est_se);
}

// [[Rcpp::export]]
List Individual_Score_Test_sp(arma::sp_mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov, arma::vec residuals)
{
    Out5 o(G.n_cols);
    check(staar_individual_score_test_sp(ctx(), vals(G), rows(G), cols(G), G.n_rows, G.n_cols, vals(Sigma_i),
                                         rows(Sigma_i), cols(Sigma_i), Sigma_iX.memptr(), Sigma_iX.n_cols, cov.memptr(),
                                         residuals.memptr(), o.Score.memptr(), o.Score_se.memptr(),
                                         o.pvalue_log.memptr(), o.Est.memptr(), o.Est_se.memptr()));
    return five(o.Score, o.Score_se, o.pvalue_log, o.Est, o.Est_se);
}

// [[Rcpp::export]]
List Individual_Score_Test_denseGRM(arma::mat G, arma::mat P, arma::vec residuals)
{
    Out5 o(G.n_cols);
    check(staar_individual_score_test_denseGRM(ctx(), G.memptr(), G.n_rows, G.n_cols, P.memptr(), residuals.memptr(),
                                               o.Score.memptr(), o.Score_se.memptr(), o.pvalue_log.memptr(),
                                               o.Est.memptr(), o.Est_se.memptr()));
    return five(o.Score, o.Score_se, o.pvalue_log, o.Est, o.Est_se);
}

// [[Rcpp::export]]
List Individual_Score_Test_sp_denseGRM(arma::sp_mat G, arma::mat P, arma::vec residuals)
{
    Out5 o(G.n_cols);
    check(staar_individual_score_test_sp_denseGRM(ctx(), vals(G), rows(G), cols(G), G.n_rows, G.n_cols, P.memptr(),
                                                  residuals.memptr(), o.Score.memptr(), o.Score_se.memptr(),
                                                  o.pvalue_log.memptr(), o.Est.memptr(), o.Est_se.memptr()));
    return five(o.Score, o.Score_se, o.pvalue_log, o.Est, o.Est_se);
}

// [[Rcpp::export]]
List Individual_Score_Test_multi(arma::mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov,
                                 arma::vec residuals, int n_pheno = 1)
{
    arma::vec Score, pvalue_log;
    Score.zeros(G.n_cols); pvalue_log.zeros(G.n_cols / n_pheno);
    check(staar_individual_score_test_multi(ctx(), G.memptr(), G.n_rows, G.n_cols, vals(Sigma_i), rows(Sigma_i),
                                            cols(Sigma_i), Sigma_iX.memptr(), Sigma_iX.n_cols, cov.memptr(),
                                            residuals.memptr(), n_pheno, Score.memptr(), pvalue_log.memptr()));
    return List::create(Named("Score") = Score, Named("pvalue_log") = pvalue_log);
}

// [[Rcpp::export]]
List Individual_Score_Test_sp_multi(arma::sp_mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov,
                                    arma::vec residuals, int n_pheno = 1)
{
    arma::vec Score, pvalue_log;
    Score.zeros(G.n_cols); pvalue_log.zeros(G.n_cols / n_pheno);
    check(staar_individual_score_test_sp_multi(ctx(), vals(G), rows(G), cols(G), G.n_rows, G.n_cols, vals(Sigma_i),
                                               rows(Sigma_i), cols(Sigma_i), Sigma_iX.memptr(), Sigma_iX.n_cols,
                                               cov.memptr(), residuals.memptr(), n_pheno, Score.memptr(),
                                               pvalue_log.memptr()));
    return List::create(Named("Score") = Score, Named("pvalue_log") = pvalue_log);
}

// [[Rcpp::export]]
List Individual_Score_Test_denseGRM_multi(arma::mat G, arma::mat P, arma::vec residuals, int n_pheno = 1)
{
    arma::vec Score, pvalue_log;
    Score.zeros(G.n_cols); pvalue_log.zeros(G.n_cols / n_pheno);
    check(staar_individual_score_test_denseGRM_multi(ctx(), G.memptr(), G.n_rows, G.n_cols, P.memptr(),
                                                     residuals.memptr(), n_pheno, Score.memptr(), pvalue_log.memptr()));
    return List::create(Named("Score") = Score, Named("pvalue_log") = pvalue_log);
}

// [[Rcpp::export]]
List Individual_Score_Test_sp_denseGRM_multi(arma::sp_mat G, arma::mat P, arma::vec residuals, int n_pheno = 1)
{
    arma::vec Score, pvalue_log;
    Score.zeros(G.n_cols); pvalue_log.zeros(G.n_cols / n_pheno);
    check(staar_individual_score_test_sp_denseGRM_multi(ctx(), vals(G), rows(G), cols(G), G.n_rows, G.n_cols,
                                                        P.memptr(), residuals.memptr(), n_pheno, Score.memptr(),
                                                        pvalue_log.memptr()));
    return List::create(Named("Score") = Score, Named("pvalue_log") = pvalue_log);
}

// [[Rcpp::export]]
List Individual_Score_Test_cond(arma::mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov, arma::mat X_adj,
                                arma::vec residuals)
{
    Out5 o(G.n_cols);
    check(staar_individual_score_test_cond(ctx(), G.memptr(), G.n_rows, G.n_cols, vals(Sigma_i), rows(Sigma_i),
                                           cols(Sigma_i), Sigma_iX.memptr(), Sigma_iX.n_cols, cov.memptr(),
                                           X_adj.memptr(), X_adj.n_cols, residuals.memptr(), o.Score.memptr(),
                                           o.Score_se.memptr(), o.pvalue_log.memptr(), o.Est.memptr(), o.Est_se.memptr()));
    return five(o.Score, o.Score_se, o.pvalue_log, o.Est, o.Est_se);
}

// [[Rcpp::export]]
List Individual_Score_Test_cond_multi(arma::mat G, arma::sp_mat Sigma_i, arma::mat Sigma_iX, arma::mat cov,
                                      arma::mat X_adj, arma::vec residuals, int n_pheno = 1)
{
    arma::vec Score, pvalue_log;
    Score.zeros(G.n_cols); pvalue_log.zeros(G.n_cols / n_pheno);
    check(staar_individual_score_test_cond_multi(ctx(), G.memptr(), G.n_rows, G.n_cols, vals(Sigma_i), rows(Sigma_i),
                                                 cols(Sigma_i), Sigma_iX.memptr(), Sigma_iX.n_cols, cov.memptr(),
                                                 X_adj.memptr(), X_adj.n_cols, residuals.memptr(), n_pheno,
                                                 Score.memptr(), pvalue_log.memptr()));
    return List::create(Named("Score") = Score, Named("pvalue_log") = pvalue_log);
}

// [[Rcpp::export]]
arma::vec Individual_Score_Test_SPA(arma::mat G, arma::mat XW, arma::mat XXWX_inv, arma::vec residuals, arma::vec muhat,
                                    double tol, int max_iter)
{
    arma::vec res;
    res.ones(G.n_cols);
    check(staar_individual_score_test_SPA(ctx(), G.memptr(), G.n_rows, G.n_cols, XW.memptr(), XXWX_inv.memptr(),
                                          XW.n_rows, residuals.memptr(), muhat.memptr(), tol, max_iter, res.memptr(),
                                          nullptr));
    return res;
}
