"""-m gpu: the CUDA path (through the C ABI, include/staar_b200.h) against the CPU oracle on the same seeded inputs
and against the committed golden vectors (outputs of the reference build).  Tolerances are BASELINE.json's:
Score / variance statistics 1e-10 relative, p-values 1e-8 relative; flips, AF, MAF bit-exact."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp

import cases
from util import assert_spa_pvalues_close, assert_stats_close, rel_err

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
GOLD = np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))


def score_floor(G, residuals):
    """Per variant: |U| below this is cancellation noise for ANY summation order — 1e-3 of that column's typical
    |U| = ||r o g||_2 (a rare variant next to a common one keeps its own, smaller, floor)."""
    G = G.toarray() if sp.issparse(G) else np.asarray(G)
    r = np.asarray(residuals)[:, None] if G.shape[0] == np.size(residuals) else None
    return 1e-3 * np.sqrt(((G * r) ** 2).sum(axis=0)) if r is not None else 0.0


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("fn", ["matrix_flip_mean", "matrix_flip_minor"])
def test_flip_dense_bit_exact(gpu_ctx, oracle, fn):
    G0 = cases.raw_dosage(777, 90)
    got, want = getattr(gpu_ctx, fn)(G0), getattr(oracle, fn)(G0)
    for k in ("Geno", "AF", "MAF"):
        assert np.array_equal(got[k], want[k], equal_nan=True), (fn, k)


def test_flip_dense_real_valued(gpu_ctx, oracle):
    """AI-weighted (non-integer) genotypes: sums are order dependent, so 1e-10 rather than bit-exact."""
    rng = np.random.default_rng(5)
    G0 = cases.raw_dosage(500, 40) * rng.uniform(0.2, 1.7, size=48)[None, :]
    got, want = gpu_ctx.matrix_flip_mean(G0), oracle.matrix_flip_mean(G0)
    for k in ("Geno", "AF", "MAF"):
        assert rel_err(got[k], want[k], 1e-300) < 1e-10


def test_flip_edge_shapes(gpu_ctx, oracle):
    for n, p in ((1, 1), (3, 1), (1, 5), (17, 2)):
        G0 = cases.raw_dosage(n, p, adversarial=False)
        got, want = gpu_ctx.matrix_flip_mean(G0), oracle.matrix_flip_mean(G0)
        for k in ("Geno", "AF", "MAF"):
            assert np.array_equal(got[k], want[k], equal_nan=True)


# ------------------------------------------------------------------ golden vectors (reference build outputs)
@pytest.mark.parametrize("name", sorted({k.split("/")[0] for k in GOLD.files}))
def test_gpu_matches_golden(gpu_ctx, oracle, name):
    fn, kw = make_golden.all_cases(oracle, GOLD)[name]
    assert make_golden.fingerprint_G(kw) == GOLD[f"{name}/__sha256__"].tobytes().decode()
    res = getattr(gpu_ctx, fn)(**kw)
    if not isinstance(res, dict):
        res = {"pvalue": res}
    want = {k: GOLD[f"{name}/{k}"] for k in res}
    if name.startswith("flip"):
        for k in want:
            assert np.array_equal(res[k], want[k], equal_nan=True), k
    elif name == "score_spa":
        assert rel_err(res["pvalue"], want["pvalue"], 1e-300) < 1e-8
    elif "multi" in name:
        assert rel_err(res["Score"], want["Score"], score_floor(kw["G"], kw["residuals"])) < 1e-10
        assert np.all(np.abs(res["pvalue_log"] - want["pvalue_log"]) <= 1e-8 * np.maximum(1.0, np.abs(want["pvalue_log"])))
    else:
        assert_stats_close(res, want, score_floor(kw["G"], kw["residuals"]))


# ------------------------------------------------------------------ score tests vs oracle, more shapes
@pytest.mark.parametrize("n,p,q", [(64, 5, 2), (1000, 130, 13), (2500, 33, 7)])
def test_score_sparse_vs_oracle(gpu_ctx, oracle, n, p, q):
    c, _ = cases.case_sparse(oracle, n=n, p=p, q=q, seed=n)
    assert_stats_close(gpu_ctx.Individual_Score_Test(**c), oracle.Individual_Score_Test(**c), score_floor(c["G"], c["residuals"]))
    cs = dict(c, G=sp.csc_matrix(c["G"]))
    assert_stats_close(gpu_ctx.Individual_Score_Test_sp(**cs), oracle.Individual_Score_Test_sp(**cs), score_floor(c["G"], c["residuals"]))


def test_score_nullmodel_cache_and_residual_swap(gpu_ctx, oracle):
    """Same Sigma_i/Sigma_iX/cov with a new residual vector must not reuse the stale residuals."""
    c, _ = cases.case_sparse(oracle, n=400, p=20, q=4, seed=2)
    a = gpu_ctx.Individual_Score_Test(**c)
    c2 = dict(c, residuals=c["residuals"][::-1].copy())
    b = gpu_ctx.Individual_Score_Test(**c2)
    assert_stats_close(b, oracle.Individual_Score_Test(**c2), score_floor(c2["G"], c2["residuals"]))
    assert_stats_close(gpu_ctx.Individual_Score_Test(**c), a)


def test_score_dense_grm_vs_oracle(gpu_ctx, oracle):
    c, _ = cases.case_dense(oracle, n=700, p=77, q=6, seed=31)
    assert_stats_close(gpu_ctx.Individual_Score_Test_denseGRM(**c), oracle.Individual_Score_Test_denseGRM(**c),
                       score_floor(c["G"], c["residuals"]))


def test_score_cond_singular_raises(gpu_ctx, oracle):
    from staarpipeline_b200.api import StaarError
    c, _ = cases.case_cond(oracle)
    A = np.column_stack([c["X_adj"], np.zeros(c["X_adj"].shape[0])])      # zero column -> exactly singular
    with pytest.raises(StaarError) as e:
        gpu_ctx.Individual_Score_Test_cond(**dict(c, X_adj=A))
    assert e.value.code == 3


def test_dimension_mismatch_raises(gpu_ctx, oracle):
    from staarpipeline_b200.api import StaarError
    c, _ = cases.case_sparse(oracle, n=100, p=4, q=3)
    with pytest.raises(StaarError):
        gpu_ctx.Individual_Score_Test(**dict(c, residuals=c["residuals"][:-1]))


def test_spa_trace_and_parity(gpu_ctx, oracle):
    c, _ = cases.case_spa(oracle, n=6000, p=48, seed=41)
    got, tr = gpu_ctx.Individual_Score_Test_SPA(**c, return_trace=True)
    want, otr = oracle.Individual_Score_Test_SPA(**c, return_trace=True)
    assert_spa_pvalues_close(got, want)
    # same branches taken and same number of Newton iterations as the reference's control flow
    assert np.array_equal(tr[:, 3], otr[:, 3])


# ------------------------------------------------------------------ packed fast path
def _packed_case(oracle, n, p, q, seed, related=True, impute="mean"):
    from staarpipeline_b200 import synth
    nm = synth.nullmodel_sparse_grm(n, seed=seed, q=q, shuffle=True) if related else synth.nullmodel_unrelated(n, seed=seed, q=q)
    G0 = cases.raw_dosage(n, p, seed)
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    flipped = getattr(oracle, f"matrix_flip_{impute}")(G0)
    want = oracle.Individual_Score_Test(flipped["Geno"], nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    want.update(AF=flipped["AF"], MAF=flipped["MAF"])
    return nm, G0, synth.pack_codes(np.asfortranarray(codes)), want, flipped["Geno"]


@pytest.mark.parametrize("n,p,q,related,impute", [(1000, 70, 13, True, "mean"), (777, 40, 5, False, "mean"),
                                                 (2049, 300, 13, True, "minor"), (255, 9, 1, True, "mean")])
def test_packed_score_vs_oracle(gpu_ctx, oracle, n, p, q, related, impute):
    from staarpipeline_b200.api import IMPUTE_MEAN, IMPUTE_MINOR
    nm, G0, packed, want, Gf = _packed_case(oracle, n, p, q, seed=n + p, related=related, impute=impute)
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    got = gpu_ctx.packed_score_host(packed, n, IMPUTE_MEAN if impute == "mean" else IMPUTE_MINOR)
    assert np.array_equal(got["AF"], want["AF"]) and np.array_equal(got["MAF"], want["MAF"])      # bit-exact
    assert_stats_close({k: got[k] for k in ("Score", "Score_se", "pvalue_log", "Est", "Est_se")},
                       {k: want[k] for k in ("Score", "Score_se", "pvalue_log", "Est", "Est_se")},
                       score_floor(Gf, nm.residuals))
    # exact-zero variance branch (monomorphic / all-missing columns) must be hit exactly as in the reference
    zero = want["Score_se"] == 0
    assert zero.any() and np.array_equal(got["Score_se"] == 0, zero)


def test_packed_analysis_filter_and_compaction(gpu_ctx, oracle):
    """The driver's MAF/MAC filter and SPA pre-filter on the device == the same filter applied on the host to the
    unfiltered results (R/Individual_Analysis.R:242,324,290-297); stable order, exact copies."""
    import torch
    from staarpipeline_b200 import api
    from staarpipeline_b200 import synth
    n, p, q = 2000, 3000, 5
    nm = synth.nullmodel_sparse_grm(n, seed=9, q=q, shuffle=True)
    G0 = cases.raw_dosage(n, p, 9)
    packed = synth.pack_codes(np.asfortranarray(np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)))
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    full = gpu_ctx.packed_score_host(packed, n, api.IMPUTE_MEAN)
    dev = torch.device("cuda:0")
    P, stride = packed.shape
    d_in = torch.from_numpy(packed).to(dev); d_cols = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    gpu_ctx.packed_reorder_device(d_in.data_ptr(), d_cols.data_ptr(), stride, n, P)
    idx = torch.full((P,), -1, dtype=torch.int64, device=dev); sel = torch.full((P,), -1, dtype=torch.int64, device=dev)
    outs = torch.zeros((7, P), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    mac_cutoff, pcut = 20.0, 0.05
    nt, nsel = gpu_ctx.packed_analysis_device(d_cols.data_ptr(), stride, n, P, api.IMPUTE_MEAN, mac_cutoff, pcut,
                                              idx.data_ptr(), [outs[i].data_ptr() for i in range(7)], sel.data_ptr())
    keep = full["MAF"] > (mac_cutoff - 0.5) / n / 2
    assert 0 < keep.sum() < P and nt == keep.sum()
    assert np.array_equal(idx.cpu().numpy()[:nt], np.flatnonzero(keep))
    res = outs.cpu().numpy()
    for i, k in enumerate(("AF", "MAF", "Score", "Score_se", "pvalue_log", "Est", "Est_se")):
        assert np.array_equal(res[i, :nt], full[k][keep], equal_nan=True), k
    want_sel = np.flatnonzero(np.exp(-full["pvalue_log"][keep]) < pcut)
    assert nsel == want_sel.size and np.array_equal(sel.cpu().numpy()[:nsel], want_sel)


def test_packed_kinship_paths(gpu_ctx, oracle):
    """Kinship term of the packed path on every route: table families (sizes 2-16, chunked above 4), a family of more
    than 16 members (generic record path), heavy missingness among relatives, mean and minor imputation."""
    from staarpipeline_b200 import synth
    from staarpipeline_b200.api import IMPUTE_MEAN, IMPUTE_MINOR
    n, q = 1500, 4
    rng = np.random.default_rng(123)
    sizes = [2] * 60 + [3] * 40 + [4] * 30 + [5] * 10 + [6] * 10 + [7, 9, 11, 13, 16, 16, 21, 40]
    perm = rng.permutation(n)
    rows, cols, vals, pos = [], [], [], 0
    for s_ in sizes:
        A = rng.normal(size=(s_, s_)); Sig = A @ A.T / s_ + np.eye(s_)
        inv = np.linalg.inv(Sig)
        idx = perm[pos:pos + s_]; pos += s_
        r, c = np.meshgrid(idx, idx, indexing="ij")
        rows.append(r.ravel()); cols.append(c.ravel()); vals.append(inv.ravel())
    for j in perm[pos:]:
        rows.append(np.array([j])); cols.append(np.array([j])); vals.append(np.array([1.0 + rng.random()]))
    Sigma_i = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    X = synth.covariates(n, rng, q)
    Sigma_iX = np.asfortranarray(Sigma_i @ X)
    cov = np.asfortranarray(np.linalg.inv(X.T @ Sigma_iX))
    res = np.asarray(Sigma_i @ rng.normal(size=n)).ravel()
    G0 = cases.raw_dosage(n, 80, seed=5)
    G0[rng.random(G0.shape) < 0.05] = np.nan                      # 5 % missing everywhere, relatives included
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    gpu_ctx.set_nullmodel_sparse(Sigma_i, Sigma_iX, cov, res)
    for impute, flag in (("mean", IMPUTE_MEAN), ("minor", IMPUTE_MINOR)):
        fl = getattr(oracle, f"matrix_flip_{impute}")(G0)
        want = oracle.Individual_Score_Test(fl["Geno"], Sigma_i, Sigma_iX, cov, res)
        got = gpu_ctx.packed_score_host(packed, n, flag)
        assert np.array_equal(got["AF"], fl["AF"]) and np.array_equal(got["MAF"], fl["MAF"])
        assert_stats_close({k: got[k] for k in want}, want, score_floor(fl["Geno"], res))


def test_full_size_chunk_vs_oracle(gpu_ctx, oracle):
    """BASELINE configs[1] sample size (n = 100,000, sparse GRM, q = 13): one chunk through the chunk-driver entry
    point (host FP64 block -> pack -> re-order -> scan / tcgen05 contraction / finaliser) against the oracle, and the
    int32 / raw-byte ingest paths against the FP64 one (bit-identical)."""
    from staarpipeline_b200 import synth
    from staarpipeline_b200.api import IMPUTE_MEAN
    n, p = 100_000, 48
    nm = synth.nullmodel_sparse_grm(n, seed=synth.BASE_SEED, q=13, shuffle=True)
    codes = synth.dosage_codes(n, 1000, p)
    G0 = synth.codes_to_dosage(codes)
    fl = oracle.matrix_flip_mean(G0)
    want = oracle.Individual_Score_Test(fl["Geno"], nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    got = gpu_ctx.individual_analysis_chunk(G0, IMPUTE_MEAN)
    assert np.array_equal(got["AF"], fl["AF"]) and np.array_equal(got["MAF"], fl["MAF"])
    assert_stats_close({k: got[k] for k in want}, want, score_floor(fl["Geno"], nm.residuals))
    # pinned host memory: a share of the variants is packed by the GPU itself, reading the host matrix over PCIe
    import torch
    pinned = torch.from_numpy(np.ascontiguousarray(G0.T)).pin_memory()          # (p, n) C-order == n x p F-order
    alt = gpu_ctx.individual_analysis_chunk(pinned.numpy().T, IMPUTE_MEAN)
    for k in got:
        assert np.array_equal(alt[k], got[k], equal_nan=True), ("pinned", k)
    ci = codes.astype(np.int32)
    for G in (np.where(ci == 3, np.int32(-2147483648), ci).astype(np.int32), np.where(codes == 3, np.uint8(255), codes).astype(np.uint8)):
        alt = gpu_ctx.individual_analysis_chunk(np.asfortranarray(G), IMPUTE_MEAN)
        for k in got:
            assert np.array_equal(alt[k], got[k], equal_nan=True), (G.dtype, k)


def test_host_packer_matches_numpy(oracle):
    from staarpipeline_b200 import api, synth
    G0 = cases.raw_dosage(333, 21)
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    assert np.array_equal(api.pack_dosage_host(G0, threads=3), synth.pack_codes(np.asfortranarray(codes)))
    with pytest.raises(api.StaarError):
        api.pack_dosage_host(G0 * 0.5)


def test_device_generator_matches_host(gpu_ctx):
    import torch
    from staarpipeline_b200 import api, synth
    n, p, first = 1234, 50, 1000
    prm = synth.variant_params(n, first, p)
    stride = api.packed_stride(n)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a).to(dev)
    th, te, tm, sa = t(prm.thr_hom), t(prm.thr_het), t(prm.thr_miss), t(prm.store_alt)
    out = torch.zeros((p, stride), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.synth_packed_device(out.data_ptr(), stride, n, first, p, synth.BASE_SEED, th.data_ptr(), te.data_ptr(),
                                tm.data_ptr(), sa.data_ptr())
    want = synth.pack_codes(synth.dosage_codes(n, first, p, params=prm), stride)
    assert np.array_equal(out.cpu().numpy(), want)
    # generated directly in a given sample order == generated in natural order, then re-ordered
    order = np.random.default_rng(0).permutation(n).astype(np.int32)
    d_order = torch.from_numpy(order).to(dev)
    gpu_ctx.synth_packed_device(out.data_ptr(), stride, n, first, p, synth.BASE_SEED, th.data_ptr(), te.data_ptr(),
                                tm.data_ptr(), sa.data_ptr(), d_order.data_ptr())
    codes = synth.dosage_codes(n, first, p, params=prm)
    assert np.array_equal(out.cpu().numpy(), synth.pack_codes(np.asfortranarray(codes[order]), stride))


def test_resident_order_and_reorder(gpu_ctx, oracle):
    """Resident columns are in the null-model sample order: families inside 16-sample words, a true permutation;
    re-ordering caller-order columns on the device + the resident entry point == the host entry point, bit for bit."""
    import torch
    from staarpipeline_b200 import api, synth
    n, p, q = 3000, 64, 6
    nm, G0, packed, want, Gf = _packed_case(oracle, n, p, q, seed=77, related=True)
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    order = gpu_ctx.nullmodel_sample_order()
    assert np.array_equal(np.sort(order), np.arange(n))
    # every related pair whose members are both in the table region shares a word
    S = sp.coo_matrix(nm.Sigma_i)
    pos = np.empty(n, dtype=np.int64); pos[order] = np.arange(n)
    off = S.row != S.col
    same_word = pos[S.row[off]] // 16 == pos[S.col[off]] // 16
    assert same_word.mean() > 0.99
    host = gpu_ctx.packed_score_host(packed, n, api.IMPUTE_MEAN)
    dev = torch.device("cuda:0")
    P, stride = packed.shape
    d_in = torch.from_numpy(packed).to(dev)
    d_out = torch.zeros_like(d_in)
    outs = torch.zeros((7, P), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.packed_reorder_device(d_in.data_ptr(), d_out.data_ptr(), stride, n, P)
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    assert np.array_equal(d_out.cpu().numpy(), synth.pack_codes(np.asfortranarray(codes[order]), stride))
    gpu_ctx.packed_score_device(d_out.data_ptr(), stride, n, P, api.IMPUTE_MEAN, [outs[i].data_ptr() for i in range(7)])
    gpu_ctx.synchronize()
    res = outs.cpu().numpy()
    for i, k in enumerate(("AF", "MAF", "Score", "Score_se", "pvalue_log", "Est", "Est_se")):
        assert np.array_equal(res[i], host[k], equal_nan=True), k


@pytest.fixture
def gpu_ctx_pair(monkeypatch):
    """a context whose packed contraction runs on CTA pairs (tcgen05 cta_group::2; selected when the context is created)"""
    from staarpipeline_b200.api import Context
    monkeypatch.setenv("STAAR_B200_CONTRACT", "pair")
    ctx = Context(0)
    yield ctx
    ctx.close()


def test_packed_contraction_cta_pair_is_exact_integer(gpu_ctx_pair, oracle):
    """the cta_group::2 variant of K2 (two CTAs share one digit stream) gives the same exact integers"""
    test_packed_contraction_is_exact_integer(gpu_ctx_pair, oracle)


def test_packed_contraction_is_exact_integer(gpu_ctx, oracle):
    """K2 (packed): the INT8-sliced tensor-core contraction must equal, bit for bit, the integer sum
    sum_j code_j * round(Y[j,c] * 2^(54-e_c)) over the RAW 2-bit codes (missing = 3); flip and imputation are
    applied analytically afterwards."""
    import math
    from staarpipeline_b200 import synth
    from staarpipeline_b200.api import IMPUTE_MEAN
    for n, p, q in ((700, 40, 5), (5000, 24, 13)):
        nm, G0, packed, want, Gf = _packed_case(oracle, n, p, q, seed=3 * n, related=True)
        gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
        gpu_ctx.packed_score_host(packed, n, IMPUTE_MEAN)
        P = packed.shape[0]
        d = gpu_ctx.packed_debug_dump(P)
        Y = np.column_stack([nm.residuals, nm.Sigma_iX, nm.Sigma_i.diagonal()])
        codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.int64)     # raw stored codes
        got = d["hi"].astype(object) * (1 << 24) + d["lo"].astype(object)
        for c in range(q + 2):
            e = math.frexp(np.abs(Y[:, c]).max())[1]
            assert d["col_scale"][c] == math.ldexp(1.0, e - 54)
            X = np.array([int(np.rint(math.ldexp(float(y), 54 - e))) for y in Y[:, c]], dtype=object)
            assert list(got[:, c]) == list((codes.astype(object) * X[:, None]).sum(axis=0)), f"column {c}"
        # slot 15: the constant-1 column (scale 1) = exact sum of the raw codes
        assert d["col_scale"][15] == 1.0 and list(got[:, 15]) == list(codes.astype(object).sum(axis=0))
        assert np.array_equal((d["flip"] & 1) != 0, oracle.matrix_flip_mean(G0)["AF"] > 0.5)


@pytest.mark.parametrize("impute", ["mean", "minor"])
def test_packed_spa_stage_vs_oracle(gpu_ctx, oracle, impute):
    """SPA stage on resident 2-bit columns (cfg5's second stage) == Individual_Score_Test_SPA of the reference on the
    flipped / imputed FP64 columns of the same variants (R/Individual_Analysis.R:290-297), with the resident block in
    the sparse null model's sample order and missing genotypes present."""
    import torch
    from staarpipeline_b200 import api, synth
    n, p, q = 5000, 64, 5
    nb = synth.nullmodel_binary(n, seed=31, q=q, intercept=-3.5)
    ns = synth.nullmodel_sparse_grm(n, seed=31, q=q, shuffle=True)
    G0 = cases.raw_dosage(n, p, 31)
    rng = np.random.default_rng(31)
    case_idx = np.flatnonzero(nb.extra["y"] == 1)
    for c in range(6):                                                      # planted signals: Newton-Raphson branch
        take = rng.choice(case_idx, size=max(2, int(len(case_idx) * (0.15 + 0.1 * c))), replace=False)
        G0[take, c] = np.where(np.isnan(G0[take, c]) | (G0[take, c] <= -1), G0[take, c], np.minimum(G0[take, c] + 1.0, 2.0))
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    P, stride = packed.shape
    gpu_ctx.set_nullmodel_sparse(ns.Sigma_i, ns.Sigma_iX, ns.cov, ns.residuals)     # fixes the resident sample order
    gpu_ctx.set_nullmodel_spa(nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat)
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(packed).to(dev); d_cols = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    gpu_ctx.packed_reorder_device(d_in.data_ptr(), d_cols.data_ptr(), stride, n, P)
    sel = np.array([0, 1, 2, 3, 4, 5, 9, 17, 18, 40, P - 1, 7], dtype=np.int64)      # any order, as a gather list
    d_sel = torch.from_numpy(sel).to(dev)
    d_p = torch.full((sel.size,), -1.0, dtype=torch.float64, device=dev)
    d_tr = torch.zeros((sel.size, 4), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    flag = api.IMPUTE_MEAN if impute == "mean" else api.IMPUTE_MINOR
    gpu_ctx.packed_spa_device(d_cols.data_ptr(), stride, n, d_sel.data_ptr(), sel.size, flag, cases.TOL_SPA,
                              cases.MAX_ITER, d_p.data_ptr(), d_tr.data_ptr())
    fl = getattr(oracle, f"matrix_flip_{impute}")(G0)
    want, otr = oracle.Individual_Score_Test_SPA(np.asfortranarray(fl["Geno"][:, sel]), nb.XW, nb.XXWX_inv, nb.residuals,
                                                 nb.muhat, cases.TOL_SPA, cases.MAX_ITER, return_trace=True)
    got = d_p.cpu().numpy()
    assert_spa_pvalues_close(got, want)
    assert np.array_equal(d_tr.cpu().numpy()[:, 3], otr[:, 3])             # same branches, same Newton iteration counts
    # and identical to the frozen-signature entry point on the same FP64 columns
    same = gpu_ctx.Individual_Score_Test_SPA(np.asfortranarray(fl["Geno"][:, sel]), nb.XW, nb.XXWX_inv, nb.residuals,
                                             nb.muhat, cases.TOL_SPA, cases.MAX_ITER)
    assert np.array_equal(got, same)


@pytest.mark.parametrize("n,q,p,planted", [(9000, 5, 10, 3), (40000, 34, 8, 3), (140000, 13, 6, 1)])
def test_spa_cluster_widths_vs_oracle(gpu_ctx, oracle, n, q, p, planted):
    """The SPA kernel splits a variant over a thread-block cluster whose width grows with n (2, 4, 8 CTAs here;
    1 in the small cases above); q = 34 exercises the second covariate slot per lane of the XW·g sweep."""
    c, _ = cases.case_spa(oracle, n=n, p=p, q=q, seed=n % 97, planted=planted)
    got, tr = gpu_ctx.Individual_Score_Test_SPA(**c, return_trace=True)
    want, otr = oracle.Individual_Score_Test_SPA(**c, return_trace=True)
    assert_spa_pvalues_close(got, want)
    assert np.array_equal(tr[:, 3], otr[:, 3])


def _rare_columns(rng, n, p, max_carriers):
    """sparse genotype columns as the rare-variant route sees them: 0/1/2 carriers plus a few mean-imputed values"""
    G = np.zeros((n, p))
    for j in range(p):
        c = int(rng.integers(0, max_carriers + 1)) if j > 2 else (0, 1, max_carriers)[j]
        idx = rng.choice(n, size=c, replace=False)
        G[idx, j] = rng.choice([1.0, 2.0, 0.0123], size=c, p=[0.8, 0.15, 0.05])
    return G


@pytest.mark.parametrize("n,p", [(1500, 24), (9000, 12)])
def test_sp_dense_grm_carrier_route_vs_oracle(gpu_ctx, oracle, n, p):
    """Individual_Score_Test_sp_denseGRM on columns with <= n/8 stored entries takes the carrier form of g'Pg
    (no P*G product); n = 9000 makes the longest carrier list (1125) span two shared-memory tiles.  Compared with the
    oracle and with the dense-G entry point (P*G route) on the same columns."""
    from staarpipeline_b200 import synth
    nd = synth.nullmodel_dense_grm(n, seed=n, q=5, markers=300)
    G = _rare_columns(np.random.default_rng(n), n, p, n // 8)
    Gs = sp.csc_matrix(G)
    got = gpu_ctx.Individual_Score_Test_sp_denseGRM(Gs, nd.P, nd.residuals)
    want = oracle.Individual_Score_Test_sp_denseGRM(Gs, nd.P, nd.residuals)
    assert_stats_close(got, want, score_floor(G, nd.residuals))
    dense = gpu_ctx.Individual_Score_Test_denseGRM(np.asfortranarray(G), nd.P, nd.residuals)
    assert_stats_close(got, dense, score_floor(G, nd.residuals))
    assert got["Score_se"][0] == 0.0 and got["pvalue_log"][0] == 0.0          # empty column: the Cov == 0 branch


def test_sp_dense_grm_multi_carrier_route_vs_oracle(gpu_ctx, oracle):
    """two traits: the k x k blocks need g_a' P g_b between DIFFERENT Kronecker columns (disjoint carrier rows)"""
    c, _ = cases.case_dense_multi(oracle, n=400, p=12, k=2)
    n = 400
    G0 = _rare_columns(np.random.default_rng(3), n, 12, n // 8)
    GK = sp.kron(sp.identity(2, format="csc"), sp.csc_matrix(G0), format="csc")
    kw = dict(c, G=GK)
    got = gpu_ctx.Individual_Score_Test_sp_denseGRM_multi(**kw)
    want = oracle.Individual_Score_Test_sp_denseGRM_multi(**kw)
    assert rel_err(got["Score"], want["Score"], score_floor(GK, kw["residuals"])) < 1e-10
    assert np.all(np.abs(got["pvalue_log"] - want["pvalue_log"]) <= 1e-8 * np.maximum(1.0, np.abs(want["pvalue_log"])))


@pytest.mark.parametrize("related", [True, False])
def test_sp_carrier_route_vs_oracle(gpu_ctx, oracle, related):
    """Individual_Score_Test_sp on columns with <= n/8 stored entries: contraction and g'Sigma_i g over the stored
    entries only (no dense copy of G).  Against the oracle and against the dense-G entry point."""
    from staarpipeline_b200 import synth
    n, p, q = 3000, 40, 7
    nm = synth.nullmodel_sparse_grm(n, seed=77, q=q, shuffle=True) if related else synth.nullmodel_unrelated(n, seed=77, q=q)
    G = _rare_columns(np.random.default_rng(77), n, p, n // 8)
    Gs = sp.csc_matrix(G)
    got = gpu_ctx.Individual_Score_Test_sp(Gs, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    want = oracle.Individual_Score_Test_sp(Gs, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    assert_stats_close(got, want, score_floor(G, nm.residuals))
    dense = gpu_ctx.Individual_Score_Test(np.asfortranarray(G), nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    assert_stats_close(got, dense, score_floor(G, nm.residuals))
    assert got["Score_se"][0] == 0.0 and got["pvalue_log"][0] == 0.0


def test_sp_multi_carrier_route_vs_oracle(gpu_ctx, oracle):
    """three traits, Kronecker-expanded sparse G as R/Individual_Analysis.R:267 builds it: k x k blocks from
    g_a' Sigma_i g_b between columns of different trait blocks"""
    from staarpipeline_b200 import synth
    n, k, pp = 900, 3, 18
    nm = synth.nullmodel_multi(n, k=k, seed=5, q=4)
    G0 = _rare_columns(np.random.default_rng(5), n, pp, n * k // 8 // 2)
    GK = sp.kron(sp.identity(k, format="csc"), sp.csc_matrix(G0), format="csc")
    got = gpu_ctx.Individual_Score_Test_sp_multi(GK, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals, k)
    want = oracle.Individual_Score_Test_sp_multi(GK, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals, k)
    assert rel_err(got["Score"], want["Score"], score_floor(GK, nm.residuals)) < 1e-10
    assert np.all(np.abs(got["pvalue_log"] - want["pvalue_log"]) <= 1e-8 * np.maximum(1.0, np.abs(want["pvalue_log"])))


def test_cond_repeated_calls_reuse_and_invalidate(gpu_ctx, oracle):
    """R/Individual_Analysis_cond.R:183-208 calls Individual_Score_Test_cond once per tested variant (p = 1) with the
    same X_adj: the operands derived from X_adj stay on the device between such calls.  Same results as one batched
    call and as the oracle; a different X_adj or residual vector must not see stale operands."""
    c, _ = cases.case_cond(oracle, n=700, p=9, q=5, n_known=3, seed=21)
    want = oracle.Individual_Score_Test_cond(**c)
    batched = gpu_ctx.Individual_Score_Test_cond(**c)
    floor = score_floor(c["G"], c["residuals"])
    assert_stats_close(batched, want, floor)
    for i in range(c["G"].shape[1]):
        one = gpu_ctx.Individual_Score_Test_cond(**dict(c, G=np.asfortranarray(c["G"][:, i:i + 1])))
        for k in want:
            assert np.array_equal(one[k].ravel(), batched[k].ravel()[i:i + 1]), (k, i)
    c2 = dict(c, X_adj=np.asfortranarray(c["X_adj"][:, 1:]))                   # drop one known variant
    assert_stats_close(gpu_ctx.Individual_Score_Test_cond(**c2), oracle.Individual_Score_Test_cond(**c2), floor)
    c3 = dict(c, residuals=c["residuals"] * 1.5)
    assert_stats_close(gpu_ctx.Individual_Score_Test_cond(**c3), oracle.Individual_Score_Test_cond(**c3), 1.5 * floor)
    assert_stats_close(gpu_ctx.Individual_Score_Test_cond(**c), want, floor)


def test_pipelined_chunks_match_synchronous(gpu_ctx, oracle):
    """staar_chunk_submit / staar_chunk_wait with two chunks in flight == staar_individual_analysis_chunk, for FP64
    (pinned and pageable), int32 and uint8 blocks; a block with a real-valued genotype falls back at wait()."""
    import torch
    from staarpipeline_b200 import synth, api
    n, p, q = 3000, 257, 6
    nm = synth.nullmodel_sparse_grm(n, seed=8, q=q, shuffle=True)
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    blocks = [np.asfortranarray(cases.raw_dosage(n, p, seed)) for seed in (1, 2, 3, 4, 5)]
    want = [gpu_ctx.individual_analysis_chunk(b, api.IMPUTE_MEAN) for b in blocks]
    pinned = [torch.from_numpy(np.ascontiguousarray(b.T)).pin_memory() for b in blocks]       # (p, n) C-order == n x p F-order
    views = [t.numpy().T for t in pinned]
    got = [None] * len(blocks)
    gpu_ctx.chunk_submit(views[0], api.IMPUTE_MEAN, 0, 0)
    for i in range(len(blocks)):
        if i + 1 < len(blocks):
            gpu_ctx.chunk_submit(views[i + 1], api.IMPUTE_MEAN, 0, (i + 1) & 1)
        got[i] = gpu_ctx.chunk_wait(i & 1)
    for g, w in zip(got, want):
        for k in w:
            assert np.array_equal(g[k], w[k], equal_nan=True), k
    # pageable FP64, int32, uint8
    b = blocks[0]
    bi = np.asfortranarray(np.where(np.isnan(b) | (b <= -1), np.iinfo(np.int32).min, b).astype(np.int32))
    bu = np.asfortranarray(np.where(np.isnan(b) | (b <= -1), 255, b).astype(np.uint8))
    for slot, blk in enumerate((b, bi)):
        gpu_ctx.chunk_submit(blk, api.IMPUTE_MEAN, 0, slot)
    for slot in (0, 1):
        g = gpu_ctx.chunk_wait(slot)
        for k in want[0]:
            assert np.array_equal(g[k], want[0][k], equal_nan=True), (slot, k)
    gpu_ctx.chunk_submit(bu, api.IMPUTE_MINOR, 0, 0)
    g = gpu_ctx.chunk_wait(0)
    w = gpu_ctx.individual_analysis_chunk(bu, api.IMPUTE_MINOR)
    for k in w:
        assert np.array_equal(g[k], w[k], equal_nan=True), k
    # real-valued genotype (in the GPU-packed part of a pinned block, and in the host part): FP64 kernels at wait()
    for col in (p - 2, 3):
        t = pinned[1].clone().pin_memory(); v = t.numpy().T
        v[7, col] = 0.37
        w = gpu_ctx.individual_analysis_chunk(v, api.IMPUTE_MEAN)
        gpu_ctx.chunk_submit(v, api.IMPUTE_MEAN, 0, 1)
        g = gpu_ctx.chunk_wait(1)
        for k in w:
            assert np.array_equal(g[k], w[k], equal_nan=True), (col, k)
    # a slot cannot be submitted twice
    gpu_ctx.chunk_submit(views[0], api.IMPUTE_MEAN, 0, 0)
    with pytest.raises(api.StaarError):
        gpu_ctx.chunk_submit(views[1], api.IMPUTE_MEAN, 0, 0)
    gpu_ctx.chunk_wait(0)


def test_resident_entry_points_argument_errors(gpu_ctx):
    """error behaviour of the resident-path entry points: state and argument errors come back as StaarError, never
    as a crash or a silent CPU path"""
    import torch
    from staarpipeline_b200 import api, synth
    n = 512
    nb = synth.nullmodel_binary(n, seed=2, q=3, intercept=-2.0)
    dev = torch.device("cuda:0")
    stride = api.packed_stride(n)
    cols = torch.zeros((4, stride), dtype=torch.uint8, device=dev)
    sel = torch.zeros(4, dtype=torch.int64, device=dev)
    pv = torch.zeros(4, dtype=torch.float64, device=dev)
    fresh = api.Context(0)
    try:
        with pytest.raises(api.StaarError):                                   # SPA operands not set
            fresh.packed_spa_device(cols.data_ptr(), stride, n, sel.data_ptr(), 4, api.IMPUTE_MEAN, 1e-4, 10, pv.data_ptr())
        fresh.set_nullmodel_spa(nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat)
        with pytest.raises(api.StaarError):                                   # n differs from the operands'
            fresh.packed_spa_device(cols.data_ptr(), stride, n + 4, sel.data_ptr(), 4, api.IMPUTE_MEAN, 1e-4, 10, pv.data_ptr())
        with pytest.raises(api.StaarError):                                   # stride not a multiple of 16
            fresh.packed_spa_device(cols.data_ptr(), stride - 8, n, sel.data_ptr(), 4, api.IMPUTE_MEAN, 1e-4, 10, pv.data_ptr())
        fresh.packed_spa_device(cols.data_ptr(), stride, n, sel.data_ptr(), 0, api.IMPUTE_MEAN, 1e-4, 10, pv.data_ptr())   # empty: no-op
        fresh.packed_spa_device(cols.data_ptr(), stride, n, sel.data_ptr(), 4, api.IMPUTE_MEAN, 1e-4, 10, pv.data_ptr())
        assert np.array_equal(pv.cpu().numpy(), np.ones(4))                   # monomorphic columns: p = 1
        with pytest.raises(api.StaarError):                                   # XW / XXWX_inv shapes disagree
            fresh.set_nullmodel_spa(nb.XW, nb.XXWX_inv[:-1], nb.residuals, nb.muhat)
        with pytest.raises(api.StaarError):                                   # chunk driver without a null model
            fresh.chunk_submit(np.zeros((n, 3), order="F"), api.IMPUTE_MEAN, 0, 0)
        with pytest.raises(api.StaarError):
            fresh.chunk_wait(0)
    finally:
        fresh.close()


@pytest.mark.parametrize("fn", ["matrix_flip_mean", "matrix_flip_minor"])
def test_flip_large_block_pipelined_transfers(gpu_ctx, oracle, fn):
    """a block above 64 MB in pageable memory takes the column-block pipeline (upload, flip kernel and download of
    different blocks in flight at once, staged through pinned rings): still bit-exact, also when the output aliases
    the input as in the Rcpp shim"""
    n, p = 20011, 530                                                       # 85 MB, odd n, several staging blocks
    G0 = cases.raw_dosage(n, p, 123)
    want = getattr(oracle, fn)(G0)
    got = getattr(gpu_ctx, fn)(G0)
    for k in ("Geno", "AF", "MAF"):
        assert np.array_equal(got[k], want[k], equal_nan=True), k


@pytest.mark.parametrize("impute", ["mean", "minor"])
def test_packed_dense_grm_vs_oracle(gpu_ctx, oracle, impute):
    """config 3 on resident 2-bit columns: flip / impute from the code counts, carrier form of g'Pg for variants with
    <= n/8 non-zero genotypes, decoded FP64 columns through P*G for the rest == matrix_flip_* followed by
    Individual_Score_Test_denseGRM of the reference (adversarial columns included: all missing, monomorphic, AF = 0.5)"""
    import torch
    from staarpipeline_b200 import api, synth
    n, p = 3000, 96
    nd = synth.nullmodel_dense_grm(n, seed=12, q=5, markers=300)
    G0 = cases.raw_dosage(n, p, 12)
    G0[:, 5:40] = np.where(np.random.default_rng(1).random((n, 35)) < 0.97, 2.0, G0[:, 5:40])     # rare after the flip
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    P_, stride = packed.shape
    gpu_ctx.set_nullmodel_dense(nd.P, nd.residuals)
    dev = torch.device("cuda:0")
    d_cols = torch.from_numpy(packed).to(dev)
    outs = torch.zeros((7, P_), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    flag = api.IMPUTE_MEAN if impute == "mean" else api.IMPUTE_MINOR
    gpu_ctx.packed_score_dense_device(d_cols.data_ptr(), stride, n, P_, flag, [outs[i].data_ptr() for i in range(7)])
    res = outs.cpu().numpy()
    fl = getattr(oracle, f"matrix_flip_{impute}")(G0)
    want = oracle.Individual_Score_Test_denseGRM(fl["Geno"], nd.P, nd.residuals)
    assert np.array_equal(res[0], fl["AF"]) and np.array_equal(res[1], fl["MAF"])
    nnz = (fl["Geno"] != 0).sum(axis=0)
    assert (nnz <= n // 8).sum() >= 10 and (nnz > n // 8).sum() >= 10                # both routes are exercised
    got = dict(zip(("Score", "Score_se", "pvalue_log", "Est", "Est_se"), res[2:]))
    assert_stats_close(got, want, score_floor(fl["Geno"], nd.residuals))


@pytest.mark.parametrize("impute", ["mean", "minor"])
def test_packed_multi_trait_vs_oracle(gpu_ctx, oracle, impute):
    """config 4 on resident 2-bit columns of G0: flip / impute from the code counts, Kronecker-expanded carrier lists on
    the device == matrix_flip_* then Individual_Score_Test_sp_multi on Diagonal(k) (x) Geno (related samples, missing
    genotypes, common and rare variants, adversarial columns)"""
    import torch
    from staarpipeline_b200 import api, synth
    n, k, pp = 900, 3, 40
    nm = synth.nullmodel_multi(n, k=k, seed=15, q=4)
    G0 = cases.raw_dosage(n, pp, 15)
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    P_, stride = packed.shape
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
    dev = torch.device("cuda:0")
    d_cols = torch.from_numpy(packed).to(dev)
    af = torch.zeros(P_, dtype=torch.float64, device=dev); maf = torch.zeros_like(af)
    score = torch.zeros(P_ * k, dtype=torch.float64, device=dev); pl = torch.zeros(P_, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    flag = api.IMPUTE_MEAN if impute == "mean" else api.IMPUTE_MINOR
    gpu_ctx.packed_score_multi_device(d_cols.data_ptr(), stride, n, P_, k, flag, af.data_ptr(), maf.data_ptr(), score.data_ptr(),
                                      pl.data_ptr())
    fl = getattr(oracle, f"matrix_flip_{impute}")(G0)
    GK = sp.kron(sp.identity(k, format="csc"), sp.csc_matrix(fl["Geno"]), format="csc")
    want = oracle.Individual_Score_Test_sp_multi(GK, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals, k)
    assert np.array_equal(af.cpu().numpy(), fl["AF"]) and np.array_equal(maf.cpu().numpy(), fl["MAF"])
    assert rel_err(score.cpu().numpy(), want["Score"], score_floor(GK, nm.residuals)) < 1e-10
    got_pl = pl.cpu().numpy()
    assert np.all(np.abs(got_pl - want["pvalue_log"]) <= 1e-8 * np.maximum(1.0, np.abs(want["pvalue_log"])))


def test_packed_conditional_test_vs_oracle(gpu_ctx, oracle):
    """conditional test on selected resident 2-bit columns (null-model sample order, missing genotypes) ==
    matrix_flip_mean then Individual_Score_Test_cond on the same variants"""
    import torch
    from staarpipeline_b200 import api, synth
    n, p, q = 2000, 60, 5
    nm = synth.nullmodel_sparse_grm(n, seed=19, q=q, shuffle=True)
    G0 = cases.raw_dosage(n, p, 19, adversarial=False)
    fl = oracle.matrix_flip_mean(G0)
    X_adj, res_c = synth.conditional_design(nm, fl["Geno"][:, :4], "optimal")       # four known variants + covariates
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    P_, stride = packed.shape
    gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, res_c)
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(packed).to(dev); d_cols = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    gpu_ctx.packed_reorder_device(d_in.data_ptr(), d_cols.data_ptr(), stride, n, P_)
    sel = np.array([10, 4, 33, 59, 21, 22, 47], dtype=np.int64)
    d_sel = torch.from_numpy(sel).to(dev)
    torch.cuda.synchronize()
    got = gpu_ctx.packed_cond(d_cols.data_ptr(), stride, n, d_sel.data_ptr(), sel.size, api.IMPUTE_MEAN, X_adj)
    want = oracle.Individual_Score_Test_cond(np.asfortranarray(fl["Geno"][:, sel]), nm.Sigma_i, nm.Sigma_iX, nm.cov, X_adj, res_c)
    assert_stats_close(got, want, score_floor(fl["Geno"][:, sel], res_c))
    again = gpu_ctx.packed_cond(d_cols.data_ptr(), stride, n, d_sel.data_ptr(), 1, api.IMPUTE_MEAN, X_adj)     # cached operands
    for k_ in want:
        assert np.array_equal(again[k_], got[k_][:1])


def test_resident_paths_odd_sizes(gpu_ctx, oracle):
    """ragged shapes for the resident dense-GRM and multi-trait paths: n not a multiple of 4 / 16 / 64, one and three
    columns (tail words of the 2-bit columns, carrier lists ending mid-word)"""
    import torch
    from staarpipeline_b200 import api, synth
    dev = torch.device("cuda:0")
    for n, p in ((1003, 3), (517, 1)):
        G0 = cases.raw_dosage(n, max(p, 6), n)[:, :p]
        codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
        packed = synth.pack_codes(np.asfortranarray(codes))
        P_, stride = packed.shape
        d_cols = torch.from_numpy(packed).to(dev)
        fl = oracle.matrix_flip_mean(G0)
        # dense GRM
        nd = synth.nullmodel_dense_grm(n, seed=n, q=3, markers=100)
        gpu_ctx.set_nullmodel_dense(nd.P, nd.residuals)
        outs = torch.zeros((7, P_), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        gpu_ctx.packed_score_dense_device(d_cols.data_ptr(), stride, n, P_, api.IMPUTE_MEAN, [outs[i].data_ptr() for i in range(7)])
        res = outs.cpu().numpy()
        want = oracle.Individual_Score_Test_denseGRM(fl["Geno"], nd.P, nd.residuals)
        assert np.array_equal(res[0], fl["AF"]) and np.array_equal(res[1], fl["MAF"])
        assert_stats_close(dict(zip(("Score", "Score_se", "pvalue_log", "Est", "Est_se"), res[2:])), want,
                           score_floor(fl["Geno"], nd.residuals))
        # two traits
        k = 2
        nm = synth.nullmodel_multi(n, k=k, seed=n, q=3)
        gpu_ctx.set_nullmodel_sparse(nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals)
        af = torch.zeros(P_, dtype=torch.float64, device=dev); maf = torch.zeros_like(af)
        score = torch.zeros(P_ * k, dtype=torch.float64, device=dev); pl = torch.zeros(P_, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        gpu_ctx.packed_score_multi_device(d_cols.data_ptr(), stride, n, P_, k, api.IMPUTE_MEAN, af.data_ptr(), maf.data_ptr(),
                                          score.data_ptr(), pl.data_ptr())
        GK = sp.kron(sp.identity(k, format="csc"), sp.csc_matrix(fl["Geno"]), format="csc")
        wantm = oracle.Individual_Score_Test_sp_multi(GK, nm.Sigma_i, nm.Sigma_iX, nm.cov, nm.residuals, k)
        assert np.array_equal(af.cpu().numpy(), fl["AF"])
        assert rel_err(score.cpu().numpy(), wantm["Score"], score_floor(GK, nm.residuals)) < 1e-10
        assert np.all(np.abs(pl.cpu().numpy() - wantm["pvalue_log"]) <= 1e-8 * np.maximum(1.0, np.abs(wantm["pvalue_log"])))


def test_config5_resident_flow_vs_reference_flow(gpu_ctx, oracle):
    """the whole config-5 chunk on the device (score test + MAC filter + SPA pre-filter, then the SPA stage on the
    selected resident columns) against the R driver's flow on the host (R/Individual_Analysis.R:186-300):
    matrix_flip_mean -> MAC filter -> Individual_Score_Test -> p < 0.05 -> Individual_Score_Test_SPA -> overwrite"""
    import torch
    from staarpipeline_b200 import api, synth
    n, p, q = 6000, 400, 5
    nb = synth.nullmodel_binary(n, seed=77, q=q, intercept=-3.0)
    G0 = cases.raw_dosage(n, p, 77)
    codes = np.where(np.isnan(G0) | (G0 <= -1), 3, G0).astype(np.uint8)
    packed = synth.pack_codes(np.asfortranarray(codes))
    P_, stride = packed.shape
    gpu_ctx.set_nullmodel_sparse(nb.Sigma_i, nb.Sigma_iX, nb.cov, nb.residuals)
    gpu_ctx.set_nullmodel_spa(nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat)
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(packed).to(dev); d_cols = torch.zeros_like(d_in)
    torch.cuda.synchronize()
    gpu_ctx.packed_reorder_device(d_in.data_ptr(), d_cols.data_ptr(), stride, n, P_)
    idx = torch.zeros(P_, dtype=torch.int64, device=dev); sel = torch.zeros(P_, dtype=torch.int64, device=dev)
    outs = torch.zeros((7, P_), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    mac_cutoff = 20.0
    nt, nsel = gpu_ctx.packed_analysis_device(d_cols.data_ptr(), stride, n, P_, api.IMPUTE_MEAN, mac_cutoff, 0.05, idx.data_ptr(),
                                              [outs[i].data_ptr() for i in range(7)], sel.data_ptr())
    cols = idx[sel[:nsel]].contiguous()
    pv = torch.zeros(max(nsel, 1), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.packed_spa_device(d_cols.data_ptr(), stride, n, cols.data_ptr(), nsel, api.IMPUTE_MEAN, cases.TOL_SPA, cases.MAX_ITER,
                              pv.data_ptr())
    got_p = np.exp(-outs[4, :nt].cpu().numpy())
    got_p[sel[:nsel].cpu().numpy()] = pv[:nsel].cpu().numpy()
    # the host flow
    fl = oracle.matrix_flip_mean(G0)
    keep = fl["MAF"] > (mac_cutoff - 0.5) / n / 2
    Gk = np.asfortranarray(fl["Geno"][:, keep])
    st = oracle.Individual_Score_Test(Gk, nb.Sigma_i, nb.Sigma_iX, nb.cov, nb.residuals)
    want_p = np.exp(-st["pvalue_log"])
    hit = want_p < 0.05
    want_p[hit] = oracle.Individual_Score_Test_SPA(np.asfortranarray(Gk[:, hit]), nb.XW, nb.XXWX_inv, nb.residuals, nb.muhat,
                                                   cases.TOL_SPA, cases.MAX_ITER)
    assert nt == keep.sum() and np.array_equal(idx[:nt].cpu().numpy(), np.flatnonzero(keep))
    assert nsel == hit.sum() and nsel >= 5 and np.array_equal(sel[:nsel].cpu().numpy(), np.flatnonzero(hit))
    assert_spa_pvalues_close(got_p[hit], want_p[hit])
    assert rel_err(got_p[~hit], want_p[~hit], 1e-300) < 1e-8
