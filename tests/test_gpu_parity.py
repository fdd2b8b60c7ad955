"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI
(ctypes -> libknn_b200.so), against the CPU oracle on the same seeded inputs, against the
committed golden vectors of the reference's own tests, and through size-independent
properties at larger sizes.  Bar: ids identical except fp64-arbitrated near-ties
(oracle/parity.py), distances within 1e-5 relative."""
import json

import numpy as np
import pytest

from oracle import flat_oracle as fo
from oracle.evaluate import Fixture, evaluate_ids
from oracle.parity import check_parity

pytestmark = pytest.mark.gpu

IP, L2 = 0, 1


@pytest.fixture(scope="module")
def knn():
    import knn_b200

    assert knn_b200._lib.load().knn_device_count() >= 1, "no CUDA device: the CUDA path cannot run"
    return knn_b200


def _data(nq, nb, d, seed, normalize=True, scale=1.0):
    rng = np.random.default_rng(seed)
    xb = (rng.standard_normal((nb, d)) * scale).astype(np.float32)
    xq = (rng.standard_normal((nq, d)) * scale).astype(np.float32)
    if normalize:
        fo.normalize_L2(xb)
        fo.normalize_L2(xq)
    return xq, xb


def _search(knn, xq, xb, k, metric, path=0, **params):
    idx = knn.IndexFlat(xb.shape[1], metric)
    idx.set_param("path", path)
    for name, v in params.items():
        idx.set_param(name, v)
    idx.train(xb)
    idx.add(xb)
    assert idx.ntotal == xb.shape[0]
    D, I = idx.search(xq, k)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (xq.shape[0], k) == I.shape
    assert D.flags.c_contiguous and I.flags.c_contiguous
    return D, I, idx


# ---------------------------------------------------------------------------------------------
def test_normalize_l2_matches_oracle(knn):
    rng = np.random.default_rng(1)
    for shape in [(17, 1024), (5, 33), (1000, 128), (3, 7)]:
        x = (rng.standard_normal(shape) * 3).astype(np.float32)
        x[1] = 0
        ref = x.copy()
        fo.normalize_L2(ref)
        got = x.copy()
        assert knn.normalize_L2(got) is None  # in place, like faiss
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7)
        assert np.array_equal(got[1], np.zeros(shape[1], np.float32))
    with pytest.raises(TypeError):
        knn.normalize_L2(x.astype(np.float64))
    with pytest.raises(ValueError):
        knn.normalize_L2(np.asfortranarray(x))


def _faiss_search_flow(knn, fx_dir, k, path=0):
    """seqvec_search/main.py:29-50 restated on top of the shim."""
    queries = np.load(fx_dir / "test.npy")
    haystack = np.load(fx_dir / "train.npy")
    knn.normalize_L2(queries)
    knn.normalize_L2(haystack)
    index = knn.IndexFlat(haystack.shape[1], knn.METRIC_INNER_PRODUCT)
    index.set_param("path", path)
    index.train(haystack)
    index.add(haystack)
    scores, result = index.search(queries, k)
    return result, scores


def test_reference_known_answer_small_random(knn, golden_dir, expected):
    """/root/reference/tests/test_main.py:10-18 through the CUDA path."""
    I, D = _faiss_search_flow(knn, golden_dir / "small-random", 5)
    auc1s, tps = evaluate_ids(Fixture(golden_dir / "small-random"), I)
    assert auc1s == [1.0, 1 / 3, 2 / 3, 0.0, 0.0, 1 / 3]
    assert tps == [1.0, 2 / 3, 2 / 3, 1.0, 1.0, 1.0]
    assert np.array_equal(I, expected["small-random.ip.k5.I"])
    np.testing.assert_allclose(D, expected["small-random.ip.k5.D"], rtol=1e-5)


def test_reference_known_answer_pfam_20_10(knn, golden_dir, expected):
    """/root/reference/tests/test_main.py:21-27 (kNN half) through the CUDA path."""
    I, D = _faiss_search_flow(knn, golden_dir / "pfam-20-10", 10)
    auc1s, tps = evaluate_ids(Fixture(golden_dir / "pfam-20-10"), I)
    assert np.mean(auc1s) == 0.871
    assert np.mean(tps) == 0.91
    assert np.array_equal(I, expected["pfam-20-10.ip.k10.I"])
    np.testing.assert_allclose(D, expected["pfam-20-10.ip.k10.D"], rtol=1e-5)


@pytest.mark.parametrize("name", ["pfam-20-10-sum", "pfam-20-dist"])
def test_regression_fixtures(knn, golden_dir, expected, name):
    summary = json.loads((golden_dir / "expected.json").read_text())
    I, D = _faiss_search_flow(knn, golden_dir / name, 13)
    auc1s, tps = evaluate_ids(Fixture(golden_dir / name), I)
    assert np.mean(auc1s) == summary[name]["mean_auc1"]
    assert np.mean(tps) == summary[name]["mean_tp"]
    assert np.array_equal(I, expected[f"{name}.ip.k13.I"])
    np.testing.assert_allclose(D, expected[f"{name}.ip.k13.D"], rtol=1e-5)


@pytest.mark.parametrize("mname,metric", [("ip", IP), ("l2", L2)])
def test_cath_search_flow(knn, golden_dir, expected, mname, metric):
    """cath/search.py:13-26 restated: all-vs-all with hits+1, both metrics, self hit dropped."""
    emb = np.load(golden_dir / "pfam-20-10/train.npy")
    x = emb
    if metric == IP:
        x = emb.copy()
        knn.normalize_L2(x)
    index = knn.IndexFlat(x.shape[1], metric)
    index.add(x)
    scores, results = index.search(x, 11)
    I, D = results[:, 1:], scores[:, 1:]
    assert np.array_equal(results[:, 0], np.arange(x.shape[0]))  # column 0 is the self hit
    ref_x = emb.copy()
    if metric == IP:
        fo.normalize_L2(ref_x)
    check_parity(np.ascontiguousarray(D), np.ascontiguousarray(I),
                 expected[f"cath-search.pfam-20-10-train.{mname}.hits10.D"],
                 expected[f"cath-search.pfam-20-10-train.{mname}.hits10.I"], ref_x, ref_x, metric, max_excused_frac=0.01)


@pytest.mark.parametrize("mname,metric", [("ip", IP), ("l2", L2)])
def test_k_larger_than_ntotal_pads(knn, golden_dir, expected, mname, metric):
    xb = np.load(golden_dir / "small-random/train.npy")
    xq = np.load(golden_dir / "small-random/test.npy")
    D, I, _ = _search(knn, xq, xb, 16, metric)
    assert np.array_equal(I, expected[f"small-random.raw.{mname}.k16.I"])
    ref_D = expected[f"small-random.raw.{mname}.k16.D"]
    assert np.array_equal(D[:, 11:], ref_D[:, 11:])  # -FLT_MAX / +FLT_MAX
    np.testing.assert_allclose(D[:, :11], ref_D[:, :11], rtol=1e-5)


def test_empty_cases(knn):
    idx = knn.IndexFlat(16, IP)
    D, I = idx.search(np.zeros((3, 16), np.float32), 4)  # empty index: all padding
    assert (I == -1).all() and (D == -np.finfo(np.float32).max).all()
    idx.add(np.ones((2, 16), np.float32))
    D, I = idx.search(np.zeros((0, 16), np.float32), 4)  # no queries
    assert D.shape == (0, 4) and I.shape == (0, 4)
    with pytest.raises(AssertionError):
        idx.add(np.ones((2, 17), np.float32))
    with pytest.raises(ValueError):
        idx.search(np.ones((2, 16), np.float32), 0)
    with pytest.raises(Exception):
        idx.search(np.ones((2, 16), np.float32), knn.MAX_K + 1)
    with pytest.raises(NotImplementedError):
        knn.IndexLSH(16, 8)


# ---------------------------------------------------------------------------------------------
EXACT_CASES = [
    # nq, nb, d, k, normalize
    (6, 11, 1024, 5, True),
    (1, 1, 8, 1, False),
    (37, 5000, 64, 10, True),
    (9, 40000, 100, 100, True),     # d padded to 128, two select segments
    (130, 3000, 1024, 1000, True),  # k close to a segment's sort capacity
    (3, 70000, 32, 2048, False),    # k = KNN_MAX_K, radix path
    (64, 4096, 1900, 13, True),     # UniRep-like width
    (20, 9000, 20, 11, False),      # amino-acid-composition-like width
]


@pytest.mark.parametrize("metric", [IP, L2])
@pytest.mark.parametrize("case", EXACT_CASES)
def test_exact_path_parity(knn, case, metric):
    nq, nb, d, k, norm = case
    xq, xb = _data(nq, nb, d, seed=nq + nb + d, normalize=norm and metric == IP, scale=2.5)
    D, I, idx = _search(knn, xq, xb, k, metric, path=1)
    assert idx.stat("path") == 1
    D_ref, I_ref = fo.knn_flat(xq, xb, k, metric, want=k + 4)  # + the next candidates: k-boundary rule
    stats = check_parity(D, I, D_ref, I_ref, xq, xb, metric, max_excused_frac=2e-3)
    assert stats["positions"] == nq * k


TENSOR_CASES = [
    # nq, nb, d, k, query_batch
    (300, 20000, 1024, 100, 16384),
    (128, 9000, 1024, 10, 16384),
    (200, 16384, 64, 5, 128),        # several query batches, smallest K loop
    (77, 12345, 96, 50, 16384),      # padded d, ragged everything
    (150, 30000, 1024, 1000, 16384),
    (64, 8192, 1280, 1, 16384),
    (513, 50000, 512, 2048, 256),    # k = KNN_MAX_K: candidate capacity 16384
]


BF16, FP16 = 1, 2  # values of the "shadow_fmt" parameter (0 = automatic)


@pytest.mark.parametrize("fmt", [BF16, FP16])
@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("metric", [IP, L2])
@pytest.mark.parametrize("case", TENSOR_CASES)
def test_tensor_path_parity(knn, case, metric, cta_group, fmt):
    """cta_group 1: one CTA per 128x256 tile; 2: CTA pairs on 256x256 tiles (tcgen05 cta_group::2).
    fmt: 16-bit format of the tensor-core operands (bf16 or fp16 shadow rows and queries)."""
    nq, nb, d, k, qb = case
    xq, xb = _data(nq, nb, d, seed=7 * nq + nb, normalize=metric == IP, scale=1.7)
    D, I, idx = _search(knn, xq, xb, k, metric, path=2, query_batch=qb, cta_group=cta_group, shadow_fmt=fmt)
    assert idx.stat("path") == 2 and idx.stat("gemm_launches") >= 1
    assert idx.stat("shadow_fmt") == fmt
    assert idx.stat("overflow_batches") == 0
    D_ref, I_ref = fo.knn_flat(xq, xb, k, metric, want=k + 4)  # + the next candidates: k-boundary rule
    # L2 on unnormalised rows: |x|^2+|y|^2 ~ 6000, so the fp32 expansion formula resolves distances to
    # ~5e-4 only - comparable to the gaps between consecutive neighbours at k >= 1000; more positions
    # are legitimately undecidable in fp32 (still arbitrated one by one in fp64 by check_parity).
    check_parity(D, I, D_ref, I_ref, xq, xb, metric, max_excused_frac=2e-3 if metric == IP else 1e-2)
    # the rerank repeats the scan kernel's arithmetic: both device paths agree bit for bit
    D1, I1, _ = _search(knn, xq, xb, k, metric, path=1)
    assert np.array_equal(I1, I)
    assert np.array_equal(D1, D)


def test_tensor_path_real_embeddings(knn, golden_dir):
    """Real (clustered) SeqVec embeddings tiled into a larger database: stresses the candidate
    margin where many scores sit close together."""
    base = np.concatenate([np.load(golden_dir / n / f) for n in ["pfam-20-10", "pfam-20-10-sum", "pfam-20-dist"]
                           for f in ["train.npy", "test.npy"]])
    rng = np.random.default_rng(5)
    xb = np.concatenate([base + rng.standard_normal(base.shape).astype(np.float32) * 0.02 * s for s in range(1, 9)])
    xq = base[:400].copy()
    fo.normalize_L2(xb)
    fo.normalize_L2(xq)
    D, I, idx = _search(knn, xq, xb, 100, IP, path=2)
    assert idx.stat("path") == 2
    D_ref, I_ref = fo.knn_flat(xq, xb, 100, IP, want=100 + 4)  # + the next candidates: k-boundary rule
    check_parity(D, I, D_ref, I_ref, xq, xb, IP, max_excused_frac=5e-3)


@pytest.mark.parametrize("path", [1, 2])
def test_exact_ties_lower_id_first(knn, path):
    """Duplicated database rows give exactly equal scores: the lower id must come first."""
    xq, xb = _data(96, 4500, 128, seed=11)
    xb = np.concatenate([xb, xb])  # row j and row j + 4500 are identical
    D, I, _ = _search(knn, xq, xb, 20, IP, path=path, tensor_min_n=1)
    D_ref, I_ref = fo.knn_flat(xq, xb, 20, IP, want=20 + 4)  # + the next candidates: k-boundary rule
    check_parity(D, I, D_ref, I_ref, xq, xb, IP, max_excused_frac=5e-3)
    pairs = I.reshape(96, 10, 2)
    assert (pairs[:, :, 1] == pairs[:, :, 0] + 4500).all()
    assert np.array_equal(D[:, 0::2], D[:, 1::2])


def test_candidate_overflow_falls_back_not_truncates(knn):
    """Every database row identical -> every score ties -> the candidate lists overflow; the
    engine must notice and still return the exact answer (ids 0..k-1)."""
    nb, d, k = 40000, 64, 10
    row = np.random.default_rng(2).standard_normal(d).astype(np.float32)
    xb = np.tile(row, (nb, 1))
    xq = np.tile(row, (128, 1))
    D, I, idx = _search(knn, xq, xb, k, IP, path=2)
    assert idx.stat("overflow_batches") >= 1
    assert np.array_equal(I, np.tile(np.arange(k), (128, 1)))


def test_overflow_repairs_only_the_queries_it_happened_to(knn):
    """A few queries sit in a huge clump of identical rows (their lists overflow), the rest of the batch is ordinary:
    only the clump queries are redone by the exact scan, and every row of the result still equals the exact path."""
    rng = np.random.default_rng(8)
    nb, d, k = 60000, 128, 20
    xb = rng.standard_normal((nb, d)).astype(np.float32)
    clump = rng.standard_normal(d).astype(np.float32)
    xb[10000:40000] = clump  # 30,000 identical rows > candidate capacity (8192)
    xq = rng.standard_normal((500, d)).astype(np.float32)
    xq[[3, 77, 400]] = clump
    D, I, idx = _search(knn, xq, xb, k, IP, path=2)
    # the 3 clump queries, plus the few whose early (looser) panel threshold still let the clump in
    assert idx.stat("overflow_batches") == 1 and 3 <= idx.stat("overflow_queries") < 100
    D1, I1, _ = _search(knn, xq, xb, k, IP, path=1)
    assert np.array_equal(I, I1) and np.array_equal(D, D1)
    assert np.array_equal(I[3], np.arange(10000, 10000 + k))  # ties: lower id first


def test_incremental_add_and_reset(knn):
    xq, xb = _data(40, 6000, 256, seed=3)
    idx = knn.IndexFlat(256, IP)
    idx.add(xb[:1000])
    idx.add(xb[1000:1001])
    idx.add(xb[1001:])
    assert idx.ntotal == 6000
    D, I = idx.search(xq, 7)
    D_ref, I_ref = fo.knn_flat(xq, xb, 7, IP, want=7 + 4)  # + the next candidates: k-boundary rule
    check_parity(D, I, D_ref, I_ref, xq, xb, IP)
    got = idx.reconstruct_n(998, 5)
    assert np.array_equal(got, xb[998:1003])
    idx.reset()
    assert idx.ntotal == 0
    idx.add(xb[:10])
    D, I = idx.search(xq, 3)
    assert I.max() < 10


def test_add_copies_caller_memory(knn):
    """faiss semantics: the drivers mutate their array after add (pfam/proteins_search.py:37,49)."""
    xq, xb = _data(5, 100, 64, seed=9)
    idx = knn.IndexFlat(64, IP)
    keep = xb.copy()
    idx.add(xb)
    xb[:] = 0
    D, I = idx.search(xq, 3)
    D_ref, I_ref = fo.knn_flat(xq, keep, 3, IP, want=3 + 4)  # + the next candidates: k-boundary rule
    check_parity(D, I, D_ref, I_ref, xq, keep, IP)


def test_input_coercion_like_faiss(knn):
    xq, xb = _data(5, 300, 48, seed=4)
    idx = knn.IndexFlat(48, L2)
    idx.add(xb.astype(np.float64))            # dtype coerced
    D, I = idx.search(np.asfortranarray(xq), 4)  # layout coerced
    D_ref, I_ref = fo.knn_flat(xq, xb, 4, L2, want=4 + 4)  # + the next candidates: k-boundary rule
    check_parity(D, I, D_ref, I_ref, xq, xb, L2)
    fp16 = xb.astype(np.float16)               # cath/search.py:40 up-casts fp16 files
    idx2 = knn.IndexFlat(48, L2)
    idx2.add(fp16.astype(np.float32))
    D, I = idx2.search(xq, 4)
    D_ref, I_ref = fo.knn_flat(xq, fp16.astype(np.float32), 4, L2, want=4 + 4)  # + the next candidates: k-boundary rule
    check_parity(D, I, D_ref, I_ref, xq, fp16.astype(np.float32), L2)


def test_write_read_index_roundtrip(knn, tmp_path):
    xq, xb = _data(8, 777, 40, seed=6)
    for metric, fourcc in [(IP, b"IxFI"), (L2, b"IxF2")]:
        idx = knn.IndexFlat(40, metric)
        idx.add(xb)
        path = tmp_path / f"flat{metric}.index"
        knn.write_index(idx, str(path))
        raw = path.read_bytes()
        assert raw[:4] == fourcc
        assert len(raw) == 4 + 4 + 8 + 8 + 8 + 1 + 4 + 8 + xb.nbytes  # header + vectors, as faiss writes it
        assert np.array_equal(np.frombuffer(raw[-xb.nbytes:], np.float32).reshape(xb.shape), xb)
        back = knn.read_index(str(path))
        assert (back.d, back.ntotal, back.metric_type) == (40, 777, metric)
        D0, I0 = idx.search(xq, 5)
        D1, I1 = back.search(xq, 5)
        assert np.array_equal(I0, I1) and np.array_equal(D0, D1)


def test_bf16_storage_index(knn):
    """Config C5: the bf16 values ARE the database; results must equal an fp32 search over the
    bf16-rounded rows."""
    import torch

    xq, xb = _data(200, 20000, 256, seed=8)
    xb_r = torch.from_numpy(xb).to(torch.bfloat16).to(torch.float32).numpy()
    for path in (1, 2):
        idx = knn.IndexFlat(256, IP, bf16_storage=True)
        idx.set_param("path", path)
        idx.add(xb)
        D, I = idx.search(xq, 50)
        D_ref, I_ref = fo.knn_flat(xq, xb_r, 50, IP, want=50 + 4)  # + the next candidates: k-boundary rule
        check_parity(D, I, D_ref, I_ref, xq, xb_r, IP, max_excused_frac=2e-3)


@pytest.mark.parametrize("metric", [IP, L2])
@pytest.mark.parametrize("nq,nb,d,k", [(1, 20000, 1024, 100), (17, 9000, 96, 10), (32, 30000, 256, 1000), (33, 12345, 1024, 5),
                                       (64, 50000, 512, 200), (5, 8192, 1900, 13), (40, 8192, 1900, 13), (64, 70000, 64, 2048),
                                       (65, 33333, 1024, 100), (128, 60001, 1024, 10), (100, 9000, 320, 1000), (127, 20000, 1900, 50),
                                       (129, 40000, 1024, 100), (256, 70001, 1024, 10), (200, 9000, 320, 1000), (255, 30011, 512, 64)])
def test_few_queries_stream_kernel(knn, nq, nb, d, k, metric):
    """Launches with <= 64 queries take the few-queries variant of the GEMM kernel (database rows as the M operand,
    queries resident in shared memory) unless the resident queries leave no room for the ring (d = 1900, nq > 32);
    65..128 queries take its CTA-pair form (tcgen05 cta_group::2, each CTA of the pair keeps half of the queries),
    129..256 its two-pair form (cluster of 4, database half-tiles TMA-multicast to both pairs).
    Same bits as the main kernel and as the exact scan, both operand formats, dense first panel included."""
    xq, xb = _data(nq, nb, d, seed=nq * 7 + d, normalize=metric == IP, scale=1.3)
    D1, I1, _ = _search(knn, xq, xb, k, metric, path=1)
    for fmt in (BF16, FP16):
        for stream in (1, 0):
            # stream_quad: the two-pair (cluster of 4, TMA multicast) form for 129..256 queries - off by default
            # (measured slower than the main kernel), exercised here so that it stays correct
            D, I, idx = _search(knn, xq, xb, k, metric, path=2, shadow_fmt=fmt, stream_kernel=stream, stream_quad=stream)
            assert idx.stat("path") == 2 and idx.stat("overflow_batches") == 0
            assert np.array_equal(I, I1) and np.array_equal(D, D1), (fmt, stream)
    D_ref, I_ref = fo.knn_flat(xq, xb, k, metric, want=k + 4)  # + the next candidates: k-boundary rule
    check_parity(D1, I1, D_ref, I_ref, xq, xb, metric, max_excused_frac=2e-3 if metric == IP else 1e-2)


def test_shadow_format_is_chosen_from_the_data(knn):
    """fp16 shadow rows while the data sits inside fp16's range (3 more mantissa bits than bf16 -> smaller error
    bound -> fewer candidates to rescore); data outside it (saturating or flushing to zero) is noticed through the
    MEASURED rounding error and the shadow rewritten in bf16.  The result is the exact one either way."""
    k = 20
    for scale, want in [(1.0, FP16), (3e5, BF16), (2e-7, BF16)]:
        xq, xb = _data(150, 12000, 128, seed=21, normalize=False, scale=scale)
        D, I, idx = _search(knn, xq, xb, k, IP, path=2)
        assert idx.stat("shadow_fmt") == want, (scale, idx.stat("shadow_fmt"))
        assert idx.stat("shadow_conversions") == (0 if want == FP16 else 1)
        assert idx.stat("overflow_batches") == 0
        D1, I1, _ = _search(knn, xq, xb, k, IP, path=1)
        assert np.array_equal(I, I1) and np.array_equal(D, D1)
    # rows that leave the range arrive later: the next search converts, once
    xq, xb = _data(100, 10000, 64, seed=22, normalize=False)
    idx = knn.IndexFlat(64, L2)
    idx.set_param("path", 2)
    idx.add(xb)
    idx.search(xq, 5)
    assert idx.stat("shadow_fmt") == FP16
    big = (xb[:500] * 1e6).astype(np.float32)
    idx.add(big)
    D, I = idx.search(xq, 5)
    assert idx.stat("shadow_fmt") == BF16 and idx.stat("shadow_conversions") == 1
    idx.add(xb[:10])
    idx.search(xq, 5)
    assert idx.stat("shadow_conversions") == 1  # stays bf16
    all_rows = np.concatenate([xb, big])
    D_ref, I_ref = fo.knn_flat(xq, all_rows, 5, L2, want=5 + 4)  # + the next candidates: k-boundary rule
    check_parity(D, I, D_ref, I_ref, xq, all_rows, L2, max_excused_frac=1e-2)
    idx.reset()
    idx.add(xb)
    idx.search(xq, 5)
    assert idx.stat("shadow_fmt") == FP16  # a fresh database decides afresh


def test_shadow_format_switch_and_forced_fp16_out_of_range(knn):
    """set_param("shadow_fmt") on a filled index rewrites the shadow from the fp32 master rows; every format gives the
    same bits (the rerank is exact).  fp16 forced onto out-of-range data saturates, the error bound explodes, the
    candidate lists overflow and the exact scan answers: slow, never wrong."""
    xq, xb = _data(130, 15000, 256, seed=23)
    idx = knn.IndexFlat(256, IP)
    idx.set_param("path", 2)
    idx.add(xb)
    res = {}
    for fmt in (FP16, BF16, FP16):
        idx.set_param("shadow_fmt", fmt)
        res[fmt] = idx.search(xq, 30)
        assert idx.stat("shadow_fmt") == fmt
    assert idx.stat("shadow_conversions") == 2
    assert np.array_equal(res[FP16][1], res[BF16][1]) and np.array_equal(res[FP16][0], res[BF16][0])
    D_ref, I_ref = fo.knn_flat(xq, xb, 30, IP, want=30 + 4)  # + the next candidates: k-boundary rule
    check_parity(res[FP16][0], res[FP16][1], D_ref, I_ref, xq, xb, IP, max_excused_frac=2e-3)
    xq, xb = _data(64, 9000, 64, seed=24, normalize=False, scale=3e5)
    D, I, idx = _search(knn, xq, xb, 7, IP, path=2, shadow_fmt=FP16)
    assert idx.stat("shadow_fmt") == FP16
    D1, I1, _ = _search(knn, xq, xb, 7, IP, path=1)
    assert np.array_equal(I, I1) and np.array_equal(D, D1)


def test_shadow_format_follows_the_search_regime(knn):
    """Automatic choice: fp16 operands where the exact rescoring dominates (k large against n), bf16 where the GEMM
    does (fp16 costs tensor-core power); the shadow rows are rewritten from the fp32 rows when a search asks for the
    other regime.  Same bits either way."""
    xq, xb = _data(120, 30000, 128, seed=27)
    idx = knn.IndexFlat(128, IP)
    idx.set_param("path", 2)
    idx.add(xb)
    idx_exact = knn.IndexFlat(128, IP)
    idx_exact.set_param("path", 1)
    idx_exact.add(xb)
    for k, want, conversions in [(2, BF16, 1), (2, BF16, 1), (5, FP16, 2), (100, FP16, 2), (3, BF16, 3)]:
        D, I = idx.search(xq, k)
        assert (idx.stat("shadow_fmt"), idx.stat("shadow_conversions")) == (want, conversions), k
        D1, I1 = idx_exact.search(xq, k)
        assert np.array_equal(I, I1) and np.array_equal(D, D1)
    big = knn.IndexFlat(128, IP)   # the expected size is known up front: no rewrite
    big.reserve(1 << 21)
    big.add(xb)
    assert big.stat("shadow_fmt") == BF16


@pytest.mark.parametrize("bits", [3, 5, 6])
def test_bf16_operands_with_fewer_mantissa_bits(knn, bits):
    """"mantissa_bits": bf16 operands rounded to fewer mantissa bits (operand bits that never toggle cost no
    multiplier power).  The rounding error is measured, the bound grows with it, the result stays exact."""
    xq, xb = _data(200, 25000, 512, seed=28)
    D, I, idx = _search(knn, xq, xb, 50, IP, path=2, shadow_fmt=BF16, mantissa_bits=bits)
    assert idx.stat("mantissa_bits") == bits
    assert bits < 5 or idx.stat("overflow_batches") == 0  # 3 bits: the bound is so wide that lists overflow -> exact scan
    D1, I1, _ = _search(knn, xq, xb, 50, IP, path=1)
    assert np.array_equal(I, I1) and np.array_equal(D, D1)
    import torch

    idx = knn.IndexFlat(512, L2, bf16_storage=True)  # queries only: the bf16 rows are the database
    idx.set_param("mantissa_bits", bits)
    with pytest.raises(ValueError):
        idx.set_param("shadow_fmt", FP16)
    idx.add(xb)
    res = {}
    for path in (2, 1):
        idx.set_param("path", path)
        res[path] = idx.search(xq, 20)
    assert np.array_equal(res[1][1], res[2][1]) and np.array_equal(res[1][0], res[2][0])
    xb_r = torch.from_numpy(xb).to(torch.bfloat16).to(torch.float32).numpy()
    D_ref, I_ref = fo.knn_flat(xq, xb_r, 20, L2, want=20 + 4)  # + the next candidates: k-boundary rule
    check_parity(res[2][0], res[2][1], D_ref, I_ref, xq, xb_r, L2, max_excused_frac=2e-3)


@pytest.mark.parametrize("fmt", [BF16, FP16])
@pytest.mark.parametrize("path", [1, 2])
def test_nan_rows_never_enter_a_result(knn, path, fmt):
    """Like faiss's heap (a NaN score never beats the heap top), a row holding a NaN is never returned - and must
    not disturb the candidate threshold of the tensor path either."""
    xq, xb = _data(100, 12000, 128, seed=26)
    bad = np.arange(0, 12000, 37)
    xb_nan = xb.copy()
    xb_nan[bad, 5] = np.nan
    D, I, _ = _search(knn, xq, xb_nan, 25, IP, path=path, shadow_fmt=fmt)
    keep = np.setdiff1d(np.arange(12000), bad)
    D_ref, I_ref = fo.knn_flat(xq, xb[keep], 25, IP, want=25 + 4)  # + the next candidates: k-boundary rule
    check_parity(D, keep_inverse(I, keep), D_ref, I_ref, xq, xb[keep], IP, max_excused_frac=2e-3)
    assert not np.isin(I, bad).any() and np.isfinite(D).all()


def keep_inverse(I, keep):
    inv = np.full(int(keep.max()) + 1, -1, dtype=np.int64)
    inv[keep] = np.arange(len(keep))
    return inv[I]


def test_select_keeps_every_real_entry_among_thousands_of_padding(knn):
    """Found by tests/test_gpu_fuzz.py: when a list longer than the in-shared-memory sort capacity (4096) holds
    FEWER than k real entries and thousands of padding ones (a shard's rescored list after the cross-shard bound has
    been applied, k = 2000), the padding must not crowd real entries out of the sort buffer."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(5)
    nl, nq, k, real = 3, 50, 2000, 300
    D = torch.full((nl, nq, k), -np.finfo(np.float32).max, device="cuda")
    I = torch.full((nl, nq, k), -1, dtype=torch.int64, device="cuda")
    vals = torch.stack([(torch.randperm(1 << 20, device="cuda", generator=g)[:nl * real] + 1).float() / (1 << 20) for _ in range(nq)])
    D[:, :, :real] = torch.sort(vals.view(nq, nl, real).permute(1, 0, 2), dim=2, descending=True)[0]  # distinct scores: no ties
    I[:, :, :real] = torch.stack([torch.randperm(100000, device="cuda", generator=g)[:nl * real].view(nl, real) for _ in range(nq)], dim=1)
    Dm, Im = knn.merge_topk(D, I, IP)
    flatD = D[:, :, :real].permute(1, 0, 2).reshape(nq, nl * real)
    flatI = I[:, :, :real].permute(1, 0, 2).reshape(nq, nl * real)
    order = torch.argsort(flatD, dim=1, descending=True, stable=True)
    assert torch.equal(Dm[:, :nl * real], torch.gather(flatD, 1, order))
    assert torch.equal(Im[:, :nl * real], torch.gather(flatI, 1, order))
    assert (Im[:, nl * real:] == -1).all()


def test_torch_device_api_and_merge(knn):
    import torch

    xq, xb = _data(256, 24000, 128, seed=12)
    dev = torch.device("cuda:0")
    tq, tb = torch.from_numpy(xq).to(dev), torch.from_numpy(xb).to(dev)
    full = knn.IndexFlat(128, IP)
    full.add(tb)
    D, I = full.search(tq, 30)
    assert D.is_cuda and I.dtype == torch.int64
    D_ref, I_ref = fo.knn_flat(xq, xb, 30, IP, want=30 + 4)  # + the next candidates: k-boundary rule
    check_parity(D.cpu().numpy(), I.cpu().numpy(), D_ref, I_ref, xq, xb, IP, max_excused_frac=2e-3)
    # row-sharded: 3 ragged shards, global ids via id_base, merged on the device
    bounds = [0, 7000, 15001, 24000]
    Ds, Is = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        shard = knn.IndexFlat(128, IP)
        shard.add(tb[a:b])
        d_, i_ = shard.search(tq, 30, id_base=a)
        Ds.append(d_)
        Is.append(i_)
    Dm, Im = knn.merge_topk(torch.stack(Ds), torch.stack(Is), IP)
    assert torch.equal(Im, I) and torch.equal(Dm, D)
    # device normalize
    t = torch.randn(100, 64, device=dev)
    ref = t.cpu().numpy().copy()
    fo.normalize_L2(ref)
    knn.normalize_L2(t)
    np.testing.assert_allclose(t.cpu().numpy(), ref, rtol=2e-6, atol=1e-7)


def test_properties_at_scale(knn):
    """All-vs-all on 60k x 1024 normalised rows, k=100, tensor path: properties that need no
    oracle (sorted, self hit first with score ~1, ids unique and in range), plus an exhaustive
    oracle check on a sample of the queries."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(1234)
    xb = torch.randn(60000, 1024, device="cuda", generator=g)
    knn.normalize_L2(xb)
    idx = knn.IndexFlat(1024, IP)
    idx.add(xb)
    D, I = idx.search(xb[:8192], 100)
    assert idx.stat("path") == 2 and idx.stat("overflow_batches") == 0
    Dn, In = D.cpu().numpy(), I.cpu().numpy()
    assert (np.diff(Dn, axis=1) <= 0).all()
    assert np.array_equal(In[:, 0], np.arange(8192))
    np.testing.assert_allclose(Dn[:, 0], 1.0, atol=1e-5)
    assert In.min() >= 0 and In.max() < 60000
    assert all(len(set(r)) == 100 for r in In[::64].tolist())
    sample = np.arange(0, 8192, 128)
    xb_h = xb.cpu().numpy()
    D_ref, I_ref = fo.knn_flat(xb_h[sample], xb_h, 100, IP, want=100 + 4)  # + the next candidates: k-boundary rule
    check_parity(Dn[sample], In[sample], D_ref, I_ref, xb_h[sample], xb_h, IP, max_excused_frac=5e-3)


@pytest.mark.parametrize("metric,normalize", [(IP, True), (L2, False)])
def test_two_phase_sharded_search_equals_single_index(knn, metric, normalize):
    """Row-sharded search with the cross-shard bound exchange (filter -> max of the per-query lower
    bounds -> finish -> merge), emulated in one process with three ragged shards whose rows have
    different norms (so their error bounds differ).  Must equal the unsharded search bit for bit."""
    import torch

    xq, xb = _data(300, 40000, 256, seed=21, normalize=normalize, scale=1.3)
    if not normalize:
        xb[25000:] *= 1.7  # the third shard gets a larger max norm -> larger eps
    dev = torch.device("cuda:0")
    tq, tb = torch.from_numpy(xq).to(dev), torch.from_numpy(xb).to(dev)
    k = 64
    full = knn.IndexFlat(256, metric)
    full.set_param("path", 2)
    full.add(tb)
    D, I = full.search(tq, k)
    bounds = [0, 9000, 25000, 40000]
    shards = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        sh = knn.IndexFlat(256, metric)
        sh.set_param("path", 2)
        sh.add(tb[a:b])
        shards.append(sh)
    j = -(-k // len(shards))
    both = [sh.search_filter(tq, k, j) for sh in shards]
    lowers = [b[0] for b in both]
    lower_k = torch.stack(lowers).max(dim=0).values               # best shard's own k-th best
    lower_j = torch.stack([b[1] for b in both]).min(dim=0).values  # every shard holds j rows at or above its j-th
    lower = torch.maximum(lower_k, lower_j)
    assert (lower >= lowers[0]).all() and (lower > lowers[0]).any()  # the exchange does raise bounds
    if normalize:  # homogeneous shards: the min-of-j-th bound is usually the tighter one
        assert (lower_j > lower_k).float().mean() > 0.5
    Ds, Is = zip(*[sh.search_finish(lower, k, id_base=a) for sh, a in zip(shards, bounds[:-1])])
    Dm, Im = knn.merge_topk(torch.stack(Ds), torch.stack(Is), metric)
    assert torch.equal(Im, I) and torch.equal(Dm, D)
    D_ref, I_ref = fo.knn_flat(xq, xb, k, metric, want=k + 4)  # + the next candidates: k-boundary rule
    check_parity(Dm.cpu().numpy(), Im.cpu().numpy(), D_ref, I_ref, xq, xb, metric, max_excused_frac=1e-2)
    # a shard on the exact path takes part too (its bound is -FLT_MAX)
    small = knn.IndexFlat(256, metric)
    small.add(tb[:100])
    lo = small.search_filter(tq, k)
    assert (lo == -np.finfo(np.float32).max).all()
    d_, i_ = small.search_finish(lower, k)
    assert i_.shape == (300, k)


@pytest.mark.parametrize("metric", [IP, L2])
def test_l2_blocked_rerank_and_overlapped_batches_change_nothing(knn, metric):
    """Rescoring-heavy shape (nq k > 4 N, rows larger than 64 MB): the rerank walks the database in L2-sized row
    ranges (grid = range x query), a call of several query batches runs the finish phase of batch b on a side stream
    under the filter of batch b + 1, and a single-batch call with k >= 256 is cut into batches for the same reason.
    None of it may change a bit of the result: compared with the exact scan and with every switch off."""
    xq, xb = _data(4500, 20000, 1024, seed=77, normalize=metric == IP, scale=1.0)
    xb[15000:15040] = xb[100]  # duplicates straddling the two row ranges (12,288 rows of 4 KB each)
    k = 300
    D1, I1, _ = _search(knn, xq[:600], xb, k, metric, path=1)
    base = None
    for switches in ({}, {"l2_blocked_rerank": 1}, {"overlap_finish": 0}, {"split_single_batch": 0},
                     {"query_batch": 1024}, {"query_batch": 1024, "overlap_finish": 0, "l2_blocked_rerank": 0}):
        D, I, idx = _search(knn, xq, xb, k, metric, path=2, **switches)
        assert idx.stat("path") == 2 and idx.stat("overflow_batches") == 0, switches
        if base is None:
            base = (D, I)
            assert np.array_equal(I[:600], I1) and np.array_equal(D[:600], D1)
        assert np.array_equal(I, base[1]) and np.array_equal(D, base[0]), switches
    D_ref, I_ref = fo.knn_flat(xq[:200], xb, k, metric, want=k + 4)
    check_parity(base[0][:200], base[1][:200], D_ref, I_ref, xq[:200], xb, metric, max_excused_frac=1e-2)
    # host path (numpy in / out, pipelined per batch) returns the same bits as the device path
    import torch

    idx = knn.IndexFlat(1024, metric)
    idx.set_param("path", 2)
    idx.set_param("query_batch", 1024)
    idx.add(xb)
    Dd, Id = idx.search(torch.from_numpy(xq).cuda(), k)
    assert np.array_equal(Id.cpu().numpy(), base[1]) and np.array_equal(Dd.cpu().numpy(), base[0])


@pytest.mark.parametrize("metric", [IP, L2])
def test_small_database_large_k_single_dense_panel(knn, metric):
    """k >= 256 of <= 16,384 rows (C2: 14,433 rows, k = 1000): the whole database is scored as one dense panel (one
    candidate slot per row) instead of letting a quarter of the second panel's scores through the threshold filter.
    Same bits as with the switch off and as the exact scan; duplicates and k close to N included."""
    xq, xb = _data(1500, 14433, 256, seed=5, normalize=metric == IP)
    xb[9000:9050] = xb[17]
    for k in (256, 1000, 2048):
        D1, I1, _ = _search(knn, xq[:300], xb, k, metric, path=1)
        for on in (1, 0):
            D, I, idx = _search(knn, xq, xb, k, metric, path=2, dense_small_db=on)
            assert idx.stat("path") == 2 and idx.stat("overflow_batches") == 0, (k, on)
            assert np.array_equal(I[:300], I1) and np.array_equal(D[:300], D1), (k, on)
            if on:
                launches_on = idx.stat("gemm_launches")
            else:
                assert idx.stat("gemm_launches") > launches_on  # several panels without the switch, one per batch with it
