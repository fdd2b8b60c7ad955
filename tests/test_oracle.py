"""CPU tests: the oracle against the reference's golden vectors and against itself."""
import json

import numpy as np
import pytest

import _c_oracle
from oracle import flat_oracle as fo
from oracle.evaluate import Fixture, evaluate_ids
from oracle.parity import ParityError, check_parity


def _search(fx_dir, k, metric=fo.METRIC_INNER_PRODUCT, normalize=True):
    xb = np.load(fx_dir / "train.npy")
    xq = np.load(fx_dir / "test.npy")
    if normalize:
        fo.normalize_L2(xq)
        fo.normalize_L2(xb)
    idx = fo.IndexFlat(xb.shape[1], metric)
    idx.train(xb)
    idx.add(xb)
    D, I = idx.search(xq, k)
    return xq, xb, D, I


def test_small_random_known_answer(golden_dir, expected):
    """/root/reference/tests/test_main.py:10-18 restated (exact list equality)."""
    xq, xb, D, I = _search(golden_dir / "small-random", 5)
    auc1s, tps = evaluate_ids(Fixture(golden_dir / "small-random"), I)
    assert auc1s == [1.0, 1 / 3, 2 / 3, 0.0, 0.0, 1 / 3]
    assert tps == [1.0, 2 / 3, 2 / 3, 1.0, 1.0, 1.0]
    assert np.array_equal(I, expected["small-random.ip.k5.I"])
    assert np.array_equal(D, expected["small-random.ip.k5.D"])
    # SURVEY.md section 8c lists these ids for the fixture
    assert I.tolist() == [[1, 5, 9, 8, 6], [1, 6, 9, 8, 2], [6, 8, 1, 9, 5], [8, 10, 4, 5, 2],
                          [6, 1, 8, 3, 10], [1, 8, 5, 9, 0]]


def test_pfam_20_10_known_answer(golden_dir, expected):
    """/root/reference/tests/test_main.py:21-27 restated (kNN half)."""
    xq, xb, D, I = _search(golden_dir / "pfam-20-10", 10)
    auc1s, tps = evaluate_ids(Fixture(golden_dir / "pfam-20-10"), I)
    assert np.mean(auc1s) == 0.871
    assert np.mean(tps) == 0.91
    assert np.array_equal(I, expected["pfam-20-10.ip.k10.I"])


@pytest.mark.parametrize("name", ["pfam-20-10-sum", "pfam-20-dist"])
def test_regression_fixtures(golden_dir, expected, name):
    summary = json.loads((golden_dir / "expected.json").read_text())
    xq, xb, D, I = _search(golden_dir / name, 13)
    auc1s, tps = evaluate_ids(Fixture(golden_dir / name), I)
    assert np.mean(auc1s) == summary[name]["mean_auc1"]
    assert np.mean(tps) == summary[name]["mean_tp"]
    assert np.array_equal(I, expected[f"{name}.ip.k13.I"])


def test_normalize_semantics():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((7, 33)).astype(np.float32)
    x[3] = 0
    y = x.copy()
    assert fo.normalize_L2(y) is None
    assert np.array_equal(y[3], np.zeros(33, np.float32))  # zero rows untouched
    n = np.linalg.norm(y.astype(np.float64), axis=1)
    assert np.allclose(np.delete(n, 3), 1.0, atol=1e-6)
    z = x.copy()
    _c_oracle.normalize_l2(z)
    assert np.allclose(z, y, rtol=1e-6, atol=1e-7)
    with pytest.raises(TypeError):
        fo.normalize_L2(x.astype(np.float64))
    with pytest.raises(ValueError):
        fo.normalize_L2(x[:, ::2])


@pytest.mark.parametrize("metric", [fo.METRIC_INNER_PRODUCT, fo.METRIC_L2])
def test_padding_when_k_exceeds_ntotal(golden_dir, expected, metric):
    xb = np.load(golden_dir / "small-random/train.npy")
    xq = np.load(golden_dir / "small-random/test.npy")
    idx = fo.IndexFlat(1024, metric)
    idx.add(xb)
    D, I = idx.search(xq, 16)
    assert (I[:, 11:] == -1).all() and (I[:, :11] >= 0).all()
    fmax = np.finfo(np.float32).max
    assert (D[:, 11:] == (-fmax if metric == fo.METRIC_INNER_PRODUCT else fmax)).all()
    m = "ip" if metric == fo.METRIC_INNER_PRODUCT else "l2"
    assert np.array_equal(I, expected[f"small-random.raw.{m}.k16.I"])
    assert sorted(I[0, :11].tolist()) == list(range(11))


@pytest.mark.parametrize("metric", [fo.METRIC_INNER_PRODUCT, fo.METRIC_L2])
@pytest.mark.parametrize("shape", [(37, 500, 64, 10), (5, 3000, 1024, 100), (64, 1500, 96, 1)])
def test_c_oracle_agrees_with_numpy_oracle(metric, shape):
    nq, nb, d, k = shape
    rng = np.random.default_rng(nq * 1000 + nb)
    xb = rng.standard_normal((nb, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    if metric == fo.METRIC_INNER_PRODUCT:
        fo.normalize_L2(xb)
        fo.normalize_L2(xq)
    D_ref, I_ref = fo.knn_flat(xq, xb, k, metric)
    D, I = _c_oracle.knn_flat(xq, xb, k, metric, nthreads=4)
    stats = check_parity(D, I, D_ref, I_ref, xq, xb, metric)
    assert stats["excused"] <= 0.01 * I.size
    D1, I1 = _c_oracle.knn_flat(xq, xb, k, metric, nthreads=1)
    assert np.array_equal(I1, I) and np.array_equal(D1, D)  # threading does not change results


def test_exact_ties_lower_id_first():
    xb = np.zeros((8, 16), np.float32)
    xb[:, 0] = [1, 2, 2, 3, 2, 1, 3, 0]
    xq = np.zeros((1, 16), np.float32)
    xq[0, 0] = 1
    for impl in (lambda: fo.knn_flat(xq, xb, 5, 0), lambda: _c_oracle.knn_flat(xq, xb, 5, 0)):
        D, I = impl()
        assert I.tolist() == [[3, 6, 1, 2, 4]]
        assert D.tolist() == [[3, 3, 2, 2, 2]]
    D, I = fo.knn_flat(xq, xb, 3, 1)  # L2: dist = (1-v)^2 -> v=1 (ids 0,5) then v=2/0 ties
    assert I.tolist() == [[0, 5, 1]]


def test_parity_checker_rejects_real_mismatch():
    rng = np.random.default_rng(3)
    xb = rng.standard_normal((300, 32)).astype(np.float32)
    xq = rng.standard_normal((4, 32)).astype(np.float32)
    D, I = fo.knn_flat(xq, xb, 5, 0)
    I_bad = I.copy()
    I_bad[2, [0, 1]] = I_bad[2, [1, 0]]
    with pytest.raises(ParityError):
        check_parity(D, I_bad, D, I, xq, xb, 0)
    D_bad = D.copy()
    D_bad[1, 3] *= 1.001
    with pytest.raises(ParityError):
        check_parity(D_bad, I, D, I, xq, xb, 0)
    assert check_parity(D, I, D, I, xq, xb, 0)["excused"] == 0


def test_parity_checker_excuses_only_fp32_near_ties():
    """Two rows whose fp64 scores differ by less than tau may swap (both orders are valid fp32 orderings); a duplicate
    id, an id out of range, a padding mismatch or too many excused positions are rejected."""
    rng = np.random.default_rng(4)
    xb = rng.standard_normal((50, 64)).astype(np.float32)
    xq = rng.standard_normal((3, 64)).astype(np.float32)
    xb[7] = xb[3]
    xb[7, 0] = np.nextafter(xb[3, 0], np.float32(10))  # one ulp apart: scores differ far below tau
    xq[0] = xb[3] * 2
    D, I = fo.knn_flat(xq, xb, 4, 0)
    assert set(I[0, :2].tolist()) == {3, 7}
    I_swapped, D_swapped = I.copy(), D.copy()
    I_swapped[0, [0, 1]] = I[0, [1, 0]]
    D_swapped[0, [0, 1]] = D[0, [1, 0]]
    stats = check_parity(D_swapped, I_swapped, D, I, xq, xb, 0)
    assert stats["excused"] == 2 and stats["rows_with_ties"] == 1
    with pytest.raises(ParityError):  # ... but not more of them than the test allows
        check_parity(D_swapped, I_swapped, D, I, xq, xb, 0, max_excused_frac=0.1)
    I_dup = I.copy()
    I_dup[1, 1] = I_dup[1, 0]
    with pytest.raises(ParityError):
        check_parity(D, I_dup, D, I, xq, xb, 0)
    I_oob = I.copy()
    I_oob[2, 3] = 50
    with pytest.raises(ParityError):
        check_parity(D, I_oob, D, I, xq, xb, 0)
    Dp, Ip = fo.knn_flat(xq, xb[:2], 4, 0)  # k > ntotal: two padded columns
    I_nopad = Ip.copy()
    I_nopad[:, 3] = 0
    with pytest.raises(ParityError):
        check_parity(Dp, I_nopad, Dp, Ip, xq, xb[:2], 0)
    with pytest.raises(ParityError):
        check_parity(D.astype(np.float64), I, D, I, xq, xb, 0)


def test_parity_checker_distance_bound_is_the_stated_1e5_relative():
    """BASELINE.json: 'distances must agree within 1e-5 relative' - no additive slack for scores away from zero;
    only below the fp32 noise floor (tau / 8 absolute) does the absolute bound take over."""
    rng = np.random.default_rng(5)
    xb = rng.standard_normal((400, 1024)).astype(np.float32)
    xq = rng.standard_normal((3, 1024)).astype(np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    D, I = fo.knn_flat(xq, xb, 6, 0)
    assert 1e-5 * D.min() > 2 * 32 * 2.0 ** -24 / 8  # every score is in the relative regime
    ok = D * np.float32(1 + 5e-6)
    assert check_parity(ok, I, D, I, xq, xb, 0)["max_rel_err_D"] == pytest.approx(5e-6, rel=0.2)
    # 4e-5 relative passed the old checker at this magnitude (its additive tau term); the stated bar rejects it
    with pytest.raises(ParityError, match="relative 1e-5"):
        check_parity(D * np.float32(1 + 4e-5), I, D, I, xq, xb, 0)
    # a score next to zero: relative error is meaningless there, the absolute tolerance applies
    xb0 = xb.copy()
    xb0[I[0, 0]] -= xq[0] * (xb0[I[0, 0]] @ xq[0])  # orthogonal to query 0: score ~ 1e-9
    D0, I0 = fo.knn_flat(xq[:1], xb0[I[0, :1]], 1, 0)
    assert abs(D0[0, 0]) < 1e-6
    check_parity(D0 + np.float32(1e-7), I0, D0, I0, xq[:1], xb0[I[0, :1]], 0)
    with pytest.raises(ParityError, match="near-zero"):
        check_parity(D0 + np.float32(1e-5), I0, D0, I0, xq[:1], xb0[I[0, :1]], 0)


def _near_tie_rows(base, n, rng, rel=1e-7):
    """n copies of `base` whose scores against any query differ by ~rel (far below tau, above fp32 resolution of the sum)."""
    rows = np.repeat(base[None], n, 0).copy()
    for i in range(n):
        rows[i, i] = rows[i, i] * np.float32(1 + rel * (i + 1) * 8)
    return rows


def test_parity_checker_k_boundary_and_cluster_rules():
    """inside a tau-cluster the id multiset must match; the cluster at the k-boundary may trade
    members with the reference's (k+1)-th candidate when that lies inside tau - and only then."""
    rng = np.random.default_rng(6)
    d = 64
    xb = rng.standard_normal((40, d)).astype(np.float32)
    xq = rng.standard_normal((1, d)).astype(np.float32)
    xq[0] = xb[5] * 3  # row 5 is the clear best hit
    xb[20:23] = _near_tie_rows(xb[5] * np.float32(0.9), 3, rng)  # a cluster of three near-identical rows right behind the best hit
    D_ref, I_ref = fo.knn_flat(xq, xb, 8, 0)
    row = I_ref[0].tolist()
    cl = [row.index(i) for i in (20, 21, 22)]
    assert max(cl) - min(cl) == 2, "the three near-ties sit next to each other"
    # k chosen so that the boundary cuts the cluster: two members inside, the third is the (k+1)-th candidate
    k = min(cl) + 2
    D, I = D_ref[:, :k].copy(), I_ref[:, :k].copy()
    I[0, k - 1], D[0, k - 1] = I_ref[0, k], D_ref[0, k]  # took the (k+1)-th instead of the k-th: a valid fp32 answer
    stats = check_parity(D, I, D_ref[:, :k + 1], I_ref[:, :k + 1], xq, xb, 0)
    assert stats["boundary_swaps"] == 1 and stats["excused"] == 1
    # the same swap with a row that is NOT inside tau of the boundary is a real error
    far = [i for i in range(40) if i not in row][0]
    I_bad = I_ref[:, :k].copy()
    I_bad[0, k - 1] = far
    with pytest.raises(ParityError):
        check_parity(D_ref[:, :k].copy(), I_bad, D_ref[:, :k + 1], I_ref[:, :k + 1], xq, xb, 0)
    # permutation inside the cluster (away from the boundary): excused, multiset unchanged
    kk = max(cl) + 2
    I_perm = I_ref[:, :kk].copy()
    I_perm[0, cl[0]], I_perm[0, cl[2]] = I_ref[0, cl[2]], I_ref[0, cl[0]]
    assert check_parity(D_ref[:, :kk].copy(), I_perm, D_ref[:, :kk + 1], I_ref[:, :kk + 1], xq, xb, 0)["excused"] == 2
    # an id from outside replacing a cluster member in the middle of the row: multiset violated
    I_out = I_ref[:, :kk].copy()
    I_out[0, cl[1]] = far
    with pytest.raises(ParityError):
        check_parity(D_ref[:, :kk].copy(), I_out, D_ref[:, :kk + 1], I_ref[:, :kk + 1], xq, xb, 0)
