"""GPU parity tests of the post-search kernels, called through the C ABI:
bit-exact against the vectors the UNMODIFIED reference functions produced (tests/golden/postproc/),
against the CPU oracle on larger seeded inputs, and through size-independent properties."""
import numpy as np
import pytest

from oracle import postproc_oracle as po
from oracle.evaluate import Fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def knn():
    import knn_b200

    assert knn_b200._lib.load().knn_device_count() >= 1, "no CUDA device: the CUDA path cannot run"
    return knn_b200


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(golden_dir / "postproc" / "postproc.npz")


def csr_lists(offsets, members):
    return [members[offsets[i]:offsets[i + 1]].tolist() for i in range(len(offsets) - 1)]


# ---- evaluate ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["small-random", "pfam-20-10", "pfam-20-dist"])
@pytest.mark.parametrize("tag", ["golden", "random"])
def test_evaluate_faiss_matches_reference(knn, g, golden_dir, name, tag):
    fx = Fixture(golden_dir / name)
    auc1s, tps = knn.evaluate_faiss(fx, g[f"evaluate.{name}.{tag}.I"])
    assert auc1s == g[f"evaluate.{name}.{tag}.auc1"].tolist()
    assert tps == g[f"evaluate.{name}.{tag}.tp"].tolist()


def test_evaluate_reference_known_answers_end_to_end(knn, golden_dir):
    """tests/test_main.py:10-18 of the reference with both the search and the evaluation on the GPU."""
    import torch

    fx = Fixture(golden_dir / "small-random")
    xb, xq = np.load(fx.train), np.load(fx.test)
    knn.normalize_L2(xq)
    knn.normalize_L2(xb)
    idx = knn.IndexFlat(1024, knn.METRIC_INNER_PRODUCT)
    idx.add(xb)
    D, I = idx.search(torch.from_numpy(xq).cuda(), 5)  # device path: I never leaves the GPU
    auc1s, tps = knn.evaluate_faiss(fx, I)
    assert auc1s == [1.0, 1 / 3, 2 / 3, 0.0, 0.0, 1 / 3]
    assert tps == [1.0, 2 / 3, 2 / 3, 1.0, 1.0, 1.0]


@pytest.mark.parametrize("nq,k,n_db,nfam", [(1000, 100, 5000, 37), (257, 1, 300, 5), (33, 1000, 2000, 3), (5, 2048, 100, 2)])
def test_evaluate_ids_matches_oracle(knn, nq, k, n_db, nfam):
    rng = np.random.default_rng(nq + k)
    df = rng.integers(0, nfam, n_db).astype(np.int32)
    qf = rng.integers(0, nfam, nq).astype(np.int32)
    I = rng.integers(0, n_db, (nq, k))
    # long leading runs for some queries, -1 padding (wraps to the last row like train_ids[-1])
    for q in range(0, nq, 3):
        same = np.flatnonzero(df == qf[q])
        I[q, :min(k, 1 + q % 70)] = rng.choice(same, min(k, 1 + q % 70))
    I[::7, -1] = -1
    lead, tp, size = knn.evaluate_ids(I, qf, df)
    rl, rt, rs = po.evaluate_counts(I, qf, df)
    assert np.array_equal(lead, rl) and np.array_equal(tp, rt) and np.array_equal(size, rs)


def test_evaluate_out_of_range_id_raises(knn):
    with pytest.raises(IndexError):
        knn.evaluate_ids(np.asarray([[0, 5]]), np.zeros(1, np.int32), np.zeros(5, np.int32))


# ---- compute_is_correct -----------------------------------------------------------------------------
def test_compute_is_correct_matches_reference(knn, g):
    out = knn.compute_is_correct(g["is_correct.results"], g["is_correct.mapping"])
    assert out.dtype == np.bool_ and out.shape == g["is_correct.out"].shape
    assert np.array_equal(out, g["is_correct.out"])


def test_compute_is_correct_matches_oracle_at_cath_size(knn):
    rng = np.random.default_rng(5)
    n, k = 14433, 10  # C2: CATH20 all-vs-all, hits=10 (cath/search.py:14)
    codes = np.cumsum(rng.integers(0, 2, (n, 4)), axis=0).astype(np.int32) % np.asarray([5, 40, 1200, 5125], np.int32)
    I = rng.integers(0, n, (n, k))
    out = knn.compute_is_correct(I, codes)
    assert np.array_equal(out, po.compute_is_correct(I, codes))
    assert out.shape == (n, 4, k)


# ---- homolog sets -----------------------------------------------------------------------------------
def test_compute_correctness_array_matches_reference(knn, g):
    hom = csr_lists(g["correctness.offsets"], g["correctness.members"])
    out = knn.compute_correctness_array(g["correctness.full"], hom)
    assert np.array_equal(out, g["correctness.out"])
    out2 = knn.compute_correctness_array(g["correctness.full"], (g["correctness.offsets"], g["correctness.members"]))
    assert np.array_equal(out2, g["correctness.out"])


def test_compute_auc1_matches_reference(knn, g):
    n = len(g["correctness.offsets"]) - 1
    target_ids = [f"P{i}" for i in range(n)]
    hom = csr_lists(g["correctness.offsets"], g["correctness.members"])
    # same sets as the generator built: names of the member rows (+ one absent name where the size says so)
    sizes = g["auc1.set_sizes"]
    hom_names = {f"P{q}": set(target_ids[v] for v in h) | ({"absent"} if sizes[q] > len(h) else set()) for q, h in enumerate(hom)}
    out = knn.compute_auc1(g["auc1.hits"], hom_names, target_ids, target_ids)
    assert out.dtype == np.float64 and np.array_equal(out, g["auc1.out"])


def test_sets_match_oracle_large(knn):
    rng = np.random.default_rng(11)
    n, k = 3000, 300
    hom = [np.unique(rng.integers(0, n, rng.integers(0, 400))).tolist() for _ in range(n)]
    hits = rng.integers(-1, n, (n, k))
    for q in range(0, n, 2):
        if hom[q]:
            r = min(k, q % 90)
            hits[q, :r] = rng.choice(hom[q], r)
    assert np.array_equal(knn.compute_correctness_array(hits, hom), po.compute_correctness_array(hits, hom))
    names = [f"P{i}" for i in range(n)]
    hn = {names[q]: set(names[v] for v in h) for q, h in enumerate(hom)}
    ref = po.compute_auc1(hits, hom, [len(h) for h in hom], n_db=n)
    assert np.array_equal(knn.compute_auc1(hits, hn, names, names), ref)


# ---- remove_self_hit --------------------------------------------------------------------------------
def test_remove_self_hit_matches_reference(knn, g):
    h, s = g["selfhit.hits_in"].copy(), g["selfhit.scores_in"].copy()
    ho, so = knn.remove_self_hit(h, s)
    assert knn.remove_self_hit.last_missing == int(g["selfhit.bogus"])
    assert np.array_equal(h, g["selfhit.hits_inplace"]) and np.array_equal(s, g["selfhit.scores_inplace"])
    assert np.array_equal(ho, g["selfhit.hits_out"]) and np.array_equal(so, g["selfhit.scores_out"])
    assert np.shares_memory(ho, h)  # views of the mutated inputs, like the reference


@pytest.mark.parametrize("n,k", [(500, 1000), (1000, 33), (64, 1), (100, 2), (77, 65)])
def test_remove_self_hit_matches_oracle(knn, n, k):
    import torch

    rng = np.random.default_rng(n * k)
    self_ids = rng.permutation(n + 10)[:n]
    hits = rng.integers(n + 10, n + 5000, (n, k))
    for r in range(n):
        if r % 5 != 4:
            hits[r, rng.integers(0, k) if r % 3 else 0] = self_ids[r]
        if r % 11 == 0 and k > 2:  # self id twice: the first occurrence counts
            hits[r, k - 1] = self_ids[r]
    scores = rng.random((n, k)).astype(np.float32)
    rh, rs = hits.copy(), scores.copy()
    eh, es, bogus = po.remove_self_hit(rh, rs, self_ids)
    th, ts = torch.from_numpy(hits).cuda(), torch.from_numpy(scores).cuda()
    oh, os_ = knn.remove_self_hit(th, ts, self_ids)
    assert knn.remove_self_hit.last_missing == bogus
    assert np.array_equal(th.cpu().numpy(), rh) and np.array_equal(ts.cpu().numpy(), rs)
    assert np.array_equal(oh.cpu().numpy(), eh) and np.array_equal(os_.cpu().numpy(), es)


def test_cath_search_flow_self_hit_first(knn):
    """cath/search.py:13-26 drops column 0 blindly; after remove_self_hit column 0 IS the query itself."""
    import torch

    rng = np.random.default_rng(3)
    x = rng.standard_normal((2000, 128)).astype(np.float32)
    x[100] = x[7]  # duplicate rows: the self hit may come second
    knn.normalize_L2(x)
    idx = knn.IndexFlat(128, knn.METRIC_INNER_PRODUCT)
    idx.add(x)
    D, I = idx.search(torch.from_numpy(x).cuda(), 11)
    knn.remove_self_hit(I, D)
    assert torch.equal(I[:, 0], torch.arange(2000, device=I.device)) and knn.remove_self_hit.last_missing == 0


# ---- prefilter writer -------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["pfam-20-10", "edge"])
def test_write_prefilter_db_matches_reference(knn, g, tag, tmp_path):
    db = tmp_path / "prefilter"
    knn.write_prefilter_db(g[f"prefilter.{tag}.hits"], db, g[f"prefilter.{tag}.queries"], g[f"prefilter.{tag}.scores"],
                           g[f"prefilter.{tag}.test_map"], g[f"prefilter.{tag}.train_map"])
    assert db.with_suffix(".dbtype").read_bytes() == b"\x07\x00\x00\x00"
    assert db.with_suffix(".0").read_bytes() == g[f"prefilter.{tag}.data"].tobytes()
    assert db.with_suffix(".index").read_bytes() == g[f"prefilter.{tag}.index"].tobytes()


@pytest.mark.parametrize("nq,k,big", [(300, 100, False), (50, 2048, False), (40, 1000, True), (3000, 7, False), (1, 1, False)])
def test_prefilter_matches_oracle(knn, nq, k, big):
    rng = np.random.default_rng(nq + k)
    n_train = 5000
    hits = rng.integers(0, n_train, (nq, k)).astype(np.int64)
    hits[rng.random((nq, k)) < 0.05] = -1
    if big:  # sections beyond the shared-memory staging size: 19-digit ids and 30-digit scores
        scores = (rng.standard_normal((nq, k)) * 1e28).astype(np.float32)
        train_map = rng.integers(10 ** 18, 2 ** 62, n_train).astype(np.int64)
    else:
        scores = rng.uniform(-1, 1, (nq, k)).astype(np.float32)
        train_map = rng.permutation(n_train * 3)[:n_train].astype(np.int64)
    test_map = rng.permutation(nq * 2)[:nq].astype(np.int64)
    queries = rng.permutation(nq).astype(np.int64)
    data, index = knn.format_prefilter_db(hits, queries, scores, test_map, train_map)
    rd, ri = po.write_prefilter_db(hits, queries, scores, test_map, train_map)
    assert data.cpu().numpy().tobytes() == rd
    assert index.cpu().numpy().tobytes() == ri


def test_prefilter_properties_at_c4_size(knn):
    """Size-independent checks at the C4 result shape (100k x 100): section lengths sum to the file size, every
    section ends with NUL, the index parses back to the offsets, line count == number of hits != -1."""
    import torch

    gen = torch.Generator(device="cuda").manual_seed(1)
    nq, k, n_train = 100_000, 100, 1_000_000
    I = torch.randint(0, n_train, (nq, k), device="cuda", generator=gen)
    I[:, -1] = -1
    D = torch.rand((nq, k), device="cuda", generator=gen)
    ident = torch.arange(n_train, device="cuda")
    data, index = knn.format_prefilter_db(I, ident[:nq], D, ident[:nq], ident)
    data, index = data.cpu().numpy(), index.cpu().numpy().tobytes()
    rows = np.asarray([[int(v) for v in line.split(b"\t")] for line in index.splitlines()], np.int64)
    assert np.array_equal(rows[:, 0], np.arange(nq))
    assert rows[0, 1] == 0 and np.array_equal(rows[1:, 1], np.cumsum(rows[:-1, 2]))
    assert rows[-1, 1] + rows[-1, 2] == data.size
    assert (data[rows[:, 1] + rows[:, 2] - 1] == 0).all()
    assert int((data == 10).sum()) == nq * (k - 1) and int((data == 0).sum()) == nq
    # spot check: the first section against the oracle
    rd, _ = po.write_prefilter_db(I[:1].cpu().numpy(), [0], D[:1].cpu().numpy(), np.arange(nq), np.arange(n_train))
    assert data[:rows[0, 2]].tobytes() == rd


def test_prefilter_nan_score_raises_like_int_of_nan(knn):
    hits = np.zeros((2, 3), np.int64)
    scores = np.zeros((2, 3), np.float32)
    scores[1, 1] = np.nan
    with pytest.raises(ValueError):
        knn.format_prefilter_db(hits, np.arange(2), scores, np.arange(2), np.arange(1))
    hits[1, 1] = -1  # a skipped hit is never converted
    knn.format_prefilter_db(hits, np.arange(2), scores, np.arange(2), np.arange(1))
