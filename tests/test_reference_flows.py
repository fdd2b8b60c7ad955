"""The reference drivers' call sequences, restated on top of the `faiss` alias package, against
the oracle (GPU).  The drivers themselves are Python glue above the boundary and are not rebuilt;
these tests show that they would run unchanged: same calls, same argument meaning, same outputs
on disk.  (The unmodified drivers were run against the oracle in tests/golden/make_golden.py.)"""
import sys
import time

import numpy as np
import pytest

from oracle import flat_oracle as fo
from oracle.parity import check_parity


def test_alias_satisfies_reference_imports():
    """CPU: what the reference touches at import time (seqvec_search/main.py:9,23)."""
    import faiss

    assert faiss.IndexLSH is not None  # evaluated in a type annotation at import
    assert callable(faiss.normalize_L2) and callable(faiss.IndexFlat)


@pytest.mark.gpu
def test_cath_search_and_save_flow(tmp_path):
    """cath/search.py:29-53: every *.npy in a directory (any width, fp16 allowed), both metrics,
    hits+1 search, self hit dropped, results written as npz."""
    import faiss

    rng = np.random.default_rng(0)
    mats = {"aac": rng.random((400, 20)).astype(np.float32),          # amino-acid composition width
            "prott5_half": rng.standard_normal((300, 1024)).astype(np.float16),  # fp16 file, up-cast at :40
            "esm": rng.standard_normal((257, 1280)).astype(np.float32),
            "unirep": rng.standard_normal((130, 1900)).astype(np.float32)}
    for name, m in mats.items():
        np.save(tmp_path / f"{name}.npy", m)

    def search(embeddings, hits=10, metric=faiss.METRIC_INNER_PRODUCT):  # cath/search.py:13-26
        if metric == faiss.METRIC_INNER_PRODUCT:
            embeddings = embeddings.copy()
            faiss.normalize_L2(embeddings)
        index = faiss.IndexFlat(embeddings.shape[1], metric)
        index.add(embeddings)
        scores, results = index.search(embeddings, hits + 1)
        return results[:, 1:], scores[:, 1:]

    for mname, metric in [("Cosine", faiss.METRIC_INNER_PRODUCT), ("Euclidean", faiss.METRIC_L2)]:
        hits, scores = {}, {}
        for file_path in sorted(tmp_path.glob("*.npy")):
            embeddings = np.load(file_path).astype(np.float32)
            start = time.time()
            hits[file_path.stem], scores[file_path.stem] = search(embeddings, metric=metric)
            tmp_path.joinpath(file_path.with_suffix(f".{mname.lower()}-search-time.txt")).write_text(str(time.time() - start))
        np.savez(tmp_path / f"hits_{mname.lower()}.npz", **hits)
        np.savez(tmp_path / f"scores_{mname.lower()}.npz", **scores)
        back = np.load(tmp_path / f"hits_{mname.lower()}.npz")
        for name, m in mats.items():
            x = m.astype(np.float32)
            if metric == fo.METRIC_INNER_PRODUCT:
                fo.normalize_L2(x)
            D_ref, I_ref = fo.knn_flat(x, x, 11, metric)
            check_parity(np.ascontiguousarray(scores[name]), np.ascontiguousarray(back[name]),
                         np.ascontiguousarray(D_ref[:, 1:]), np.ascontiguousarray(I_ref[:, 1:]), x, x, metric,
                         max_excused_frac=1e-2)


@pytest.mark.gpu
def test_pfam_proteins_search_flat_flow(tmp_path):
    """pfam/proteins_search.py:17-57, mode 'flat': normalise in place, IndexFlat IP, train, add,
    write_index, all-vs-all k=1000, save scores/hits; consumer expects column 0 = self
    (pfam/proteins.py:85-122)."""
    import faiss

    rng = np.random.default_rng(1)
    centers = rng.standard_normal((40, 1024)).astype(np.float32)
    emb = (centers[rng.integers(0, 40, 9000)] + 0.35 * rng.standard_normal((9000, 1024))).astype(np.float32)
    np.save(tmp_path / "full_sequences.npy", emb.astype(np.float16))

    embeddings = np.load(tmp_path / "full_sequences.npy").astype(np.float32)
    faiss.normalize_L2(embeddings)
    index = faiss.IndexFlat(embeddings.shape[1], faiss.METRIC_INNER_PRODUCT)
    index.train(embeddings)
    index.add(embeddings)
    index_file = tmp_path / "full_sequences_flat.index"
    faiss.write_index(index, str(index_file))
    assert index_file.stat().st_size == 45 + embeddings.nbytes  # what the driver prints as "Index: ..."
    flat_scores, flat_hits = index.search(embeddings, 1000)
    np.save(tmp_path / "full_sequences_flat_scores.npy", flat_scores)
    np.save(tmp_path / "full_sequences_flat_hits.npy", flat_hits)

    hits = np.load(tmp_path / "full_sequences_flat_hits.npy")
    assert hits.shape == (9000, 1000) and hits.dtype == np.int64
    assert np.array_equal(hits[:, 0], np.arange(9000))  # remove_self_hit finds nothing to fix
    x = np.load(tmp_path / "full_sequences.npy").astype(np.float32)
    fo.normalize_L2(x)
    sample = np.arange(0, 9000, 45)
    D_ref, I_ref = fo.knn_flat(x[sample], x, 1000, fo.METRIC_INNER_PRODUCT)
    check_parity(flat_scores[sample], hits[sample], D_ref, I_ref, x[sample], x, fo.METRIC_INNER_PRODUCT,
                 max_excused_frac=5e-3)
    # main.py:131-132: a pre-built index can be read back and searched
    again = faiss.read_index(str(index_file))
    D2, I2 = again.search(embeddings[:64], 13)
    assert np.array_equal(I2, hits[:64, :13])


@pytest.mark.gpu
def test_faiss_search_flow_mutates_inputs_like_the_reference(golden_dir):
    """seqvec_search/main.py:29-50 normalises queries and haystack IN PLACE; callers see it."""
    import faiss

    queries = np.load(golden_dir / "pfam-20-dist/test.npy")
    haystack = np.load(golden_dir / "pfam-20-dist/train.npy")
    q0 = queries.copy()
    faiss.normalize_L2(queries)
    faiss.normalize_L2(haystack)
    assert not np.array_equal(q0, queries)
    np.testing.assert_allclose(np.linalg.norm(queries, axis=1), 1.0, atol=1e-5)
    index = faiss.IndexFlat(haystack.shape[1], faiss.METRIC_INNER_PRODUCT)
    index.train(haystack)
    index.add(haystack)
    scores, result = index.search(queries, 13)
    D_ref, I_ref = fo.knn_flat(queries, haystack, 13, 0)
    check_parity(scores, result, D_ref, I_ref, queries, haystack, 0)


# ---- knn_b200.drivers: the same driver functions with one upload instead of three round trips -------------
@pytest.mark.gpu
def test_drivers_search_equals_stepwise_flow_and_golden(golden_dir, expected):
    """knn_b200.drivers.search == cath.search.search run step by step through the alias (bit-identical), and
    matches the (D, I) the UNMODIFIED reference driver returned over the oracle (golden)."""
    import faiss

    from knn_b200 import drivers

    emb = np.load(golden_dir / "pfam-20-10/train.npy")
    for mname, metric in [("ip", faiss.METRIC_INNER_PRODUCT), ("l2", faiss.METRIC_L2)]:
        before = emb.copy()
        I, D = drivers.search(emb, hits=10, metric=metric)
        assert np.array_equal(emb, before)  # the reference normalises a copy
        assert I.shape == (200, 10) and not I.flags.c_contiguous  # views, like results[:, 1:]
        e = emb.copy()
        if metric == faiss.METRIC_INNER_PRODUCT:
            faiss.normalize_L2(e)
        index = faiss.IndexFlat(e.shape[1], metric)
        index.add(e)
        s, r = index.search(e, 11)
        assert np.array_equal(I, r[:, 1:]) and np.array_equal(D, s[:, 1:])
        check_parity(np.ascontiguousarray(D), np.ascontiguousarray(I), expected[f"cath-search.pfam-20-10-train.{mname}.hits10.D"],
                     expected[f"cath-search.pfam-20-10-train.{mname}.hits10.I"], e, e, metric)


@pytest.mark.gpu
def test_drivers_faiss_search_reference_known_answers(golden_dir, expected):
    """tests/test_main.py:10-27 of the reference through knn_b200.drivers.faiss_search + knn_b200.evaluate_faiss."""
    import knn_b200
    from knn_b200 import drivers
    from oracle.evaluate import Fixture

    fx = Fixture(golden_dir / "small-random")
    queries, haystack = np.load(fx.test), np.load(fx.train)
    q0 = queries.copy()
    results, scores, seconds = drivers.faiss_search(haystack, queries, 5)
    assert seconds >= 0 and not np.array_equal(q0, queries)  # normalised in place (main.py:31)
    np.testing.assert_allclose(np.linalg.norm(haystack, axis=1), 1.0, atol=1e-5)  # main.py:34
    assert np.array_equal(results, expected["small-random.ip.k5.I"])
    auc1s, tps = knn_b200.evaluate_faiss(fx, results)
    assert auc1s == [1.0, 1 / 3, 2 / 3, 0.0, 0.0, 1 / 3] and tps == [1.0, 2 / 3, 2 / 3, 1.0, 1.0, 1.0]
    fx = Fixture(golden_dir / "pfam-20-10")
    results, scores, _ = drivers.faiss_search(np.load(fx.train), np.load(fx.test), 10)
    auc1s, tps = knn_b200.evaluate_faiss(fx, results)
    assert np.mean(auc1s) == 0.871 and np.mean(tps) == 0.91
    # a ready index as haystack (main.py:40-41)
    index = knn_b200.IndexFlat(1024, 0)
    hay = np.load(fx.train)
    knn_b200.normalize_L2(hay)
    index.add(hay)
    r2, s2, _ = drivers.faiss_search(index, np.load(fx.test), 10)
    assert np.array_equal(r2, results) and np.array_equal(s2, scores)


@pytest.mark.gpu
def test_drivers_search_and_save_and_proteins_search(tmp_path):
    from knn_b200 import drivers, read_index

    rng = np.random.default_rng(4)
    (tmp_path / "cath").mkdir()
    mats = {"aac": rng.random((300, 20)).astype(np.float32), "t5_half": rng.standard_normal((200, 1024)).astype(np.float16)}
    for name, m in mats.items():
        np.save(tmp_path / "cath" / f"{name}.npy", m)
    drivers.search_and_save(tmp_path / "cath")
    for metric_name, metric in [("cosine", 0), ("euclidean", 1)]:
        hits = np.load(tmp_path / "cath" / f"hits_{metric_name}.npz")
        scores = np.load(tmp_path / "cath" / f"scores_{metric_name}.npz")
        for name, m in mats.items():
            I, D = drivers.search(m.astype(np.float32), metric=metric)
            assert np.array_equal(hits[name], I) and np.array_equal(scores[name], D)
            assert float((tmp_path / "cath" / f"{name}.{metric_name}-search-time.txt").read_text()) >= 0
    emb = rng.standard_normal((3000, 1024)).astype(np.float32)
    (tmp_path / "pfam").mkdir()
    np.save(tmp_path / "pfam" / "full_sequences.npy", emb.astype(np.float16))
    scores, hits = drivers.proteins_search(tmp_path / "pfam", "flat", k=1000)
    assert np.array_equal(np.load(tmp_path / "pfam" / "full_sequences_flat_hits.npy"), hits)
    assert np.array_equal(hits[:, 0], np.arange(3000)) and hits.shape == (3000, 1000)
    x = emb.astype(np.float16).astype(np.float32)
    fo.normalize_L2(x)
    D_ref, I_ref = fo.knn_flat(x[:50], x, 1000, 0)
    check_parity(scores[:50], hits[:50], D_ref, I_ref, x[:50], x, 0, max_excused_frac=5e-3)
    again = read_index(str(tmp_path / "pfam" / "full_sequences_flat.index"))
    assert again.ntotal == 3000
    with pytest.raises(ValueError):
        drivers.proteins_search(tmp_path / "pfam", "hnsw")


@pytest.mark.gpu
def test_create_index_flat_then_search_and_check_auc1(golden_dir, tmp_path, expected):
    """/root/reference/tests/test_utils.py:17-22 (`create_index.main(["--dir", ..., "--index", ...])`, file exists) plus
    the "Search and check AUC1" the reference left as a TODO: the file is loaded the way seqvec_search/main.py:131-132
    does (`read_index`), handed to `faiss_search` as a ready index, and must reproduce the known answers of
    tests/test_main.py:26-27 (0.871 / 0.91) - identical to searching `train.npy` directly."""
    import knn_b200
    from knn_b200 import drivers
    from oracle.evaluate import Fixture

    fx = Fixture(golden_dir / "pfam-20-10")
    out = tmp_path / "index.bin"
    drivers.create_index(["--dir", str(golden_dir / "pfam-20-10"), "--index", str(out)])
    assert out.exists()
    assert out.read_bytes()[:4] == b"IxFI" and out.stat().st_size == 45 + 200 * 1024 * 4
    knn_index = knn_b200.read_index(str(out))
    assert (knn_index.ntotal, knn_index.d, knn_index.metric_type) == (200, 1024, 0)
    results, scores, _ = drivers.faiss_search(knn_index, np.load(fx.test), 10)
    auc1s, tps = knn_b200.evaluate_faiss(fx, results)
    assert np.mean(auc1s) == 0.871 and np.mean(tps) == 0.91
    assert np.array_equal(results, expected["pfam-20-10.ip.k10.I"])
    direct, direct_scores, _ = drivers.faiss_search(np.load(fx.train), np.load(fx.test), 10)
    assert np.array_equal(results, direct) and np.array_equal(scores, direct_scores)
    with pytest.raises(NotImplementedError, match="exact"):
        drivers.create_index(["--dir", str(golden_dir / "pfam-20-10"), "--index", str(out), "--kind", "lsh"])
    with pytest.raises(NotImplementedError, match="exact"):
        knn_b200.IndexLSH(1024, 1024)
