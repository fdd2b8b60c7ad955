#!/usr/bin/env python3
"""Generates tests/golden/ from the reference tree.  Run ONCE in the build container
(`python tests/golden/make_golden.py`); the GPU box has no /root/reference, so the outputs
are committed.

What it does
1. copies the reference's own test fixtures (data files, not source):
   test-data/{small-random,pfam-20-10,pfam-20-10-sum,pfam-20-dist}/{train,test}.npy + id JSONs;
   checks that small-random regenerates bit-exactly from its recipe
   (test-data/small-random/generate_arrays.py:5-7, numpy seed 7);
2. imports the UNMODIFIED reference drivers (seqvec_search.main.faiss_search /
   evaluate_faiss, cath.search.search) with ``faiss`` bound to oracle/flat_oracle.py (the
   real faiss-cpu wheel is not installable offline) and empty stand-ins for the plotting
   imports, runs the reference's two known-answer tests (tests/test_main.py:10-27) and
   asserts their constants -> this is what pins the oracle;
3. stores the (D, I) the reference drivers returned, plus oracle outputs for cases the
   reference never tests (L2, k > ntotal, k = 13 default), as golden vectors in
   ``expected.npz`` / ``expected.json``.
"""
import json
import shutil
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))

from oracle import flat_oracle  # noqa: E402

FIXTURES = ["small-random", "pfam-20-10", "pfam-20-10-sum", "pfam-20-dist"]


def copy_fixtures():
    for name in FIXTURES:
        dst = HERE / name
        dst.mkdir(exist_ok=True)
        for f in ["train.npy", "test.npy", "train.json", "test.json", "ids_to_family.json"]:
            shutil.copyfile(REF / "test-data" / name / f, dst / f)
            (dst / f).chmod(0o644)
    np.random.seed(7)
    test = np.random.rand(6, 1024).astype(np.float32)
    train = np.random.rand(11, 1024).astype(np.float32)
    assert np.array_equal(test, np.load(HERE / "small-random/test.npy"))
    assert np.array_equal(train, np.load(HERE / "small-random/train.npy"))


def import_reference():
    """Bind `faiss` to the oracle and stub the imports that are missing in this image."""
    sys.modules["faiss"] = flat_oracle
    for missing in ["matplotlib", "matplotlib.pyplot", "h5py", "humanize", "seaborn"]:
        if missing not in sys.modules:
            try:
                __import__(missing)
            except ImportError:
                m = types.ModuleType(missing)
                m.__path__ = []
                sys.modules[missing] = m
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "rcParams"):  # stand-in: seqvec_search/utils.py:18 writes rcParams at import
        mpl.rcParams = {}
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
        mpl.pyplot.Figure = object
    sys.path.insert(0, str(REF))
    import seqvec_search.main as ref_main  # noqa
    import cath.search as ref_cath  # noqa
    from seqvec_search.data import LoadedData  # noqa

    return ref_main, ref_cath, LoadedData


def main():
    copy_fixtures()
    ref_main, ref_cath, LoadedData = import_reference()
    expected = {}
    summary = {}

    # --- reference test 1: tests/test_main.py:10-18 ----------------------------------
    data = LoadedData.from_options(path=HERE / "small-random", hits=5)
    queries = np.load(str(data.test))
    results, scores, _ = ref_main.faiss_search(np.load(str(data.train)), queries, data.hits)
    auc1s, tps = ref_main.evaluate_faiss(data, results)
    assert auc1s == [1.0, 1 / 3, 2 / 3, 0.0, 0.0, 1 / 3], auc1s
    assert tps == [1.0, 2 / 3, 2 / 3, 1.0, 1.0, 1.0], tps
    expected["small-random.ip.k5.I"] = results
    expected["small-random.ip.k5.D"] = scores
    summary["small-random"] = {"k": 5, "auc1s": auc1s, "tps": tps}

    # --- reference test 2 (kNN half): tests/test_main.py:21-27 -----------------------
    data = LoadedData.from_options(path=HERE / "pfam-20-10", hits=10)
    queries = np.load(str(data.test))
    results, scores, _ = ref_main.faiss_search(np.load(str(data.train)), queries, data.hits)
    auc1s, tps = ref_main.evaluate_faiss(data, results)
    assert np.mean(auc1s) == 0.871 and np.mean(tps) == 0.91, (np.mean(auc1s), np.mean(tps))
    expected["pfam-20-10.ip.k10.I"] = results
    expected["pfam-20-10.ip.k10.D"] = scores
    summary["pfam-20-10"] = {"k": 10, "mean_auc1": float(np.mean(auc1s)), "mean_tp": float(np.mean(tps))}

    # --- regression values at the default k=13 (seqvec_search/constants.py:3) --------
    for name in ["pfam-20-10-sum", "pfam-20-dist"]:
        data = LoadedData.from_options(path=HERE / name)
        queries = np.load(str(data.test))
        results, scores, _ = ref_main.faiss_search(np.load(str(data.train)), queries, data.hits)
        auc1s, tps = ref_main.evaluate_faiss(data, results)
        expected[f"{name}.ip.k13.I"] = results
        expected[f"{name}.ip.k13.D"] = scores
        summary[name] = {"k": 13, "mean_auc1": float(np.mean(auc1s)), "mean_tp": float(np.mean(tps))}

    # --- cath.search.search (cath/search.py:13-26): all-vs-all, both metrics ----------
    emb = np.load(HERE / "pfam-20-10/train.npy")
    for mname, metric in [("ip", flat_oracle.METRIC_INNER_PRODUCT), ("l2", flat_oracle.METRIC_L2)]:
        I, D = ref_cath.search(emb, hits=10, metric=metric)
        expected[f"cath-search.pfam-20-10-train.{mname}.hits10.I"] = np.ascontiguousarray(I)
        expected[f"cath-search.pfam-20-10-train.{mname}.hits10.D"] = np.ascontiguousarray(D)

    # --- cases the reference never tests (oracle-only golden vectors) -----------------
    xb = np.load(HERE / "small-random/train.npy")
    xq = np.load(HERE / "small-random/test.npy")
    for mname, metric in [("ip", 0), ("l2", 1)]:
        idx = flat_oracle.IndexFlat(1024, metric)
        idx.add(xb)
        D, I = idx.search(xq, 16)  # k > ntotal = 11 -> -1 padding
        expected[f"small-random.raw.{mname}.k16.I"] = I
        expected[f"small-random.raw.{mname}.k16.D"] = D

    np.savez(HERE / "expected.npz", **expected)
    (HERE / "expected.json").write_text(json.dumps(summary, indent=1))
    print(json.dumps(summary, indent=1))
    print("wrote", len(expected), "arrays")


if __name__ == "__main__":
    main()
