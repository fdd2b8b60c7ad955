#!/usr/bin/env python3
"""Generates tests/golden/postproc/ by running the UNMODIFIED reference post-processing
functions in the build container.  The GPU box has no
/root/reference, so the outputs are committed; nothing of the reference's source is.

  seqvec_search.main.evaluate_faiss            imported (main.py:53-82)
  seqvec_search.mmseqs.write_prefilter_db      imported (_write_prefilter_db.py:52-97)
  cath.cath.compute_is_correct                 \\  these live in script files that load datasets at
  pfam.proteins.remove_self_hit                 > import time: the function's source is cut out of the
  pfam.proteins.compute_correctness_array      /  reference file with `ast` here, at generation time,
  pfam.proteins_shared.compute_auc1               and executed with synthetic globals

The inputs are synthetic (seeded) or the reference's own fixtures; inputs and outputs are stored
together in postproc.npz so that the tests need nothing else.
"""
import ast
import sys
import tempfile
from pathlib import Path
from typing import Dict, List, Set, Tuple

import numpy
import numpy as np
from numpy import ndarray

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(HERE))

import make_golden  # noqa: E402  (faiss -> oracle binding and import stubs)


def extract(path: Path, name: str, glb: dict):
    """Compile one top-level function of a reference file without importing the file."""
    tree = ast.parse(path.read_text())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[node], type_ignores=[])
    exec(compile(mod, str(path), "exec"), glb)
    return glb[name]


def main():
    out = {}
    rng = np.random.default_rng(20261018)
    ref_main, _, LoadedData = make_golden.import_reference()
    exp = np.load(HERE / "expected.npz")

    # ---- evaluate_faiss on the reference fixtures, golden I and random I -----------------------
    for name, key in [("small-random", "small-random.ip.k5.I"), ("pfam-20-10", "pfam-20-10.ip.k10.I"),
                      ("pfam-20-dist", "pfam-20-dist.ip.k13.I")]:
        data = LoadedData.from_options(path=HERE / name, hits=5)
        for tag, res in [("golden", exp[key]),
                         ("random", rng.integers(0, len(data.train_ids), size=(len(data.test_ids), 17)))]:
            auc1s, tps = ref_main.evaluate_faiss(data, res)
            out[f"evaluate.{name}.{tag}.I"] = np.asarray(res, np.int64)
            out[f"evaluate.{name}.{tag}.auc1"] = np.asarray(auc1s, np.float64)
            out[f"evaluate.{name}.{tag}.tp"] = np.asarray(tps, np.float64)

    # ---- compute_is_correct (cath/cath.py:76-84) ------------------------------------------------
    n, hits = 300, 12
    codes = rng.integers(1, 4, size=(n, 4))
    mapping_array = np.asarray([tuple(".".join(str(v) for v in row[:l + 1]) for l in range(4)) for row in codes])
    results = rng.integers(0, n, size=(n, hits))
    results[5, 3] = -1  # numpy fancy indexing wraps
    f = extract(REF / "cath/cath.py", "compute_is_correct", dict(numpy=numpy, ndarray=ndarray, mapping_array=mapping_array))
    out["is_correct.mapping"] = mapping_array
    out["is_correct.results"] = results
    out["is_correct.out"] = f(results)

    # ---- compute_correctness_array (pfam/proteins.py:201-207) ----------------------------------
    n, hits = 200, 25
    homologs = [sorted(set(rng.integers(0, n, size=rng.integers(0, 30)).tolist())) for _ in range(n)]
    full = rng.integers(-1, n, size=(n, hits))
    f = extract(REF / "pfam/proteins.py", "compute_correctness_array",
                dict(numpy=numpy, ndarray=ndarray, tqdm=lambda x: x, homologous_proteins_int=homologs))
    out["correctness.full"] = full
    out["correctness.offsets"] = np.cumsum([0] + [len(h) for h in homologs]).astype(np.int64)
    out["correctness.members"] = np.asarray([v for h in homologs for v in h], np.int64)
    out["correctness.out"] = f(full)

    # ---- compute_auc1 (pfam/proteins_shared.py:139-157) ----------------------------------------
    target_ids = [f"P{i}" for i in range(n)]
    queries = [f"P{i}" for i in range(n)]
    hom_names = {q: set(target_ids[v] for v in h) | ({"absent_" + q} if i % 7 == 0 else set())
                 for i, (q, h) in enumerate(zip(queries, homologs))}
    hits_arr = np.asarray([[h[j % len(h)] if h and j < (q % 5) else rng.integers(0, n) for j in range(hits)]
                           for q, h in enumerate(homologs)])
    hits_arr[3, 0] = -1
    f = extract(REF / "pfam/proteins_shared.py", "compute_auc1", dict(numpy=numpy, ndarray=ndarray, Dict=Dict, Set=Set, List=List))
    out["auc1.hits"] = hits_arr
    out["auc1.set_sizes"] = np.asarray([len(hom_names[q]) for q in queries], np.int64)
    out["auc1.out"] = f(hits_arr, hom_names, queries, target_ids)

    # ---- remove_self_hit (pfam/proteins.py:85-122) ---------------------------------------------
    n, k = 64, 9
    self_ids = np.arange(0, n)
    hits0 = np.stack([rng.permutation(n + 5)[:k] for _ in range(n)])
    for r in range(n):  # most rows: self first; some: self elsewhere; some: absent
        if r % 4 != 3 and r not in hits0[r]:
            hits0[r, 0] = r
        elif r % 4 == 3:
            hits0[r][hits0[r] == r] = n + 6
    for r in range(0, n, 5):
        pos = list(hits0[r]).index(r) if r in hits0[r] else None
        if pos is not None:
            hits0[r, [pos, (r // 5) % k]] = hits0[r, [(r // 5) % k, pos]]
    scores0 = rng.random((n, k)).astype(np.float32)
    h, s = hits0.copy(), scores0.copy()
    import io, contextlib
    f = extract(REF / "pfam/proteins.py", "remove_self_hit",
                dict(numpy=numpy, ndarray=ndarray, Tuple=Tuple, original_full_sequences_ids=list(range(n)), subsampler=slice(None, None, 1)))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ho, so = f(h, s)
    out["selfhit.hits_in"] = hits0
    out["selfhit.scores_in"] = scores0
    out["selfhit.hits_inplace"] = h
    out["selfhit.scores_inplace"] = s
    out["selfhit.hits_out"] = np.ascontiguousarray(ho)
    out["selfhit.scores_out"] = np.ascontiguousarray(so)
    out["selfhit.bogus"] = np.asarray(int(buf.getvalue().split("There are ")[1].split(" ")[0]))

    # ---- write_prefilter_db (_write_prefilter_db.py:52-97) -------------------------------------
    from seqvec_search.mmseqs import write_prefilter_db

    def run_writer(tag, hits, queries, scores, test_map, train_map):
        with tempfile.TemporaryDirectory() as td:
            db = Path(td) / "prefilter"
            write_prefilter_db(hits, db, queries, scores, test_map, train_map)
            out[f"prefilter.{tag}.hits"] = hits
            out[f"prefilter.{tag}.queries"] = queries
            out[f"prefilter.{tag}.scores"] = scores
            out[f"prefilter.{tag}.test_map"] = test_map
            out[f"prefilter.{tag}.train_map"] = train_map
            out[f"prefilter.{tag}.data"] = np.frombuffer(db.with_suffix(".0").read_bytes(), np.uint8)
            out[f"prefilter.{tag}.index"] = np.frombuffer(db.with_suffix(".index").read_bytes(), np.uint8)
            assert db.with_suffix(".dbtype").read_bytes() == b"\x07\x00\x00\x00"

    I = exp["pfam-20-10.ip.k10.I"]
    D = exp["pfam-20-10.ip.k10.D"]
    run_writer("pfam-20-10", I, np.arange(I.shape[0]), D, rng.permutation(I.shape[0]).astype(np.int64),
               (rng.permutation(200) * 37).astype(np.int64))
    # edge cases: missing hits, negative / huge / tiny / infinite scores, 19-digit ids, L2-sized distances
    nq, k = 40, 14
    hits = rng.integers(0, 50, size=(nq, k)).astype(np.int64)
    hits[rng.random((nq, k)) < 0.15] = -1
    hits[7] = -1
    scores = (rng.standard_normal((nq, k)) * 10.0 ** rng.integers(-3, 36, size=(nq, k))).astype(np.float32)
    scores[0, :8] = [0.0, -0.0, 0.28999999, 0.29, -0.005, 0.999999, 1e30, -1e30]
    scores[1, :6] = [np.inf, -np.inf, 3.4e38, -3.4e38, 1e-45, 123456.789]
    scores[2, :4] = [2.0 ** 63 / 100, 2.0 ** 64 / 100, -(2.0 ** 63) / 100, 9.2e16]
    train_map = rng.integers(0, 2 ** 62, size=50).astype(np.int64)
    train_map[:3] = [0, 9223372036854775807, 10 ** 18]
    test_map = rng.integers(0, 10 ** 9, size=nq).astype(np.int64)
    with np.errstate(all="ignore"):
        run_writer("edge", hits, rng.permutation(nq).astype(np.int64), scores, test_map, train_map)

    np.savez_compressed(HERE / "postproc" / "postproc.npz", **out)
    print("wrote", len(out), "arrays,", (HERE / "postproc" / "postproc.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
