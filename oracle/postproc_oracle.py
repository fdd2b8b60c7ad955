"""CPU restatement of what the reference does with (D, I) right after the flat search
(SURVEY.md section 8, rows f3 and f4).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and the cpu_baseline leg of the benches, never by the product path.

Every function is a plain-Python / numpy loop that follows the reference line by line:

  evaluate_counts          seqvec_search/main.py:53-82   (evaluate_faiss + evaluate)
  compute_is_correct       cath/cath.py:76-84
  compute_correctness_array pfam/proteins.py:201-207
  compute_auc1             pfam/proteins_shared.py:139-157
  remove_self_hit          pfam/proteins.py:85-122
  write_prefilter_db       seqvec_search/mmseqs/_write_prefilter_db.py:52-97

Pinned status: tests/golden/make_golden_postproc.py runs the UNMODIFIED reference functions in
the build container (the functions that live in script files with import-time side effects are
extracted from the reference file with ``ast`` at generation time and executed, not copied) and
stores their outputs under tests/golden/postproc/; tests/test_postproc_oracle.py checks this
restatement against those vectors.  ``evaluate`` is additionally pinned by the reference's own
known-answer test (tests/test_main.py:17-18,26-27).

Python / numpy indexing semantics the reference relies on implicitly are kept: a label gather
with id -1 wraps to the last row (``train_ids[-1]``, ``mapping_array[-1]``).
"""
from __future__ import annotations

import numpy as np


# ---- seqvec_search/main.py:53-82 ---------------------------------------------------------------
def evaluate_counts(results, query_family, db_family):
    """Integer form of ``evaluate``: per query the length of the leading run of hits of the
    query's family (AUC1 numerator), the number of hits of that family (TP numerator) and the
    family's size in the database (the common denominator, main.py:68).

    results: (nq, k) ints; query_family: (nq,) family code per query; db_family: (N,) per row.
    """
    db_family = np.asarray(db_family)
    sizes = {}
    for f in db_family.tolist():
        sizes[f] = sizes.get(f, 0) + 1
    lead, tp, size = [], [], []
    for q, row in enumerate(np.asarray(results).tolist()):
        correct = int(query_family[q])
        fam = [int(db_family[i]) for i in row]  # negative ids wrap like the list indexing of main.py:58
        tp.append(sum(f == correct for f in fam))
        run = 0
        for f in fam:
            if f == correct:
                run += 1
            else:
                break
        lead.append(run)
        size.append(sizes.get(correct, 0))
    return np.asarray(lead, np.int32), np.asarray(tp, np.int32), np.asarray(size, np.int32)


def evaluate(results, query_family, db_family):
    """(auc1s, tps) as Python float lists, exactly main.py:80-81 (int / int true division)."""
    lead, tp, size = evaluate_counts(results, query_family, db_family)
    return [int(a) / int(s) for a, s in zip(lead, size)], [int(t) / int(s) for t, s in zip(tp, size)]


# ---- cath/cath.py:76-84 ------------------------------------------------------------------------
def compute_is_correct(results, mapping_array):
    """queries -> levels -> hits: out[q, l, h] = mapping[q, l] == mapping[results[q, h], l]."""
    mapping_array = np.asarray(mapping_array)
    return np.asarray([(mapping_array[q] == mapping_array[row]).T for q, row in zip(range(len(results)), results)])


# ---- pfam/proteins.py:201-207 ------------------------------------------------------------------
def compute_correctness_array(full, homologous_int):
    """out[q, h] = full[q, h] in set(homologous_int[q]) (plain value membership, no wrap)."""
    out = []
    for q, hits in enumerate(np.asarray(full).tolist()):
        s = set(int(v) for v in homologous_int[q])
        out.append([h in s for h in hits])
    return np.asarray(out, dtype=bool).reshape(len(full), -1)


# ---- pfam/proteins_shared.py:139-157 -----------------------------------------------------------
def compute_auc1(hits, homologous_int, set_sizes, n_db=None):
    """Leading run of hits that are homologs of the query, divided by max(len(set), 1).
    ``homologous_int[q]`` holds the database row numbers of the query's homolog set that exist
    in the database; ``set_sizes[q]`` is len(homologous_proteins[query]) (it may count names
    absent from target_ids).  Negative hit ids wrap like target_ids[hit] (proteins_shared.py:152)
    when the database size ``n_db`` is given."""
    out = []
    for q, row in enumerate(np.asarray(hits).tolist()):
        s = set(int(v) for v in homologous_int[q])
        run = 0
        for h in row:
            if h < 0 and n_db is not None:
                h += n_db
            if h in s:
                run += 1
            else:
                break
        out.append(run / max(int(set_sizes[q]), 1))
    return np.asarray(out)


# ---- pfam/proteins.py:85-122 -------------------------------------------------------------------
def remove_self_hit(hits, scores, self_ids):
    """In place, like the reference: where the first hit is not the query itself, the self hit
    (first occurrence; the LAST column when absent) is rotated to column 0.  Returns the views
    hits[:, 1:], scores[:, 1:] and the number of rows whose self hit was missing."""
    self_ids = np.asarray(self_ids)
    bogus = 0
    for r in np.argwhere(hits[:, 0] != self_ids)[:, 0]:
        sid = self_ids[r]
        row = list(hits[r])
        if sid in row:
            index = row.index(sid)
        else:
            index = len(row) - 1
            bogus += 1
        hits[r, 0], hits[r, 1:index + 1] = hits[r, index].copy(), hits[r, 0:index].copy()
        scores[r, 0], scores[r, 1:index + 1] = scores[r, index].copy(), scores[r, 0:index].copy()
    return hits[:, 1:], scores[:, 1:], bogus


# ---- seqvec_search/mmseqs/_write_prefilter_db.py:52-97 -----------------------------------------
def score_to_int(score_f32: np.float32) -> int:
    """int(numpy.clip(score, -1e30, 1e30) * 100) with float32 arithmetic throughout - what
    numpy >= 2 (NEP 50: the Python-int bounds and the factor 100 are weak scalars) computes at
    _write_prefilter_db.py:76,88 for the float32 scores faiss returns."""
    s = np.float32(score_f32)
    lim = np.float32(10 ** 30)
    if s != s:
        raise ValueError("cannot convert float NaN to integer")
    c = np.minimum(np.maximum(s, -lim), lim)
    with np.errstate(over="ignore"):
        v = np.float32(c * np.float32(100))
    return int(v)


def write_prefilter_db(hits, queries, scores, test_map, train_map):
    """Returns (data_bytes, index_bytes) of the `.0` and `.index` files (the `.dbtype` file is the
    constant b"\\x07\\x00\\x00\\x00", _write_prefilter_db.py:66)."""
    data, index = bytearray(), bytearray()
    offset = 0
    for query, hit_entry, score_entry in zip(np.asarray(queries).tolist(), np.asarray(hits).tolist(), np.asarray(scores, np.float32)):
        length = 0
        for hit, score in zip(hit_entry, score_entry):
            if hit == -1:
                continue
            line = f"{int(train_map[hit])}\t{score_to_int(score)}\t0\n".encode()
            length += len(line)
            data += line
        data += b"\0"
        length += 1
        index += f"{int(test_map[query])}\t{offset}\t{length}\n".encode()
        offset += length
    return bytes(data), bytes(index)
