"""Device-side versions of what the reference's drivers do with (D, I) after index.search
(SURVEY.md section 8 rows f3, f4).  Same names and argument meaning as the reference functions
(paths relative to /root/reference), results bit-identical to them:

  evaluate_faiss / evaluate_ids   seqvec_search/main.py:53-82
  compute_is_correct              cath/cath.py:76-84
  compute_correctness_array       pfam/proteins.py:201-207
  compute_auc1                    pfam/proteins_shared.py:139-157
  remove_self_hit                 pfam/proteins.py:85-122
  write_prefilter_db              seqvec_search/mmseqs/_write_prefilter_db.py:52-97

Inputs may be numpy arrays or CUDA torch tensors (e.g. the (D, I) a device-path search just
returned: nothing leaves the GPU between the search and these kernels).  The string -> integer
label encoding the reference does implicitly with dicts stays on the host (it is O(N) set-up,
not per-hit work); every per-hit loop runs in libknn_b200.so.  No CPU fallback.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import _lib
from .index import _default_device, _is_torch, _torch_stream

_ERR_TEXT = {1: "cannot convert float NaN to integer", 2: "index out of bounds (hit id outside the label / id table)",
             4: "cannot convert float infinity to integer"}


def _dev(x, dtype, device):
    """numpy / torch -> contiguous CUDA tensor of `dtype` on cuda:device (no copy when it already is one)."""
    import torch

    if _is_torch(x):
        return x.to(device=torch.device("cuda", device), dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), device=torch.device("cuda", device)).to(dtype).contiguous()


def _device_of(*xs) -> int:
    for x in xs:
        if _is_torch(x) and x.is_cuda:
            return x.device.index
    return _default_device()


def _raise_flags(err) -> None:
    e = int(err.item())
    if e & 1:
        raise ValueError(_ERR_TEXT[1])
    if e & 4:
        raise OverflowError(_ERR_TEXT[4])
    if e & 2:
        raise IndexError(_ERR_TEXT[2])


def _out(t, like):
    return t if _is_torch(like) else t.cpu().numpy()


# ---- seqvec_search/main.py:53-82 ---------------------------------------------------------------
def evaluate_ids(results, query_family, db_family):
    """Integer core of ``evaluate``: (lead, tp, family_size) per query, int32.  AUC1 = lead / family_size,
    TP = tp / family_size (main.py:80-81)."""
    import torch

    dev = _device_of(results)
    I = _dev(results, torch.int64, dev)
    if I.dim() != 2:
        raise ValueError("results must be (nq, k)")
    qf = _dev(query_family, torch.int32, dev)
    df = _dev(db_family, torch.int32, dev)
    nq, k = I.shape
    if qf.numel() != nq:
        raise ValueError("one family label per query expected")
    lead = torch.empty(nq, dtype=torch.int32, device=I.device)
    tp = torch.empty(nq, dtype=torch.int32, device=I.device)
    err = torch.zeros(1, dtype=torch.int32, device=I.device)
    _lib.check(_lib.load().knn_eval_family_dev(nq, k, I.data_ptr(), qf.data_ptr(), df.data_ptr(), df.numel(), lead.data_ptr(),
                                               tp.data_ptr(), err.data_ptr(), _torch_stream(dev)))
    _raise_flags(err)
    # Counter(family of every database row) (main.py:68), looked up per query
    sizes = torch.bincount(df.to(torch.int64), minlength=int(qf.max().item()) + 1 if nq else 0)[qf.to(torch.int64)].to(torch.int32)
    return _out(lead, results), _out(tp, results), _out(sizes, results)


def evaluate_faiss(data, results):
    """Drop-in for seqvec_search.main.evaluate_faiss: ``data`` needs .test_ids, .train_ids and .ids_to_family;
    returns (auc1s, tps) as lists of Python floats, bit-identical to the reference's int / int divisions."""
    fams = {}
    code = lambda name: fams.setdefault(data.ids_to_family[name], len(fams))  # noqa: E731
    df = np.asarray([code(i) for i in data.train_ids], np.int32)
    qf = np.asarray([code(i) for i in data.test_ids], np.int32)
    res = results if _is_torch(results) else np.asarray(results)
    lead, tp, size = evaluate_ids(res, qf[:len(res)], df)
    if _is_torch(lead):
        lead, tp, size = lead.cpu().numpy(), tp.cpu().numpy(), size.cpu().numpy()
    # a query whose family is absent from the database raises KeyError in the reference (main.py:80)
    if (size == 0).any():
        raise KeyError(data.ids_to_family[data.test_ids[int(np.argmax(size == 0))]])
    return [int(a) / int(s) for a, s in zip(lead, size)], [int(t) / int(s) for t, s in zip(tp, size)]


# ---- cath/cath.py:76-84 ------------------------------------------------------------------------
def encode_levels(mapping_array) -> np.ndarray:
    """(N, levels) array of label strings -> int32 codes, equal strings <=> equal codes per column."""
    m = np.asarray(mapping_array)
    return np.stack([np.unique(m[:, l], return_inverse=True)[1] for l in range(m.shape[1])], axis=1).astype(np.int32)


def compute_is_correct(results, mapping_array):
    """queries -> levels -> hits boolean array; ``mapping_array`` as built by cath_shared.load_mapping (strings)
    or already encoded int32 codes."""
    import torch

    m = mapping_array
    if not _is_torch(m):
        m = np.asarray(m)
        if m.dtype.kind not in "iu":
            m = encode_levels(m)
    dev = _device_of(results, m)
    I = _dev(results, torch.int64, dev)
    M = _dev(m, torch.int32, dev)
    nq, k = I.shape
    levels = M.shape[1]
    out = torch.empty((nq, levels, k), dtype=torch.uint8, device=I.device)
    err = torch.zeros(1, dtype=torch.int32, device=I.device)
    _lib.check(_lib.load().knn_eval_levels_dev(nq, k, I.data_ptr(), M.data_ptr(), levels, M.shape[0], out.data_ptr(),
                                               err.data_ptr(), _torch_stream(dev)))
    _raise_flags(err)
    return _out(out.to(torch.bool), results)


# ---- pfam/proteins.py:201-207, pfam/proteins_shared.py:139-157 ---------------------------------
def sets_to_csr(sets):
    """list of iterables of database row numbers -> (offsets int64 (n+1), members int64 sorted within a set)."""
    rows = [np.unique(np.fromiter(s, dtype=np.int64, count=len(s))) for s in sets]
    offsets = np.zeros(len(rows) + 1, np.int64)
    np.cumsum([len(r) for r in rows], out=offsets[1:])
    members = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    return offsets, members.astype(np.int64)


def _eval_sets(hits, offsets, members, n_db_wrap, want_correct, want_lead):
    import torch

    dev = _device_of(hits)
    I = _dev(hits, torch.int64, dev)
    off = _dev(offsets, torch.int64, dev)
    mem = _dev(members, torch.int64, dev)
    nq, k = I.shape
    if off.numel() != nq + 1:
        raise ValueError("one homolog set per query expected")
    if mem.numel() == 0:
        mem = torch.zeros(1, dtype=torch.int64, device=I.device)
    correct = torch.empty((nq, k), dtype=torch.uint8, device=I.device) if want_correct else None
    lead = torch.empty(nq, dtype=torch.int32, device=I.device) if want_lead else None
    _lib.check(_lib.load().knn_eval_sets_dev(nq, k, I.data_ptr(), off.data_ptr(), mem.data_ptr(), int(n_db_wrap),
                                             correct.data_ptr() if want_correct else None,
                                             lead.data_ptr() if want_lead else None, _torch_stream(dev)))
    return correct, lead


def compute_correctness_array(full, homologous_proteins_int):
    """out[q, h] = full[q, h] in homologous_proteins_int[q]; the sets may be given as a list of iterables (as
    the reference's global is) or as a ready (offsets, members) CSR pair."""
    import torch

    off, mem = homologous_proteins_int if isinstance(homologous_proteins_int, tuple) else sets_to_csr(homologous_proteins_int)
    correct, _ = _eval_sets(full, off, mem, 0, True, False)
    return _out(correct.to(torch.bool), full)


def compute_auc1(hits, homologous_proteins, queries, target_ids):
    """Same signature as pfam.proteins_shared.compute_auc1: ``homologous_proteins`` maps a query name to the set
    of homologous protein names, ``target_ids[i]`` is the name of database row i.  Returns float64 AUC1s."""
    pos = {}
    for i, name in enumerate(target_ids):
        pos.setdefault(name, []).append(i)  # duplicate names: every row carrying the name is a member
    sets, sizes = [], []
    for q in queries:
        names = homologous_proteins[q]
        sets.append([i for n in names for i in pos.get(n, ())])
        sizes.append(max(len(names), 1))
    off, mem = sets_to_csr(sets)
    _, lead = _eval_sets(hits, off, mem, len(target_ids), False, True)
    return lead.cpu().numpy().astype(np.int64) / np.asarray(sizes, np.int64)


# ---- pfam/proteins.py:85-122 -------------------------------------------------------------------
def remove_self_hit(hits, scores, self_ids=None):
    """In place like the reference (the self hit is rotated to column 0), returns the views
    ``hits[:, 1:], scores[:, 1:]``.  ``self_ids`` defaults to 0..n-1 (the reference's
    ``numpy.arange(len(ids))[subsampler]``).  The number of rows without a self hit is in
    ``remove_self_hit.last_missing``."""
    import torch

    dev = _device_of(hits, scores)
    on_dev = _is_torch(hits)
    I = hits if on_dev else torch.as_tensor(hits).to(torch.device("cuda", dev))
    D = scores if _is_torch(scores) else torch.as_tensor(scores).to(torch.device("cuda", dev))
    if I.dtype != torch.int64 or D.dtype != torch.float32 or not I.is_contiguous() or not D.is_contiguous():
        raise TypeError("remove_self_hit works in place: contiguous int64 hits and float32 scores expected")
    nq, k = I.shape
    sid = _dev(self_ids, torch.int64, dev) if self_ids is not None else None
    missing = torch.zeros(1, dtype=torch.int64, device=I.device)
    _lib.check(_lib.load().knn_remove_self_hit_dev(nq, k, I.data_ptr(), D.data_ptr(), sid.data_ptr() if sid is not None else None,
                                                   missing.data_ptr(), _torch_stream(dev)))
    remove_self_hit.last_missing = int(missing.item())
    if not on_dev:  # mutate the caller's arrays like the reference does
        hits[...] = I.cpu().numpy()
        scores[...] = D.cpu().numpy()
    return hits[:, 1:], scores[:, 1:]


remove_self_hit.last_missing = 0


# ---- seqvec_search/mmseqs/_write_prefilter_db.py:52-97 -----------------------------------------
def format_prefilter_db(hits, queries, scores, test_faiss_to_mmseqs, train_faiss_to_mmseqs, clip: bool = True):
    """The bytes of the `.0` (data) and `.index` files as two uint8 CUDA tensors."""
    import torch

    dev = _device_of(hits, scores)
    I = _dev(hits, torch.int64, dev)
    D = _dev(scores, torch.float32, dev)
    if I.shape != D.shape or I.dim() != 2:
        raise ValueError("hits and scores must both be (nq, k)")
    Q = _dev(queries, torch.int64, dev)
    tm = _dev(test_faiss_to_mmseqs, torch.int64, dev)
    rm = _dev(train_faiss_to_mmseqs, torch.int64, dev)
    nq, k = I.shape
    if Q.numel() != nq:
        raise ValueError("one query id per row of hits expected")
    lib = _lib.load()
    s = _torch_stream(dev)
    sec_off = torch.empty(nq + 1, dtype=torch.int64, device=I.device)
    idx_off = torch.empty(nq + 1, dtype=torch.int64, device=I.device)
    err = torch.zeros(1, dtype=torch.int32, device=I.device)
    args = (nq, k, I.data_ptr(), D.data_ptr(), Q.data_ptr(), tm.data_ptr(), tm.numel(), rm.data_ptr(), rm.numel(), 1 if clip else 0)
    _lib.check(lib.knn_prefilter_measure_dev(*args, sec_off.data_ptr(), idx_off.data_ptr(), err.data_ptr(), s))
    _raise_flags(err)
    data = torch.empty(int(sec_off[-1].item()), dtype=torch.uint8, device=I.device)
    index = torch.empty(int(idx_off[-1].item()), dtype=torch.uint8, device=I.device)
    _lib.check(lib.knn_prefilter_emit_dev(*args, sec_off.data_ptr(), idx_off.data_ptr(), data.data_ptr(), index.data_ptr(),
                                          err.data_ptr(), s))
    _raise_flags(err)
    return data, index


def write_prefilter_db(hits, prefilter_db, queries, scores, test_faiss_to_mmseqs, train_faiss_to_mmseqs, clip: bool = True):
    """Same signature and files as seqvec_search.mmseqs.write_prefilter_db: `<prefilter_db>.dbtype`, `.0`, `.index`."""
    prefilter_db = Path(prefilter_db)
    data, index = format_prefilter_db(hits, queries, scores, test_faiss_to_mmseqs, train_faiss_to_mmseqs, clip)
    prefilter_db.with_suffix(".dbtype").write_bytes(b"\x07\x00\x00\x00")
    with prefilter_db.with_suffix(".0").open("wb") as f:
        f.write(memoryview(data.cpu().numpy()))
    with prefilter_db.with_suffix(".index").open("wb") as f:
        f.write(memoryview(index.cpu().numpy()))
