"""Parity rule between an engine's (D, I) and the oracle's (TEST INFRASTRUCTURE ONLY).

BASELINE.json: "neighbour IDs must be identical to the reference's faiss IndexFlatIP on
the same inputs, except at exact-score ties, and distances must agree within 1e-5
relative."  Two fp32 evaluations of the same inner product that sum in a different order
differ by accumulation noise (SURVEY.md section 7 "Parity definition": two CPU fp32 orders
already swap 9/512 rows at N=200k), so an "exact-score tie" is decided in fp64:

* ids must match position by position;
* a position where they differ is excused iff the fp64 scores of the two ids involved
  differ by at most ``tau = 2*sqrt(d)*2^-24 * scale`` (scale = |x||y| for IP,
  |x|^2+|y|^2 for L2), i.e. both orders are valid orderings up to fp32 noise;
* no id may appear twice in a row, padding (-1) must match exactly;
* ``|D - D_ref| <= 1e-5*|D_ref| + tau`` everywhere (tau covers scores near zero where a
  relative bound is meaningless).

The number of excused positions is returned so that tests can bound it.
"""
from __future__ import annotations

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


class ParityError(AssertionError):
    pass


def tie_tolerance(d: int) -> float:
    return 2.0 * np.sqrt(d) * 2.0 ** -24


def check_parity(D, I, D_ref, I_ref, xq, xb, metric, *, rtol=1e-5, max_excused_frac=None):
    """Raise ParityError unless (D, I) matches (D_ref, I_ref) under the rule above.

    ``xq``/``xb`` are the fp32 matrices actually searched (after any normalisation).
    Returns a dict with the number of excused tie positions.
    """
    D = np.asarray(D)
    I = np.asarray(I)
    D_ref = np.asarray(D_ref)
    I_ref = np.asarray(I_ref)
    if D.shape != D_ref.shape or I.shape != I_ref.shape or D.shape != I.shape:
        raise ParityError("shape mismatch %s %s vs %s %s" % (D.shape, I.shape, D_ref.shape, I_ref.shape))
    if D.dtype != np.float32 or I.dtype != np.int64:
        raise ParityError("dtype mismatch: D %s I %s" % (D.dtype, I.dtype))
    nq, k = I.shape
    d = xq.shape[1]
    tau_unit = tie_tolerance(d)
    xq64 = None
    excused = 0
    qn = np.sqrt((xq.astype(np.float64) ** 2).sum(1))
    bn_max = float(np.sqrt((xb.astype(np.float64) ** 2).sum(1)).max()) if xb.shape[0] else 0.0

    pad = I_ref < 0
    if not np.array_equal(pad, I < 0):
        raise ParityError("padding positions differ")
    if pad.any() and not np.array_equal(D[pad], D_ref[pad]):
        raise ParityError("padding distance values differ")

    # --- distances -------------------------------------------------------------------
    if metric == METRIC_INNER_PRODUCT:
        scale = qn[:, None] * bn_max
    else:
        scale = (qn[:, None] ** 2 + bn_max ** 2)
    tol = rtol * np.abs(D_ref.astype(np.float64)) + tau_unit * scale
    bad = (np.abs(D.astype(np.float64) - D_ref.astype(np.float64)) > tol) & ~pad
    if bad.any():
        r, c = np.argwhere(bad)[0]
        raise ParityError(
            "distance mismatch at (%d,%d): %r vs ref %r (tol %g); %d bad of %d"
            % (r, c, D[r, c], D_ref[r, c], tol[r, c], int(bad.sum()), bad.size)
        )

    # --- ids ---------------------------------------------------------------------------
    diff_rows = np.flatnonzero((I != I_ref).any(axis=1))
    for r in diff_rows:
        ours = I[r][~pad[r]]
        ref = I_ref[r][~pad[r]]
        if len(set(ours.tolist())) != ours.size:
            raise ParityError("row %d: duplicate ids" % r)
        if ours.min() < 0 or ours.max() >= xb.shape[0]:
            raise ParityError("row %d: id out of range" % r)
        if xq64 is None:
            xq64 = xq.astype(np.float64)
        pos = np.flatnonzero(ours != ref)
        a = xb[ours[pos]].astype(np.float64)
        b = xb[ref[pos]].astype(np.float64)
        q = xq64[r]
        if metric == METRIC_INNER_PRODUCT:
            sa, sb = a @ q, b @ q
            sc = qn[r] * bn_max
        else:
            sa = ((a - q) ** 2).sum(1)
            sb = ((b - q) ** 2).sum(1)
            sc = qn[r] ** 2 + bn_max ** 2
        tau = tau_unit * sc
        worst = np.abs(sa - sb).max()
        if worst > tau:
            p = pos[np.abs(sa - sb).argmax()]
            raise ParityError(
                "row %d pos %d: id %d vs ref %d, fp64 scores differ by %g > tau %g"
                % (r, p, ours[p], ref[p], worst, tau)
            )
        excused += pos.size
    frac = excused / max(1, I.size)
    if max_excused_frac is not None and frac > max_excused_frac:
        raise ParityError("too many tie-excused positions: %d of %d" % (excused, I.size))
    return {"excused": excused, "positions": int(I.size), "rows_with_ties": int(diff_rows.size)}
