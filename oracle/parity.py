"""Parity rule between an engine's (D, I) and the oracle's (TEST INFRASTRUCTURE ONLY).

BASELINE.json: "neighbour IDs must be identical to the reference's faiss IndexFlatIP on
the same inputs, except at exact-score ties, and distances must agree within 1e-5
relative."  Two fp32 evaluations of the same inner product that sum in a different order
differ by accumulation noise (SURVEY.md section 7 "Parity definition": two CPU fp32 orders
already swap 9/512 rows at N=200k), so an "exact-score tie" is decided in fp64.  The rule
is SURVEY.md section 8(c)'s, all of it:

* ids must match position by position;
* a position where they differ is excused iff the fp64 scores of the two ids involved
  differ by at most ``tau = 2*sqrt(d)*2^-24 * scale`` (scale = |x||y| for IP,
  |x|^2+|y|^2 for L2), i.e. both orders are valid orderings up to fp32 noise;
* the reference row is cut into tau-clusters (maximal runs of consecutive entries whose fp64
  scores are within tau of their neighbour): inside a cluster the id MULTISET must match;
* the cluster that touches the k-boundary may exchange members with the reference's
  (k+1)-th, (k+2)-th ... candidates (pass a reference with more than k columns:
  ``flat_oracle.knn_flat(..., want=k+e)``) as long as those lie inside tau of the cluster;
  without the extra columns a foreign id at the boundary is arbitrated against the id it
  displaced;
* no id may appear twice in a row, padding (-1) must match exactly;
* ``|D - D_ref| <= max(1e-5*|D_ref|, tau/8)``: strictly the stated relative bound - no
  additive slack - wherever it is physically meaningful.  ``tau/8`` (= 8 ulp of a
  scale-sized float at d = 1024) is the noise floor of an fp32 sum of d products: two valid
  fp32 evaluations of the same score (another summation order; for L2 the reference's own
  ``|x|^2+|y|^2-2<x,y>``, which cancels) differ by that much however small the score is,
  so below ``|D_ref| = 12500*tau`` (0.048 for unit vectors at d = 1024; C4's scores are
  0.13 and up) the relative bound cannot be what BASELINE.json means and the floor applies.

Returned: the number of excused positions (tests bound it) and ``max_rel_err_D``, the
largest relative distance error over the positions the relative bound applies to.
"""
from __future__ import annotations

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


class ParityError(AssertionError):
    pass


RUN_LOG = []  # one dict per successful check_parity call of this process (tests/conftest.py prints the totals)


def tie_tolerance(d: int) -> float:
    return 2.0 * np.sqrt(d) * 2.0 ** -24


def _scores64(q64, rows, metric):
    rows = rows.astype(np.float64)
    if metric == METRIC_INNER_PRODUCT:
        return rows @ q64
    return ((rows - q64) ** 2).sum(1)


def check_parity(D, I, D_ref, I_ref, xq, xb, metric, *, rtol=1e-5, max_excused_frac=None):
    """Raise ParityError unless (D, I) matches (D_ref, I_ref) under the rule above.

    ``xq``/``xb`` are the fp32 matrices actually searched (after any normalisation).
    ``D_ref``/``I_ref`` may carry more columns than ``D``/``I``: the reference's next-best
    candidates beyond k, used for the k-boundary rule.
    """
    D = np.asarray(D)
    I = np.asarray(I)
    D_ref = np.asarray(D_ref)
    I_ref = np.asarray(I_ref)
    if D.shape != I.shape or D_ref.shape != I_ref.shape or D.ndim != 2 or D_ref.ndim != 2:
        raise ParityError("shape mismatch %s %s vs %s %s" % (D.shape, I.shape, D_ref.shape, I_ref.shape))
    if D.shape[0] != D_ref.shape[0] or D_ref.shape[1] < D.shape[1]:
        raise ParityError("shape mismatch %s %s vs %s %s" % (D.shape, I.shape, D_ref.shape, I_ref.shape))
    if D.dtype != np.float32 or I.dtype != np.int64:
        raise ParityError("dtype mismatch: D %s I %s" % (D.dtype, I.dtype))
    nq, k = I.shape
    D_more, I_more = D_ref[:, k:], I_ref[:, k:]  # the reference's candidates beyond k (may be empty)
    D_ref, I_ref = D_ref[:, :k], I_ref[:, :k]
    d = xq.shape[1]
    tau_unit = tie_tolerance(d)
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
        scale = qn * bn_max
    else:
        scale = qn ** 2 + bn_max ** 2
    tau_q = (tau_unit * scale)[:, None]  # absolute tie tolerance per query
    ref64 = D_ref.astype(np.float64)
    err = np.abs(D.astype(np.float64) - ref64)
    floor = tau_q / 8.0  # fp32 accumulation noise floor (see the module docstring)
    relative = (rtol * np.abs(ref64) >= floor) & ~pad  # the stated 1e-5 relative bound applies
    near_zero = ~relative & ~pad
    bad = (relative & (err > rtol * np.abs(ref64))) | (near_zero & (err > floor))
    if bad.any():
        r, c = np.argwhere(bad)[0]
        raise ParityError(
            "distance mismatch at (%d,%d): %r vs ref %r (%s bound); %d bad of %d"
            % (r, c, D[r, c], D_ref[r, c], "relative 1e-5" if relative[r, c] else "near-zero tau/8", int(bad.sum()), bad.size)
        )
    max_rel = float((err[relative] / np.abs(ref64[relative])).max()) if relative.any() else 0.0

    # --- ids ---------------------------------------------------------------------------
    diff_rows = np.flatnonzero((I != I_ref).any(axis=1))
    boundary_swaps = 0
    for r in diff_rows:
        valid = ~pad[r]
        ours = I[r][valid]
        ref = I_ref[r][valid]
        if len(set(ours.tolist())) != ours.size:
            raise ParityError("row %d: duplicate ids" % r)
        if ours.min() < 0 or ours.max() >= xb.shape[0]:
            raise ParityError("row %d: id out of range" % r)
        q64 = xq[r].astype(np.float64)
        tau = float(tau_q[r, 0])
        s_ref = _scores64(q64, xb[ref], metric)
        s_ours = _scores64(q64, xb[ours], metric)
        # (1) position-wise: the two ids at a differing position are fp64-indistinguishable within tau
        pos = np.flatnonzero(ours != ref)
        gap = np.abs(s_ours[pos] - s_ref[pos])
        if gap.max() > tau:
            p = pos[gap.argmax()]
            raise ParityError(
                "row %d pos %d: id %d vs ref %d, fp64 scores differ by %g > tau %g"
                % (r, p, ours[p], ref[p], gap.max(), tau)
            )
        # (2) tau-clusters of the reference row: the id multiset inside a cluster must match
        n = ref.size
        cuts = np.flatnonzero(np.abs(np.diff(s_ref)) > tau) + 1
        starts = np.concatenate([[0], cuts])
        ends = np.concatenate([cuts, [n]])
        for a, b in zip(starts, ends):
            if a > pos.max() or b <= pos.min():
                continue
            mine, theirs = set(ours[a:b].tolist()), set(ref[a:b].tolist())
            if mine == theirs:
                continue
            if b != n or valid.sum() != k:
                raise ParityError("row %d: ids of the tau-cluster at positions [%d,%d) differ: %s vs ref %s"
                                  % (r, a, b, sorted(mine - theirs), sorted(theirs - mine)))
            # (3) the cluster at the k-boundary may trade members with the reference's next candidates inside tau
            foreign = np.asarray(sorted(mine - theirs), dtype=np.int64)
            s_f = _scores64(q64, xb[foreign], metric)
            more = I_more[r][I_more[r] >= 0] if I_more.shape[1] else np.empty(0, np.int64)
            s_more = _scores64(q64, xb[more], metric) if more.size else np.empty(0)
            # the reference's candidates beyond k that are chained to the boundary cluster within tau
            allowed, last, chained = set(), s_ref[n - 1], 0
            for idx, sc in zip(more.tolist(), s_more.tolist()):
                if abs(sc - last) > tau:
                    break
                allowed.add(idx)
                last = sc
                chained += 1
            for f, sf in zip(foreign.tolist(), s_f.tolist()):
                if f in allowed:
                    continue
                # not among the extra columns supplied: acceptable only if the chain ran through ALL of them (the
                # cluster continues beyond what the reference listed) and the id itself continues it
                if chained < more.size or abs(sf - last) > tau:
                    raise ParityError("row %d: id %d entered the top-%d but is not inside the tau-cluster at the "
                                      "reference's k-boundary (fp64 score %g, boundary %g, tau %g)" % (r, f, k, sf, last, tau))
            boundary_swaps += len(foreign)
        excused += pos.size
    frac = excused / max(1, I.size)
    if max_excused_frac is not None and frac > max_excused_frac:
        raise ParityError("too many tie-excused positions: %d of %d" % (excused, I.size))
    stats = {"excused": excused, "positions": int(I.size), "rows_with_ties": int(diff_rows.size),
             "boundary_swaps": boundary_swaps, "max_rel_err_D": max_rel}
    RUN_LOG.append(stats)
    return stats
