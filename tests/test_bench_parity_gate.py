"""The gate behind bench.py's `parity_vs_fp64` record (CPU): it must pass the engine's legitimate answers - identical
lists, swaps of fp32 near-ties - and fail every real defect: a wrong id, a dropped neighbour, a duplicate, a wrong
global id offset (what a mis-sharded merge would produce), a distance off by more than 1e-5 relative."""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench  # noqa: E402


def _reference(ns=8, n=4000, d=64, k=20, extra=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    xb = torch.nn.functional.normalize(torch.randn(n, d, generator=g, dtype=torch.float64), dim=1)
    xq = torch.nn.functional.normalize(torch.randn(ns, d, generator=g, dtype=torch.float64), dim=1)
    s = xq @ xb.T
    best_s, best_i = torch.topk(s, k + extra, dim=1)
    return s, best_s, best_i, k, d


def test_gate_passes_the_exact_answer_and_fp32_rounding_of_distances():
    s, best_s, best_i, k, d = _reference()
    D, I = best_s[:, :k].float(), best_i[:, :k].clone()
    rec = bench.compare_with_fp64(D, I, best_s, best_i, k, d)
    assert rec["ok"] and rec["id_mismatches_beyond_tau"] == 0 and rec["excused"] == 0 and rec["max_rel_err_D"] < 1e-6


def test_gate_excuses_only_near_ties():
    s, best_s, best_i, k, d = _reference(seed=1)
    tau = 2.0 * d ** 0.5 * 2.0 ** -24
    # make positions 3 and 4 of query 0 a near-tie in the fp64 reference, then report them swapped
    best_s[0, 4] = best_s[0, 3] - tau / 4
    D, I = best_s[:, :k].float(), best_i[:, :k].clone()
    I[0, 3], I[0, 4] = best_i[0, 4], best_i[0, 3]
    D[0, 3], D[0, 4] = best_s[0, 4].float(), best_s[0, 3].float()
    rec = bench.compare_with_fp64(D, I, best_s, best_i, k, d)
    assert rec["ok"] and rec["excused"] == 2
    # the same swap between scores that are NOT a near-tie is a mismatch
    best_s[0, 4] = best_s[0, 3] - 100 * tau
    D[0, 3], D[0, 4] = best_s[0, 4].float(), best_s[0, 3].float()
    rec = bench.compare_with_fp64(D, I, best_s, best_i, k, d)
    assert not rec["ok"] and rec["id_mismatches_beyond_tau"] == 2


def test_gate_fails_real_defects():
    s, best_s, best_i, k, d = _reference(seed=2)
    D, I = best_s[:, :k].float(), best_i[:, :k].clone()
    bad = I.clone()
    bad[2, 5] = (bad[2, 5] + 1234) % 4000              # a wrong neighbour
    assert not bench.compare_with_fp64(D, bad, best_s, best_i, k, d)["ok"]
    shifted = I + 7                                      # a wrong global id base (mis-sharded merge)
    assert not bench.compare_with_fp64(D, shifted, best_s, best_i, k, d)["ok"]
    dropped = torch.cat([I[:, 1:], best_i[:, k:k + 1]], 1)  # the best hit lost, everything moved up
    Dd = torch.cat([D[:, 1:], best_s[:, k:k + 1].float()], 1)
    rec = bench.compare_with_fp64(Dd, dropped, best_s, best_i, k, d)
    assert not rec["ok"] and rec["id_mismatches_beyond_tau"] > 0
    dup = I.clone()
    dup[1, 7] = dup[1, 6]
    rec = bench.compare_with_fp64(D, dup, best_s, best_i, k, d)
    assert not rec["ok"] and rec["rows_with_duplicate_ids"] == 1
    off = D.clone()
    off[3, 0] *= 1 + 3e-5                                # a distance off by more than 1e-5 relative
    rec = bench.compare_with_fp64(off, I, best_s, best_i, k, d)
    assert not rec["ok"] and rec["max_rel_err_D"] > 1e-5 and rec["id_mismatches_beyond_tau"] == 0
