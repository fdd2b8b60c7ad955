"""The post-processing oracle (oracle/postproc_oracle.py) against vectors produced by the
UNMODIFIED reference functions (tests/golden/make_golden_postproc.py) - CPU only."""
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from oracle import postproc_oracle as po  # noqa: E402
from oracle.evaluate import Fixture  # noqa: E402

GOLD = REPO / "tests" / "golden"


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD / "postproc" / "postproc.npz")


def family_codes(fx):
    fams = sorted(set(fx.ids_to_family.values()))
    code = {f: i for i, f in enumerate(fams)}
    q = np.asarray([code[fx.ids_to_family[i]] for i in fx.test_ids], np.int32)
    d = np.asarray([code[fx.ids_to_family[i]] for i in fx.train_ids], np.int32)
    return q, d


@pytest.mark.parametrize("name", ["small-random", "pfam-20-10", "pfam-20-dist"])
@pytest.mark.parametrize("tag", ["golden", "random"])
def test_evaluate_matches_reference(g, name, tag):
    fx = Fixture(GOLD / name)
    qf, df = family_codes(fx)
    auc1s, tps = po.evaluate(g[f"evaluate.{name}.{tag}.I"], qf, df)
    assert auc1s == g[f"evaluate.{name}.{tag}.auc1"].tolist()  # bit-identical float64
    assert tps == g[f"evaluate.{name}.{tag}.tp"].tolist()


def test_evaluate_known_answers_of_the_reference_test():
    # tests/test_main.py:17-18 of the reference
    fx = Fixture(GOLD / "small-random")
    qf, df = family_codes(fx)
    I = np.load(GOLD / "expected.npz")["small-random.ip.k5.I"]
    auc1s, tps = po.evaluate(I, qf, df)
    assert auc1s == [1.0, 1 / 3, 2 / 3, 0.0, 0.0, 1 / 3]
    assert tps == [1.0, 2 / 3, 2 / 3, 1.0, 1.0, 1.0]


def test_compute_is_correct(g):
    out = po.compute_is_correct(g["is_correct.results"], g["is_correct.mapping"])
    assert out.shape == g["is_correct.out"].shape and np.array_equal(out, g["is_correct.out"])


def csr_lists(offsets, members):
    return [members[offsets[i]:offsets[i + 1]].tolist() for i in range(len(offsets) - 1)]


def test_compute_correctness_array(g):
    hom = csr_lists(g["correctness.offsets"], g["correctness.members"])
    out = po.compute_correctness_array(g["correctness.full"], hom)
    assert np.array_equal(out, g["correctness.out"])


def test_compute_auc1(g):
    hom = csr_lists(g["correctness.offsets"], g["correctness.members"])
    out = po.compute_auc1(g["auc1.hits"], hom, g["auc1.set_sizes"], n_db=len(hom))
    assert np.array_equal(out, g["auc1.out"])


def test_remove_self_hit(g):
    h, s = g["selfhit.hits_in"].copy(), g["selfhit.scores_in"].copy()
    ho, so, bogus = po.remove_self_hit(h, s, np.arange(h.shape[0]))
    assert bogus == int(g["selfhit.bogus"]) and bogus > 0
    assert np.array_equal(h, g["selfhit.hits_inplace"]) and np.array_equal(s, g["selfhit.scores_inplace"])
    assert np.array_equal(ho, g["selfhit.hits_out"]) and np.array_equal(so, g["selfhit.scores_out"])


@pytest.mark.parametrize("tag", ["pfam-20-10", "edge"])
def test_write_prefilter_db(g, tag):
    data, index = po.write_prefilter_db(g[f"prefilter.{tag}.hits"], g[f"prefilter.{tag}.queries"], g[f"prefilter.{tag}.scores"],
                                        g[f"prefilter.{tag}.test_map"], g[f"prefilter.{tag}.train_map"])
    assert data == g[f"prefilter.{tag}.data"].tobytes()
    assert index == g[f"prefilter.{tag}.index"].tobytes()
