"""AUC1 / TP evaluation restated for the tests (TEST INFRASTRUCTURE ONLY).

Follows seqvec_search/main.py:53-82 (``evaluate_faiss`` + ``evaluate``) and
seqvec_search/data.py:38-54 (``LoadedData.from_options``): the reference's two known-answer
tests (tests/test_main.py:10-27) assert on these numbers, which are functions of the
neighbour ids ``I`` only.
"""
from __future__ import annotations

import json
from collections import Counter
from pathlib import Path


class Fixture:
    def __init__(self, path: Path):
        path = Path(path)
        self.path = path
        self.train = path / "train.npy"
        self.test = path / "test.npy"
        self.train_ids = json.loads((path / "train.json").read_text())
        self.test_ids = json.loads((path / "test.json").read_text())
        self.ids_to_family = json.loads((path / "ids_to_family.json").read_text())


def evaluate_ids(fx: Fixture, results):
    """results: (nq, k) int array of database row numbers -> (auc1s, tps) lists."""
    family_sizes = dict(Counter(fx.ids_to_family[i] for i in fx.train_ids))
    auc1s, tps = [], []
    for key, row in enumerate(results):
        name = fx.test_ids[key]
        matches = [fx.train_ids[i] for i in row]
        correct = fx.ids_to_family[name]
        tp = sum(fx.ids_to_family[i] == correct for i in matches)
        auc1 = 0
        for i in matches:
            if fx.ids_to_family[i] == correct:
                auc1 += 1
            else:
                break
        auc1s.append(auc1 / family_sizes[correct])
        tps.append(tp / family_sizes[correct])
    return auc1s, tps
