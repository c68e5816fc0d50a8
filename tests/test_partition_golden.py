"""Bit-exact host topology: cgl_gan_b200.partition against the reference's own
allocate_dataset / init_groups (lifted verbatim and run by tests/golden/make_golden.py). CPU only."""
import json
import os
from random import Random

import numpy as np
import pytest

from cgl_gan_b200 import partition as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARTS = json.load(open(os.path.join(GOLD, "partitions.json")))


@pytest.mark.parametrize("style", ["cgl", "fl", "fl2d", "cap"])
def test_allocate_dataset_bit_exact(style):
    labels = np.array(PARTS["labels"])
    rd = Random(); rd.seed(20211212)
    for iid in (0, 1, 2):       # one continuing rd stream across the iid loop, like the reference's __main__
        test_idx, parts = P.allocate_dataset(labels, iid, 10, 10, 100, rd, style=style)
        gold = PARTS[style][str(iid)]
        assert parts == gold["parts"], (style, iid)
        if gold["test"] is not None:
            assert test_idx[:len(gold["test"])] == gold["test"]
        else:
            assert test_idx is None


def test_gmm_default_partition_sizes():
    """Known-answer values of the repo-default CGLGAN 2DMG run."""
    labels = np.load(os.path.join(GOLD, "gmm_labels.npy")).astype(np.int64)
    g = PARTS["gmm_default"]
    assert [int((labels == c).sum()) for c in range(10)] == g["class_counts"] == \
        [10106, 10028, 9807, 9961, 9935, 10065, 9991, 10069, 9963, 10075]
    rd = Random(); rd.seed(20211212)
    for iid in (0, 1, 2):
        _, parts = P.allocate_dataset(labels, iid, 10, 10, 10000, rd, style="cgl")
        assert [len(p) for p in parts] == g["sizes"][str(iid)]
        for p in parts:                      # CGLGAN/2DMG/main.py:461
            rd.sample(range(len(p)), 100)
    assert g["sizes"]["1"] == [8000, 29941, 4000, 3000, 2000, 7000, 30125, 9000, 1000, 1000]
    assert g["sizes"]["2"][-1] == 10074      # the tensor-form scan never hands out the last sample


def test_gmm_class_draws_match_reference():
    """The vectorised 2-D mixture draws the same class sequence as the reference's per-point loop."""
    labels = np.load(os.path.join(GOLD, "gmm_labels.npy")).astype(np.int64)
    data, targets = P.gmm_labels_and_data(10, 10000)
    assert np.array_equal(targets.numpy().astype(np.int64), labels)
    assert data.shape == (100000, 2) and abs(float(data.norm(dim=1).mean()) - 1.0) < 1e-3


@pytest.mark.parametrize("frac,tag", [(0.2, "frac02"), (1, "frac1")])
def test_init_groups_bit_exact(frac, tag):
    gold = PARTS["init_groups_" + tag]
    xs = [np.eye(10, dtype=np.int64)[i] * 100 for i in range(10)]
    groups, _ = P.init_groups(10, xs, frac, max_groups=40)
    assert groups == gold["onehot"]
    xs2 = [np.array(x) for x in gold["random_freq"]]
    groups, choose_r = P.init_groups(10, xs2, frac, max_groups=40)
    assert groups == gold["random"]
    assert choose_r == [0 in g for g in groups]


def test_assignment_blocks():
    cl, sl = P.assign_clients(10, 5)
    assert cl == [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9]] and sl[3] == [1]
    cl, sl = P.assign_clients(10, 3)      # 10 // 3 = 3 per server, worker 9 is served by nobody
    assert cl == [[0, 1, 2], [3, 4, 5], [6, 7, 8]] and sl[9] == []
