"""Committed fixtures (tests/golden/): the reference's own known answers (reference_goldens.json, transcribed from its ScalaTest
specs) and the oracle's frozen outputs for a set of small cases covering every BASELINE configuration (oracle_fixtures.npz, made by
tests/golden/make_fixtures.py).  CPU: the oracle reproduces both.  GPU: the CUDA path, through the C ABI, meets the same bits / bars."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, HERE)
from cases import CASES  # noqa: E402

from oracle import reference as ref  # noqa: E402

FIX = np.load(os.path.join(HERE, "oracle_fixtures.npz"))
GOLDENS = json.load(open(os.path.join(HERE, "reference_goldens.json")))["goldens"]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def check(name, got_shape, got, bar):
    assert tuple(got_shape) == tuple(FIX[name + "/shape"].tolist()), name
    want = FIX[name + "/strict"]
    if bar == "exact":
        assert np.array_equal(bits(got), bits(want)), name
    elif bar == "ulp2":
        d = np.minimum(ref.ulp_distance(got, want), ref.ulp_distance(got, FIX[name + "/contracted"]))
        assert ((d <= 2) | (np.abs(got - want) <= 2e-7)).all(), (name, int(d.max()))
    else:
        scale = float(np.abs(want).max()) or 1.0
        assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= 1e-5 * scale, name


def test_every_case_has_a_fixture_and_every_golden_a_case():
    assert {k.split("/")[0] for k in FIX.files} == set(CASES)
    for g in GOLDENS:
        assert g["case"] in CASES and g["source"].count(":") == 1


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_its_fixture(name):
    build, bar = CASES[name]
    t = build(ref.Tensor)
    got = t.flat_array()
    assert np.array_equal(bits(got), bits(FIX[name + "/strict"])), name  # the oracle itself must not drift: bit for bit
    assert tuple(t.shape) == tuple(FIX[name + "/shape"].tolist())


@pytest.mark.parametrize("g", GOLDENS, ids=[g["case"] for g in GOLDENS])
def test_oracle_meets_the_references_known_answers(g):
    t = CASES[g["case"]][0](ref.Tensor)
    if "toString" in g:
        assert t.to_string() == g["toString"], g["source"]
    else:
        got, want = t.flat_array(), np.array(g["values"], np.float32)
        assert ((ref.ulp_distance(got, want) <= 2) | (np.abs(got - want) <= 2e-7)).all(), g["source"]


@pytest.fixture(scope="module")
def cuda():
    from compute.scala_b200 import cuda as c

    c.init()
    yield c
    c.synchronize()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_meets_the_fixture(cuda, name):
    build, bar = CASES[name]
    t = build(cuda.Tensor)
    check(name, t.shape, t.flatArray(), bar)


@pytest.mark.gpu
@pytest.mark.parametrize("g", GOLDENS, ids=[g["case"] for g in GOLDENS])
def test_cuda_meets_the_references_known_answers(cuda, g):
    t = CASES[g["case"]][0](cuda.Tensor)
    if "toString" in g:
        assert t.toString() == g["toString"], g["source"]
    else:
        got, want = t.flatArray(), np.array(g["values"], np.float32)
        assert ((ref.ulp_distance(got, want) <= 2) | (np.abs(got - want) <= 2e-7)).all(), g["source"]
