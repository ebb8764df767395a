"""Randomised GPU-vs-oracle parity sweep (tools/fuzz_parity.py): random shapes, feature counts, levels, scale factors, thresholds,
single frames and small host batches, three kinds of lapping area.  Three seeds x 20 configurations run in the suite; the tool
takes any other seed (profiles/*_fuzz.txt hold the builder's larger sweeps)."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [12345, 777, 20261018])
def test_random_configurations(seed):
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tools", "fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(20, seed) == 0


@pytest.mark.parametrize("seed", [5, 4711])
def test_random_knn2_sizes(seed):
    """tools/knn2_fuzz.py: the brute-force 2-NN (tensor-core matcher by default) against the oracle on random, ragged and tiny
    (nq, ndb) pairs with duplicated rows (ties -> lower index) and exact / near copies of database rows as queries."""
    spec = importlib.util.spec_from_file_location("knn2_fuzz", os.path.join(ROOT, "tools", "knn2_fuzz.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(40, seed) == 0
