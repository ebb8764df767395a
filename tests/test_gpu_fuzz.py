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
