"""Randomised GPU-vs-oracle parity sweep (tools/fuzz_parity.py): random shapes, feature counts, levels, scale factors, thresholds,
single frames and small host batches.  The committed seeds are a smoke-sized subset; run the tool with other seeds for more."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_configurations():
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tools", "fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(10, 12345) == 0
