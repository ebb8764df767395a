"""Regenerates tests/golden/bow_golden.json from the CPU oracle (run from the repo root: python tests/golden/make_bow_golden.py).
The fixture freezes the oracle's answers for the bag-of-words transform and the vocabulary-guided searches so that neither
the oracle nor the CUDA path can drift unnoticed; the reference ships no fixtures for this path."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_lib, test_bow_cpu  # noqa: E402

CASES = [dict(name="k4_L3", seed=501, k=4, L=3, levelsup=2, n_a=300, n_b=280),
         dict(name="k10_L4", seed=502, k=10, L=4, levelsup=3, n_a=1200, n_b=1100),
         dict(name="k6_L3_root", seed=503, k=6, L=3, levelsup=4, n_a=200, n_b=220)]

if __name__ == "__main__":
    o = oracle_lib.load()
    for c in CASES:
        def kf_frame(P, fva, fvb):
            return o.search_by_bow_kf_frame(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], fvb, -1, 0.6, True)

        def kf_kf(P, fva, fvb):
            return o.search_by_bow_kf_kf(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], P["valid_b"], fvb, 0.8, True)

        def tri(P, fva, fvb, seed):
            kpa, kpb, F, ep, scale, sigma2, fa, fb, sa, sb = test_bow_cpu.tri_inputs(P, seed)
            return o.search_for_triangulation(kpa, P["desc_a"], fa, sa, fva, kpb, P["desc_b"], fb, sb, fvb, F, ep, scale, sigma2)

        c["expect"] = test_bow_cpu.run_golden_case(c, o.vocabulary, kf_frame, kf_kf, tri)
        print(c["name"], c["expect"])
    json.dump(dict(cases=CASES), open(os.path.join(ROOT, "tests", "golden", "bow_golden.json"), "w"), indent=1)
