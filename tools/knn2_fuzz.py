"""Randomised parity sweep of the brute-force 2-NN (orbx_knn2, host buffers) against the CPU oracle: ragged and tiny sizes,
duplicated rows (ties -> lower index), queries equal to database rows.  usage: knn2_fuzz.py [n_cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wut_cuda_orb_slam3_b200 as orbx
from tests import oracle_lib


def run(n_cases, seed):
    rng = np.random.default_rng(seed)
    o = oracle_lib.load()
    matcher = orbx.ORBmatcher()
    bad = 0
    for case in range(n_cases):
        nq = int(rng.choice([1, 2, 7, 15, 16, 17, 31, 32, 33, 255, 256, 257, int(rng.integers(1, 900))]))
        ndb = int(rng.choice([1, 2, 7, 8, 9, 127, 128, 129, 255, 256, 257, int(rng.integers(1, 20000)), int(rng.integers(1, 300000))]))
        db = rng.integers(0, 256, (ndb, 32), dtype=np.uint8)
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        for _ in range(min(ndb // 3, 50)):                         # duplicated rows: ties
            db[int(rng.integers(0, ndb))] = db[int(rng.integers(0, ndb))]
        for i in range(0, nq, 3):                                  # near and exact copies of database rows
            q[i] = db[int(rng.integers(0, ndb))]
            if i % 2:
                q[i, int(rng.integers(0, 32))] ^= np.uint8(1 << int(rng.integers(0, 8)))
        idx, dist = matcher.knn2(q, db)
        ridx, rdist = o.knn2(q, db)
        if not (np.array_equal(idx, ridx) and np.array_equal(dist, rdist)):
            bad += 1
            print("MISMATCH case", case, dict(nq=nq, ndb=ndb), flush=True)
    print("knn2 fuzz: %d cases, %d mismatching (env %s)" % (n_cases, bad, {k: v for k, v in os.environ.items() if k.startswith("ORBX_")}))
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 1) else 0)
