"""CPU tests of the bag-of-words oracle: the C++ restatement (oracle/orb_oracle.cpp) against independent numpy / pure-Python
restatements of the DBoW2 rules on small cases, and the committed fixture tests/golden/bow_golden.json."""
import json
import os
import zlib

import numpy as np
import pytest

from tests import bow_synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def ham(a, b):
    return int(POP[np.bitwise_xor(a, b)].sum())


def py_transform(parent, desc, weights, L, feature, levelsup):
    """TemplatedVocabulary::transform (one feature) in plain Python."""
    n = len(parent)
    children = [[] for _ in range(n)]
    for i in range(1, n):
        children[parent[i]].append(i)
    words = {}
    for i in range(1, n):
        if not children[i]:
            words[i] = len(words)
    nid_level = L - levelsup
    nid = 0
    node, level = 0, 0
    while True:
        level += 1
        kids = children[node]
        node = kids[0]
        best = ham(feature, desc[node])
        for c in kids[1:]:
            d = ham(feature, desc[c])
            if d < best:
                best, node = d, c
        if level == nid_level:
            nid = node
        if not children[node]:
            break
    return words[node], weights[node], nid


@pytest.mark.parametrize("k,L,levelsup", [(4, 3, 1), (3, 4, 2), (10, 3, 4), (5, 2, 0)])
def test_transform_matches_python(oracle, k, L, levelsup):
    parent, desc, weights = bow_synth.make_vocab(7 + k, k, L)
    feats = bow_synth.make_features(11, desc, parent, 150)
    voc = oracle.vocabulary(parent, desc, weights, k, L)
    word, w, node = voc.transform_features(feats, levelsup)
    for i in range(len(feats)):
        assert (int(word[i]), float(w[i]), int(node[i])) == py_transform(parent, desc, weights, L, feats[i], levelsup), i


@pytest.mark.parametrize("weighting,scoring", [(0, 0), (1, 0), (2, 0), (3, 0), (0, 1), (0, 5)])
def test_compute_bow_matches_python(oracle, weighting, scoring):
    k, L, levelsup = 4, 3, 2
    parent, desc, weights = bow_synth.make_vocab(21, k, L, zero_weight_frac=0.2)
    feats = bow_synth.make_features(22, desc, parent, 300)
    voc = oracle.vocabulary(parent, desc, weights, k, L, scoring, weighting)
    (ids, vals), (nodes, offs, idx) = voc.transform(feats, levelsup)
    bow, fv = {}, {}
    for i, f in enumerate(feats):
        word, w, nid = py_transform(parent, desc, weights, L, f, levelsup)
        if w > 0:
            if weighting in (0, 1):
                bow[word] = bow.get(word, 0.0) + w if word in bow else w
            else:
                bow.setdefault(word, w)
            fv.setdefault(nid, []).append(i)
    keys = sorted(bow)
    must = scoring != 5
    if weighting in (0, 1) and bow and not must:
        nd = float(len(bow))
        for kk in keys:
            bow[kk] /= nd
    if must:
        norm = 0.0
        if scoring != 1:
            for kk in keys:
                norm += abs(bow[kk])
        else:
            for kk in keys:
                norm += bow[kk] * bow[kk]
            norm = norm ** 0.5
        if norm > 0:
            for kk in keys:
                bow[kk] /= norm
    assert ids.tolist() == keys
    assert vals.tolist() == [bow[kk] for kk in keys]            # bit-identical doubles
    assert nodes.tolist() == sorted(fv)
    for j, nd_ in enumerate(nodes):
        assert idx[offs[j]:offs[j + 1]].tolist() == fv[int(nd_)]


def py_search_kf_frame(P, fva, fvb, ratio, check_ori, nleft=-1):
    """SearchByBoW(KF, Frame) in plain Python (mono and two-fisheye variants)."""
    da, db = P["desc_a"], P["desc_b"]
    match = [-1] * len(db)
    hist = [[] for _ in range(30)]
    nm = 0
    common = sorted(set(fva) & set(fvb))
    for node in common:
        for ia in fva[node]:
            if not P["valid_a"][ia]:
                continue
            b1 = b2 = b1r = b2r = 256
            j1 = j1r = -1
            for j in fvb[node]:
                if match[j] >= 0:
                    continue
                d = ham(da[ia], db[j])
                if nleft == -1 or j < nleft:
                    if d < b1:
                        b2, b1, j1 = b1, d, j
                    elif d < b2:
                        b2 = d
                else:
                    if d < b1r:
                        b2r, b1r, j1r = b1r, d, j
                    elif d < b2r:
                        b2r = d
            if b1 <= 50:
                acc = []
                if np.float32(b1) < np.float32(ratio) * np.float32(b2):
                    acc.append(j1)
                if nleft != -1 and b1r <= 50:
                    acc.append(j1r)
                for j in acc:
                    match[j] = ia
                    rot = np.float32(P["angle_a"][ia]) - np.float32(P["angle_b"][j])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360))
                    b = int(np.floor(float(np.float32(rot * np.float32(1.0 / 30))) + 0.5))
                    hist[0 if b == 30 else b].append(j)
                    nm += 1
    if check_ori:
        sizes = [len(h) for h in hist]
        m1 = m2 = m3 = 0; i1 = i2 = i3 = -1
        for i, s in enumerate(sizes):
            if s > m1:
                m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
            elif s > m2:
                m3, m2, i3, i2 = m2, s, i2, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < np.float32(0.1) * np.float32(m1):
            i2 = i3 = -1
        elif m3 < np.float32(0.1) * np.float32(m1):
            i3 = -1
        for i in range(30):
            if i in (i1, i2, i3):
                continue
            for j in hist[i]:
                match[j] = -1
                nm -= 1
    return match, nm


def fv_dict(fv):
    nodes, offs, idx = fv
    return {int(n): idx[offs[i]:offs[i + 1]].tolist() for i, n in enumerate(nodes)}


@pytest.mark.parametrize("nleft,check_ori,ratio", [(-1, True, 0.6), (-1, False, 0.9), (120, True, 0.75)])
def test_search_by_bow_kf_frame_matches_python(oracle, nleft, check_ori, ratio):
    k, L = 4, 3
    parent, desc, weights = bow_synth.make_vocab(31, k, L)
    voc = oracle.vocabulary(parent, desc, weights, k, L)
    P = bow_synth.make_pair(32, desc, parent, 260, 240)
    _, fva = voc.transform(P["desc_a"], 2)
    _, fvb = voc.transform(P["desc_b"], 2)
    got, nm = oracle.search_by_bow_kf_frame(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], fvb, nleft, ratio,
                                            check_ori)
    exp, nm_exp = py_search_kf_frame(P, fv_dict(fva), fv_dict(fvb), ratio, check_ori, nleft)
    assert got.tolist() == exp and nm == nm_exp
    assert nm > 20                      # the planted correspondences are found


def test_l1_score(oracle):
    k, L = 4, 3
    parent, desc, weights = bow_synth.make_vocab(41, k, L)
    voc = oracle.vocabulary(parent, desc, weights, k, L)
    P = bow_synth.make_pair(42, desc, parent, 200, 200)
    a, _ = voc.transform(P["desc_a"], 2)
    b, _ = voc.transform(P["desc_b"], 2)
    s = oracle.bow_score_l1(a, b)
    da, db_ = dict(zip(a[0].tolist(), a[1].tolist())), dict(zip(b[0].tolist(), b[1].tolist()))
    acc = 0.0
    for w in sorted(set(da) & set(db_)):
        acc += abs(da[w] - db_[w]) - abs(da[w]) - abs(db_[w])
    assert s == -acc / 2.0 and 0.0 <= s <= 1.0
    assert oracle.bow_score_l1(a, a) == pytest.approx(1.0, abs=1e-12)


GOLD_PATH = os.path.join(ROOT, "tests", "golden", "bow_golden.json")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def run_golden_case(c, make_voc, kf_frame, kf_kf, tri):
    """Shared by the CPU (oracle) and GPU (C ABI) golden tests: recompute every recorded quantity of case c."""
    parent, desc, weights = bow_synth.make_vocab(c["seed"], c["k"], c["L"])
    voc = make_voc(parent, desc, weights, c["k"], c["L"])
    P = bow_synth.make_pair(c["seed"] + 1, desc, parent, c["n_a"], c["n_b"])
    out = {}
    word, w, node = voc.transform_features(P["desc_a"], c["levelsup"])
    out["word_crc"], out["weight_crc"], out["node_crc"] = crc(word), crc(w), crc(node)
    (ids, vals), fva = voc.transform(P["desc_a"], c["levelsup"])
    _, fvb = voc.transform(P["desc_b"], c["levelsup"])
    out["bow_ids_crc"], out["bow_vals_crc"] = crc(ids), crc(vals)
    out["fv_crc"] = crc(np.concatenate([np.asarray(x, np.int64).ravel() for x in fva]))
    m, nm = kf_frame(P, fva, fvb)
    out["kf_frame"] = [crc(m), int(nm)]
    m, nm = kf_kf(P, fva, fvb)
    out["kf_kf"] = [crc(m), int(nm)]
    m, nm = tri(P, fva, fvb, c["seed"] + 2)
    out["tri"] = [crc(m), int(nm)]
    return out


def tri_inputs(P, seed):
    rng = np.random.default_rng(seed)
    kpa = bow_synth.make_keypoints(seed + 1, len(P["desc_a"]), P["angle_a"])
    kpb = bow_synth.make_keypoints(seed + 2, len(P["desc_b"]), P["angle_b"])
    F, ep = bow_synth.fundamental(seed + 3)
    scale = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    sigma2 = (scale * scale * np.float32(40.0)).astype(np.float32)       # wide gate so that a good share passes
    free_a = 1 - P["valid_a"]; free_b = 1 - P["valid_b"]
    free_a = (rng.random(len(free_a)) < 0.7).astype(np.uint8); free_b = (rng.random(len(free_b)) < 0.7).astype(np.uint8)
    st_a = (rng.random(len(free_a)) < 0.4).astype(np.uint8); st_b = (rng.random(len(free_b)) < 0.4).astype(np.uint8)
    return kpa, kpb, F, ep, scale, sigma2, free_a, free_b, st_a, st_b


def test_golden_oracle(oracle):
    gold = json.load(open(GOLD_PATH))
    for c in gold["cases"]:
        def kf_frame(P, fva, fvb):
            return oracle.search_by_bow_kf_frame(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], fvb, -1, 0.6, True)

        def kf_kf(P, fva, fvb):
            return oracle.search_by_bow_kf_kf(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], P["valid_b"], fvb, 0.8, True)

        def tri(P, fva, fvb, seed):
            kpa, kpb, F, ep, scale, sigma2, fa, fb, sa, sb = tri_inputs(P, seed)
            return oracle.search_for_triangulation(kpa, P["desc_a"], fa, sa, fva, kpb, P["desc_b"], fb, sb, fvb, F, ep, scale, sigma2)

        got = run_golden_case(c, oracle.vocabulary, kf_frame, kf_kf, tri)
        assert got == c["expect"], c["name"]
