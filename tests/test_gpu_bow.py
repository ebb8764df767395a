"""GPU parity tests of the bag-of-words transform and the vocabulary-guided searches:
the CUDA path through the C ABI against the CPU oracle, bit-exact (integer distances, match indices, identical doubles)."""
import json
import os

import numpy as np
import pytest

import wut_cuda_orb_slam3_b200 as orbx
from tests import bow_synth, test_bow_cpu

pytestmark = pytest.mark.gpu


def fv_tuple(fv):
    return (fv.node_ids, fv.offsets, fv.indices)


@pytest.fixture(scope="module")
def vocab_k10_L4():
    parent, desc, weights = bow_synth.make_vocab(601, 10, 4)
    return parent, desc, weights


@pytest.mark.parametrize("k,L,levelsup,n", [(10, 4, 4, 3000), (10, 4, 2, 1000), (4, 3, 1, 257), (20, 2, 1, 500), (3, 7, 4, 800)])
def test_transform_matches_oracle(oracle, k, L, levelsup, n):
    parent, desc, weights = bow_synth.make_vocab(700 + k + L, k, L)
    feats = bow_synth.make_features(5, desc, parent, n)
    voc = orbx.ORBVocabulary(parent, desc, weights, k, L)
    ovoc = oracle.vocabulary(parent, desc, weights, k, L)
    assert voc.info()["n_nodes"] == len(parent) and voc.info()["n_words"] == k ** L
    w, wt, nd = voc.transform_features(feats, levelsup)
    ow, owt, ond = ovoc.transform_features(feats, levelsup)
    assert np.array_equal(w, ow) and np.array_equal(wt, owt) and np.array_equal(nd, ond)
    (ids, vals), fv = voc.transform(feats, levelsup)
    (oids, ovals), ofv = ovoc.transform(feats, levelsup)
    assert np.array_equal(ids, oids) and np.array_equal(vals, ovals)        # identical doubles
    for a, b in zip(fv_tuple(fv), ofv):
        assert np.array_equal(a, b)


def test_orbvoc_shape_k10_L6(oracle):
    """The ORBvoc.txt shape: k = 10, L = 6 (1,111,111 nodes), levelsup = 4 as Frame::ComputeBoW uses."""
    parent, desc, weights = bow_synth.make_vocab(611, 10, 6, dup_frac=0.01)
    feats = bow_synth.make_features(6, desc, parent, 2000, noise_bits=30)
    voc = orbx.ORBVocabulary(parent, desc, weights, 10, 6)
    ovoc = oracle.vocabulary(parent, desc, weights, 10, 6)
    (ids, vals), fv = voc.transform(feats, 4)
    (oids, ovals), ofv = ovoc.transform(feats, 4)
    assert np.array_equal(ids, oids) and np.array_equal(vals, ovals)
    for a, b in zip(fv_tuple(fv), ofv):
        assert np.array_equal(a, b)
    assert 20 <= len(fv.node_ids) <= 100            # nodes two levels below the root
    s = voc.score((ids, vals), (ids, vals))
    assert s == oracle.bow_score_l1((ids, vals), (ids, vals))


@pytest.mark.parametrize("weighting,scoring", [(0, 0), (1, 0), (2, 0), (3, 0), (0, 1), (1, 5)])
def test_weighting_and_normalisation(oracle, weighting, scoring):
    parent, desc, weights = bow_synth.make_vocab(621, 5, 3, zero_weight_frac=0.2)
    feats = bow_synth.make_features(7, desc, parent, 700)
    voc = orbx.ORBVocabulary(parent, desc, weights, 5, 3, scoring, weighting)
    ovoc = oracle.vocabulary(parent, desc, weights, 5, 3, scoring, weighting)
    (ids, vals), fv = voc.transform(feats, 2)
    (oids, ovals), ofv = ovoc.transform(feats, 2)
    assert np.array_equal(ids, oids) and np.array_equal(vals, ovals)
    for a, b in zip(fv_tuple(fv), ofv):
        assert np.array_equal(a, b)


def test_text_file_loader(oracle, tmp_path):
    parent, desc, weights = bow_synth.make_vocab(631, 4, 3)
    path = tmp_path / "voc.txt"
    bow_synth.write_vocab_text(path, parent, desc, weights, 4, 3)
    voc = orbx.ORBVocabulary.loadFromTextFile(path)
    ref = orbx.ORBVocabulary(parent, desc, weights, 4, 3)
    assert voc.info() == ref.info()
    feats = bow_synth.make_features(8, desc, parent, 400)
    a = voc.transform(feats, 2); b = ref.transform(feats, 2)
    assert np.array_equal(a[0][0], b[0][0]) and np.array_equal(a[0][1], b[0][1])
    with pytest.raises(orbx.OrbxError):
        orbx.ORBVocabulary.loadFromTextFile(tmp_path / "missing.txt")


@pytest.mark.parametrize("n_a,n_b,levelsup,nleft,ratio,check_ori", [
    (2000, 2000, 2, -1, 0.6, True), (2000, 1900, 2, -1, 0.9, True), (1000, 1200, 3, -1, 0.75, False),
    (1500, 1500, 2, 800, 0.7, True), (600, 500, 4, -1, 0.9, True), (300, 40, 1, -1, 0.9, True)])
def test_search_by_bow_kf_frame(oracle, vocab_k10_L4, n_a, n_b, levelsup, nleft, ratio, check_ori):
    parent, desc, weights = vocab_k10_L4
    voc = orbx.ORBVocabulary(parent, desc, weights, 10, 4)
    P = bow_synth.make_pair(800 + n_a + levelsup, desc, parent, n_a, n_b)
    _, fva = voc.transform(P["desc_a"], levelsup)
    _, fvb = voc.transform(P["desc_b"], levelsup)
    mb, ma, nm = orbx.search_by_bow(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], fvb, nleft_b=nleft,
                                    nnratio=ratio, check_orientation=check_ori)
    omb, onm = oracle.search_by_bow_kf_frame(P["desc_a"], P["angle_a"], P["valid_a"], fv_tuple(fva), P["desc_b"], P["angle_b"],
                                             fv_tuple(fvb), nleft, ratio, check_ori)
    assert np.array_equal(mb, omb) and nm == onm
    assert nm == int((mb >= 0).sum())
    if nleft < 0:
        sel = ma >= 0
        assert np.array_equal(mb[ma[sel]], np.flatnonzero(sel))          # the inverse map is consistent
    if n_a >= 1000:
        assert nm > 50


@pytest.mark.parametrize("n_a,n_b,levelsup,ratio,check_ori", [(2000, 2000, 2, 0.6, True), (1800, 2000, 2, 0.8, True),
                                                              (900, 1000, 3, 0.9, False), (500, 500, 4, 0.8, True)])
def test_search_by_bow_kf_kf(oracle, vocab_k10_L4, n_a, n_b, levelsup, ratio, check_ori):
    parent, desc, weights = vocab_k10_L4
    voc = orbx.ORBVocabulary(parent, desc, weights, 10, 4)
    P = bow_synth.make_pair(900 + n_a + levelsup, desc, parent, n_a, n_b)
    _, fva = voc.transform(P["desc_a"], levelsup)
    _, fvb = voc.transform(P["desc_b"], levelsup)
    mb, ma, nm = orbx.search_by_bow(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], fvb, valid_b=P["valid_b"],
                                    kf_kf=True, nnratio=ratio, check_orientation=check_ori)
    oma, onm = oracle.search_by_bow_kf_kf(P["desc_a"], P["angle_a"], P["valid_a"], fv_tuple(fva), P["desc_b"], P["angle_b"],
                                          P["valid_b"], fv_tuple(fvb), ratio, check_ori)
    assert np.array_equal(ma, oma) and nm == onm
    assert nm == int((ma >= 0).sum()) and nm > 30
    sel = ma >= 0
    assert len(set(ma[sel].tolist())) == int(sel.sum())                 # vbMatched2: a KF2 feature is used at most once
    assert P["valid_b"][ma[sel]].all() and P["valid_a"][sel].all()


@pytest.mark.parametrize("only_stereo,coarse,check_ori", [(False, False, True), (True, False, True), (False, True, True),
                                                          (False, False, False)])
def test_search_for_triangulation(oracle, vocab_k10_L4, only_stereo, coarse, check_ori):
    parent, desc, weights = vocab_k10_L4
    voc = orbx.ORBVocabulary(parent, desc, weights, 10, 4)
    P = bow_synth.make_pair(1000 + only_stereo + 2 * coarse, desc, parent, 2000, 2000)
    _, fva = voc.transform(P["desc_a"], 2)
    _, fvb = voc.transform(P["desc_b"], 2)
    kpa, kpb, F, ep, scale, sigma2, fa, fb, sa, sb = test_bow_cpu.tri_inputs(P, 77)
    ma, nm = orbx.search_for_triangulation(kpa, P["desc_a"], fa, sa, fva, kpb, P["desc_b"], fb, sb, fvb, F, ep, scale, sigma2,
                                           only_stereo=only_stereo, coarse=coarse, check_orientation=check_ori)
    oma, onm = oracle.search_for_triangulation(kpa, P["desc_a"], fa, sa, fv_tuple(fva), kpb, P["desc_b"], fb, sb, fv_tuple(fvb), F, ep,
                                               scale, sigma2, only_stereo, coarse, check_ori)
    assert np.array_equal(ma, oma) and nm == onm
    assert nm == int((ma >= 0).sum()) and nm > (5 if not coarse else 100)
    sel = ma >= 0
    assert fa[sel].all() and fb[ma[sel]].all()
    if only_stereo:
        assert sa[sel].all() and sb[ma[sel]].all()


def test_edge_cases(oracle, vocab_k10_L4):
    parent, desc, weights = vocab_k10_L4
    voc = orbx.ORBVocabulary(parent, desc, weights, 10, 4)
    P = bow_synth.make_pair(1100, desc, parent, 300, 300)
    _, fva = voc.transform(P["desc_a"], 2)
    _, fvb = voc.transform(P["desc_b"], 2)
    # nothing valid -> no matches
    mb, ma, nm = orbx.search_by_bow(P["desc_a"], P["angle_a"], np.zeros(300, np.uint8), fva, P["desc_b"], P["angle_b"], fvb)
    assert nm == 0 and (mb < 0).all() and (ma < 0).all()
    # disjoint node sets -> no matches
    odd = orbx.FeatureVector(fvb.node_ids + 100000, fvb.offsets, fvb.indices)
    mb, ma, nm = orbx.search_by_bow(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], odd)
    assert nm == 0 and (mb < 0).all()
    # empty sets
    empty = orbx.FeatureVector(np.zeros(0, np.uint32), np.zeros(1, np.int32), np.zeros(0, np.uint32))
    mb, ma, nm = orbx.search_by_bow(np.zeros((0, 32), np.uint8), np.zeros(0, np.float32), np.zeros(0, np.uint8), empty, P["desc_b"],
                                    P["angle_b"], fvb)
    assert nm == 0 and len(ma) == 0 and (mb < 0).all()
    (ids, vals), fv = voc.transform(np.zeros((0, 32), np.uint8), 2)
    assert len(ids) == 0 and len(fv.node_ids) == 0
    # malformed feature vectors are rejected, not dereferenced
    bad = orbx.FeatureVector(fvb.node_ids[::-1].copy(), fvb.offsets, fvb.indices)
    with pytest.raises(orbx.OrbxError):
        orbx.search_by_bow(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], bad)
    big = orbx.FeatureVector(fvb.node_ids, fvb.offsets, fvb.indices + 1000)
    with pytest.raises(orbx.OrbxError):
        orbx.search_by_bow(P["desc_a"], P["angle_a"], P["valid_a"], fva, P["desc_b"], P["angle_b"], big)
    # a single node holding everything (levelsup >= L): the whole-set scan of the brute-force definition
    _, ra = voc.transform(P["desc_a"], 4); _, rb = voc.transform(P["desc_b"], 4)
    assert ra.node_ids.tolist() == [0]
    mb, ma, nm = orbx.search_by_bow(P["desc_a"], P["angle_a"], P["valid_a"], ra, P["desc_b"], P["angle_b"], rb, nnratio=0.9)
    omb, onm = oracle.search_by_bow_kf_frame(P["desc_a"], P["angle_a"], P["valid_a"], fv_tuple(ra), P["desc_b"], P["angle_b"],
                                             fv_tuple(rb), -1, 0.9, True)
    assert np.array_equal(mb, omb) and nm == onm and nm > 20


def test_golden_fixture_through_c_abi():
    """The committed known answers (tests/golden/bow_golden.json) reproduced by the CUDA path alone (no oracle code runs)."""
    gold = json.load(open(test_bow_cpu.GOLD_PATH))
    for c in gold["cases"]:
        def make_voc(parent, desc, weights, k, L):
            return orbx.ORBVocabulary(parent, desc, weights, k, L)

        def wrap(fv):
            return fv if isinstance(fv, orbx.FeatureVector) else orbx.FeatureVector(*fv)

        def kf_frame(P, fva, fvb):
            mb, _, nm = orbx.search_by_bow(P["desc_a"], P["angle_a"], P["valid_a"], wrap(fva), P["desc_b"], P["angle_b"], wrap(fvb),
                                           nnratio=0.6, check_orientation=True)
            return mb, nm

        def kf_kf(P, fva, fvb):
            _, ma, nm = orbx.search_by_bow(P["desc_a"], P["angle_a"], P["valid_a"], wrap(fva), P["desc_b"], P["angle_b"], wrap(fvb),
                                           valid_b=P["valid_b"], kf_kf=True, nnratio=0.8, check_orientation=True)
            return ma, nm

        def tri(P, fva, fvb, seed):
            kpa, kpb, F, ep, scale, sigma2, fa, fb, sa, sb = test_bow_cpu.tri_inputs(P, seed)
            return orbx.search_for_triangulation(kpa, P["desc_a"], fa, sa, wrap(fva), kpb, P["desc_b"], fb, sb, wrap(fvb), F, ep, scale, sigma2)

        class V:          # adapt ORBVocabulary to the (ids, vals), (nodes, offs, idx) shape run_golden_case expects
            def __init__(self, v): self.v = v
            def transform_features(self, d, lu): return self.v.transform_features(d, lu)
            def transform(self, d, lu):
                bow, fv = self.v.transform(d, lu)
                return bow, fv_tuple(fv)

        got = test_bow_cpu.run_golden_case(c, lambda *a: V(make_voc(*a)), kf_frame, kf_kf, tri)
        assert got == c["expect"], c["name"]
