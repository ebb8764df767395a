"""Seeded synthetic vocabularies and key-frame feature sets for the bag-of-words / guided-search tests (ORBvoc.txt and the
datasets are not available offline).  numpy only; used by the CPU and the GPU tests and by tests/golden/make_bow_golden.py."""
import numpy as np

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


def flip_bits(rng, desc, nbits):
    """desc [n, 32] u8 with `nbits[i]` random bits flipped in row i."""
    bits = np.unpackbits(desc, axis=1)
    for i in range(len(desc)):
        pos = rng.choice(256, size=int(nbits[i]), replace=False)
        bits[i, pos] ^= 1
    return np.packbits(bits, axis=1)


def make_vocab(seed, k, L, zero_weight_frac=0.05, dup_frac=0.05):
    """A full k-ary tree with L levels below the root, in the node order loadFromTextFile would create from a
    breadth-first text file: parent[], descriptors[n,32], weights[n].  Children drift from their parent by fewer bits
    the deeper they are; a few siblings are exact duplicates (ties -> the first child must win) and a few leaves have
    weight 0 ("stopped" words)."""
    rng = np.random.default_rng(seed)
    parent = [np.array([-1], np.int64)]
    desc = [np.zeros((1, 32), np.uint8)]
    level_ids = np.array([0], np.int64)
    level_desc = desc[0]
    drift = [128, 64, 48, 32, 24, 16, 12, 8, 6, 4]
    next_id = 1
    for lev in range(1, L + 1):
        n_p = len(level_ids)
        kids = np.repeat(level_desc, k, axis=0)
        for c0 in range(0, len(kids), 65536):           # Bernoulli bit flips, chunked to bound memory
            m = rng.random((min(65536, len(kids) - c0), 256)) < drift[min(lev - 1, 9)] / 256.0
            kids[c0:c0 + len(m)] ^= np.packbits(m, axis=1)
        if k >= 3:
            dup = np.flatnonzero(rng.random(n_p) < dup_frac)
            kids[dup * k + 2] = kids[dup * k + 1]
        parent.append(np.repeat(level_ids, k))
        desc.append(kids)
        level_ids = np.arange(next_id, next_id + n_p * k, dtype=np.int64)
        next_id += n_p * k
        level_desc = kids
    parent = np.concatenate(parent); desc = np.concatenate(desc)
    weights = rng.uniform(0.1, 9.0, len(parent))
    weights[rng.choice(level_ids, max(1, int(zero_weight_frac * len(level_ids))), replace=False)] = 0.0
    return parent.astype(np.int32), desc.astype(np.uint8), weights.astype(np.float64)


def write_vocab_text(path, parent, desc, weights, k, L, scoring=0, weighting=0):
    """The ORBvoc.txt format read by TemplatedVocabulary::loadFromTextFile (TemplatedVocabulary.h:1337-1420)."""
    has_child = np.zeros(len(parent), bool)
    has_child[parent[1:]] = True
    with open(path, "w") as f:
        f.write("%d %d %d %d\n" % (k, L, scoring, weighting))
        for i in range(1, len(parent)):
            f.write("%d %d %s %.17g\n" % (parent[i], 0 if has_child[i] else 1, " ".join(str(int(b)) for b in desc[i]), weights[i]))


def make_features(seed, voc_desc, parent, n, noise_bits=10):
    """n descriptors near random leaves of the vocabulary."""
    rng = np.random.default_rng(seed)
    has_child = np.zeros(len(parent), bool)
    has_child[parent[1:]] = True
    leaves = np.flatnonzero(~has_child)
    pick = rng.choice(leaves, n)
    return flip_bits(rng, voc_desc[pick], rng.integers(0, noise_bits + 1, n))


def make_pair(seed, voc_desc, parent, n_a, n_b, match_frac=0.6, rot=37.0):
    """Two feature sets with planted correspondences: B holds noisy copies of a fraction of A (few flipped bits, some exact
    duplicates so that ties and failing ratio tests occur) plus unrelated features; angles follow a dominant rotation with
    outliers; map-point validity flags are random."""
    rng = np.random.default_rng(seed)
    desc_a = make_features(seed + 1, voc_desc, parent, n_a)
    n_m = int(match_frac * min(n_a, n_b))
    src = rng.choice(n_a, n_m, replace=False)
    copies = flip_bits(rng, desc_a[src], rng.integers(0, 40, n_m))
    dup = rng.random(n_m) < 0.1
    copies[dup] = desc_a[src[dup]]
    rest = make_features(seed + 2, voc_desc, parent, n_b - n_m)
    desc_b = np.concatenate([copies, rest])
    perm = rng.permutation(n_b)
    desc_b = desc_b[perm]
    angle_a = rng.uniform(0, 360, n_a).astype(np.float32)
    angle_b = rng.uniform(0, 360, n_b).astype(np.float32)
    inv = np.empty(n_b, np.int64); inv[perm] = np.arange(n_b)
    good = rng.random(n_m) < 0.8
    ab = (angle_a[src] - np.float32(rot) + rng.normal(0, 4, n_m).astype(np.float32)) % np.float32(360)
    ab = np.where(ab >= 360, 0, ab).astype(np.float32)
    angle_b[inv[:n_m][good]] = ab[good]
    valid_a = (rng.random(n_a) < 0.85).astype(np.uint8)
    valid_b = (rng.random(n_b) < 0.85).astype(np.uint8)
    return dict(desc_a=desc_a, desc_b=desc_b, angle_a=angle_a, angle_b=angle_b, valid_a=valid_a, valid_b=valid_b)


def make_keypoints(seed, n, angles, cols=752, rows=480, nlevels=8):
    rng = np.random.default_rng(seed)
    kp = np.zeros(n, KP_DTYPE)
    kp["x"] = rng.uniform(20, cols - 20, n).astype(np.float32)
    kp["y"] = rng.uniform(20, rows - 20, n).astype(np.float32)
    kp["octave"] = rng.integers(0, nlevels, n)
    kp["angle"] = angles
    kp["size"] = 31
    kp["class_id"] = -1
    return kp


def fundamental(seed, fx=435.2, fy=435.2, cx=367.2, cy=252.2):
    """F12 = K1^-T [t12]x R12 K2^-1 for a small random relative pose (float32, as the caller of the C ABI would pass it)."""
    rng = np.random.default_rng(seed)
    w = rng.normal(0, 0.05, 3)
    th = np.linalg.norm(w)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx
    t = rng.normal(0, 0.3, 3)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
    F = np.linalg.inv(K).T @ tx @ R @ np.linalg.inv(K)
    ep = np.array([cx + rng.normal(0, 120), cy + rng.normal(0, 90)], np.float32)
    return F.astype(np.float32), ep
