"""TEST INFRASTRUCTURE -- Python restatement of ORBMatcher.search_by_BoW_kf_f / search_by_BoW_kf_kf
(reference ORBMatcher.py:12-213) with the reference's own per-pair Python Hamming distance.  Pinned against vectors produced
by the reference's ORBMatcher class itself (tests/golden/make_golden.py::matcher_case -> matcher_small.npz)."""
import types

import numpy as np

TH_LOW, HISTO_LENGTH = 50, 30


def distance(a, b):                      # ORBMatcher.py:12-14
    return sum(bin(int(v)).count("1") for v in np.bitwise_xor(a, b))


def _three(rot_hist):
    return np.argsort([len(h) for h in rot_hist])[::-1][:3]


def _bin(rot):
    if rot < 0.0:
        rot += 360.0
    b = round(rot * (1.0 / HISTO_LENGTH))
    return 0 if b == HISTO_LENGTH else b


def _merge(fa, fb):
    """Nodes present in both ordered feature vectors, in order (the reference's two-iterator merge)."""
    ia, ib = iter(fa), iter(fb)
    try:
        a, b = next(ia), next(ib)
        while True:
            if a == b:
                yield a
                a, b = next(ia), next(ib)
            elif a < b:
                a = next(ia)
            else:
                b = next(ib)
    except StopIteration:
        return


def bow_kf_f(kf, frame, nnratio=1, check_ori=True):
    mps = kf.get_map_point_matches()
    out = [None] * frame.N
    n = 0
    hist = [[] for _ in range(HISTO_LENGTH)]
    for node in _merge(kf.mFeatVec, frame.mFeatVec):
        for ik in kf.mFeatVec[node]:
            mp = mps[ik]
            if not mp or mp.is_bad():
                continue
            b1, bf, b2 = 256, -1, 256
            for jf in frame.mFeatVec[node]:
                if out[jf]:
                    continue
                d = distance(kf.mDescriptors[ik], frame.mDescriptors[jf])
                if d < b1:
                    b2, b1, bf = b1, d, jf
                elif d < b2:
                    b2 = d
            if b1 <= TH_LOW and float(b1) < nnratio * float(b2):
                out[bf] = mp
                if check_ori:
                    hist[_bin(kf.mvKeysUn[ik].angle - frame.mvKeys[bf].angle)].append(bf)
                n += 1
    if check_ori:
        keep = _three(hist)
        for i in range(HISTO_LENGTH):
            if i not in keep:
                for idx in hist[i]:
                    out[idx] = None
                    n -= 1
    return n, out


def bow_kf_kf(k1, k2, nnratio=1, check_ori=True):
    m1, m2 = k1.get_map_point_matches(), k2.get_map_point_matches()
    out = [None] * len(m1)
    used = [False] * len(m2)
    n = 0
    hist = [[] for _ in range(HISTO_LENGTH)]
    for node in _merge(k1.mFeatVec, k2.mFeatVec):
        for i1 in k1.mFeatVec[node]:
            if not m1[i1] or m1[i1].is_bad():
                continue
            b1, bi, b2 = 256, -1, 256
            for i2 in k2.mFeatVec[node]:
                if used[i2] or not m2[i2] or m2[i2].is_bad():
                    continue
                d = distance(k1.mDescriptors[i1], k2.mDescriptors[i2])
                if d < b1:
                    b2, b1, bi = b1, d, i2
                elif d < b2:
                    b2 = d
            if b1 < TH_LOW and b1 < nnratio * b2:
                out[i1] = m2[bi]
                used[bi] = True
                if check_ori:
                    hist[_bin(k1.mvKeysUn[i1].angle - k2.mvKeysUn[bi].angle)].append(i1)
                n += 1
    if check_ori:
        keep = _three(hist)
        for i in range(HISTO_LENGTH):
            if i not in keep:
                for idx in hist[i]:
                    out[idx] = None
                    n -= 1
    return n, out


# ---------------------------------------------------------------- synthetic scene for the tests / golden vectors
class FakeMapPoint:
    def __init__(self, uid, bad):
        self.uid, self._bad = uid, bad

    def is_bad(self):
        return self._bad


def make_case(seed=3, n=320, n_nodes=6):
    """Two 'keyframes' observing the same scene: descriptors of B are noisy copies of A's (random permutation), both carry a
    FeatureVector over `n_nodes` vocabulary nodes, some features have no map point, some map points are bad."""
    from collections import OrderedDict
    rng = np.random.default_rng(seed)
    dA = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    twins = rng.choice(n, n // 3, replace=False)        # near-duplicate descriptors: the best / second-best ratio test bites
    dA[twins] = dA[(twins + 1) % n] ^ np.packbits(rng.random((len(twins), 256)) < 0.04, axis=1, bitorder="little")
    perm = rng.permutation(n)
    flips = np.packbits(rng.random((n, 256)) < rng.choice([0.03, 0.12, 0.3], size=(n, 1)), axis=1, bitorder="little")
    dB = (dA ^ flips)[perm]
    node_of_A = rng.integers(0, n_nodes, n)
    node_of_B = np.where(rng.random(n) < 0.9, node_of_A[perm], rng.integers(0, n_nodes, n))

    def featvec(nodes):
        fv = {}
        for i, nd in enumerate(nodes.tolist()):
            fv.setdefault(nd * 7 + 3, []).append(i)
        return OrderedDict(sorted(fv.items()))
    angA = rng.uniform(0, 360, n).astype(np.float32)
    angB = ((angA[perm] + rng.choice([0.0, 40.0, 80.0, 120.0, 200.0], size=n, p=[0.5, 0.2, 0.15, 0.1, 0.05]) + rng.normal(0, 2, n)) % 360).astype(np.float32)

    def side(desc, nodes, ang, off):
        mps = [None if rng.random() < 0.15 else FakeMapPoint(off + i, bool(rng.random() < 0.05)) for i in range(n)]
        keys = [types.SimpleNamespace(angle=float(a)) for a in ang]
        return types.SimpleNamespace(mDescriptors=desc, mFeatVec=featvec(nodes), mvKeysUn=keys, mvKeys=keys, N=n,
                                     get_map_point_matches=lambda m=mps: m, mps=mps)
    return side(dA, node_of_A, angA, 0), side(dB, node_of_B, angB, 10000)
