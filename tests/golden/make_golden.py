"""Regenerates tests/golden/*.npz.  Runs ONLY in the build container (needs /root/reference):

  * extractor vectors come from the reference's own ORBextractor.cpp, compiled unmodified against
    oracle/cvshim with the canonical flags/allocator (oracle/_ref/libref_canonical.so, see oracle/Makefile);
  * stereo vectors come from the reference's own Frame class, imported unchanged from /root/reference/Frame.py
    and constructed exactly like Tracking.grab_image_stereo does (Tracking.py:95-112), on top of those
    extractors -- so mvuRight / mvDepth are produced by Frame.compute_stereo_matches itself.

Usage: python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import cv2  # noqa: E402
from Frame import Frame  # noqa: E402  (the reference's own class)

from oracle.refext import RefExtractor  # noqa: E402
from pyorbslam_b200.synthetic import make_kitti_like_pair, make_stereo_pair, pair_digest  # noqa: E402


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def frame_args(fx, fy, cx, cy, W, H):   # Tracking.py:97-109 with zero distortion (bounds = image rectangle)
    return [fx, fy, cx, cy, 1.0 / fx, 1.0 / fy, 64.0 / W, 48.0 / H, 0.0, float(W), 0.0, float(H), 48, 64]


def stereo_case(name, idx, H, W, params, fx, fy, cx, cy, mbf, gen=make_stereo_pair):
    L, R = gen(idx, H, W)
    eL, eR = RefExtractor(*params), RefExtractor(*params)
    mK = np.eye(3, dtype=np.float32)
    mK[0, 0], mK[1, 1], mK[0, 2], mK[1, 2] = fx, fy, cx, cy
    f = Frame(L, R, 0.0, eL, eR, None, mK, np.zeros((4, 1), np.float32), mbf, mbf * 35 / fx, frame_args(fx, fy, cx, cy, W, H))
    kL = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in f.mvKeys], np.float32)
    kR = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in f.mvKeysRight], np.float32)
    uR = np.array([float(v) for v in f.mvuRight], np.float64)
    dep = np.array([float(v) for v in f.mvDepth], np.float64)
    # element types as the reference leaves them (0 = python int -1, 1 = np.float32 scalar, 2 = python float: the disparity <= 0 branch)
    codes = {int: 0, np.float32: 1, float: 2}
    tu = np.array([codes[type(v)] for v in f.mvuRight], np.int8)
    td = np.array([codes[type(v)] for v in f.mvDepth], np.int8)
    np.savez_compressed(os.path.join(HERE, name), idx=idx, H=H, W=W, params=np.array(params, np.float64),
                        fx=fx, fy=fy, cx=cx, cy=cy, mbf=mbf, image_digest=pair_digest(L, R),
                        kpsL=kL, descL=f.mDescriptors, kpsR=kR, descR=f.mDescriptorsRight, uRight=uR, depth=dep,
                        uRight_type=tu, depth_type=td)
    print(name, "N", f.N, "matched", int((uR >= 0).sum()), "types", np.bincount(tu, minlength=3).tolist())


def main():
    # config 1: the bundled fixture image (8-bit gray 1226x370), test.py parameters
    src = os.path.join(REF, "pyORBExtractor", "kitti06-436.png")
    img = cv2.imread(src, cv2.IMREAD_UNCHANGED)
    assert img.shape == (370, 1226) and img.dtype == np.uint8
    np.save(os.path.join(HERE, "kitti06-436.gray.npy"), img)   # raw pixels; avoids needing a PNG decoder on the box
    e = RefExtractor(2000, 1.2, 8, 20, 7)
    k, d = e.extract_arrays(img)
    pyr = e.GetImagePyramid()
    np.savez_compressed(os.path.join(HERE, "kitti06_extract.npz"), kps=k, desc=d,
                        level_sizes=np.array([p.shape for p in pyr]), pyramid_view_sha=np.array([sha(p) for p in pyr]))
    print("kitti06", len(k))
    # config 2: KITTI00-02 camera (configs/KITTI00-02.yaml:7-24)
    stereo_case("stereo_kitti_shape.npz", 0, 376, 1241, (2000, 1.2, 8, 20, 7), 718.856, 718.856, 607.1928, 185.2157, 386.1448)
    stereo_case("stereo_small.npz", 7, 240, 640, (1000, 1.2, 6, 20, 7), 500.0, 500.0, 320.0, 120.0, 100.0)
    # the bench's default scenes (KITTI-like road scene, pyorbslam_b200/synthetic.py:make_kitti_like_pair)
    stereo_case("stereo_kitti_like.npz", 1000, 376, 1241, (2000, 1.2, 8, 20, 7), 718.856, 718.856, 607.1928, 185.2157, 386.1448, gen=make_kitti_like_pair)
    # config 4 (digests only: the arrays would be ~0.5 MB); input = left view of synthetic pair 4 at 2560x1440
    big, _ = make_stereo_pair(4, 1440, 2560)
    e4 = RefExtractor(8000, 1.2, 12, 20, 7)
    k4, d4 = e4.extract_arrays(big)
    np.savez_compressed(os.path.join(HERE, "hires_extract_digest.npz"), image_sha=sha(big), n=len(k4), kps_sha=sha(k4),
                        desc_sha=sha(d4), per_level=np.bincount(k4[:, 5].astype(int), minlength=12))
    print("hires", len(k4))
    bow_case()
    matcher_case()


def bow_case():
    """the reference's own pyDBoW classes on a small synthetic vocabulary."""
    import tempfile
    from pyDBoW.TemplatedVocabulary import TemplatedVocabulary   # the reference's own class
    from oracle.bow_py import make_vocab_text
    text = make_vocab_text()
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(text)
        path = f.name
    voc = TemplatedVocabulary(k=5, L=3, weighting="TF_IDF", scoring="L1_NORM")      # System.py:38
    assert voc.load_from_text_file(path)
    os.unlink(path)
    L, _ = make_stereo_pair(7, 240, 640)
    _, desc = RefExtractor(1000, 1.2, 6, 20, 7).extract_arrays(L)
    desc = desc[:400]
    out = {}
    for lu in (4, 2, 1):
        bv, fv = voc.transform(desc, lu)
        out[f"bv_keys_{lu}"] = np.array(list(bv.keys()), np.int64)
        out[f"bv_vals_{lu}"] = np.array(list(bv.values()), np.float64)
        out[f"fv_keys_{lu}"] = np.array(list(fv.keys()), np.int64)
        out[f"fv_lens_{lu}"] = np.array([len(v) for v in fv.values()], np.int64)
        out[f"fv_idx_{lu}"] = np.array([i for v in fv.values() for i in v], np.int64)
    np.savez_compressed(os.path.join(HERE, "bow_small.npz"), vocab_sha=hashlib.sha256(text.encode()).hexdigest(), desc=desc,
                        n_nodes=len(voc.nodes), n_words=len(voc.words), **out)
    print("bow", len(voc.nodes), "nodes", len(voc.words), "words", {k: len(v) for k, v in out.items() if k.startswith("bv_keys")})


def matcher_case():
    """SURVEY.md 8(f) rank 1 (first piece): the reference's own ORBMatcher.search_by_BoW_* on a synthetic two-view case."""
    from ORBMatcher import ORBMatcher     # the reference's own class
    from oracle.matcher_py import make_case
    out = {}
    for tag, (ratio, ori) in {"a": (0.7, True), "b": (1, True), "c": (0.9, False)}.items():
        A, B = make_case()
        m = ORBMatcher(ratio, ori)
        n1, v1 = m.search_by_BoW_kf_f(A, B)
        n2, v2 = m.search_by_BoW_kf_kf(A, B)
        out[f"kf_f_n_{tag}"] = n1
        out[f"kf_f_{tag}"] = np.array([-1 if p is None else p.uid for p in v1], np.int64)
        out[f"kf_kf_n_{tag}"] = n2
        out[f"kf_kf_{tag}"] = np.array([-1 if p is None else p.uid for p in v2], np.int64)
    # the two frame projection searches, on top of the reference's own Frame.get_features_in_area / assign_features_to_grid
    from oracle.matcher_py import make_projection_case

    def ref_assign(f):
        f.pos_in_grid = lambda kps, _f=f: Frame.pos_in_grid(_f, kps)
        Frame.assign_features_to_grid(f)
        return f.mGrid
    for tag, (ratio, ori, th, motion) in {"p": (0.9, True, 15, (0.05, 0.0, 0.3)), "q": (0.8, False, 7, (0.0, 0.02, -0.9)),
                                          "r": (1, True, 10, (0.4, 0.0, 0.0))}.items():
        m = ORBMatcher(ratio, ori)
        cur, last, local = make_projection_case(get_area=Frame.get_features_in_area, assign=ref_assign, motion=motion)
        out[f"f_f_n_{tag}"] = m.search_by_projection_f_f(cur, last, th)
        out[f"f_f_{tag}"] = np.array([-1 if p is None else p.uid for p in cur.mvpMapPoints], np.int64)
        cur, last, local = make_projection_case(get_area=Frame.get_features_in_area, assign=ref_assign, motion=motion)
        out[f"f_p_n_{tag}"] = m.search_by_projection_f_p(cur, local, float(th) / 5)
        out[f"f_p_{tag}"] = np.array([-1 if p is None else p.uid for p in cur.mvpMapPoints], np.int64)
    np.savez_compressed(os.path.join(HERE, "matcher_small.npz"), **out)
    print("matcher", {k: int(v) for k, v in out.items() if "_n_" in k})


if __name__ == "__main__":
    main()
