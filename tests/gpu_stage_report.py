"""Stage-by-stage comparison of the CUDA path with the CPU oracle (a debugging aid, run on the GPU box:
`python tests/gpu_stage_report.py`).  Prints where the first divergence appears."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
from pyorbslam_b200 import ORBextractor  # noqa: E402
from pyorbslam_b200.stereo import stereo_resident  # noqa: E402
from pyorbslam_b200.synthetic import make_stereo_pair  # noqa: E402


def compare(img, params, tag):
    g = ORBextractor(*params)
    o = O.OracleExtractor(*params)
    t = time.time(); kg, dg = g.extract_arrays(img); tg = time.time() - t
    t = time.time(); kg, dg = g.extract_arrays(img); tg2 = time.time() - t
    ko, do = o.extract_arrays(img)
    print(f"[{tag}] {img.shape} {params}: gpu N={len(kg)} oracle N={len(ko)}  first call {tg*1e3:.1f} ms, second {tg2*1e3:.1f} ms")
    ok = True
    for l in range(params[2]):
        w, h = o.level_size(l)
        raw_ok = np.array_equal(g.level_image(l), o.level_bordered(l)[19:19 + h, 19:19 + w])
        view_ok = np.array_equal(g.GetImagePyramid()[l], o.GetImagePyramid()[l])
        blur_ok = True
        ob = o.level_blurred(l)
        if ob.any():
            blur_ok = np.array_equal(g.level_image(l, True), ob)
        cg, co = g.level_candidates(l), o.level_candidates(l)
        cand_ok = cg.shape == co.shape and np.array_equal(cg, co)
        lk = o.level_keypoints(l)
        gk = kg[kg[:, 5] == l]
        print(f"   level {l}: {w}x{h} raw {raw_ok} view {view_ok} blur {blur_ok} cand {len(cg)}/{len(co)} {cand_ok}  kps {len(gk)}/{len(lk)}")
        if not cand_ok and len(cg) and len(co):
            m = min(len(cg), len(co))
            d = np.nonzero((cg[:m] != co[:m]).any(1))[0]
            print("      first cand diff at", d[:3], cg[d[:3]].tolist(), co[d[:3]].tolist())
        ok &= raw_ok and view_ok and blur_ok and cand_ok
    same_k = kg.shape == ko.shape and np.array_equal(kg.view(np.uint32), ko.view(np.uint32))
    same_d = dg.shape == do.shape and np.array_equal(dg, do)
    print(f"   keypoints bit-exact {same_k}, descriptors bit-exact {same_d}")
    if not same_k and kg.shape == ko.shape:
        d = np.nonzero((kg.view(np.uint32) != ko.view(np.uint32)).any(1))[0]
        print("      kp diffs", len(d), "first", d[:3], kg[d[:3]].tolist(), ko[d[:3]].tolist())
        for c, name in enumerate(["x", "y", "size", "angle", "resp", "oct"]):
            print("        col", name, int((kg[:, c].view(np.uint32) != ko[:, c].view(np.uint32)).sum()))
    if not same_d and dg.shape == do.shape:
        print("      desc rows differing", int((dg != do).any(1).sum()))
    return ok and same_k and same_d, g, o, (kg, dg), (ko, do)


def main():
    img = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kitti06-436.gray.npy"))
    allok = compare(img, (2000, 1.2, 8, 20, 7), "kitti06")[0]
    L, R = make_stereo_pair(0)
    okL, gL, oL, (kL, dL), (okLk, okLd) = compare(L, (2000, 1.2, 8, 20, 7), "synthL")
    okR, gR, oR, (kR, dR), (okRk, okRd) = compare(R, (2000, 1.2, 8, 20, 7), "synthR")
    allok &= okL and okR
    t = time.time(); uR, dep, mi = stereo_resident(gL, gR, 386.1448, 718.856); ts = time.time() - t
    ou, od, oi, _ = O.stereo(okLk[:, [0, 1, 5]], okLd, okRk[:, [0, 1, 5]], okRd, oL.sf, oL.isf, oL.GetImagePyramid(), oR.GetImagePyramid(), 386.1448, 718.856)
    if len(uR) == len(ou):
        print(f"[stereo] {ts*1e3:.1f} ms matched gpu {(uR>=0).sum()} oracle {(ou>=0).sum()} idx equal {np.array_equal(mi, oi)} "
              f"uR bit-exact {np.array_equal(uR.view(np.uint32), ou.view(np.uint32))} depth bit-exact {np.array_equal(dep.view(np.uint32), od.view(np.uint32))}")
        allok &= np.array_equal(mi, oi) and np.array_equal(uR.view(np.uint32), ou.view(np.uint32))
    else:
        print("[stereo] length mismatch", len(uR), len(ou)); allok = False
    rng = np.random.default_rng(5)
    allok &= compare(rng.integers(0, 256, (376, 1241), dtype=np.uint8), (2000, 1.2, 8, 20, 7), "noise")[0]
    allok &= compare(make_stereo_pair(5, 240, 320)[0], (500, 1.3, 5, 15, 5), "small")[0]
    allok &= compare(make_stereo_pair(4, 1440, 2560)[0], (8000, 1.2, 12, 20, 7), "hires")[0]
    print("ALL OK" if allok else "MISMATCHES PRESENT")


if __name__ == "__main__":
    main()
