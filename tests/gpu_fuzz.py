"""Randomised parity check of the CUDA path against the CPU oracle: random image sizes, contents (synthetic scenes, noise, blurred
noise, low-contrast scenes that send most cells through the minThFAST retry) and extractor parameters (1-8 levels, scale factors
1.1-2.0, quotas 50-4000, several threshold pairs); extractor outputs bit-exact, then stereo (uRight, depth, match index) bit-exact
or both sides raising IndexError.  `python tests/gpu_fuzz.py [seed] [cases]` on a GPU box; tests/test_gpu_parity.py runs a small
batch.  1110 single-image cases (seeds 1, 7, 11, 21) and 85 batch-engine cases (seeds 3, 22) passed with 0 mismatches at the end of round 1;
round 2: 1680 single-image and 100 batch-engine cases, 0 mismatches (the last 460 on the final FAST kernel, 60 of them with
B200ORB_FAST_LC=128, i.e. through the row-band path)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
from pyorbslam_b200 import ORBextractor  # noqa: E402
from pyorbslam_b200.stereo import stereo_resident  # noqa: E402
from pyorbslam_b200.synthetic import make_kitti_like_pair, make_stereo_pair  # noqa: E402


def run(seed=0, ncase=100, verbose=True):
    rng = np.random.default_rng(seed)
    bad = refused = 0
    t0 = time.time()
    for c in range(ncase):
        H = int(rng.integers(90, 700)); W = int(rng.integers(max(100, H // 2 + 1), 1400))
        nlev = int(rng.integers(1, 9)); sf = float(rng.choice([1.1, 1.2, 1.2, 1.25, 1.3, 1.41, 1.5, 1.8, 2.0]))
        while min(H, W) / sf ** (nlev - 1) < 70: nlev -= 1
        nf = int(rng.choice([50, 300, 1000, 2000, 4000])); ini = int(rng.choice([20, 20, 12, 30, 7])); mn = int(rng.choice([7, 7, 5, 3, 12]))
        kind = rng.integers(0, 5)
        if kind == 4:
            L, R = make_kitti_like_pair(int(rng.integers(0, 10 ** 6)), H, W)      # the bench's scenes
        elif kind == 0:
            L, R = make_stereo_pair(int(rng.integers(0, 10 ** 6)), H, W)
        elif kind == 1:
            L = rng.integers(0, 256, (H, W), dtype=np.uint8); R = np.roll(L, -int(rng.integers(1, 40)), axis=1)
        elif kind == 2:
            L = O.blur7(rng.integers(0, 256, (H, W), dtype=np.uint8)); R = np.roll(L, -int(rng.integers(1, 40)), axis=1)
        else:
            base = make_stereo_pair(int(rng.integers(0, 10 ** 6)), H, W)
            L = (base[0] // 4 + 90).astype(np.uint8); R = (base[1] // 4 + 90).astype(np.uint8)       # low contrast: many minThFAST retries
        L, R = np.ascontiguousarray(L), np.ascontiguousarray(R)
        params = (nf, sf, nlev, ini, mn)
        try:
            gL, gR = ORBextractor(*params), ORBextractor(*params)
            oL, oR = O.OracleExtractor(*params), O.OracleExtractor(*params)
            kgl, dgl = gL.extract_arrays(L); kgr, dgr = gR.extract_arrays(R)
            kol, dol = oL.extract_arrays(L); kor, dor = oR.extract_arrays(R)
            ok = (kgl.shape == kol.shape and np.array_equal(kgl.view(np.uint32), kol.view(np.uint32)) and np.array_equal(dgl, dol) and
                  kgr.shape == kor.shape and np.array_equal(kgr.view(np.uint32), kor.view(np.uint32)) and np.array_equal(dgr, dor))
            if ok and len(kol) and len(kor):
                try:
                    ou, od, oi, _ = O.stereo(kol[:, [0, 1, 5]], dol, kor[:, [0, 1, 5]], dor, oL.sf, oL.isf, oL.GetImagePyramid(), oR.GetImagePyramid(), 386.1448, 718.856)
                    oerr = None
                except IndexError as e:
                    oerr = e
                try:
                    gu, gd, gi = stereo_resident(gL, gR, 386.1448, 718.856)
                    gerr = None
                except IndexError as e:
                    gerr = e
                if (oerr is None) != (gerr is None):
                    ok = False
                elif oerr is None:
                    ok = np.array_equal(gu.view(np.uint32), ou.view(np.uint32)) and np.array_equal(gd.view(np.uint32), od.view(np.uint32)) and np.array_equal(gi, oi)
        except ValueError as e:
            refused += 1
            if verbose:
                print("case", c, (H, W), params, "refused:", e)
            continue
        if not ok:
            bad += 1
            print("MISMATCH case", c, (H, W), params, "kind", int(kind), len(kgl), len(kol))
    if verbose:
        print("fuzz done:", ncase, "cases,", bad, "mismatches,", refused, "refused,", round(time.time() - t0, 1), "s")
    return ncase, bad, refused


def run_batch(seed=0, ncase=10, pairs=9, verbose=True):
    """The batched engine (many images per launch sequence) against the one-image API on random
    geometries: keypoints, descriptors, uRight, depth, match index of every pair identical."""
    import torch
    from pyorbslam_b200 import StereoFrontend
    rng = np.random.default_rng(seed)
    bad = done = 0
    for c in range(ncase):
        H = int(rng.integers(120, 500)); W = int(rng.integers(H, 1300))
        nlev = int(rng.integers(2, 9)); sf = float(rng.choice([1.15, 1.2, 1.3, 1.5, 2.0]))
        while min(H, W) / sf ** (nlev - 1) < 70: nlev -= 1
        params = (int(rng.choice([300, 1000, 2000])), sf, nlev, int(rng.choice([20, 12])), int(rng.choice([7, 5])))
        try:
            fe = StereoFrontend(*params, H, W, pairs)
            eL, eR = ORBextractor(*params, reuse_identical_input=False), ORBextractor(*params, reuse_identical_input=False)
        except ValueError:
            continue
        gen = make_kitti_like_pair if c % 2 else make_stereo_pair
        ps = [gen(int(rng.integers(0, 10 ** 6)), H, W) for _ in range(pairs)]
        out = fe.run(torch.from_numpy(np.stack([p[0] for p in ps])).cuda(), torch.from_numpy(np.stack([p[1] for p in ps])).cuda(), 386.1448, 718.856)
        torch.cuda.synchronize()
        nk = out["nkp"].cpu().numpy()
        ok = True
        for i, (L, R) in enumerate(ps):
            kL, dL = eL.extract_arrays(L); kR, dR = eR.extract_arrays(R)
            try:
                u, d, m = stereo_resident(eL, eR, 386.1448, 718.856)
            except IndexError:
                continue                        # the batch engine flags these pairs instead of raising; not compared here
            nl, nr = int(nk[0, i]), int(nk[1, i])
            ok = ok and (nl, nr) == (len(kL), len(kR))
            if not ok:
                break
            ok = ok and np.array_equal(out["kps"][0, i, :nl].cpu().numpy().view(np.uint32), kL.view(np.uint32))
            ok = ok and np.array_equal(out["kps"][1, i, :nr].cpu().numpy().view(np.uint32), kR.view(np.uint32))
            ok = ok and np.array_equal(out["desc"][0, i, :nl].cpu().numpy(), dL) and np.array_equal(out["desc"][1, i, :nr].cpu().numpy(), dR)
            ok = ok and np.array_equal(out["uRight"][i, :nl].cpu().numpy().view(np.uint32), u.view(np.uint32))
            ok = ok and np.array_equal(out["depth"][i, :nl].cpu().numpy().view(np.uint32), d.view(np.uint32))
            ok = ok and np.array_equal(out["matchIdx"][i, :nl].cpu().numpy(), m)
        done += 1
        if not ok:
            bad += 1
            print("BATCH MISMATCH case", c, (H, W), params)
        del fe
    if verbose:
        print("batch fuzz done:", done, "cases,", bad, "mismatches")
    return done, bad


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[3] == "batch":
        sys.exit(1 if run_batch(int(sys.argv[1]), int(sys.argv[2]))[1] else 0)
    _, nbad, _ = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 100)
    sys.exit(1 if nbad else 0)
