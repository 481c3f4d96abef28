"""Round-2 parity additions (B200 box), all through the C ABI:
  * the batched octree's global node scratch (per-level quota too large for shared memory) with several pairs per launch,
  * range errors of the throughput API (the reference raises where a SAD window leaves the pyramid view, Frame.py:230-250),
  * element types of mvuRight / mvDepth (np.float32 scalars under NumPy >= 2, Frame.py:269-278),
  * caller-supplied output tensors are validated before raw pointers are taken from them."""
import os

import numpy as np
import pytest

import oracle as O
from pyorbslam_b200 import ORBextractor
from pyorbslam_b200.stereo import compute_stereo_matches
from pyorbslam_b200.synthetic import make_stereo_pair

pytestmark = pytest.mark.gpu


def _same(kg, dg, ko, do):
    assert kg.shape == ko.shape, (kg.shape, ko.shape)
    assert np.array_equal(kg.view(np.uint32), ko.view(np.uint32))
    assert dg.shape == do.shape and np.array_equal(dg, do)


def test_batch_octree_global_node_scratch_many_pairs_vs_oracle():
    """nfeatures = 30000 over 4 levels: the level-0 quota (9.6 k nodes x 80 B) does not fit in shared memory, so k_octree keeps its
    node arrays in the global scratch block [slot][level].  Five pairs per launch: every slot must use its own block
    (ORBextractor.cpp:539-762 per image; round-1 indexed the blocks with the wrong grid dimension)."""
    import torch
    from pyorbslam_b200 import StereoFrontend
    params = (30000, 1.2, 4, 20, 7)
    H, W, n = 300, 620, 5
    rng = np.random.default_rng(5)
    pairs = [make_stereo_pair(40 + i, H, W) for i in range(n - 1)]
    noise = rng.integers(0, 256, (H, W), dtype=np.uint8)          # far more corners than the textured scenes: deep tree
    pairs.append((noise, np.roll(noise, -7, axis=1)))
    fe = StereoFrontend(*params, H, W, n)
    out = fe.run(torch.from_numpy(np.stack([p[0] for p in pairs])).cuda(), torch.from_numpy(np.stack([p[1] for p in pairs])).cuda(), 100.0, 500.0)
    fe.check_status(n)
    nkp = out["nkp"].cpu().numpy()
    kps, desc = out["kps"].cpu().numpy(), out["desc"].cpu().numpy()
    single = ORBextractor(*params)
    for i in range(n):
        for side in (0, 1):
            ko, do = O.OracleExtractor(*params).extract_arrays(pairs[i][side])
            m = int(nkp[side, i])
            _same(kps[side, i, :m], desc[side, i, :m], ko, do)
        ks, ds = single.extract_arrays(pairs[i][0])                  # the single-image object (S = 1) agrees as well
        _same(ks, ds, kps[0, i, :int(nkp[0, i])], desc[0, i, :int(nkp[0, i])])
    assert nkp[0, n - 1] > 5000                                       # the noise pair really is a deep tree


RANGE_PARAMS = (1000, 2.6, 3, 20, 7)       # scale 2.6: a right keypoint one octave above the left one can sit < 10 px from the
RANGE_SEEDS = (8, 9, 11)                   # level's left edge, so the 21-px SAD strip leaves the view; seed 9 does (found with the oracle)


def test_batch_range_error_is_reported_per_pair_like_the_single_frame_api():
    import torch
    from pyorbslam_b200 import StereoFrontend
    from pyorbslam_b200.stereo import stereo_resident
    H, W = 240, 640
    pairs = [make_stereo_pair(s, H, W) for s in RANGE_SEEDS]
    # oracle and single-frame API: pair 1 raises, pairs 0 and 2 do not
    expect = []
    for L, R in pairs:
        oL, oR = O.OracleExtractor(*RANGE_PARAMS), O.OracleExtractor(*RANGE_PARAMS)
        kL, dL = oL.extract_arrays(L)
        kR, dR = oR.extract_arrays(R)
        try:
            ou, od, oi, _ = O.stereo(kL[:, [0, 1, 5]], dL, kR[:, [0, 1, 5]], dR, oL.sf, oL.isf, oL.GetImagePyramid(), oR.GetImagePyramid(), 100.0, 500.0)
            expect.append((False, ou))
        except IndexError:
            expect.append((True, None))
        gL, gR = ORBextractor(*RANGE_PARAMS), ORBextractor(*RANGE_PARAMS)
        gL.extract_arrays(L)
        gR.extract_arrays(R)
        if expect[-1][0]:
            with pytest.raises(IndexError):
                stereo_resident(gL, gR, 100.0, 500.0)
        else:
            stereo_resident(gL, gR, 100.0, 500.0)
    assert [e[0] for e in expect] == [False, True, False]
    left = torch.from_numpy(np.stack([p[0] for p in pairs]))
    right = torch.from_numpy(np.stack([p[1] for p in pairs]))
    fe = StereoFrontend(*RANGE_PARAMS, H, W, 2)                       # 2-pair engine: chunks (0, 1) and (2)
    with pytest.raises(IndexError):
        fe.run_host(left, right, 100.0, 500.0)
    assert (fe.last_pair_status != 0).tolist() == [False, True, False]
    host = fe.last_out                                                 # every output did reach host memory
    for i in (0, 2):
        n = int(host["nkp"][0, i])
        assert np.array_equal(host["uRight"][i, :n].numpy().view(np.uint32), expect[i][1].view(np.uint32))
    # the flags are cleared per call: a clean job after a flagged one succeeds
    ok = fe.run_host(left[[0, 2]].contiguous(), right[[0, 2]].contiguous(), 100.0, 500.0)
    assert int(fe.last_pair_status.sum()) == 0 and int(ok["nkp"][0, 1]) == int(host["nkp"][0, 2])
    # device API: asynchronous, the caller asks
    fe3 = StereoFrontend(*RANGE_PARAMS, H, W, 3)
    fe3.run(left.cuda(), right.cuda(), 100.0, 500.0)
    with pytest.raises(IndexError):
        fe3.check_status(3)
    assert (fe3.last_pair_status != 0).tolist() == [False, True, False]
    fe3.run(left[[0, 2]].contiguous().cuda(), right[[0, 2]].contiguous().cuda(), 100.0, 500.0)
    assert fe3.check_status(2).sum() == 0


def test_mvuright_mvdepth_element_types_match_the_reference(golden_dir):
    """Frame.py:269-278 under NumPy >= 2 leaves np.float32 scalars for matches and the int -1 elsewhere (golden: types recorded
    from the reference's own Frame by tests/golden/make_golden.py)."""
    from test_gpu_parity import _FakeFrame
    g = np.load(os.path.join(golden_dir, "stereo_small.npz"))
    L, R = make_stereo_pair(int(g["idx"]), int(g["H"]), int(g["W"]))
    params = (1000, 1.2, 6, 20, 7)
    f = _FakeFrame(L, R, ORBextractor(*params), ORBextractor(*params), float(g["mbf"]), float(g["fx"]))
    compute_stereo_matches(f)
    codes = {int: 0, np.float32: 1, float: 2}
    tu = np.array([codes[type(v)] for v in f.mvuRight], np.int8)
    td = np.array([codes[type(v)] for v in f.mvDepth], np.int8)
    assert np.array_equal(tu, g["uRight_type"]) and np.array_equal(td, g["depth_type"])
    assert (tu == 1).sum() > 100 and (tu == 0).sum() > 0
    i = int(np.argmax(tu == 1))
    assert type(f.mvuRight[i]) is np.float32 and type(f.mvDepth[i]) is np.float32
    # what the type buys: a downstream expression such as abs(ur - mvuRight[i]) (ORBMatcher.py:356-358) stays float32
    assert type(np.float32(3.0) - f.mvuRight[i]) is np.float32


def test_caller_supplied_outputs_are_validated():
    import torch
    from pyorbslam_b200 import StereoFrontend
    fe = StereoFrontend(500, 1.2, 4, 20, 7, 160, 320, 2)
    L, R = make_stereo_pair(9, 160, 320)
    left, right = torch.from_numpy(np.stack([L, L])).cuda(), torch.from_numpy(np.stack([R, R])).cuda()
    good = fe.alloc_outputs(2)
    fe.run(left, right, 100.0, 300.0, out=good)
    small = fe.alloc_outputs(1)
    with pytest.raises(ValueError):
        fe.run(left, right, 100.0, 300.0, out=small)                 # undersized: would be written out of bounds
    bad = dict(good)
    bad["uRight"] = good["uRight"].double()
    with pytest.raises(ValueError):
        fe.run(left, right, 100.0, 300.0, out=bad)
    with pytest.raises(ValueError):
        fe.run(left, right, 100.0, 300.0, out=fe.alloc_outputs(2, pinned_host=True))    # host tensors for the device API
    with pytest.raises(ValueError):
        fe.run_host(left.cpu(), right.cpu(), 100.0, 300.0, out=good)                      # device tensors for the host API
    del bad["nkp"]
    with pytest.raises(ValueError):
        fe.run(left, right, 100.0, 300.0, out=bad)


def test_multi_gpu_runner_equals_single_engine():
    """StereoFrontendMulti (one engine + host thread per device, frames sharded by contiguous ranges, all writing one set of
    result arrays -- SURVEY.md 8e) against one engine's run_host.  On a one-GPU box both engines live on device 0, which still
    exercises the sharding, the shard offsets of the C ABI and the concurrent host threads; with two GPUs it uses both."""
    import torch
    from pyorbslam_b200 import StereoFrontend, StereoFrontendMulti, _lib
    from pyorbslam_b200.synthetic import make_kitti_like_pair
    params = (800, 1.2, 5, 20, 7)
    H, W, n = 200, 480, 7
    pairs = [make_kitti_like_pair(60 + i, H, W) for i in range(n)]
    left = torch.from_numpy(np.stack([p[0] for p in pairs])).pin_memory()
    right = torch.from_numpy(np.stack([p[1] for p in pairs])).pin_memory()
    one = StereoFrontend(*params, H, W, 3).run_host(left, right, 120.0, 400.0)
    devs = [0, 1] if _lib.device_count() >= 2 else [0, 0]
    for devices in (devs, [0, 0, 0]):
        multi = StereoFrontendMulti(*params, H, W, 2, devices=devices)
        assert [s[1:] for s in multi.shards(n)] == [tuple(map(int, (-(-r * n // len(devices)), min(-(-(r + 1) * n // len(devices)), n)))) for r in range(len(devices))]
        got = multi.run_host(left, right, 120.0, 400.0)
        assert torch.equal(got["nkp"], one["nkp"]) and int(got["nkp"].min()) > 100
        for i in range(n):
            nl, nr = int(one["nkp"][0, i]), int(one["nkp"][1, i])
            assert torch.equal(got["kps"][0, i, :nl], one["kps"][0, i, :nl]) and torch.equal(got["kps"][1, i, :nr], one["kps"][1, i, :nr])
            assert torch.equal(got["desc"][0, i, :nl], one["desc"][0, i, :nl]) and torch.equal(got["desc"][1, i, :nr], one["desc"][1, i, :nr])
            for key in ("uRight", "depth", "matchIdx"):
                assert torch.equal(got[key][i, :nl], one[key][i, :nl])
        assert int(multi.last_pair_status.sum()) == 0
        multi.close()


def test_fast_row_band_path_vs_oracle(monkeypatch):
    """k_fast_cells keeps a work list of half a cell's pixels; a cell with more passing pixels is walked in bands of rows (all bands
    scored first, then listed again, cut at the threshold, suppressed and emitted in row order).  B200ORB_FAST_LC (read when an
    extractor plans its geometry) shrinks the list to 64 entries so that ordinary cells take that path as well: same keypoints and
    descriptors as the oracle on a textured scene, on noise (every cell dense) and on a low-contrast scene (minThFAST retries)."""
    from pyorbslam_b200.synthetic import make_kitti_like_pair
    monkeypatch.setenv("B200ORB_FAST_LC", "64")
    rng = np.random.default_rng(17)
    H, W = 230, 410
    scene = make_kitti_like_pair(5, H, W)[0]
    imgs = [scene, rng.integers(0, 256, (H, W), dtype=np.uint8), (scene // 4 + 90).astype(np.uint8)]
    params = (1500, 1.2, 5, 20, 7)
    for img in imgs:
        ko, do = O.OracleExtractor(*params).extract_arrays(img)
        kg, dg = ORBextractor(*params).extract_arrays(img)
        _same(kg, dg, ko, do)
    monkeypatch.delenv("B200ORB_FAST_LC")
    kg, dg = ORBextractor(*params).extract_arrays(imgs[1])          # and the default capacity on the dense image
    ko, do = O.OracleExtractor(*params).extract_arrays(imgs[1])
    _same(kg, dg, ko, do)


@pytest.mark.parametrize("lanes", ["1", "2"])
def test_run_host_lanes_and_half_chunks_equal_device_path(monkeypatch, lanes):
    """run_host alternates chunks between two compute lanes (own stream + workspace) and, for engines of 16 pairs or more, cuts the
    job into half-capacity chunks with a ramp at both ends; B200ORB_HOST_LANES=1 keeps one lane and full chunks.  Either way every
    output equals what the device API gives chunk by chunk (70 pairs through a 16-pair engine: ramp 2, 4 | 8 ... | 4, 2)."""
    import torch
    from pyorbslam_b200 import StereoFrontend
    from pyorbslam_b200.synthetic import make_kitti_like_pair
    monkeypatch.setenv("B200ORB_HOST_LANES", lanes)
    params = (400, 1.2, 4, 20, 7)
    H, W, n, P = 120, 320, 70, 16
    base = [make_kitti_like_pair(80 + i, H, W) for i in range(7)]
    left = torch.from_numpy(np.stack([np.roll(base[i % 7][0], 5 * (i // 7), axis=1) for i in range(n)])).pin_memory()
    right = torch.from_numpy(np.stack([np.roll(base[i % 7][1], 5 * (i // 7), axis=1) for i in range(n)])).pin_memory()
    fe = StereoFrontend(*params, H, W, P)
    host = fe.run_host(left, right, 90.0, 300.0)
    ref = StereoFrontend(*params, H, W, P)
    for c in range(0, n, P):
        m = min(P, n - c)
        dev = ref.run(left[c:c + m].cuda(), right[c:c + m].cuda(), 90.0, 300.0)
        ref.check_status(m)
        dev = {k: v.cpu() for k, v in dev.items()}
        assert torch.equal(host["nkp"][:, c:c + m], dev["nkp"][:, :m]), c
        for i in range(m):
            nl, nr = int(dev["nkp"][0, i]), int(dev["nkp"][1, i])
            for side, k in ((0, nl), (1, nr)):
                assert torch.equal(host["kps"][side, c + i, :k], dev["kps"][side, i, :k]) and torch.equal(host["desc"][side, c + i, :k], dev["desc"][side, i, :k]), (c, i)
            for key in ("uRight", "depth", "matchIdx"):
                assert torch.equal(host[key][c + i, :nl], dev[key][i, :nl]), (key, c, i)
    assert int(host["nkp"].min()) > 50


def test_extractors_of_different_geometry_interleave():
    """The dynamic shared-memory ceiling of a kernel is a per-device attribute, not a per-object one: an extractor with a small
    octree / FAST footprint planned AFTER a large one must not break the large one's next launch (the attribute used to be set to
    each plan's own need).  Also: objects keep their own tensor maps, tables and workspaces when used alternately."""
    big_p, small_p = (30000, 1.2, 4, 20, 7), (300, 1.5, 3, 20, 7)
    rng = np.random.default_rng(23)
    noise = rng.integers(0, 256, (300, 620), dtype=np.uint8)
    scene = make_stereo_pair(12, 150, 260)[0]
    big, small = ORBextractor(*big_p), ORBextractor(*small_p)
    for rnd in range(2):
        a = np.roll(noise, 3 * rnd, axis=1)
        b = np.roll(scene, 2 * rnd, axis=0)
        _same(*big.extract_arrays(a), *O.OracleExtractor(*big_p).extract_arrays(a))
        _same(*small.extract_arrays(b), *O.OracleExtractor(*small_p).extract_arrays(b))
