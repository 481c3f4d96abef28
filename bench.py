#!/usr/bin/env python
"""bench.py -- stereo frames/s of the B200 stereo ORB front-end (extract L + extract R + stereo match).

  python bench.py [--gpus N --steps K --warmup W]            our arm (one process per GPU under torchrun for N > 1)
  python bench.py --impl reference [--steps K --warmup W]    the reference's CPU implementation on the host cores

Workload = BASELINE.json configs[2]: a batch of synthetic KITTI-shape stereo pairs (1241x376, 2000 features,
8 levels, 1.2, FAST 20/7), `--pairs` pairs per GPU per step (weak scaling: frames are independent, every rank
owns its own batch, no data-path collective).  One step = one pass of the hot path over that batch.
Scenes: `--scenes kitti_like` (default; road scenes tuned to the corner / retry / match statistics of the one real KITTI
frame the reference ships) or `layered` (round 1's harsher generator: more FAST candidates, heavy occlusion).
  value : frames/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : frames/s through the host-buffer C-ABI call (pinned H2D of the images + D2H of every result inside the timing)
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORB = dict(nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7)   # configs/KITTI00-02.yaml:36-50
H, W = 376, 1241
MBF, FX = 386.1448, 718.856                                                       # configs/KITTI00-02.yaml:7,24
METRIC, UNIT = "stereo_frames_per_sec", "frames/s"
STREAMS = 1
WORKLOAD_NAME = "BASELINE.json configs[2]: batch of synthetic KITTI00-02-shape stereo pairs, extract L+R + compute_stereo_matches"


# ------------------------------------------------------------------------------------------------ geometry / bytes
WORKLOADS = {
    # BASELINE.json configs[2] (the metric's configuration): KITTI00-02 camera and ORB settings
    "kitti": dict(H=376, W=1241, orb=dict(nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7), pairs=4096, chunk=128,
                  name="BASELINE.json configs[2]: batch of synthetic KITTI00-02-shape stereo pairs, extract L+R + compute_stereo_matches"),
    # BASELINE.json configs[3]: bandwidth stress (not the headline metric; run with --no-cpu-baseline, the CPU path needs ~10 s per frame)
    "hires": dict(H=1440, W=2560, orb=dict(nfeatures=8000, scaleFactor=1.2, nlevels=12, iniThFAST=20, minThFAST=7), pairs=256, chunk=64,
                  name="BASELINE.json configs[3]: synthetic 2560x1440 stereo pairs, 8000 features, 12 levels"),
}


def level_sizes(Hh=None, Ww=None, nlevels=None, scale=None):
    Hh, Ww = Hh or H, Ww or W
    nlevels, scale = nlevels or ORB["nlevels"], scale or ORB["scaleFactor"]
    s, out = np.float32(1.0), []
    for l in range(nlevels):
        if l:
            s = np.float32(np.float64(s) * np.float64(np.float32(scale)))
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(Ww) * inv)), int(np.rint(np.float32(Hh) * inv))))
    return out


def algorithmic_bytes(K=2000):
    """bytes each kernel must move per IMAGE (per PAIR for stereo), and B_frame."""
    lv = level_sizes()
    S0 = lv[0][0] * lv[0][1]
    s = [w * h for w, h in lv]
    b = [(w + 38) * (h + 38) for w, h in lv]
    S, Sb = sum(s), sum(b)
    per_image = {
        "border0": S0 + b[0],
        "resize_chain": (S - s[-1]) + (Sb - b[0]),
        "blur": 2 * S,
        "fast_cells": S,
        "octree": None,                      # data dependent: 4 B per FAST candidate in, 4 B per keypoint out (filled at run time)
        "orient_describe": S + 56 * K,
    }
    stereo_per_pair = 2 * 56 * K + 352 * K + 8 * K
    B_img = S0 + Sb + (S - s[-1]) + S + 2 * S + S + 56 * K
    return per_image, stereo_per_pair, 2 * B_img + stereo_per_pair


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_worker(args):
    """One stereo frame the way the reference computes it: ORBextractor.operator_kd x2 (+ tuple -> cv2.KeyPoint lists,
    Frame.py:114-121) + GetImagePyramid x2 (Frame.py:59-60) + compute_stereo_matches (Frame.py:161-279)."""
    L, R, use_ref = args
    import oracle as O
    from oracle import refext, stereo_py
    global _W_EXT
    if "_W_EXT" not in globals():
        p = (ORB["nfeatures"], ORB["scaleFactor"], ORB["nlevels"], ORB["iniThFAST"], ORB["minThFAST"])
        mk = (lambda: refext.RefExtractor(*p, kind="shipped")) if use_ref else (lambda: O.OracleExtractor(*p))
        _W_EXT = (mk(), mk())
    eL, eR = _W_EXT
    t0 = time.perf_counter()
    tl, dl = eL.operator_kd(L)
    tr, dr = eR.operator_kd(R)
    try:
        import cv2
        kl = [cv2.KeyPoint(*k) for k in tl]
        kr = [cv2.KeyPoint(*k) for k in tr]
        keysL = [(k.pt[0], k.pt[1], k.octave) for k in kl]
        keysR = [(k.pt[0], k.pt[1], k.octave) for k in kr]
    except ImportError:
        keysL = [(k[0], k[1], k[5]) for k in tl]
        keysR = [(k[0], k[1], k[5]) for k in tr]
    pl, pr = eL.GetImagePyramid(), eR.GetImagePyramid()
    t1 = time.perf_counter()
    u, d = stereo_py.stereo_matches(keysL, dl, keysR, dr, eL.GetScaleFactors(), eL.GetInverseScaleFactors(), pl, pr, MBF, np.float32(FX))
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, sum(1 for v in u if v != -1)


class CpuReference:
    """The reference CPU path on the host cores.  Extractor = the reference's own ORBextractor.cpp compiled unmodified
    (oracle/_ref/libref_shipped.so, -O3) when present, else the oracle port; stereo = Python restatement of
    Frame.compute_stereo_matches with the reference's per-keypoint Python structure (oracle/stereo_py.py)."""

    def __init__(self, workers=None):
        from oracle import refext
        import oracle as O
        O.build()
        self.use_ref = refext.available("shipped")
        if self.use_ref:
            refext._lib("shipped")      # map the compiled reference into THIS process too (the forked workers inherit the mapping),
                                        # so a loaded-library record of the arm shows oracle/_ref/libref_shipped.so
        self.workers = workers or max(1, min(os.cpu_count() or 1, 32))
        self.pool = mp.get_context("fork").Pool(self.workers)
        self.pairs = [make_pair(i, H, W) for i in range(min(self.workers, 8))]   # a few distinct frames, reused round-robin

    def step(self, n_pairs):
        work = [(self.pairs[i % len(self.pairs)][0], self.pairs[i % len(self.pairs)][1], self.use_ref) for i in range(n_pairs)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, work, chunksize=1)
        dt = time.perf_counter() - t0
        return dt, res

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self, n_pairs, res):
        te = float(np.mean([r[0] for r in res]))
        ts = float(np.mean([r[1] for r in res]))
        return (f"kind=port because compute_stereo_matches is Python and cannot travel (restated in oracle/stereo_py.py); the extractor half "
                f"is {'the reference ORBextractor.cpp itself, compiled unmodified' if self.use_ref else 'the oracle port'}. "
                f"{n_pairs} synthetic {W}x{H} pairs per step over {self.workers} worker processes; per frame on one core: "
                f"extract L+R + KeyPoint lists + pyramids {te*1e3:.0f} ms "
                f"({'reference ORBextractor.cpp compiled -O3 against oracle/cvshim scalar primitives' if self.use_ref else 'oracle port'}), "
                f"compute_stereo_matches {ts*1e3:.0f} ms (Python restatement, reference structure)")


SCENES = "kitti_like"


def make_pair(idx, Hh, Ww):
    from pyorbslam_b200 import synthetic
    return (synthetic.make_kitti_like_pair if SCENES == "kitti_like" else synthetic.make_stereo_pair)(idx, Hh, Ww)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference()
    n = 2 * ref.workers
    for _ in range(args.warmup):
        ref.step(min(n, ref.workers))
    total_t, res = 0.0, []
    for _ in range(args.steps):
        dt, r = ref.step(n)
        total_t += dt
        res = r
    ref.close()
    fps = n * args.steps / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(n, None),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": ref.workers, "kind": "port",
                         "sample": ref.describe(n, res)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(pairs, chunk):
    return {"workload": WORKLOAD_NAME, "scenes": SCENES,
            "image": [H, W], **ORB, "bf": MBF, "fx": FX, "pairs_per_gpu_per_step": pairs, "chunk_pairs": chunk, "concurrent_streams": STREAMS,
            "l2": "inputs of one step exceed the 126 MB L2 (0.93 MB per pair)", "parallelism": "frames sharded per GPU, no collective"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, so the pinned staging buffers are first-touched
    on that NUMA node (matters once several ranks stream 30+ GB/s each through the host)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception as e:      # affinity is an optimisation only
        return f"unavailable: {e}"


# ------------------------------------------------------------------------------------------------ our arm
def run_b200_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (rank 0 of a 1-GPU run only), before this process creates a CUDA context
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = CpuReference()
        n = 2 * ref.workers
        ref.step(ref.workers)                      # warm the pool / page in the libraries
        dt, res = ref.step(n)
        ref.close()
        cpu_baseline = {"value": n / dt, "unit": UNIT, "cores": ref.workers, "kind": "port",
                        "sample": ref.describe(n, res),
                        "one_core_frames_per_sec": 1.0 / float(np.mean([r[0] + r[1] for r in res]))}

    import torch
    import torch.distributed as dist
    from pyorbslam_b200 import StereoFrontend, _lib

    if _lib.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: pyorbslam_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)    # pinned host buffers should live on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)     # fd 1 already points at stderr (claim_stdout), so NCCL_DEBUG lines cannot reach stdout

    B, P = args.pairs, args.chunk
    nb = args.base_pairs
    # synthetic batch: `nb` distinct scenes per rank, frame i = scene (i % nb) rolled horizontally by 9 * (i // nb) px
    # in BOTH views (disparities unchanged; every frame lands differently on the FAST cell grid)
    base = [make_pair(1000 * rank + i, H, W) for i in range(nb)]
    bl = torch.from_numpy(np.stack([p[0] for p in base])).to(dev)
    br = torch.from_numpy(np.stack([p[1] for p in base])).to(dev)
    left = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    right = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    for g in range(0, B, nb):
        m = min(nb, B - g)
        left[g:g + m] = torch.roll(bl[:m], shifts=9 * (g // nb), dims=2)
        right[g:g + m] = torch.roll(br[:m], shifts=9 * (g // nb), dims=2)
    del bl, br

    # `--streams` front-ends (each with its own workspace) run consecutive chunks concurrently on their own CUDA
    # streams, so one chunk's latency-bound kernels (octree, small pyramid levels) fill the SMs the other leaves idle
    NS = max(1, args.streams)
    fes = [StereoFrontend(ORB["nfeatures"], ORB["scaleFactor"], ORB["nlevels"], ORB["iniThFAST"], ORB["minThFAST"], H, W, P, device=local)
           for _ in range(NS)]
    fe = fes[0]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NS)] if NS > 1 else [torch.cuda.current_stream(dev)]
    chunks = [(c, min(P, B - c)) for c in range(0, B, P)]
    outs = [fe.alloc_outputs(n) for _, n in chunks]

    def step():
        if NS == 1:
            for (c, n), o in zip(chunks, outs):
                fe.run(left[c:c + n], right[c:c + n], MBF, FX, out=o)
            return
        main = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(main)
        for k, ((c, n), o) in enumerate(zip(chunks, outs)):
            with torch.cuda.stream(streams[k % NS]):
                fes[k % NS].run(left[c:c + n], right[c:c + n], MBF, FX, out=o)
        for s in streams:
            main.wait_stream(s)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # ---- the timed region: EXACTLY `steps` steps of the production schedule, no profiling events between the kernels ----
    clocks = ClockSampler(local)
    l0 = _lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.kernel_launches() - l0
    clk = clocks.stop()
    # ---- per-stage pass (not part of `value`): the same chunks again, one at a time on ONE stream, with CUDA events recorded between
    # the stages on the launching stream -- with `--streams` > 1 the timed region above overlaps consecutive chunks, so the stage
    # times below add up to more than ms_per_step / chunks ----
    fe.profile(True, max_calls=args.steps * len(chunks))
    for _ in range(args.steps):
        for (c, n), o in zip(chunks, outs):
            fe.run(left[c:c + n], right[c:c + n], MBF, FX, out=o)
    barrier()
    stage_ms, prof_calls, prof_pairs = {}, 0, 0
    for f in fes[:1]:
        sm_, pc_, pp_ = f.profile_read()
        f.profile(False)
        for k_, v_ in sm_.items():
            stage_ms[k_] = stage_ms.get(k_, 0.0) + v_
        prof_calls += pc_
        prof_pairs += pp_
    last_fe = fe
    ncand = last_fe.candidate_count(2 * chunks[-1][1]) / (2 * chunks[-1][1])
    nkp = float(torch.cat([o["nkp"].float().flatten() for o in outs]).mean())
    valid = [torch.arange(fe.capacity, device=dev)[None, :] < o["nkp"][0][:, None] for o in outs]
    matched = float(sum(int(((o["uRight"] >= 0) & v).sum()) for o, v in zip(outs, valid))) / B
    ham_ok = float(sum(int(((o["matchIdx"] >= 0) & v).sum()) for o, v in zip(outs, valid))) / B
    # the same first chunk under the opt-in upstream-ORB-SLAM2 view (true level images instead of the reference's sheared caster
    # view): how many matches the scenes give when the SAD windows are not sheared -- a scene statistic, not a result
    fe.set_stereo_options(dense_pyramid=True)
    c0, n0 = chunks[0]
    od = fe.run(left[c0:c0 + n0], right[c0:c0 + n0], MBF, FX)
    matched_dense = float(int(((od["uRight"] >= 0) & valid[0]).sum())) / n0
    fe.set_stereo_options()
    del od

    # ---- checker leg: a few frames of the timed batch itself against the CPU oracle (bit-exact keypoints, descriptors, matches) ----
    oracle_check = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O                    # the oracle is the checker here, never the thing measured
        prm = (ORB["nfeatures"], ORB["scaleFactor"], ORB["nlevels"], ORB["iniThFAST"], ORB["minThFAST"])
        picked, same_all = [0, B // 2 + 3, B - 1], True
        for i in picked:
            ci, k = i // P, i % P
            o = outs[ci]
            Li, Ri = left[i].cpu().numpy(), right[i].cpu().numpy()
            oL, oR = O.OracleExtractor(*prm), O.OracleExtractor(*prm)
            kL, dL = oL.extract_arrays(Li)
            kR, dR = oR.extract_arrays(Ri)
            nl, nr = int(o["nkp"][0, k]), int(o["nkp"][1, k])
            ok = (nl, nr) == (len(kL), len(kR))
            if ok:
                ok = (np.array_equal(o["kps"][0, k, :nl].cpu().numpy().view(np.uint32), kL.view(np.uint32)) and np.array_equal(o["desc"][0, k, :nl].cpu().numpy(), dL)
                      and np.array_equal(o["kps"][1, k, :nr].cpu().numpy().view(np.uint32), kR.view(np.uint32)) and np.array_equal(o["desc"][1, k, :nr].cpu().numpy(), dR))
            if ok:
                ou, od, oi, _ = O.stereo(kL[:, [0, 1, 5]], dL, kR[:, [0, 1, 5]], dR, oL.sf, oL.isf, oL.GetImagePyramid(), oR.GetImagePyramid(), MBF, FX)
                ok = (np.array_equal(o["uRight"][k, :nl].cpu().numpy().view(np.uint32), ou.view(np.uint32)) and np.array_equal(o["depth"][k, :nl].cpu().numpy().view(np.uint32), od.view(np.uint32))
                      and np.array_equal(o["matchIdx"][k, :nl].cpu().numpy(), oi))
            same_all = same_all and bool(ok)
        oracle_check = {"frames_of_the_timed_batch": picked, "bit_exact_vs_oracle": same_all}

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t)
    value = world * B * args.steps / (ms_max / 1e3)

    # ---- end to end through the host-buffer C-ABI call ----
    Be = min(B, args.e2e_pairs)
    lh = torch.empty((Be, H, W), dtype=torch.uint8, pin_memory=True)
    rh = torch.empty((Be, H, W), dtype=torch.uint8, pin_memory=True)
    lh.copy_(left[:Be])
    rh.copy_(right[:Be])
    oh = fe.alloc_outputs(Be, pinned_host=True)
    for _ in range(2):
        fe.run_host(lh, rh, MBF, FX, out=oh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fe.run_host(lh, rh, MBF, FX, out=oh)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * args.steps / float(te)
    d2h = sum(v.numel() * v.element_size() for v in oh.values())
    # the e2e result must equal the device-resident one (same frames): every output array of the first chunk, not just the counts
    n0c = chunks[0][1]
    same = bool((oh["nkp"][:, :n0c] == outs[0]["nkp"].cpu()).all())
    valid0 = (torch.arange(fe.capacity)[None, :] < oh["nkp"][0, :n0c][:, None])
    for key in ("uRight", "depth", "matchIdx"):
        same = same and bool(((oh[key][:n0c] == outs[0][key].cpu()) | ~valid0).all())
    for side in (0, 1):
        vs = (torch.arange(fe.capacity)[None, :] < oh["nkp"][side, :n0c][:, None])[..., None]
        same = same and bool(((oh["kps"][side, :n0c] == outs[0]["kps"][side].cpu()) | ~vs).all())
        same = same and bool(((oh["desc"][side, :n0c] == outs[0]["desc"][side].cpu()) | ~vs).all())

    # ---- the host-copy ceiling of this box: the same bytes per chunk (pinned H2D of both views + D2H of every output array) with
    # no kernel at all, uploads and downloads on their own streams, all ranks at once.  e2e cannot exceed it; e2e / ceiling says how
    # much of what the host <-> device links deliver the pipeline uses ----
    fe.set_copy_only(True)                       # the very same C-ABI call: same copies, streams, events and buffers, no kernels
    try:
        fe.run_host(lh, rh, MBF, FX, out=oh)
        barrier()
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fe.run_host(lh, rh, MBF, FX, out=oh)
            torch.cuda.synchronize()
            tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            best = float(tc) if best is None else min(best, float(tc))
            barrier()
    finally:
        fe.set_copy_only(False)
    tc = best
    ceiling_value = world * Be * args.steps / float(tc)

    # ---- latency of the reference-compatible single-frame API (config 2: what one Frame.__init__ costs) ----
    dropin = None
    if rank == 0:
        from pyorbslam_b200 import ORBextractor
        from pyorbslam_b200.stereo import stereo_resident
        prm = (ORB["nfeatures"], ORB["scaleFactor"], ORB["nlevels"], ORB["iniThFAST"], ORB["minThFAST"])
        eL, eR = (ORBextractor(*prm, device=local, reuse_identical_input=False) for _ in range(2))   # time real extractions
        L0, R0 = left[0].cpu().numpy(), right[0].cpu().numpy()

        def one_frame():
            kl, dl = eL.operator_kd(L0)                       # Frame.ExtractORB(0) incl. the 6-tuple list
            kr, dr = eR.operator_kd(R0)
            pl, pr = eL.GetImagePyramid(), eR.GetImagePyramid()   # Frame.py:59-60
            return stereo_resident(eL, eR, MBF, FX)           # Frame.compute_stereo_matches

        def one_frame_arrays():                               # the same C-ABI calls without building 2 x 2000 Python tuples
            eL.extract_arrays(L0)
            eR.extract_arrays(R0)
            eL.GetImagePyramid(), eR.GetImagePyramid()
            return stereo_resident(eL, eR, MBF, FX)

        def timed(fn, n=30):
            for _ in range(5):
                fn()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            return 1e3 * (time.perf_counter() - t0) / n
        dropin = {"ms_per_frame": timed(one_frame), "ms_per_frame_array_api": timed(one_frame_arrays),
                  "ms_extract_one_image_array_api": timed(lambda: eL.extract_arrays(L0)),
                  "ms_tuple_list_one_image": timed(lambda: eL.operator_kd(L0)) - timed(lambda: eL.extract_arrays(L0)),
                  "what": "ms_per_frame = ORBextractor.operator_kd x2 (tuple lists) + GetImagePyramid x2 + compute_stereo_matches through the drop-in "
                          "Python API, one pair at a time (CUDA graph per image, results in one pinned copy, pyramid downloaded in the background); "
                          "_array_api = the same calls returning arrays (no 2 x 2000 Python tuples, which cost more than the GPU work)"}

    # ---- the other BASELINE.json configurations (not the headline metric; recorded so that the driver's line carries them) ----
    other_configs = None
    if rank == 0 and world == 1 and not args.skip_other_configs:
        other_configs = measure_other_configs(local, dev, not args.no_cpu_baseline)

    # ---- BoW transform of one frame's descriptors (Frame.compute_BoW), GPU vs the Python restatement ----
    bow_extra = None
    if rank == 0 and not args.no_cpu_baseline:
        import types
        from oracle.bow_py import Vocabulary, make_vocab_text          # checker / CPU baseline leg only
        from pyorbslam_b200.bow import GpuVocabulary
        voc = Vocabulary.from_text(make_vocab_text(seed=3, k=10, L=4, p_early_leaf=0.0, p_zero_weight=0.0))
        model = types.SimpleNamespace(L=voc.L, nodes=[types.SimpleNamespace(children=voc.children[i], descriptor=None if i == 0 else voc.desc[i],
                                                                           weight=voc.weight[i], word_id=voc.word_id[i]) for i in range(len(voc.children))])
        gv = GpuVocabulary(model, device=local)
        _, desc0 = eL.operator_kd(L0)
        for _ in range(3):
            gv.transform(desc0, 4)
        t0 = time.perf_counter()
        for _ in range(20):
            gbv, gfv = gv.transform(desc0, 4)
        gpu_ms = 1e3 * (time.perf_counter() - t0) / 20
        ns = 200
        t0 = time.perf_counter()
        voc.transform(desc0[:ns], 4)
        cpu_ms = 1e3 * (time.perf_counter() - t0) * len(desc0) / ns
        obv, ofv = voc.transform(desc0[:ns], 4)
        sbv, sfv = gv.transform(desc0[:ns].copy(), 4)
        bow_extra = {"features": int(len(desc0)), "vocabulary": "synthetic k=10 L=4 (11 110 nodes)", "gpu_ms_per_frame": gpu_ms,
                     "cpu_python_ms_per_frame_extrapolated": cpu_ms, "cpu_sample_features": ns,
                     "identical_on_sample": list(obv.items()) == list(sbv.items()) and list(ofv.items()) == list(sfv.items())}

    # ---- the two projection searches (TrackWithMotionModel / SearchLocalPoints), GPU path vs the Python restatement ----
    proj_extra = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import matcher_py as MP                              # scene generator + checker / CPU baseline leg only
        from pyorbslam_b200.matcher import install_matcher

        class _Matcher:
            def __init__(self, r, o):
                self.mfNNratio, self.mbCheckOrientation = r, o
        install_matcher(_Matcher)
        mt = _Matcher(0.9, True)
        npts = 2000

        def uid(fr):
            return [-1 if p is None else p.uid for p in fr.mvpMapPoints]
        res = {}
        for name in ("f_f", "f_p"):
            def run(gpu):
                cur, last, loc = MP.make_projection_case(seed=5, n=npts)
                t0 = time.perf_counter()
                if name == "f_f":
                    n = mt.search_by_projection_f_f(cur, last, 15) if gpu else MP.projection_f_f(cur, last, 15, True)
                else:
                    n = mt.search_by_projection_f_p(cur, loc, 3.0) if gpu else MP.projection_f_p(cur, loc, 3.0, 0.9)
                return 1e3 * (time.perf_counter() - t0), n, uid(cur)
            run(True)
            g = min(run(True) for _ in range(3))
            c = run(False)
            res[name] = {"gpu_ms": g[0], "cpu_python_ms": c[0], "matches": int(g[1]), "identical": g[1] == c[1] and g[2] == c[2]}
        proj_extra = {"map_points": npts, "features": npts, "what": "one call of ORBMatcher.search_by_projection_f_f / _f_p on a synthetic "
                      "two-frame scene; CPU = Python restatement with a table popcount (the reference's bin().count distance is slower)", **res}

    if rank == 0:
        per_image, stereo_pp, B_frame = algorithmic_bytes(ORB["nfeatures"])
        per_image["octree"] = 4 * ncand + 4 * nkp
        peak, peak_src = measured_peak()
        kernels = {}
        launches_per_call = fe.stage_launches()
        for k, tot in stage_ms.items():
            bytes_total = (stereo_pp * prof_pairs) if k == "stereo" else (per_image[k] * 2 * prof_pairs)
            gbs = bytes_total / (tot / 1e3) / 1e9 if tot > 0 else 0.0
            kernels[k] = {"ms_total": tot, "share": tot / max(sum(stage_ms.values()), 1e-9), "ms_per_chunk": tot / max(prof_calls, 1),
                          "launches_per_chunk": launches_per_call[k], "avg_launch_ms": tot / max(prof_calls * launches_per_call[k], 1),
                          "algorithmic_bytes_per_launch": bytes_total / max(prof_calls * launches_per_call[k], 1), "achieved_gbs": gbs, "frac": gbs / peak}
        dom = max(stage_ms, key=stage_ms.get)
        try:       # static ncu figures of the same kernels (profiles/traffic.json, from the committed --set full capture)
            tj_all = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            for k in kernels:
                kernels[k]["ncu"] = {"dram_bytes_per_launch_group": tj_all.get(k), "issue_slots_busy_pct": tj_all.get("_issue_slots_busy_pct", {}).get(k),
                                     "alu_pipe_pct": tj_all.get("_alu_pipe_pct", {}).get(k), "l1_data_pipe_pct": tj_all.get("_l1_data_pipe_pct", {}).get(k)}
        except Exception:
            pass
        traffic, issue_pct, alu_pct = None, None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):       # static numbers from the committed ncu --set full capture of the same kernels
            try:
                tj = json.load(open(tp))
                traffic = tj.get(dom)
                issue_pct = tj.get("_issue_slots_busy_pct", {}).get(dom)
                alu_pct = tj.get("_alu_pipe_pct", {}).get(dom)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(B, P), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * Be * H * W, "d2h_bytes_per_step": d2h, "pairs_per_step": Be,
                    "matches_device_resident_result": same,
                    "copy_ceiling": {"value": ceiling_value, "unit": UNIT, "e2e_over_ceiling": e2e_value / ceiling_value,
                                     "h2d_gbs": ceiling_value * 2 * H * W / 1e9, "d2h_gbs": ceiling_value * d2h / Be / 1e9,
                                     "what": "the same pinned H2D + D2H bytes per chunk with no kernels, all ranks concurrently: what the "
                                             "host <-> device links of this box deliver for this traffic pattern"}},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                         "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                         "avg_launch_ms": kernels[dom]["avg_launch_ms"], "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"],
                         "note": "byte-granular integer kernel: ncu shows it bound by instruction issue / the integer ALU pipe (half rate), not by HBM; "
                                 "avg_launch_ms comes from a separate pass with CUDA events between the stages, after the timed region",
                         "issue_slots_busy_pct_ncu": issue_pct, "alu_pipe_busy_pct_ncu": alu_pct},
            "roofline_whole_path": {"B_frame_bytes": B_frame, "achieved": B_frame * (value / world) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": B_frame * (value / world) / 1e9 / peak},
            "kernels": kernels,
            "workload_stats": {"keypoints_per_image": nkp, "fast_candidates_per_image": ncand, "stereo_matches_per_pair": matched,
                               "hamming_accepted_per_pair": ham_ok, "stereo_matches_per_pair_true_level_images": matched_dense,
                               "note": "every Hamming-accepted keypoint runs the full 11-shift SAD slide (the kernel's work); the reference then "
                                       "rejects ~30 % of them because its pyramid view is sheared (SURVEY.md F6): the windows it compares come from "
                                       "38 px further along per row, where the disparity differs.  With true level images (opt-in flag) the same "
                                       "scenes match like upstream ORB-SLAM2 on KITTI",
                               "workspace_bytes": sum(f.workspace_bytes() for f in fes), "rank0_cpu_affinity": numa if isinstance(numa, str) else f"{len(numa)} cpus: {numa[0]}-{numa[-1]}"},
        }
        line["oracle_check"] = oracle_check
        line["dropin_single_frame_latency"] = dropin
        line["other_baseline_configs"] = other_configs
        line["bow_transform_8f_rank2"] = bow_extra
        line["projection_search_8f_rank1"] = proj_extra
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_other_configs(local, dev, with_cpu):
    """BASELINE.json configs[0], [3], [4] next to the headline configs[2]: timings only (parity of each is a -m gpu test)."""
    import torch
    from pyorbslam_b200 import ORBextractor, StereoFrontend
    from pyorbslam_b200.stereo import stereo_host
    from pyorbslam_b200.synthetic import make_kitti_like_pair
    out = {}

    def timed(fn, n, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / n

    # configs[0]: pyORBExtractor/test.py on the bundled kitti06-436.png (stored as raw gray pixels), 20 iterations like test.py:28-38
    fixture = os.path.join(ROOT, "tests", "golden", "kitti06-436.gray.npy")
    kitti_prm = (2000, 1.2, 8, 20, 7)
    if os.path.exists(fixture):
        img = np.load(fixture)
        e = ORBextractor(*kitti_prm, device=local, reuse_identical_input=False)
        n = len(e.extract_arrays(img)[0])
        out["config0_fixture_image"] = {"image": list(img.shape), "keypoints": n, "iterations": 20,
                                        "ms_per_iter_operator_kd": timed(lambda: e.operator_kd(img), 20),
                                        "ms_per_iter_arrays": timed(lambda: e.extract_arrays(img), 20),
                                        "what": "test.py's loop: operator_kd (6-tuple list + descriptors) per iteration; _arrays = the same call returning arrays"}
    # configs[3]: 2560x1440, 8000 features, 12 levels
    try:
        Hh, Ww, Ph, nbh, Bh = 1440, 2560, 16, 2, 64
        prm = (8000, 1.2, 12, 20, 7)
        base = [make_kitti_like_pair(5000 + i, Hh, Ww) for i in range(nbh)]
        fe = StereoFrontend(*prm, Hh, Ww, Ph, device=local)
        lh = torch.empty((Bh, Hh, Ww), dtype=torch.uint8, pin_memory=True)
        rh = torch.empty((Bh, Hh, Ww), dtype=torch.uint8, pin_memory=True)
        for g in range(0, Bh, nbh):
            for j in range(nbh):
                lh[g + j] = torch.from_numpy(np.roll(base[j][0], 9 * (g // nbh), axis=1))
                rh[g + j] = torch.from_numpy(np.roll(base[j][1], 9 * (g // nbh), axis=1))
        ld, rd = lh.to(dev), rh.to(dev)
        outs = [fe.alloc_outputs(Ph) for _ in range(0, Bh, Ph)]

        def step():
            for k, c in enumerate(range(0, Bh, Ph)):
                fe.run(ld[c:c + Ph], rd[c:c + Ph], MBF, FX, out=outs[k])
        ms = timed(step, 3)
        oh = fe.alloc_outputs(Bh, pinned_host=True)
        ms_e2e = timed(lambda: fe.run_host(lh, rh, MBF, FX, out=oh), 3, warm=1)
        nk = float(oh["nkp"].float().mean())
        out["config3_hires"] = {"image": [Hh, Ww], "nfeatures": 8000, "nlevels": 12, "pairs_per_step": Bh, "chunk_pairs": Ph,
                                "frames_per_sec_device_resident": Bh / (ms / 1e3), "frames_per_sec_e2e": Bh / (ms_e2e / 1e3),
                                "keypoints_per_image": nk, "scenes": "kitti_like"}
        del fe, ld, rd, lh, rh, oh, outs
        torch.cuda.empty_cache()
    except Exception as ex:       # a side measurement must not cost the headline line
        out["config3_hires"] = {"error": repr(ex)}
    # configs[4]: compute_stereo_matches only, 1k..16k keypoints per image (synthetic keypoints / descriptors on real pyramids)
    try:
        L, R = make_kitti_like_pair(12, H, W)
        eL, eR = ORBextractor(*kitti_prm, device=local), ORBextractor(*kitti_prm, device=local)
        eL.extract_arrays(L)
        eR.extract_arrays(R)
        pyrL, pyrR = eL.GetImagePyramid(), eR.GetImagePyramid()
        sf, isf = np.array(eL.GetScaleFactors(), np.float32), np.array(eL.GetInverseScaleFactors(), np.float32)
        quota = np.array(eL.features_per_level(), np.float64)
        sweep = []
        for n in (1000, 2000, 4000, 8000, 16000):
            rng = np.random.default_rng(n)
            octv = rng.choice(len(sf), size=n, p=quota / quota.sum())
            s_ = sf[octv]
            lx = rng.integers(19, (W / s_ - 20).astype(int)).astype(np.float32)
            ly = rng.integers(19, (H / s_ - 20).astype(int)).astype(np.float32)
            disp = rng.integers(1, 80, n).astype(np.float32)
            kL = np.stack([np.where(octv > 0, lx * s_, lx), np.where(octv > 0, ly * s_, ly), octv.astype(np.float32)], 1).astype(np.float32)
            rx = np.maximum(lx - np.floor(disp / s_), 19).astype(np.float32)
            kR = np.stack([np.where(octv > 0, rx * s_, rx), kL[:, 1], octv.astype(np.float32)], 1).astype(np.float32)
            dL = rng.integers(0, 256, (n, 32), dtype=np.uint8)
            dR = dL ^ np.packbits(rng.random((n, 256)) < 0.1, axis=1, bitorder="little")
            perm = rng.permutation(n)
            kR, dR = kR[perm], dR[perm]
            gu = stereo_host(kL, dL, kR, dR, sf, isf, pyrL, pyrR, MBF, FX, device=local)[0]
            row = {"keypoints": n, "gpu_ms": timed(lambda: stereo_host(kL, dL, kR, dR, sf, isf, pyrL, pyrR, MBF, FX, device=local), 5, warm=1),
                   "matches": int((gu >= 0).sum())}
            if with_cpu:          # CPU leg: the Python restatement on a bounded sample of the left keypoints (its cost is linear in them)
                from oracle import stereo_py
                m = min(n, 96)
                keysR = [(float(a), float(b), int(c)) for a, b, c in kR]
                keysL = [(float(a), float(b), int(c)) for a, b, c in kL[:m]]
                t0 = time.perf_counter()
                cu, _ = stereo_py.stereo_matches(keysL, dL[:m], keysR, dR, sf.tolist(), isf.tolist(), pyrL, pyrR, MBF, np.float32(FX))
                row["cpu_python_ms_extrapolated"] = 1e3 * (time.perf_counter() - t0) * n / m
                row["cpu_sample_left_keypoints"] = m
                row["identical_on_sample"] = bool(np.array_equal(np.array([float(v) for v in cu], np.float32), gu[:m]))
            sweep.append(row)
        out["config4_stereo_only_sweep"] = {"what": "b200orb_stereo_host (uploads the caller's keypoints, descriptors and pyramid views, K7 + K8, downloads) vs the "
                                            "Python restatement of Frame.compute_stereo_matches (row index built over all right keypoints, then a bounded "
                                            "sample of left keypoints, extrapolated linearly)", "rows": sweep}
    except Exception as ex:
        out["config4_stereo_only_sweep"] = {"error": repr(ex)}
    return out


def run_multi_runner(args):
    """`--impl multi --gpus N`: ONE process drives N GPUs through the library's own runner (pyorbslam_b200.StereoFrontendMulti: one
    engine + host thread per GPU, contiguous frame ranges, one set of result arrays).  The driver's N > 1 runs use torchrun (one
    process per GPU); this leg shows the same sharding as a library feature.  value = e2e frames/s from pinned host buffers."""
    import torch
    from pyorbslam_b200 import StereoFrontendMulti, _lib
    G = max(1, min(args.gpus, _lib.device_count()))
    fe = StereoFrontendMulti(ORB["nfeatures"], ORB["scaleFactor"], ORB["nlevels"], ORB["iniThFAST"], ORB["minThFAST"], H, W, args.chunk, devices=list(range(G)))
    n = args.e2e_pairs * G
    nb = args.base_pairs
    base = [make_pair(i, H, W) for i in range(nb)]
    lh = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
    rh = torch.empty((n, H, W), dtype=torch.uint8, pin_memory=True)
    for g in range(0, n, nb):
        m = min(nb, n - g)
        lh[g:g + m] = torch.from_numpy(np.stack([np.roll(p[0], 9 * (g // nb), axis=1) for p in base[:m]]))
        rh[g:g + m] = torch.from_numpy(np.stack([np.roll(p[1], 9 * (g // nb), axis=1) for p in base[:m]]))
    out = fe.alloc_outputs(n)
    for _ in range(max(args.warmup, 2)):
        fe.run_host(lh, rh, MBF, FX, out=out)
    l0 = _lib.kernel_launches()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fe.run_host(lh, rh, MBF, FX, out=out)
    dt = time.perf_counter() - t0
    d2h = sum(v.numel() * v.element_size() for v in out.values())
    emit({"impl": "multi", "metric": METRIC, "value": n * args.steps / dt, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": max(args.warmup, 2),
          "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
          "config": workload_config(args.e2e_pairs, args.chunk),
          "e2e": {"value": n * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * n * H * W, "d2h_bytes_per_step": d2h, "pairs_per_step": n},
          "gpu_launches": _lib.kernel_launches() - l0, "shards": fe.shards(n),
          "what": "one process, one host thread + engine per GPU (StereoFrontendMulti.run_host), wall clock around the call"})


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner at WARN level, nvcc-built
    helpers, ...) write to file descriptor 1 directly, so fd 1 is pointed at stderr for the whole run and the JSON line
    goes to a saved duplicate of the original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "multi"])
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS), help="kitti = the headline metric's configuration")
    ap.add_argument("--pairs", type=int, default=None, help="stereo pairs per GPU per step (kitti: 4096 = BASELINE.json configs[2])")
    ap.add_argument("--chunk", type=int, default=None, help="pairs per kernel-sequence launch (kitti: 128)")
    ap.add_argument("--base-pairs", type=int, default=16, help="distinct synthetic scenes per rank")
    ap.add_argument("--e2e-pairs", type=int, default=2048, help="pairs per end-to-end step, capped at --pairs; the same at every --gpus "
                    "(pinned host memory: 0.93 MB in + 0.25 MB out per pair and rank)")
    ap.add_argument("--streams", type=int, default=2, help="front-ends running consecutive chunks concurrently (own workspace + stream each)")
    ap.add_argument("--scenes", default="kitti_like", choices=["kitti_like", "layered"], help="synthetic scene generator (pyorbslam_b200/synthetic.py)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-other-configs", action="store_true", help="skip the config 1 / 4 / 5 side measurements (fixture image, hires, stereo-only sweep)")
    args = ap.parse_args()
    global STREAMS, H, W, ORB, WORKLOAD_NAME, SCENES
    STREAMS = args.streams
    SCENES = args.scenes
    wl = WORKLOADS[args.workload]
    H, W, ORB, WORKLOAD_NAME = wl["H"], wl["W"], wl["orb"], wl["name"]
    args.pairs = args.pairs or wl["pairs"]
    args.chunk = args.chunk or wl["chunk"]
    args.e2e_pairs = min(args.e2e_pairs, args.pairs)
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "multi":
        run_multi_runner(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
