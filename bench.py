#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native ORB front-end.

Metric (BASELINE.json): ORB extraction frames/s at 640x480, 1000 features per frame (TUM RGB-D settings:
scale 1.2, 8 levels, FAST 20/7), whole job over N GPUs.  One "step" = one pass of the extractor over one
batch of B synthetic frames per GPU.  Frames shard independently across GPUs (no collective on the data
path; torch.distributed is used only for the timing barrier and the max-over-ranks reduction).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B]          native arm (CUDA, C ABI)
  python bench.py --impl reference ...                                      reference CPU arm (oracle/_ref)

value   : device-resident throughput (inputs already in HBM), CUDA events on the extractor's stream
e2e     : same metric through the host-pointer C-ABI call (pinned host frames in, keypoints/descriptors out,
          H2D and D2H inside the timed region)
roofline: dominant stage by live CUDA-event time; algorithmic bytes of that stage / its duration vs the
          measured HBM peak (MEASURED_PEAKS.json)
cpu_baseline: the reference's own ORBextractor.cc (oracle/_ref, all host cores) on a bounded sample
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from tools.synth import synth_batch  # noqa: E402

WIDTH, HEIGHT, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH = 640, 480, 1000, 1.2, 8, 20, 7
WORKLOAD = "ORBextractor 640x480 grayscale, nFeatures=1000, scale 1.2, 8 levels, FAST 20/7 (TUM RGB-D settings)"
METRIC = "ORB extract frames/sec @640x480 1000 feat"
MASKED = False
# BASELINE.json configs that are a single-GPU extraction workload: c1 is the one the headline metric is quoted on (default);
# c3 / c5 are selectable for the record (python bench.py --workload c5), they are not the driver's bench line.
WORKLOADS = {
    "c1": (640, 480, 1000, 1024, False, WORKLOAD, METRIC),
    "c3": (752, 480, 2000, 512, False, "ORBextractor 752x480 grayscale (EuRoC-shape monocular initialiser frames), nFeatures=2000, scale 1.2, 8 levels, FAST 20/7",
           "ORB extract frames/sec @752x480 2000 feat"),
    "c5": (1920, 1080, 1000, 64, True, "batched 1920x1080 multi-sequence ORB extraction with dynamic-mask keypoint culling (detect -> MovingKeyPoints -> ProcessDesp), "
           "nFeatures=1000, scale 1.2, 8 levels, FAST 20/7", "masked ORB extract frames/sec @1920x1080 1000 feat"),
}


PAIR_WORKLOADS = {
    # name: (width, height, nfeatures, pairs per GPU per step, workload text, metric)
    "c2": (640, 480, 1000, 256, "ORBmatcher::SearchForInitialization between two synthetic 1000-keypoint 640x480 frames (frame B = frame A warped by (+7, -4) px, 2 deg), "
           "ORBmatcher(0.9, true), window 100: 64x48 grid build + windowed 256-bit Hamming search + ratio test + ordered resolve + rotation histogram, per frame pair",
           "SearchForInitialization frame pairs/sec @1000x1000 keypoints, window 100"),
    "c4": (1241, 376, 2000, 64, "KITTI stereo 1241x376: ORB extraction of the left and the right image (nFeatures=2000, scale 1.2, 8 levels, FAST 20/7) + Frame::ComputeStereoMatches "
           "(row-band Hamming + 11-shift SAD + parabola + median cut), per stereo pair", "stereo pairs/sec @1241x376 2000 feat (2 extractions + ComputeStereoMatches)"),
}


def level_pixels():
    s, tot, px = 1.0, 0, []
    f = np.float32(1.0)
    for l in range(NLEVELS):
        inv = np.float32(1.0) / f
        w, h = int(np.rint(np.float32(WIDTH) * inv)), int(np.rint(np.float32(HEIGHT) * inv))
        px.append(w * h)
        f = np.float32(np.float64(f) * np.float64(np.float32(SCALE)))
    return px


def stage_bytes(n_kp, n_cand):
    """Algorithmic bytes per FRAME of each stage."""
    P = level_pixels()
    sp = sum(P)
    return {
        "pyr_resize": sum(P[l - 1] + P[l] for l in range(1, NLEVELS)),      # read P(l-1), write P(l)
        "fast_cells": sp + 4 * n_cand,                                      # read every level once, write candidates
        "octree_sort": n_cand * (4 + 4 + 8 + 4),                            # read slots, write ordered + sorted key + packed
        "octree_tree": n_cand * 12 + 4 * n_kp,                              # read sorted keys + packed, write level keypoints
        "gauss7": 2 * sp,                                                   # read + write every level
        "orient_describe": n_kp * (749 + 512 + 60),                         # disc + samples + 28 B kp + 32 B descriptor
    }


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def cpu_threads_run(make_worker, seconds_budget=None, threads=None, steps=None, units_per_thread_step=None):
    """Runs one independent worker per host thread (make_worker(t) -> callable(i) doing one unit of work; ctypes calls release the GIL).
    Either for ~seconds_budget (cpu_baseline leg) or for `steps` steps of threads * units_per_thread_step units (--impl reference)."""
    threads = threads or max(1, os.cpu_count() or 1)
    workers = [make_worker(t) for t in range(threads)]
    done = [0] * threads; errors = []; start_evt = threading.Event()

    def body(t, deadline_box, quota):
        start_evt.wait()
        i = t
        while True:
            if quota is not None and done[t] >= quota:
                break
            if quota is None and time.perf_counter() >= deadline_box[0]:
                break
            try:
                workers[t](i)
            except BaseException as e:
                errors.append(repr(e)); break
            done[t] += 1; i += threads

    def run(quota, budget):
        for t in range(threads):
            done[t] = 0
        box = [0.0]; start_evt.clear()
        ths = [threading.Thread(target=body, args=(t, box, quota)) for t in range(threads)]
        for th in ths:
            th.start()
        t0 = time.perf_counter(); box[0] = t0 + (budget or 1e9)
        start_evt.set()
        for th in ths:
            th.join()
        if errors:
            raise RuntimeError("CPU reference worker failed: " + errors[0])
        return sum(done), time.perf_counter() - t0

    run(1, None)
    if steps is None:
        n, dt = run(None, seconds_budget)
        return {"units": n, "seconds": dt, "rate": n / dt, "threads": threads, "step_times": None}
    st, total = [], 0
    for _ in range(steps):
        n, dt = run(units_per_thread_step, None)
        st.append(dt); total += n
    return {"units": total, "seconds": sum(st), "rate": total / sum(st), "threads": threads, "step_times": st}


def cpu_reference_run(seconds_budget, frames, threads=None, steps=None, frames_per_thread_step=None):
    """Times the reference's own CPU extractor (oracle/_ref; falls back to the port if _ref is absent) with one
    independent extractor per host thread.  Either runs for ~seconds_budget (cpu_baseline leg) or for `steps`
    steps of threads*frames_per_thread_step frames each (--impl reference)."""
    import oracle
    kind = "ref" if oracle.have_ref() else "port"
    threads = threads or max(1, os.cpu_count() or 1)
    exts = [oracle.Extractor(kind, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH) for _ in range(threads)]
    nframes = len(frames)
    if MASKED:      # config c5: the reference's two-stage Amos path per frame
        from tools.synth import synth_mask
        masks = [synth_mask(i, WIDTH, HEIGHT) for i in range(min(nframes, 8))]
        lab = np.ones((HEIGHT, WIDTH), np.float64); ids = np.zeros(1, np.int32); rm = np.zeros(1, np.int32)

        def one(ext, i):
            kp, counts = ext.detect(frames[i % nframes])
            kp, counts, _ = ext.moving_keypoints(masks[i % len(masks)], lab, ids, rm, kp, counts)
            ext.process_desp(kp, counts)
    else:
        def one(ext, i):
            ext.extract(frames[i % nframes])
    done = [0] * threads
    errors = []
    start_evt = threading.Event()

    def worker(t, deadline_box, quota):
        start_evt.wait()
        i = t
        while True:
            if quota is not None and done[t] >= quota:
                break
            if quota is None and time.perf_counter() >= deadline_box[0]:
                break
            try:
                one(exts[t], i)
            except BaseException as e:      # a dead worker must not read as "0 frames/s"
                errors.append(repr(e))
                break
            done[t] += 1
            i += threads

    def run(quota, budget):
        for t in range(threads):
            done[t] = 0
        box = [0.0]
        start_evt.clear()
        ths = [threading.Thread(target=worker, args=(t, box, quota)) for t in range(threads)]
        for th in ths:
            th.start()
        t0 = time.perf_counter(); box[0] = t0 + (budget or 1e9)
        start_evt.set()
        for th in ths:
            th.join()
        if errors:
            raise RuntimeError("CPU reference worker failed: " + errors[0])
        return sum(done), time.perf_counter() - t0

    run(1, None)      # warm-up: one frame per thread
    if steps is None:
        n, dt = run(None, seconds_budget)
        return {"frames": n, "seconds": dt, "fps": n / dt, "threads": threads, "kind": "reference" if kind == "ref" else "port", "step_times": None}
    step_times, total = [], 0
    for _ in range(steps):
        n, dt = run(frames_per_thread_step, None)
        step_times.append(dt); total += n
    return {"frames": total, "seconds": sum(step_times), "fps": total / sum(step_times), "threads": threads,
            "kind": "reference" if kind == "ref" else "port", "step_times": step_times}


def matcher_bench(orbx, torch, ext, frames, device):
    """Second half of BASELINE's metric: Hamming matches/s.  (a) batched brute-force best/second-best over P frame pairs of
    1000 x 1000 descriptors resident in HBM (distance evaluations/s, pairs/s); (b) ORBmatcher::SearchForInitialization through
    the host-pointer C-ABI call on two extracted frames (pairs/s, H2D/D2H of keypoints + descriptors included)."""
    P, NQ = 256, 1000
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    q = torch.randint(0, 256, (P, NQ, 32), dtype=torch.uint8, device="cuda", generator=g)
    t = torch.randint(0, 256, (P, NQ, 32), dtype=torch.uint8, device="cuda", generator=g)
    bi = torch.empty((P, NQ), dtype=torch.int32, device="cuda"); bd = torch.empty_like(bi); sd = torch.empty_like(bi)
    m = orbx.ORBmatcher(0.9, True, device=device)
    st = torch.cuda.ExternalStream(m.stream)
    call = lambda: m.match_bruteforce_batch_device(P, q.data_ptr(), NQ, t.data_ptr(), NQ, bi.data_ptr(), bd.data_ptr(), sd.data_ptr())
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(st)
    for _ in range(reps):
        call()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kp_a, d_a = ext(frames[0]); kp_b, d_b = ext(frames[1])
    sf = ext.GetScaleFactors()
    FA, FB = orbx.FrameView(kp_a, d_a, WIDTH, HEIGHT, sf), orbx.FrameView(kp_b, d_b, WIDTH, HEIGHT, sf)
    prev = np.stack([kp_a["x"], kp_a["y"]], 1)
    for _ in range(3):
        m.SearchForInitialization(FA, FB, prev, 100)
    t0 = time.perf_counter(); reps2 = 50
    for _ in range(reps2):
        m.SearchForInitialization(FA, FB, prev, 100)
    dt = (time.perf_counter() - t0) / reps2
    # the same pair as device-resident Frames (keypoints / descriptors stay in HBM after extraction, grid built once per frame):
    # frame build = undistortion (TUM1 coefficients) + RGB-D stereo + 64x48 grid straight from the extractor's device result
    cam = orbx.Camera.make(517.306408, 516.469215, 318.643040, 255.313989, 0.262383, -0.953104, -0.005358, 0.002628, 1.163314, 40.0)
    cam0 = orbx.Camera.make(517.306408, 516.469215, 318.643040, 255.313989, bf=40.0)
    depth = np.full(len(kp_b), 2.0, np.float32)
    DA = orbx.Frame(device).assign_host(kp_a, d_a, sf, cam0, HEIGHT, WIDTH); DB = orbx.Frame(device)
    for _ in range(3):
        DB.assign(ext, cam, HEIGHT, WIDTH, depth)
    t0 = time.perf_counter()
    for _ in range(reps2):
        DB.assign(ext, cam, HEIGHT, WIDTH, depth)
    dt_build = (time.perf_counter() - t0) / reps2
    DB.assign(ext, cam0, HEIGHT, WIDTH)
    for _ in range(3):
        m.SearchForInitialization(DA, DB, prev, 100)
    t0 = time.perf_counter()
    for _ in range(reps2):
        m.SearchForInitialization(DA, DB, prev, 100)
    dt_dev = (time.perf_counter() - t0) / reps2
    # bag of words at the size of ORBvoc.txt (k = 10, L = 6: 1 111 110 nodes, 10^6 words; random node descriptors -- the timing does not
    # depend on their values): Frame::ComputeBoW = transform(descriptors, mBowVec, mFeatVec, 4), then SearchByBoW(KeyFrame, Frame)
    rngv = np.random.default_rng(5)
    nn = (10 ** 7 - 10) // 9
    ids = np.arange(1, nn + 1, dtype=np.int64)
    vparent = ((ids - 1) // 10).astype(np.int32); vleaf = (ids > (10 ** 6 - 10) // 9).astype(np.uint8)
    V = orbx.ORBVocabulary(10, 6, vparent, vleaf, rngv.integers(0, 256, (nn, 32), dtype=np.uint8), np.where(vleaf > 0, rngv.uniform(0.5, 9.0, nn), 0.0), device=device)
    for _ in range(3):
        fa = V.transform(d_a, 4); fb = V.transform(d_b, 4)
    t0 = time.perf_counter()
    for _ in range(reps2):
        fb = V.transform(d_b, 4)
    dt_bow = (time.perf_counter() - t0) / reps2
    va = np.ones(len(kp_a), np.uint8)
    mb = orbx.ORBmatcher(0.7, True, device=device)
    for _ in range(3):
        mb.SearchByBoW(0, kp_a, d_a, va, fa, kp_b, d_b, None, fb)
    t0 = time.perf_counter()
    for _ in range(reps2):
        nbow = mb.SearchByBoW(0, kp_a, d_a, va, fa, kp_b, d_b, None, fb)[0]
    dt_sbow = (time.perf_counter() - t0) / reps2
    bow = {"compute_bow_us": dt_bow * 1e6, "compute_bow_config": "orbx_vocabulary_transform, %d descriptors, k=10 L=6 tree (%d nodes) resident in HBM/L2, levelsup 4, host pointers in/out" % (len(d_b), nn),
           "search_by_bow_us": dt_sbow * 1e6, "search_by_bow_config": "KeyFrame x Frame form, %d x %d features in %d / %d nodes, host pointers in/out" % (len(kp_a), len(kp_b), len(fa["fv_nodes"]), len(fb["fv_nodes"])),
           "search_by_bow_matches": int(nbow)}
    return {"bow": bow, "frame_build_us": dt_build * 1e6, "frame_build_config": "orbx_frame_assign on the extractor's device result: %d keypoints, 5-coefficient undistortion, RGB-D stereo, grid" % len(kp_b),
            "search_for_initialization_device_frames_pairs_per_s": 1.0 / dt_dev,
            "bruteforce_pairs_per_s": P / (ms * 1e-3), "hamming_distances_per_s": P * NQ * NQ / (ms * 1e-3),
            "bruteforce_config": "%d pairs x %d x %d descriptors per launch, device-resident" % (P, NQ, NQ),
            "search_for_initialization_pairs_per_s": 1.0 / dt, "search_for_initialization_config": "host-pointer call, window 100, one pair per call (latency-bound)"}


def _peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return peak, ("measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)")


def _timed_device_steps(torch, dist, world, stream, K, W, fn):
    """W warm-up calls, then K calls bracketed by CUDA events on `stream` with a barrier + synchronize on both sides; max over ranks (ms)."""
    for _ in range(max(W, 3)):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        fn()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def _c2_frames(P, rank, world=1):
    """P frame pairs: frame A synthetic (distinct seed per pair), frame B = A warped by (+7, -4) px, 2 deg."""
    from tools.synth import synth_batch_distinct, warp_affine_nn
    A = synth_batch_distinct(P, WIDTH, HEIGHT, seed0=2000 + rank * P, workers=max(1, min(32, (os.cpu_count() or 1) // max(world, 1))))
    return A, np.stack([warp_affine_nn(a, 7, -4, 2.0) for a in A])


def _c2_pairs(orbx, device, A, Bf):
    """both frames of every pair extracted on the GPU (setup, outside the timed region)"""
    P = len(A)
    E = orbx.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=device)
    ka, da, ca = E.extract_batch(A); kb, db, cb = E.extract_batch(Bf)
    sf = E.GetScaleFactors()
    pairs = [(ka[p][:ca[p]].copy(), da[p][:ca[p]].copy(), kb[p][:cb[p]].copy(), db[p][:cb[p]].copy()) for p in range(P)]
    return pairs, sf


def bench_c2(args, rank, local_rank, world):
    """BASELINE config 2: SearchForInitialization frame pairs/s (the "Hamming matches/sec" half of the metric), P pairs per call."""
    import ctypes as C
    N, K, W, P = args.gpus, args.steps, max(args.warmup, 0), (args.batch or PAIR_WORKLOADS["c2"][3])
    A, Bf = _c2_frames(P, rank, world)                   # before CUDA is initialised (the generator forks)
    import torch
    import torch.distributed as dist
    orbx = importlib.import_module("amos-slam_b200")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pairs, sf = _c2_pairs(orbx, local_rank, A, Bf)
    from importlib import import_module
    mm = import_module("amos-slam_b200._matcher")
    M = orbx.ORBmatcher(0.9, True, device=local_rank)
    L = M._lib
    views1 = [orbx.FrameView(p[0], p[1], WIDTH, HEIGHT, sf) for p in pairs]; views2 = [orbx.FrameView(p[2], p[3], WIDTH, HEIGHT, sf) for p in pairs]
    v1 = (mm._FrameViewC * P)(*[v.c() for v in views1]); v2 = (mm._FrameViewC * P)(*[v.c() for v in views2])
    n1 = np.array([len(p[0]) for p in pairs]); o1 = np.concatenate([[0], np.cumsum(n1)])
    prev0 = np.concatenate([np.stack([p[0]["x"], p[0]["y"]], 1) for p in pairs]).astype(np.float32)
    prev = prev0.copy(); m12 = np.zeros(int(o1[-1]), np.int32); nm = np.zeros(P, np.int32)
    pp = (C.c_void_p * P)(*[prev.ctypes.data + 8 * int(o1[p]) for p in range(P)]); mp = (C.c_void_p * P)(*[m12.ctypes.data + 4 * int(o1[p]) for p in range(P)])
    cam0 = orbx.Camera.make(517.306408, 516.469215, 318.643040, 255.313989, bf=40.0)
    D1 = [orbx.Frame(local_rank).assign_host(p[0], p[1], sf, cam0, HEIGHT, WIDTH) for p in pairs]; D2 = [orbx.Frame(local_rank).assign_host(p[2], p[3], sf, cam0, HEIGHT, WIDTH) for p in pairs]
    a1 = (C.c_void_p * P)(*[f._h.value for f in D1]); a2 = (C.c_void_p * P)(*[f._h.value for f in D2])
    nmv = nm.ctypes.data_as(C.c_void_p)

    def step_device():       # frames resident in HBM (keypoints, descriptors, grids); vbPrevMatched in, matches out
        prev[:] = prev0
        orbx._check(L.orbx_search_for_initialization_frames_batch(M._h, P, a1, a2, pp, mp, 100, nmv))

    def step_host():         # host frame views in (keypoints + descriptors uploaded every call), matches out
        prev[:] = prev0
        orbx._check(L.orbx_search_for_initialization_batch(M._h, P, v1, v2, pp, mp, 100, nmv))

    stream = torch.cuda.ExternalStream(M.stream)
    sampler = ClockSampler(local_rank); sampler.start()
    # device leg: NS matcher handles, one host thread each (Tracking and LocalMapping call the reference's matcher from their own threads as well), every handle
    # with an even share of the step's pairs: one handle's host side (argument staging, result copy-back: ~350 us of a 256-pair call) overlaps the other's
    # kernels.  Timed with CUDA events: start on the first handle's stream with the GPU idle, one end event per handle's stream after its last call.
    NS = max(1, min(args.streams, P))
    Ms = [M] + [orbx.ORBmatcher(0.9, True, device=local_rank) for _ in range(NS - 1)]
    cut = [P * i // NS for i in range(NS + 1)]
    shares = [((C.c_void_p * (cut[i + 1] - cut[i]))(*[f._h.value for f in D1[cut[i]:cut[i + 1]]]), (C.c_void_p * (cut[i + 1] - cut[i]))(*[f._h.value for f in D2[cut[i]:cut[i + 1]]]),
               (C.c_void_p * (cut[i + 1] - cut[i]))(*[prev.ctypes.data + 8 * int(o1[p]) for p in range(cut[i], cut[i + 1])]),
               (C.c_void_p * (cut[i + 1] - cut[i]))(*[m12.ctypes.data + 4 * int(o1[p]) for p in range(cut[i], cut[i + 1])]), C.c_void_p(nm.ctypes.data + 4 * cut[i]), cut[i + 1] - cut[i]) for i in range(NS)]
    mstreams = [torch.cuda.ExternalStream(m_.stream) for m_ in Ms]

    def share_steps(i, n_steps, end_event, errs):
        try:
            sa1, sa2, spp, smp, snm, n = shares[i]
            lo, hi = int(o1[cut[i]]), int(o1[cut[i + 1]])
            for _ in range(n_steps):
                prev[lo:hi] = prev0[lo:hi]
                orbx._check(L.orbx_search_for_initialization_frames_batch(Ms[i]._h, n, sa1, sa2, spp, smp, 100, snm))
            if end_event is not None:
                end_event.record(mstreams[i])
        except Exception as e:                                   # noqa: BLE001
            errs.append(e)

    def run_shares(n_steps, ends):
        errs = []
        ths = [threading.Thread(target=share_steps, args=(i, n_steps, ends[i] if ends else None, errs)) for i in range(NS)]
        for t in ths: t.start()
        for t in ths: t.join()
        if errs:
            raise errs[0]

    l0 = sum(m_.launch_count for m_ in Ms)
    if NS == 1:
        ms = _timed_device_steps(torch, dist, world, stream, K, W, step_device)
    else:
        run_shares(max(W, 3), None)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); ends = [torch.cuda.Event(enable_timing=True) for _ in range(NS)]
        e0.record(mstreams[0])
        run_shares(K, ends)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e) for e in ends)
    launches = sum(m_.launch_count for m_ in Ms) - l0
    nm_dev = nm.copy(); m12_dev = m12.copy()
    # stage breakdown (a separate pass with events between the five launches)
    orbx._check(L.orbx_matcher_profile_enable(M._h, 1))
    for _ in range(5):
        step_device()
    st = (C.c_double * 5)(); nc = C.c_int()
    orbx._check(L.orbx_matcher_profile_collect(M._h, 5, st, C.byref(nc)))
    orbx._check(L.orbx_matcher_profile_enable(M._h, 0))
    stage_names = ["grid_build", "window_count", "scan", "window_fill", "resolve"]
    stage_ms = {k: st[i] / max(nc.value, 1) for i, k in enumerate(stage_names)}
    # e2e: wall clock around K synchronous host-view calls
    Ke = max(3, min(K, 20))
    for _ in range(3):
        step_host()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        step_host()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True; sampler.join(timeout=2)
    assert np.array_equal(nm, nm_dev) and np.array_equal(m12, m12_dev), "host-view and device-frame calls disagree"
    # work per pair: candidate distances actually evaluated = sum of the list lengths (one call of the windows API gives them)
    tq = np.concatenate([[len(p[0])] for p in pairs])
    lv0 = sum(int((p[0]["octave"] == 0).sum()) for p in pairs)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0].item()), float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # distances per pair, measured once with the single-pair windows API (counts = GetFeaturesInArea result sizes of the level-0 features)
    ndist = 0
    for p in range(min(P, 16)):
        q = pairs[p][0]["octave"] == 0
        xy = np.stack([pairs[p][0]["x"][q], pairs[p][0]["y"][q]], 1).astype(np.float32)
        ndist += sum(len(c) for c in M.GetFeaturesInArea(D2[p], xy, 100.0, 0, 0))
    ndist_pair = ndist / float(min(P, 16))
    value = N * P * K / (ms * 1e-3); e2e = N * P * Ke / (ms_e2e * 1e-3)
    peak, peak_src = _peaks()
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    nkeys = float(np.mean([len(p[0]) + len(p[2]) for p in pairs]))
    bytes_of = {"grid_build": 28 * nkeys / 2 + 4 * (nkeys / 2 + 3073), "window_count": 28 * ndist_pair + 8 * lv0 / P, "scan": 8 * nkeys / 2,
                "window_fill": (28 + 32 + 4) * ndist_pair + 32 * lv0 / P, "resolve": 4 * ndist_pair + 24 * nkeys / 2}
    achieved = bytes_of[dom] * P / (stage_ms[dom] * 1e-3) / 1e9
    h2d = int(sum((len(p[0]) + len(p[2])) * 60 + len(p[0]) * 8 for p in pairs) + P * 200); d2h = int(sum(len(p[0]) * 12 for p in pairs) + 8 * P)
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": N, "steps": K, "warmup": max(W, 3), "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": P, "keypoints_per_frame": nkeys / 2, "level0_queries_per_pair": lv0 / P, "hamming_distances_per_pair": ndist_pair,
                       "parallelism": "frame pairs sharded over %d GPU(s), no collective; %d matcher handle(s) (host threads) per GPU, one batched call per handle and step on an even share of the pairs" % (N, NS),
                       "l2_policy": "per-step working set is %.1f MB of keypoints + descriptors (< L2): the path is latency / issue bound, not HBM bound; every step re-reads the same resident frames" % (P * nkeys * 60 / 1e6)},
            "hamming_distances_per_s": value * ndist_pair,
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "note": "orbx_search_for_initialization_batch on host frame views: keypoints, descriptors and vbPrevMatched uploaded, matches / vbPrevMatched / counts downloaded, every step; "
                            "bound by the host's memory bandwidth, not by the GPU or the link: the caller's pageable arrays (128 KB per pair) are staged into the pinned upload mirror by 8 threads and read again by the DMA "
                            "(1 / 2 / 4 matcher handles on as many host threads measured 93 / 91 / 96 k pairs/s on the 16-vCPU box); frames kept on the device (orbx_frame_*) skip all of it: `value`"},
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": {"grid_build": "k_grid_build_pairs", "window_count": "k_window_search_pairs<false>", "scan": "k_scan_counts_pairs", "window_fill": "k_window_search_pairs<true>",
                                                   "resolve": "k_resolve_init_pairs"}[dom],
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_of[dom] * P, "ms_per_launch": stage_ms[dom], "stage_ms_per_step": stage_ms,
                         "reading": "matching is integer-pipe / latency work on a few MB: the HBM fraction is reported because the contract asks for it; the pipe utilisation of the Hamming kernels is in profiles/ (ncu smsp__inst_executed_pipe_*)"},
            "matches_per_pair": float(nm.mean())}
    if N == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = _c2_cpu(pairs, sf, args.cpu_seconds, None, None)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _c2_cpu(pairs, sf, seconds, steps, per_thread):
    """The reference's own SearchForInitialization body (oracle/_ref, else the port) on all host cores, one matcher per thread."""
    import oracle
    kind = "ref" if oracle.have_ref() else "port"
    fd = [(oracle.FrameData(p[0], p[1], WIDTH, HEIGHT, sf), oracle.FrameData(p[2], p[3], WIDTH, HEIGHT, sf), np.stack([p[0]["x"], p[0]["y"]], 1).astype(np.float32)) for p in pairs[:64]]

    def make(t):
        m = oracle.Matcher(kind, 0.9, True)
        return lambda i: m.search_for_initialization(fd[i % len(fd)][0], fd[i % len(fd)][1], fd[i % len(fd)][2], 100)
    res = cpu_threads_run(make, seconds_budget=seconds, steps=steps, units_per_thread_step=per_thread)
    res["kind"] = "reference" if kind == "ref" else "port"
    if steps is not None:
        return res
    return {"value": res["rate"], "unit": "pairs/s", "cores": res["threads"], "kind": res["kind"],
            "sample": "%d pairs in %.1f s, one independent ORBmatcher per thread on %d threads (frames already extracted), CPU: %s" % (res["units"], res["seconds"], res["threads"], cpu_model())}


def _c4_pairs(P, rank):
    from tools.synth import synth_batch_distinct, stereo_right_from_left
    Ls = synth_batch_distinct(P, WIDTH, HEIGHT, seed0=3000 + rank * P)
    Rs = np.stack([stereo_right_from_left(Ls[b], 3000 + rank * P + b + 1) for b in range(P)])
    return Ls, Rs


BF_KITTI = 386.1448          # Examples/Stereo/KITTI00-02.yaml:25


def bench_c4(args, rank, local_rank, world):
    """BASELINE config 4: KITTI stereo pairs/s = left + right extraction (two handles, two streams) + Frame::ComputeStereoMatches, B pairs per step."""
    import ctypes as C
    Ls, Rs = _c4_pairs(args.batch or PAIR_WORKLOADS["c4"][3], rank)       # before CUDA is initialised (the generator forks)
    import torch
    import torch.distributed as dist
    orbx = importlib.import_module("amos-slam_b200")
    N, K, W, B = args.gpus, args.steps, max(args.warmup, 0), len(Ls)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    EL = orbx.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank); ER = orbx.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank)
    M = orbx.ORBmatcher(device=local_rank); Lb = M._lib
    cap = EL.max_keypoints(HEIGHT, WIDTH); ER.max_keypoints(HEIGHT, WIDTH)
    hL = torch.from_numpy(Ls).pin_memory(); hR = torch.from_numpy(Rs).pin_memory(); dL = hL.cuda(); dR = hR.cuda()
    dev = [(torch.empty((B, cap, 28), dtype=torch.uint8, device="cuda"), torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")) for _ in range(2)]
    hst = [(torch.empty((B, cap, 28), dtype=torch.uint8).pin_memory(), torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()) for _ in range(2)]
    d_ur = torch.empty((B, cap), dtype=torch.float32, device="cuda"); d_dep = torch.empty_like(d_ur)
    h_ur = torch.empty((B, cap), dtype=torch.float32).pin_memory(); h_dep = torch.empty_like(h_ur).pin_memory()

    def step_device():
        for E, d, o in ((EL, dL, dev[0]), (ER, dR, dev[1])):
            E.extract_batch_raw(d.data_ptr(), B, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT, o[0].data_ptr(), o[1].data_ptr(), cap, o[2].data_ptr(), device=True)
        orbx._check(Lb.orbx_compute_stereo_matches_batch_device(M._h, EL._h, ER._h, B, cap, 0.0, BF_KITTI, C.c_void_p(d_ur.data_ptr()), C.c_void_p(d_dep.data_ptr())))

    errs = []

    def host_extract(E, h, o):
        try:
            torch.cuda.set_device(local_rank)
            E.extract_batch_raw(h.data_ptr(), B, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT, o[0].data_ptr(), o[1].data_ptr(), cap, o[2].data_ptr(), device=False)
        except BaseException as ex:
            errs.append(repr(ex))

    def step_host():         # the reference runs the two extractions in two threads (src/Frame.cc:165-173), then ComputeStereoMatches
        ths = [threading.Thread(target=host_extract, args=(EL, hL, hst[0])), threading.Thread(target=host_extract, args=(ER, hR, hst[1]))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        if errs:
            raise SystemExit("e2e worker failed: " + errs[0])
        orbx._check(Lb.orbx_compute_stereo_matches_batch(M._h, EL._h, ER._h, B, cap, 0.0, BF_KITTI, C.c_void_p(h_ur.data_ptr()), C.c_void_p(h_dep.data_ptr())))

    stream = torch.cuda.ExternalStream(M.stream)          # the matcher's stream waits for both extractors' streams: its events cover the whole step
    sampler = ClockSampler(local_rank); sampler.start()
    l0 = EL.launch_count + ER.launch_count + M.launch_count
    ms = _timed_device_steps(torch, dist, world, stream, K, W, step_device)
    launches = EL.launch_count + ER.launch_count + M.launch_count - l0
    ur_dev = d_ur.cpu().numpy().copy()
    # stage breakdown: extractor stages of the left handle ALONE (serialised pass: with the right handle running beside it on its own stream the stage
    # events of the left one absorb the other's kernels), x 2 for the two images; then the matcher launches of a whole step
    torch.cuda.synchronize()
    EL.profile_enable(True)
    for _ in range(3):
        EL.extract_batch_raw(dL.data_ptr(), B, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT, dev[0][0].data_ptr(), dev[0][1].data_ptr(), cap, dev[0][2].data_ptr(), device=True)
    torch.cuda.synchronize()
    ext_ms, ncalls = EL.profile_collect(); EL.profile_enable(False)
    orbx._check(Lb.orbx_matcher_profile_enable(M._h, 1))
    for _ in range(3):
        step_device()
    st = (C.c_double * 2)(); nc = C.c_int()
    orbx._check(Lb.orbx_matcher_profile_collect(M._h, 2, st, C.byref(nc))); orbx._check(Lb.orbx_matcher_profile_enable(M._h, 0))
    stage_ms = {("extract_" + k): 2 * v / max(ncalls, 1) for k, v in ext_ms.items()}                     # x2: left and right image
    stage_ms["stereo_match"] = st[0] / max(nc.value, 1); stage_ms["stereo_median_cut"] = st[1] / max(nc.value, 1)
    Ke = max(3, min(K, 10))
    for _ in range(2):
        step_host()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        step_host()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True; sampler.join(timeout=2)
    assert np.array_equal(h_ur.numpy(), ur_dev), "host and device stereo paths disagree"
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0].item()), float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_kp = float(hst[0][2].float().mean().item()); matched = float((h_ur.numpy() >= 0).sum(1).mean())
    value = N * B * K / (ms * 1e-3); e2e = N * B * Ke / (ms_e2e * 1e-3)
    peak, peak_src = _peaks()
    n_cand = float(sum(len(EL.debug_level_candidates(0, l)) for l in range(NLEVELS)))
    sb = {("extract_" + k): 2 * v for k, v in stage_bytes(n_kp, n_cand).items()}
    band = 4.0 * 1.2 ** 3 + 1.0                                                   # rows a right keypoint of a middle level enters (k_stereo_rows)
    sb["stereo_match"] = 2 * n_kp * 60 + 2 * 4 * n_kp * band + n_kp * (n_kp * band / HEIGHT) * 28 + matched * (64 + 11 * 121 * 2); sb["stereo_median_cut"] = n_kp * 12
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    nl = (NLEVELS - 1) if dom == "extract_pyr_resize" else 1
    achieved = sb[dom] * B / (stage_ms[dom] * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": N, "steps": K, "warmup": max(W, 3), "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": [WIDTH, HEIGHT], "nfeatures": NFEAT, "stereo_pairs_per_gpu_per_step": B, "keypoints_per_image": n_kp, "stereo_matches_per_pair": matched,
                       "parallelism": "stereo pairs sharded over %d GPU(s), no collective; left / right extractor handles on two streams, matcher stream behind both" % N,
                       "l2_policy": "working set per step (%.0f MB of images, ~%.1f GB of pyramids + scratch) exceeds the 126 MB L2" % (2 * B * WIDTH * HEIGHT / 1e6, 2 * B * 6.3e-3 * WIDTH * HEIGHT / 307200.0)},
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * B * WIDTH * HEIGHT), "d2h_bytes_per_step": int(2 * (B * cap * 60 + 4 * B) + 8 * B * cap), "steps": Ke,
                    "note": "two host threads call orbx_extract_batch (left / right images, pinned host in, keypoints + descriptors out), then orbx_compute_stereo_matches_batch (mvuRight / mvDepth out)"},
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": sb[dom] * B / nl, "ms_per_launch": stage_ms[dom] / nl, "stage_ms_per_step": stage_ms}}
    if N == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = _c4_cpu(Ls, Rs, args.cpu_seconds, None, None)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _c4_cpu(Ls, Rs, seconds, steps, per_thread):
    """The reference's own code per stereo pair: ORBextractor on the left and on the right image + Frame::ComputeStereoMatches (oracle/_ref, else the
    port).  One pair per thread (the reference itself uses two threads per pair, src/Frame.cc:165-173; with every core busy the throughput is the same)."""
    import oracle
    kind = "ref" if oracle.have_ref() else "port"
    n = min(len(Ls), 16)

    def make(t):
        el = oracle.Extractor(kind, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH); er = oracle.Extractor(kind, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH); m = oracle.Matcher(kind)

        def one(i):
            kl, dl = el.extract(Ls[i % n]); kr, dr = er.extract(Rs[i % n])
            m.compute_stereo_matches(el, er, kl, dl, kr, dr, 0.0, BF_KITTI)
        return one
    res = cpu_threads_run(make, seconds_budget=seconds, steps=steps, units_per_thread_step=per_thread)
    res["kind"] = "reference" if kind == "ref" else "port"
    if steps is not None:
        return res
    return {"value": res["rate"], "unit": "pairs/s", "cores": res["threads"], "kind": res["kind"],
            "sample": "%d stereo pairs in %.1f s, one pair at a time per thread on %d threads, CPU: %s" % (res["units"], res["seconds"], res["threads"], cpu_model())}


def reference_pairs(args, defP):
    """--impl reference for the pair workloads: the reference's own CPU code on all host cores, same metric / config keys."""
    K, W = args.steps, max(args.warmup, 0)
    threads = max(1, os.cpu_count() or 1)
    if args.workload == "c2":
        # the frames are extracted with the CPU reference as well (no GPU on this arm)
        import oracle
        from tools.synth import synth_batch_distinct, warp_affine_nn
        kind = "ref" if oracle.have_ref() else "port"
        A = synth_batch_distinct(32, WIDTH, HEIGHT, seed0=2000)
        E = oracle.Extractor(kind, NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
        pairs = []
        for a in A:
            ka, da = E.extract(a); kb, db = E.extract(warp_affine_nn(a, 7, -4, 2.0))
            pairs.append((ka, da, kb, db))
        sf = np.cumprod(np.concatenate([[1.0], np.full(NLEVELS - 1, SCALE)])).astype(np.float32)
        per = 512
        res = _c2_cpu(pairs, sf, None, W + K, per)
    else:
        Ls, Rs = _c4_pairs(16, 0)
        per = 16
        res = _c4_cpu(Ls, Rs, None, W + K, per)
    st = res["step_times"][W:]
    rate = threads * per * K / sum(st)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "pairs/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * sum(st) / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": rate, "unit": "pairs/s", "cores": threads, "kind": res["kind"],
                             "sample": "%d steps x %d threads x %d pairs, one independent worker per thread, CPU: %s" % (K, threads, per, cpu_model())},
            "e2e": {"value": rate, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def spread_device(local_rank, world):
    """Device of a rank when the box shows MORE GPUs than there are ranks: alternate between the lower and the upper half of the index
    range instead of taking 0 .. world-1.  Boards group GPUs by PCIe switch in index order, and on this pool GPUs 0-3 share one host
    uplink (4 GPUs: 116 GB/s together, profiles/r02d_h2d_matrix.jsonl) while 0,1,4,5 or 4-7 reach 4 x 55 GB/s -- the end-to-end figure
    at N = 2 / 4 is bound by exactly that link.  With as many ranks as GPUs (or fewer visible devices) this is the identity."""
    try:
        import torch
        n = torch.cuda.device_count()
    except Exception:
        return local_rank
    if n <= world or n < 2 or local_rank // 2 >= n // 2:
        return local_rank
    return (local_rank % 2) * (n // 2) + local_rank // 2


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs that are local to GPU `index` (its PCIe root's NUMA node) BEFORE the pinned host buffers are
    allocated, so first-touch places them on the right socket: with 8 ranks feeding 8 GPUs, frames crossing the inter-socket link
    are what limits the end-to-end figure.  Best effort: any failure leaves the affinity untouched."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=10).stdout.strip()
        bdf = out.lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return "%s -> cpus %s" % (bdf, cpus)
    except Exception as e:
        return "unbound (%s)" % type(e).__name__
    return "unbound"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (default: 1024 for c1 = the batch SURVEY.md 8d names, 512 for c3, 64 for c5)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c1", choices=sorted(list(WORKLOADS) + list(PAIR_WORKLOADS)), help="BASELINE.json config (c1 = headline metric)")
    ap.add_argument("--streams", type=int, default=2, help="extractor handles (camera streams) per GPU; the step's frames are split evenly between them")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-pointer leg (the line is then not a bench value)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl != "reference":
        local_rank = spread_device(local_rank, world)            # every use below is "the CUDA device of this rank"
    global WIDTH, HEIGHT, NFEAT, WORKLOAD, METRIC, MASKED
    if args.workload in PAIR_WORKLOADS:
        WIDTH, HEIGHT, NFEAT, defP, WORKLOAD, METRIC = PAIR_WORKLOADS[args.workload]
        if args.impl == "reference":
            if rank == 0:
                reference_pairs(args, defP)
            return
        (bench_c2 if args.workload == "c2" else bench_c4)(args, rank, local_rank, world)
        return
    WIDTH, HEIGHT, NFEAT, defB, MASKED, WORKLOAD, METRIC = WORKLOADS[args.workload]
    N, K, W, B = args.gpus, args.steps, max(args.warmup, 0), (args.batch or defB)
    # masks travel packed (1 bit per pixel, packed by worker threads of the library inside the host-pointer call) unless ORBX_HOST_PACK=0;
    # the default thread count leaves every feeding thread of every rank its share of the host cores
    if MASKED and "ORBX_HOST_PACK" not in os.environ:
        spare = (os.cpu_count() or 8) // max(1, world * max(1, args.streams)) - 1      # host cores per feeding thread, minus the feeder itself
        # packing costs host cycles: with fewer than 3 spare cores per handle (8 ranks on a 32-vCPU box) the byte masks are faster
        # (profiles/r02g: N = 8 packed 35.6 k vs bytes 45.7 k frames/s; N = 4 packed 41.4 k vs bytes 27.6 k)
        os.environ["ORBX_HOST_PACK"] = str(min(8, spare)) if spare >= 3 else "0"
    hostpack = MASKED and int(os.environ.get("ORBX_HOST_PACK", "0")) > 0
    link_bytes_per_frame = WIDTH * HEIGHT + ((HEIGHT * ((WIDTH + 31) // 32) * 4) if hostpack else (WIDTH * HEIGHT if MASKED else 0))
    config = {"workload": WORKLOAD, "frame": [WIDTH, HEIGHT], "nfeatures": NFEAT, "frames_per_gpu_per_step": B,
              "parallelism": "frames sharded over %d GPU(s), no collective; %d extractor handles (streams) per GPU sharing the step's frames evenly" % (N, max(1, args.streams)),
              "device_map": "rank r -> CUDA device %s" % ("r" if spread_device(1, world) == 1 else "(r % 2) * (visible / 2) + r // 2: ranks spread over both halves of the board (GPUs 0-3 share a host uplink)"),
              "l2_policy": "working set per step (%.0f MB of frames, ~%.1f GB of pyramid+scratch) exceeds the 126 MB L2" % (B * WIDTH * HEIGHT / 1e6, B * 6.3e-3 * WIDTH * HEIGHT / 307200.0)}

    if args.impl == "reference":
        if rank != 0:
            return
        frames = synth_batch(64, WIDTH, HEIGHT, seed0=0, distinct=16 if WIDTH * HEIGHT < 1000000 else 4)
        threads = max(1, os.cpu_count() or 1)
        # frames per thread per step: ~1.5 s of work per step on one core, so that thread start / join and the slowest thread's tail are
        # noise (8 frames per step read 27 % low against the continuous cpu_baseline leg in round 1)
        FPT = max(8, int(round(128 * 307200.0 / (WIDTH * HEIGHT))))
        res = cpu_reference_run(None, frames, threads=threads, steps=W + K, frames_per_thread_step=FPT)
        st = res["step_times"][W:]
        fps = threads * FPT * K / sum(st)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": N, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * sum(st) / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": res["kind"],
                                 "sample": "%d steps x %d threads x %d frames, one independent extractor per thread, CPU: %s" % (K, threads, FPT, cpu_model())},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # B frames from B distinct seeds, generated by a process pool BEFORE CUDA is initialised in this process (the pool forks)
    from tools.synth import synth_batch_distinct
    frames = synth_batch_distinct(B, WIDTH, HEIGHT, seed0=1000 + B * rank, workers=max(1, min(32, (os.cpu_count() or 1) // max(world, 1))))
    import torch
    import torch.distributed as dist
    orbx = importlib.import_module("amos-slam_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ext = orbx.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank)
    cap = ext.max_keypoints(HEIGHT, WIDTH)
    h_frames = torch.from_numpy(frames).pin_memory()
    d_frames = h_frames.cuda(non_blocking=False)
    d_kp = torch.empty((B, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
    h_kp = torch.empty((B, cap, 28), dtype=torch.uint8).pin_memory()
    h_desc = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
    h_counts = torch.zeros((B,), dtype=torch.int32).pin_memory()
    stream = torch.cuda.ExternalStream(ext.stream)

    if MASKED:
        from tools.synth import synth_mask
        h_masks = torch.from_numpy(np.stack([synth_mask(1000 * rank + b, WIDTH, HEIGHT) for b in range(B)])).pin_memory()
        d_masks = h_masks.cuda()
        d_culled = torch.zeros((B,), dtype=torch.int32, device="cuda"); h_culled = torch.zeros((B,), dtype=torch.int32).pin_memory()

        def step_device():
            ext.extract_masked_batch_raw_device(d_frames.data_ptr(), d_masks.data_ptr(), B, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT, WIDTH, WIDTH * HEIGHT,
                                                d_kp.data_ptr(), d_desc.data_ptr(), cap, d_counts.data_ptr(), d_culled.data_ptr())

        def step_host():
            import ctypes as C
            orbx._check(ext._lib.orbx_extract_masked_batch(ext._h, C.c_void_p(h_frames.data_ptr()), C.c_void_p(h_masks.data_ptr()), B, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT,
                                                           WIDTH, WIDTH * HEIGHT, C.c_void_p(h_kp.data_ptr()), C.c_void_p(h_desc.data_ptr()), cap,
                                                           C.c_void_p(h_counts.data_ptr()), C.c_void_p(h_culled.data_ptr())))
    else:
        def step_device():
            ext.extract_batch_raw(d_frames.data_ptr(), B, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT, d_kp.data_ptr(), d_desc.data_ptr(), cap, d_counts.data_ptr(), device=True)

        def step_host():
            ext.extract_batch_raw(h_frames.data_ptr(), B, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT, h_kp.data_ptr(), h_desc.data_ptr(), cap, h_counts.data_ptr(), device=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(W, 3)):
        step_device()
    torch.cuda.synchronize()
    # Two extractor handles per GPU ("one CUDA stream per camera stream, one extractor handle per stream"), each taking half of the
    # step's frames: the kernels of the two streams interleave on the SMs, which fills the tails / latency-bound stages of either one.
    NS = max(1, min(args.streams, B))
    exts = [ext] + [orbx.ORBextractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, device=local_rank) for _ in range(NS - 1)]
    streams = [torch.cuda.ExternalStream(e.stream) for e in exts]
    bounds = [(i * B // NS, (i + 1) * B // NS - i * B // NS) for i in range(NS)]          # (first frame, frame count) per handle

    def step_device2():
        for e, (b0, nb) in zip(exts, bounds):
            fo, ko, do_, co = b0 * WIDTH * HEIGHT, b0 * cap * 28, b0 * cap * 32, b0 * 4
            if MASKED:
                e.extract_masked_batch_raw_device(d_frames.data_ptr() + fo, d_masks.data_ptr() + fo, nb, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT, WIDTH, WIDTH * HEIGHT,
                                                  d_kp.data_ptr() + ko, d_desc.data_ptr() + do_, cap, d_counts.data_ptr() + co, d_culled.data_ptr() + co)
            else:
                e.extract_batch_raw(d_frames.data_ptr() + fo, nb, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT, d_kp.data_ptr() + ko, d_desc.data_ptr() + do_, cap,
                                    d_counts.data_ptr() + co, device=True)

    for _ in range(max(W, 3)):
        step_device2()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank); sampler.start()
    l0 = sum(e.launch_count for e in exts)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); ends = [torch.cuda.Event(enable_timing=True) for _ in exts]
    e0.record(streams[0])
    for st in streams[1:]:
        st.wait_event(e0)                         # every stream starts from the same point in time
    for _ in range(K):
        step_device2()
    for ev, st in zip(ends, streams):
        ev.record(st)
    barrier()
    ms = max(e0.elapsed_time(ev) for ev in ends)
    launches = sum(e.launch_count for e in exts) - l0
    # stage breakdown: a second, untimed-for-the-headline pass with the stages serialised on one stream and CUDA events between them
    ext.profile_enable(True)
    for _ in range(max(3, min(K, 10))):
        step_device()
    stage_ms, ncalls = ext.profile_collect()
    ext.profile_enable(False)
    n_kp = float(d_counts.float().mean().item())
    if ext.check_overflow():
        raise SystemExit("internal overflow flag set")

    # ---- e2e through the host-pointer C-ABI call: pinned host frames in, host keypoints / descriptors out, every step ----
    # Two extractor handles per GPU ("one CUDA stream per camera stream, one extractor handle per stream"): each is driven by its own
    # host thread through the synchronous call on its half of the step's batch, so one stream's transfers overlap the other's kernels.
    Ke = 0 if args.no_e2e else max(3, min(K, 10))
    ms_e2e = 0.0
    if Ke:
        import ctypes as C
        n_streams = NS
        halves = bounds

        def host_call(e, b0, nb):
            fo, ko, do_, co = b0 * WIDTH * HEIGHT, b0 * cap * 28, b0 * cap * 32, b0 * 4
            if MASKED:
                orbx._check(e._lib.orbx_extract_masked_batch(e._h, C.c_void_p(h_frames.data_ptr() + fo), C.c_void_p(h_masks.data_ptr() + fo), nb, HEIGHT, WIDTH, WIDTH,
                                                             WIDTH * HEIGHT, WIDTH, WIDTH * HEIGHT, C.c_void_p(h_kp.data_ptr() + ko), C.c_void_p(h_desc.data_ptr() + do_), cap,
                                                             C.c_void_p(h_counts.data_ptr() + co), C.c_void_p(h_culled.data_ptr() + co)))
            else:
                e.extract_batch_raw(h_frames.data_ptr() + fo, nb, HEIGHT, WIDTH, WIDTH, WIDTH * HEIGHT, h_kp.data_ptr() + ko, h_desc.data_ptr() + do_, cap,
                                    h_counts.data_ptr() + co, device=False)

        go = threading.Barrier(n_streams + 1); done = threading.Barrier(n_streams + 1); errs = []

        def worker(i):
            try:
                torch.cuda.set_device(local_rank)
                for _ in range(2):
                    host_call(exts[i], *halves[i])
                go.wait()
                for _ in range(Ke):
                    host_call(exts[i], *halves[i])
            except BaseException as ex:      # surface worker failures instead of reporting a bogus time
                errs.append(repr(ex)); go.abort(); done.abort(); return
            done.wait()

        ths = [threading.Thread(target=worker, args=(i,)) for i in range(n_streams)]
        for th in ths:
            th.start()
        try:
            go.wait()
            t0 = time.perf_counter()
            done.wait()
            ms_e2e = 1e3 * (time.perf_counter() - t0)          # both host calls are synchronous: wall clock covers H2D + kernels + D2H of every step
        except threading.BrokenBarrierError:
            pass
        for th in ths:
            th.join()
        if errs:
            raise SystemExit("e2e worker failed: " + errs[0])
        if world > 1:
            dist.barrier()
    sampler.stop_flag = True; sampler.join(timeout=2)

    # the links the e2e leg rides on: pinned host -> device copy rate of every rank's GPU, all ranks copying AT THE SAME TIME (barrier
    # before each repetition), so that at N > 1 the figure includes the contention on the host side (memory, root complexes) that the
    # e2e leg also sees; reported as the mean per GPU of the slowest repetition-best across ranks
    pcie_h2d = None; pcie_h2d_bi = None
    if Ke:
        hp = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); dp = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        half = hp.numel() // 2
        s2 = [torch.cuda.Stream(), torch.cuda.Stream()]
        best = 0.0
        for rep in range(4):                                      # two copy streams, like the e2e leg's two handles; best of 3 after one warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i, st in enumerate(s2):
                st.wait_event(e0)
                with torch.cuda.stream(st):
                    for _ in range(2):
                        dp[i * half:(i + 1) * half].copy_(hp[i * half:(i + 1) * half], non_blocking=True)
            for st in s2:
                torch.cuda.current_stream().wait_stream(st)
            e1.record(); torch.cuda.synchronize()
            if rep:
                best = max(best, 2 * hp.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        # the same with the result traffic running against it: a third stream copies device -> host in the e2e leg's proportion (1 : 5)
        hq = torch.empty(52 << 20, dtype=torch.uint8).pin_memory(); dq = torch.empty(52 << 20, dtype=torch.uint8, device="cuda")
        s3 = torch.cuda.Stream(); best_bi = 0.0
        for rep in range(3):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s3.wait_event(e0)
            with torch.cuda.stream(s3):
                for _ in range(2):
                    hq.copy_(dq, non_blocking=True)
            for i, st in enumerate(s2):
                st.wait_event(e0)
                with torch.cuda.stream(st):
                    for _ in range(2):
                        dp[i * half:(i + 1) * half].copy_(hp[i * half:(i + 1) * half], non_blocking=True)
            for st in s2:
                torch.cuda.current_stream().wait_stream(st)
            e1.record(); torch.cuda.synchronize()
            if rep:
                best_bi = max(best_bi, 2 * hp.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        del hp, dp, hq, dq
        tb = torch.tensor([best, best_bi], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        pcie_h2d = float(tb[0].item()) / world; pcie_h2d_bi = float(tb[1].item()) / world

    matcher_line = matcher_bench(orbx, torch, ext, frames, local_rank) if (rank == 0 and args.workload == "c1") else None

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0].item()), float(t[1].item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = N * B * K / (ms * 1e-3)
    e2e = N * B * Ke / (ms_e2e * 1e-3) if Ke else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    n_cand = float(sum(len(ext.debug_level_candidates(0, l)) for l in range(NLEVELS)))   # frame 0 of the last batch
    sb = stage_bytes(n_kp, n_cand)
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    launches_of = {"pyr_resize": NLEVELS - 1}
    dom_ms_per_launch = stage_ms[dom] / max(ncalls, 1) / launches_of.get(dom, 1)
    dom_bytes_per_launch = sb[dom] * B / launches_of.get(dom, 1)
    achieved = dom_bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9
    total_stage = sum(stage_ms.values())
    # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum per frame from the committed ncu --set full
    # capture (profiles/traffic.json, written by tools/profile_digest.py), scaled to this launch's frame count
    kernel_of = {"pyr_resize": "k_pyr_resize_t", "fast_cells": "k_fast_cells", "octree_sort": "k_octree_fused", "octree_tree": "k_octree_tree_par",
                 "gauss7": "k_gauss7", "orient_describe": "k_orient_describe"}
    traffic = None; ncu_note = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        kname = kernel_of[dom] if kernel_of[dom] in tj["dram_bytes_per_frame"] else kernel_of[dom] + "_t"      # templated kernels appear with a _t suffix in newer captures
        traffic = tj["dram_bytes_per_frame"][kname] * B / launches_of.get(dom, 1)
        ncu_note = {"source": "profiles/" + tj["source"], "issue_active_pct": tj["issue_active_pct"][kname], "dram_throughput_pct": tj["dram_throughput_pct"][kname],
                    "reading": "the kernel is bound by instruction issue (integer byte work), not by HBM: traffic ~= algorithmic bytes, DRAM a few % busy"}
    except Exception:
        pass
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": N, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(B * link_bytes_per_frame), "d2h_bytes_per_step": int(B * cap * 60 + 4 * B + 4), "steps": Ke,
                    "host_input_bytes_per_step": int(B * WIDTH * HEIGHT * (2 if MASKED else 1)),
                    "mask_transport": (("packed to 1 bit per pixel on the host by %s worker threads per handle (the masks only matter as != 0), 1/8 of the bytes on the link" % os.environ.get("ORBX_HOST_PACK")) if (MASKED and hostpack) else ("bytes" if MASKED else None)),
                    "streams_per_gpu": NS,
                    "h2d_GBps_per_gpu": (e2e / N * link_bytes_per_frame / 1e9) if e2e else None, "pcie_h2d_peak_GBps": pcie_h2d, "pcie_h2d_with_d2h_GBps": pcie_h2d_bi, "pcie_probe": "pinned 256 MiB H2D on two streams per GPU, all %d ranks concurrently, mean per GPU; _with_d2h: a third stream copies results device -> host at the e2e leg's 1:5 ratio" % world,
                    "pcie_frac": (e2e / N * link_bytes_per_frame / 1e9 / pcie_h2d_bi) if (e2e and pcie_h2d_bi) else None,   # against the probe WITH result traffic
                    "note": "one host thread per extractor handle, each calling the synchronous host-pointer batch API on its share of the step's frames"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "ncu": ncu_note,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms_per_launch,
                         "stage_ms_per_step": {k: v / max(ncalls, 1) for k, v in stage_ms.items()},
                         "stage_share": {k: (v / total_stage if total_stage else 0.0) for k, v in stage_ms.items()},
                         "stage_note": "octree_sort = the whole quadtree stage (k_octree_fused: gather, sort and tree in one launch); octree_tree is non-zero only when the sort + tree pair runs (ORBX_QT_FUSED=0 or more than ~9000 features); on the masked path gauss7 includes the closing of the masks and the culling",
                         "whole_step_algorithmic_GBps": (2 * sum(level_pixels()) + 60 * n_kp) * B * K / (ms * 1e-3) / 1e9},
            "keypoints_per_frame": n_kp, "matcher": matcher_line, "host_affinity_rank0": numa}
    if N == 1 and not args.no_cpu_baseline:
        res = cpu_reference_run(args.cpu_seconds, frames[:64])
        line["cpu_baseline"] = {"value": res["fps"], "unit": "frames/s", "cores": res["threads"], "kind": res["kind"],
                                "sample": "%d frames in %.1f s, one independent extractor per thread on %d threads, CPU: %s%s" % (
                                    res["frames"], res["seconds"], res["threads"], cpu_model(),
                                    "; NOTE: the shim's erode/dilate is an unoptimised 729-tap loop, so this c5 CPU figure is not representative of OpenCV's morphology" if MASKED else "")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
