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
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS), help="BASELINE.json config (c1 = headline metric)")
    ap.add_argument("--streams", type=int, default=2, help="extractor handles (camera streams) per GPU; the step's frames are split evenly between them")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-pointer leg (the line is then not a bench value)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    global WIDTH, HEIGHT, NFEAT, WORKLOAD, METRIC, MASKED
    WIDTH, HEIGHT, NFEAT, defB, MASKED, WORKLOAD, METRIC = WORKLOADS[args.workload]
    N, K, W, B = args.gpus, args.steps, max(args.warmup, 0), (args.batch or defB)
    config = {"workload": WORKLOAD, "frame": [WIDTH, HEIGHT], "nfeatures": NFEAT, "frames_per_gpu_per_step": B,
              "parallelism": "frames sharded over %d GPU(s), no collective; %d extractor handles (streams) per GPU sharing the step's frames evenly" % (N, max(1, args.streams)),
              "l2_policy": "working set per step (%.0f MB of frames, ~%.1f GB of pyramid+scratch) exceeds the 126 MB L2" % (B * WIDTH * HEIGHT / 1e6, B * 6.3e-3 * WIDTH * HEIGHT / 307200.0)}

    if args.impl == "reference":
        if rank != 0:
            return
        frames = synth_batch(64, WIDTH, HEIGHT, seed0=0, distinct=16 if WIDTH * HEIGHT < 1000000 else 4)
        threads = max(1, os.cpu_count() or 1)
        FPT = 8            # frames per thread per step: long enough that thread start/join is noise, short enough for K steps in seconds
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

    import torch
    import torch.distributed as dist
    orbx = importlib.import_module("amos-slam_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    frames = synth_batch(B, WIDTH, HEIGHT, seed0=1000 * rank, distinct=24 if WIDTH * HEIGHT < 1000000 else 8)
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
    kernel_of = {"pyr_resize": "k_pyr_resize_w", "fast_cells": "k_fast_cells", "octree_sort": "k_octree_sort", "octree_tree": "k_octree_tree_par",
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
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(B * WIDTH * HEIGHT * (2 if MASKED else 1)), "d2h_bytes_per_step": int(B * cap * 60 + 4 * B + 4), "steps": Ke,
                    "streams_per_gpu": NS,
                    "h2d_GBps_per_gpu": (e2e / N * WIDTH * HEIGHT * (2 if MASKED else 1) / 1e9) if e2e else None, "pcie_h2d_peak_GBps": pcie_h2d, "pcie_h2d_with_d2h_GBps": pcie_h2d_bi, "pcie_probe": "pinned 256 MiB H2D on two streams per GPU, all %d ranks concurrently, mean per GPU; _with_d2h: a third stream copies results device -> host at the e2e leg's 1:5 ratio" % world,
                    "pcie_frac": (e2e / N * WIDTH * HEIGHT * (2 if MASKED else 1) / 1e9 / pcie_h2d_bi) if (e2e and pcie_h2d_bi) else None,   # against the probe WITH result traffic
                    "note": "one host thread per extractor handle, each calling the synchronous host-pointer batch API on its share of the step's frames"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "ncu": ncu_note,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms_per_launch,
                         "stage_ms_per_step": {k: v / max(ncalls, 1) for k, v in stage_ms.items()},
                         "stage_share": {k: (v / total_stage if total_stage else 0.0) for k, v in stage_ms.items()},
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
