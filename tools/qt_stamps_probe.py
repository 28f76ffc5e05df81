"""Phase times of the quadtree kernels for ONE 640 x 480 frame (the latency form Tracking calls), from clock64 stamps compiled into a
PROBE build of the library (-DORBX_QT_STAMPS -> amos-slam_b200/liborbx_b200_stamps.so; the product library carries no stamps).
usage: python tools/qt_stamps_probe.py   (builds the probe library when it is missing; run on the GPU box)"""
import ctypes as C, importlib, os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tools.synth import synth_frame
orbx = importlib.import_module("amos-slam_b200")
PROBE = os.path.join(orbx.HERE, "liborbx_b200_stamps.so")


def build_probe():
    srcs = [os.path.join(orbx.CSRC, s) for s in orbx.SOURCES]
    subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + orbx.NVCC_FLAGS + ["-DORBX_QT_STAMPS", "-o", PROBE] + srcs)


if __name__ == "__main__":
    if "--build" in sys.argv or not os.path.exists(PROBE):
        build_probe()
        if "--build" in sys.argv:
            sys.exit(0)
    orbx.LIB_PATH = PROBE
    L = orbx.lib()
    L.orbx_probe_qt_stamps.argtypes = [C.c_void_p, C.c_int]
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    for seed in (0, 1, 2):
        A = synth_frame(seed, 640, 480)
        for _ in range(5):
            kp, _d = E(A)
        st = np.zeros(64, np.int64)
        orbx._check(L.orbx_probe_qt_stamps(st.ctypes.data, 64))
        lv0 = int((kp["octave"] == 0).sum())
        fused = os.environ.get("ORBX_QT_FUSED", "1") != "0"
        if fused:
            names = {0: "start", 1: "count scan", 3: "gather + codes", 4: "bucket sort", 16: "dd + histograms", 17: "nodes after the full passes", 24: "  round 0: rank", 25: "  round 0: children",
                     26: "  round 0: scan + cut", 18: "  round 0: emit", 30: "final: compact", 31: "final: rank", 39: "final: select"}
            idx = [0, 1, 3, 4, 5, 6, 7, 16, 17, 24, 25, 26, 18, 19, 20, 21, 22, 23, 30, 31, 39]
        else:
            names = {0: "start", 1: "count scan", 2: "gather", 3: "codes", 10: "write-out", 16: "tree start", 17: "codes to smem", 18: "roots", 38: "loop end", 39: "select"}
            idx = list(range(0, 11)) + list(range(16, 40))
        print("seed %d: %d keypoints (%d on level 0); cycles between stamps of the level-0 CTA (1 cycle = ~0.52 ns at 1.92 GHz)" % (seed, len(kp), lv0))
        prev = None
        for i in idx:
            if st[i] == 0 or st[i] < st[0]:
                continue
            if fused:
                lab = names.get(i, "radix pass %d" % (i - 5) if i < 16 else "largest-first round %d" % (i - 18))
            else:
                lab = names.get(i, "radix pass %d" % (i - 4) if i < 10 else "round %d" % (i - 20))
            if prev is not None and i not in (0,) and not (i == 16 and not fused) and st[i] >= prev:
                print("  %-28s %7d cycles  %6.2f us" % (lab, st[i] - prev, (st[i] - prev) * 0.52e-3))
            prev = st[i]
        if fused:
            print("  fused kernel total %.2f us" % ((st[39] - st[0]) * 0.52e-3))
            continue
        print("  sort kernel total %.2f us, tree kernel total %.2f us" % ((st[10] - st[0]) * 0.52e-3, (st[39] - st[16]) * 0.52e-3))
