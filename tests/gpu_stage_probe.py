"""Stage-by-stage GPU vs oracle probe (run on the GPU box: python tests/gpu_stage_probe.py)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
orbx = importlib.import_module("amos-slam_b200")

def main():
    P = oracle.Extractor('port')
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    ok = True
    for seed in range(3):
        img = oracle.synth_frame(seed)
        kp_o, d_o = P.extract(img)
        kp_g, d_g = E(img)
        print("seed", seed, "n oracle", len(kp_o), "n gpu", len(kp_g))
        for l in range(8):
            po = P.pyramid_level(l)
            pg = E.debug_pyramid_level(0, l, po.shape)
            co = P.level_candidates(l)
            cg = E.debug_level_candidates(0, l)
            same_c = len(co) == len(cg) and np.array_equal(co['x'], cg['x']) and np.array_equal(co['y'], cg['y']) and np.array_equal(co['response'], cg['response'])
            print("  level", l, "pyr diff", int((po != pg).sum()), "cand", len(co), len(cg), "equal", same_c)
            ok &= (po == pg).all() and same_c
        if len(kp_o) == len(kp_g):
            for f in kp_o.dtype.names:
                eq = np.array_equal(kp_o[f], kp_g[f])
                if not eq:
                    bad = np.nonzero(kp_o[f] != kp_g[f])[0]
                    print("  field", f, "mismatch at", bad[:10], kp_o[f][bad[:5]], kp_g[f][bad[:5]])
                ok &= eq
            dd = (d_o != d_g).any(1).sum()
            print("  desc rows differing", int(dd))
            ok &= dd == 0
        else:
            ok = False
    print("ALL OK" if ok else "MISMATCH")
    # quick batch timing
    B = 64
    imgs = np.stack([oracle.synth_frame(100 + i) for i in range(B)])
    kp, desc, counts = E.extract_batch(imgs)
    t = time.time(); kp, desc, counts = E.extract_batch(imgs); dt = time.time() - t
    print("batch", B, "host e2e ms", dt * 1e3, "fps", B / dt, "counts", counts[:8], "overflow", E.check_overflow())
    k0, d0 = P.extract(imgs[5])
    print("batch frame 5 equal:", np.array_equal(kp[5][:counts[5]], k0), np.array_equal(desc[5][:counts[5]], d0))

if __name__ == "__main__":
    main()
