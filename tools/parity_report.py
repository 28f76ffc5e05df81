#!/usr/bin/env python3
"""Parity counters of SURVEY.md 8(d) "Reporting": the CUDA path through the C ABI against the reference's own code
(oracle/_ref: the reference sources compiled against the OpenCV-free shim; the port restatement if _ref is absent) on many
seeded frames per configuration.  Run on a B200:  python tools/parity_report.py [--frames N] > gpurun_out/parity.md
Test infrastructure: uses oracle/ as the checker only."""
import argparse, importlib, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle                                   # noqa: E402
import match_cases as mc                        # noqa: E402
import frame_cases as fc                        # noqa: E402
import bow_cases as bc                          # noqa: E402
from tools.synth import synth_frame, warp_affine_nn   # noqa: E402

orbx = importlib.import_module("amos-slam_b200")
FIELDS = ("x", "y", "size", "angle", "response", "octave", "class_id")


def extractor_rows(name, w, h, nfeat, frames, kind):
    E = orbx.ORBextractor(nfeat, 1.2, 8, 20, 7); O = oracle.Extractor(kind, nfeat, 1.2, 8, 20, 7)
    tot = dict(frames=0, kp_ref=0, kp_gpu=0, count_mismatch=0, desc_bytes_diff=0, angle_max_abs=0.0, sincos_sensitive=0, sincos_sensitive_desc_diff=0, angle_diff_kp=0, **{f: 0 for f in FIELDS})
    t_ref = 0.0
    for s in range(frames):
        img = synth_frame(1000 + s, w, h)
        kg, dg = E(img)
        t0 = time.perf_counter(); kr, dr = O.extract(img); t_ref += time.perf_counter() - t0
        tot["frames"] += 1; tot["kp_ref"] += len(kr); tot["kp_gpu"] += len(kg)
        if len(kg) != len(kr):
            tot["count_mismatch"] += 1; continue
        for f in FIELDS:
            tot[f] += int((kg[f] != kr[f]).sum())
        tot["desc_bytes_diff"] += int((dg != dr).sum())
        # keypoints whose rotation is sincos-SENSITIVE: glibc's sincosf (what the reference calls at :181) differs from the correctly rounded
        # sin / cos of angle * factorPI.  They are listed separately: a right-in-the-maths sin/cos may flip their sample coordinates.
        rad = kr["angle"] * np.float32(np.pi / 180.0)
        ls, lc = np.zeros_like(rad), np.zeros_like(rad); oracle.port_lib().port_libm_sincosf(np.ascontiguousarray(rad), ls, lc, len(rad))
        sens = (ls != np.sin(rad.astype(np.float64)).astype(np.float32)) | (lc != np.cos(rad.astype(np.float64)).astype(np.float32))
        tot["sincos_sensitive"] += int(sens.sum()); tot["sincos_sensitive_desc_diff"] += int((dg[sens] != dr[sens]).any(axis=1).sum())
        tot["angle_diff_kp"] += int((kg["angle"] != kr["angle"]).sum())
        if len(kg):
            tot["angle_max_abs"] = max(tot["angle_max_abs"], float(np.abs(kg["angle"] - kr["angle"]).max()))
    tot["ref_ms_per_frame_1core"] = 1e3 * t_ref / max(frames, 1)
    return name, tot


def random_config_rows(n_cfg, kind, seed=77):
    """Randomised extractor configurations (image size, feature count, scale factor, levels, FAST thresholds, row padding): every field of every keypoint
    and every descriptor byte against the reference build; configurations the reference itself cannot run (a level too small for its cell grid) are skipped."""
    rng = np.random.default_rng(seed)
    out = dict(configs=0, skipped=0, frames=0, keypoints=0, diff_fields=0, diff_desc=0, count_mismatch=0, worst="")
    tries = 0
    while out["configs"] < n_cfg and tries < 4 * n_cfg:
        tries += 1
        w = int(rng.integers(96, 1400)); h = int(rng.integers(96, 900)); nf = int(rng.choice([150, 500, 1000, 2000, 3000]))
        sf = float(rng.choice([1.1, 1.2, 1.25, 1.3, 1.5, 2.0])); nl = int(rng.integers(1, 9)); ini = int(rng.choice([10, 20, 30, 45])); mn = int(rng.choice([3, 5, 7, 12]))
        wl, hl = w / sf ** (nl - 1), h / sf ** (nl - 1)                    # smallest level: needs a cell grid and round(width / height) >= 1 quadtree roots (:719)
        if mn > ini or min(wl, hl) < 64 or (wl - 32) / (hl - 32) < 0.6:
            out["skipped"] += 1; continue
        try:
            E = orbx.ORBextractor(nf, sf, nl, ini, mn); O = oracle.Extractor(kind, nf, sf, nl, ini, mn)
        except Exception:
            out["skipped"] += 1; continue
        out["configs"] += 1
        for s in range(2):
            img = synth_frame(7000 + 10 * out["configs"] + s, w, h)
            if s == 1:                                                     # a padded row stride (cv::Mat ROI)
                buf = np.zeros((h, w + 13), np.uint8); buf[:, :w] = img; img_g = buf[:, :w]
            else:
                img_g = img
            try:
                kg, dg = E(img_g)
            except orbx.OrbxError:
                out["skipped"] += 1; break
            kr, dr = O.extract(img)
            out["frames"] += 1; out["keypoints"] += len(kr)
            if len(kg) != len(kr):
                out["count_mismatch"] += 1; out["worst"] = "%dx%d nf=%d sf=%g nl=%d th=%d/%d" % (w, h, nf, sf, nl, ini, mn); continue
            d = sum(int((kg[f] != kr[f]).sum()) for f in FIELDS)
            out["diff_fields"] += d; out["diff_desc"] += int((dg != dr).sum())
            if d:
                out["worst"] = "%dx%d nf=%d sf=%g nl=%d th=%d/%d" % (w, h, nf, sf, nl, ini, mn)
    return out


def matcher_rows(pairs, kind):
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    sf = E.GetScaleFactors()
    out = dict(pairs=0, init_queries=0, init_idx_diff=0, init_count_diff=0, projf_idx_diff=0, projf_count_diff=0, projp_idx_diff=0, projp_count_diff=0, matches=0)
    for s in range(pairs):
        A = synth_frame(2000 + s, 640, 480); B = warp_affine_nn(A, 7, -4, 2.0)
        ka, da = E(A); kb, db = E(B)
        pi = mc.projection_inputs(ka, kb, seed=s + 1)
        FA, FB = orbx.FrameView(ka, da, 640, 480, sf), orbx.FrameView(kb, db, 640, 480, sf, u_right=pi["u_right"])
        oFA, oFB = oracle.FrameData(ka, da, 640, 480, sf), oracle.FrameData(kb, db, 640, 480, sf, u_right=pi["u_right"])
        prev = np.stack([ka["x"], ka["y"]], 1)
        g = orbx.ORBmatcher(0.9, True).SearchForInitialization(FA, FB, prev, 100)
        o = oracle.Matcher(kind, 0.9, True).search_for_initialization(oFA, oFB, prev, 100)
        out["pairs"] += 1; out["init_queries"] += len(ka); out["matches"] += int(o[0])
        out["init_idx_diff"] += int((g[1] != o[1]).sum()) + int((g[2] != o[2]).any(1).sum()); out["init_count_diff"] += int(g[0] != o[0])
        uv, iz = mc.project(pi["xyz"])
        for th, mono in mc.PROJ_FRAME_CASES:
            g = orbx.ORBmatcher(0.9, True).SearchByProjectionFrame(FB, uv, iz, ka["octave"], ka["angle"], da, pi["valid"], pi["obs"], pi["occ"], th, False, False, 40.0)
            o = oracle.Matcher("port", 0.9, True).search_by_projection_frame_port(oFB, uv, iz, ka["octave"], ka["angle"], da, pi["valid"], pi["obs"], pi["occ"], th, False, False, 40.0)
            out["projf_idx_diff"] += int((g[1] != o[1]).sum()); out["projf_count_diff"] += int(g[0] != o[0]); out["matches"] += int(o[0])
        for th in mc.PROJ_POINT_CASES:
            g = orbx.ORBmatcher(0.8, True).SearchByProjectionPoints(FB, pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], da, pi["obs"], pi["occ"], th)
            o = oracle.Matcher(kind, 0.8, True).search_by_projection_points(oFB, pi["tuv"], pi["tur"], pi["lvl"], pi["vc"], da, pi["obs"], pi["occ"], th)
            out["projp_idx_diff"] += int((g[1] != o[1]).sum()); out["projp_count_diff"] += int(g[0] != o[0]); out["matches"] += int(o[0])
    return out


def stereo_rows(pairs, kind):
    out = dict(pairs=0, left_keypoints=0, matched=0, u_right_diff=0, depth_diff=0)
    EL, ER = orbx.ORBextractor(2000, 1.2, 8, 20, 7), orbx.ORBextractor(2000, 1.2, 8, 20, 7)
    OL, OR = oracle.Extractor(kind, 2000, 1.2, 8, 20, 7), oracle.Extractor(kind, 2000, 1.2, 8, 20, 7)
    for s in range(pairs):
        L, R = mc.stereo_pair(seed=3 + 2 * s)
        kl, dl = EL(L); kr, dr = ER(R)
        okl, odl = OL.extract(L); okr, odr = OR.extract(R)
        ur, dep = orbx.ORBmatcher().ComputeStereoMatches(EL, ER, kl, dl, kr, dr, 0.0, mc.BF_KITTI)
        our, odep = oracle.Matcher(kind).compute_stereo_matches(OL, OR, okl, odl, okr, odr, 0.0, mc.BF_KITTI)
        out["pairs"] += 1; out["left_keypoints"] += len(okl); out["matched"] += int((our >= 0).sum())
        if len(ur) != len(our):
            out["u_right_diff"] += max(len(ur), len(our)); continue
        out["u_right_diff"] += int((ur != our).sum()); out["depth_diff"] += int((dep != odep).sum())
    return out


def frame_rows(frames, kind):
    out = dict(frames=0, keypoints=0, keys_un_diff=0, u_right_diff=0, depth_diff=0, grid_diff=0, bounds_diff=0)
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7); F = orbx.Frame(); dimg = fc.depth_image()
    for s in range(frames):
        name = ("tum1", "tum2", "four", "tum3")[s % 4]
        k, d = E(synth_frame(3000 + s, 640, 480))
        F.assign(E, fc.cam_struct(orbx, name), 480, 640, dimg)
        o = oracle.frame_build(kind, k, fc.CAMS[name], fc.BF, 480, 640, dimg)
        ku, ur, dp, b = F.read(); cs, en = F.grid()
        out["frames"] += 1; out["keypoints"] += len(k)
        out["keys_un_diff"] += int((ku != o["keys_un"]).sum()); out["u_right_diff"] += int((ur != o["u_right"]).sum()); out["depth_diff"] += int((dp != o["depth"]).sum())
        out["grid_diff"] += int(not (np.array_equal(cs, o["cell_start"]) and np.array_equal(en, o["entries"]))); out["bounds_diff"] += int((b != o["bounds"]).sum())
    return out


def bow_rows(pairs, kind):
    out = dict(frames=0, features=0, word_diff=0, node_diff=0, bow_id_diff=0, bow_val_diff=0, fv_diff=0, bow_entries=0, sb_calls=0, sb_matches=0, sb_idx_diff=0, sb_count_diff=0)
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    pool = np.concatenate([E(synth_frame(500 + s, 640, 480))[1] for s in range(6)])
    voc = bc.build_vocabulary(pool, 10, 3)
    V = orbx.ORBVocabulary(10, 3, *voc); O = oracle.Vocabulary(kind, 10, 3, *voc)
    keys = ("word", "node", "bow_ids", "bow_vals", "fv_nodes", "fv_offsets", "fv_idx")
    for s in range(pairs):
        A = synth_frame(4000 + s, 640, 480); B = warp_affine_nn(A, 7, -4, 2.0)
        ka, da = E(A); kb, db = E(B)
        res = []
        for d in (da, db):
            g, o = V.transform(d, 1), O.transform(d, 1)
            out["frames"] += 1; out["features"] += len(d); out["bow_entries"] += len(o["bow_ids"])
            out["word_diff"] += int((g["word"] != o["word"]).sum()); out["node_diff"] += int((g["node"] != o["node"]).sum())
            same_len = len(g["bow_ids"]) == len(o["bow_ids"])
            out["bow_id_diff"] += int((g["bow_ids"] != o["bow_ids"]).sum()) if same_len else max(len(g["bow_ids"]), len(o["bow_ids"]))
            out["bow_val_diff"] += int((g["bow_vals"] != o["bow_vals"]).sum()) if same_len else max(len(g["bow_ids"]), len(o["bow_ids"]))
            out["fv_diff"] += int(not all(np.array_equal(g[k], o[k]) for k in ("fv_nodes", "fv_offsets", "fv_idx")))
            res.append(o)
        va, vb = bc.validity(len(ka), s + 1), bc.validity(len(kb), s + 100)
        for kfkf in (0, 1):
            for nn, ori in bc.MATCH_VARIANTS:
                g = orbx.ORBmatcher(nn, ori).SearchByBoW(kfkf, ka, da, va, res[0], kb, db, vb if kfkf else None, res[1])
                o = oracle.search_by_bow(kind, nn, ori, kfkf, ka, da, va, res[0], kb, db, vb, res[1])
                out["sb_calls"] += 1; out["sb_matches"] += int(o[0]); out["sb_count_diff"] += int(g[0] != o[0])
                out["sb_idx_diff"] += int((g[1] != o[1]).sum()) + int((g[2] != o[2]).sum())
    # CPU time of the reference's transform at the size of ORBvoc.txt (k = 10, L = 6, random node descriptors)
    rng = np.random.default_rng(5); nn_ = (10 ** 7 - 10) // 9; ids = np.arange(1, nn_ + 1, dtype=np.int64)
    leaf = (ids > (10 ** 6 - 10) // 9).astype(np.uint8)
    big = oracle.Vocabulary(kind, 10, 6, ((ids - 1) // 10).astype(np.int32), leaf, rng.integers(0, 256, (nn_, 32), dtype=np.uint8), np.where(leaf > 0, rng.uniform(0.5, 9.0, nn_), 0.0))
    big.transform(da, 4)
    t0 = time.perf_counter()
    for _ in range(5):
        big.transform(da, 4)
    out["ref_transform_ms"] = 1e3 * (time.perf_counter() - t0) / 5
    return out


def rank3_rows(pairs):
    """The rank-3 matchers and the batched descriptor selection against the port oracle (itself pinned to the reference bodies by the golden tests)."""
    out = dict(pairs=0, kf=[0, 0], kfp=[0, 0], sim3=[0, 0], fuse=[0, 0], tri=[0, 0], dist=[0, 0])
    E = orbx.ORBextractor(1000, 1.2, 8, 20, 7); sf = E.GetScaleFactors()
    pool = np.concatenate([E(synth_frame(500 + s, 640, 480))[1] for s in range(6)])
    voc = bc.build_vocabulary(pool, 10, 3); V = orbx.ORBVocabulary(10, 3, *voc)
    P = oracle.Matcher("port", 0.9, True)
    for s in range(pairs):
        A = synth_frame(5000 + s, 640, 480); B = warp_affine_nn(A, 7, -4, 2.0)
        ka, da = E(A); kb, db = E(B)
        pi = mc.projection_inputs(ka, kb, seed=s + 1)
        FA, FB = orbx.FrameView(ka, da, 640, 480, sf), orbx.FrameView(kb, db, 640, 480, sf, u_right=pi["u_right"])
        oFA, oFB = oracle.FrameData(ka, da, 640, 480, sf), oracle.FrameData(kb, db, 640, 480, sf, u_right=pi["u_right"])
        uv, _ = mc.project(pi["xyz"])
        kf = mc.keyframe_inputs(ka, kb, pi, seed=s + 9)
        for th, od, ori in mc.KF_CASES:
            g = orbx.ORBmatcher(0.9, ori).SearchByProjectionKeyFrame(FB, uv, kf["lvl"], ka["angle"], da, kf["valid"], kf["occ"], th, od)
            o = oracle.Matcher("port", 0.9, ori).search_by_projection_keyframe_port(oFB, uv, kf["lvl"], ka["angle"], da, kf["valid"], kf["occ"], th, od)
            out["kf"][0] += int(o[0]); out["kf"][1] += int(g[0] != o[0]) + int((g[1] != o[1]).sum())
        kp = mc.keyframe_points_inputs(ka, kb, pi, seed=s + 13)
        for th in mc.KFP_CASES:
            g = orbx.ORBmatcher().SearchByProjectionKeyFramePoints(FB, kp["uv"], kp["lvl"], da, kp["valid"], kp["kf_matched"], th)
            o = P.search_by_projection_keyframe_points_port(oFB, kp["uv"], kp["lvl"], da, kp["valid"], kp["kf_matched"], th)
            out["kfp"][0] += int(o[0]); out["kfp"][1] += int(g[0] != o[0]) + int((g[1] != o[1]).sum())
        s1, s2, _ = mc.sim3_inputs(ka, kb, seed=s + 21)
        for th in mc.SIM3_TH:
            g = orbx.ORBmatcher().SearchBySim3(FA, FB, s1["uv"], s1["lvl"], da, s1["valid"], s2["uv"], s2["lvl"], db, s2["valid"], th)
            o = P.search_by_sim3_port(oFA, oFB, s1["uv"], s1["lvl"], da, s1["valid"], s2["uv"], s2["lvl"], db, s2["valid"], th)
            out["sim3"][0] += int(o[0]); out["sim3"][1] += int(g[0] != o[0]) + int((g[1] != o[1]).sum())
        fu = mc.fuse_inputs(ka, kb, pi, seed=s + 29)
        for th in mc.FUSE_TH:
            for sim3 in (0, 1):
                g = orbx.ORBmatcher().FuseSearch(FB, fu["uv"], None if sim3 else fu["ur"], fu["lvl"], da, fu["valid"], fu["inv_sigma2"], th)
                o = P.fuse_search_port(oFB, fu["uv"], None if sim3 else fu["ur"], fu["lvl"], da, fu["valid"], fu["inv_sigma2"], th)
                out["fuse"][0] += int((o >= 0).sum()); out["fuse"][1] += int((g != o).sum())
        fa, fb = V.transform(da, 1), V.transform(db, 1)
        t = bc.tri_inputs(len(ka), len(kb), seed=s + 31)
        for ori, st in bc.TRI_VARIANTS:
            g = orbx.ORBmatcher(0.6, ori).SearchForTriangulation(ka, da, t["free1"], t["ur1"], fa, kb, db, t["free2"], t["ur2"], fb, bc.TRI_F12, (339.292, 265.63), t["sf"], t["sigma2"], st)
            o = oracle.search_for_triangulation("port", ori, ka, da, t["free1"], t["ur1"], fa, kb, db, t["free2"], t["ur2"], fb, bc.TRI_F12, (339.292, 265.63), bc.TRI_CAM, t["sf"], t["sigma2"], st)
            out["tri"][0] += int(o[0]); out["tri"][1] += int(g[0] != o[0]) + int((g[1] != o[1]).sum())
        di = bc.distinctive_inputs(np.concatenate([da, db]), seed=s + 41)
        g = orbx.ORBmatcher().ComputeDistinctiveDescriptors(di["f_offsets"], di["f_desc"]); o = oracle.distinctive_descriptors("port", di["f_offsets"], di["f_desc"])
        out["dist"][0] += len(o); out["dist"][1] += int((g != o).sum())
        out["pairs"] += 1
    return out


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=100); a = ap.parse_args()
    kind = "ref" if oracle.have_ref() else "port"
    oracle.build_port()
    print("# Parity report (GPU through the C ABI vs `oracle/%s`)\n" % ("_ref: the reference's own sources" if kind == "ref" else "port"))
    print("Seeded synthetic frames (`tools/synth.py`, seeds 1000+i); every number below is a count of DIFFERENCES unless named otherwise.\n")
    print("## Extractor `operator()` (a1-a10)\n")
    print("| config | frames | keypoints (ref) | keypoints (gpu) | frames with different count | x | y | size | angle | response | octave | descriptor bytes | max abs angle diff | keypoints with ANY angle difference (would be listed as angle-sensitive) | sincos-sensitive keypoints (glibc sincosf != correctly rounded) | of those: descriptor differs | ref ms/frame (1 core) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for name, w, h, nf, n in (("C1 640x480/1000", 640, 480, 1000, a.frames), ("C3 752x480/2000", 752, 480, 2000, max(a.frames // 2, 1)),
                              ("C4 1241x376/2000", 1241, 376, 2000, max(a.frames // 4, 1)), ("C5 1920x1080/1000", 1920, 1080, 1000, max(a.frames // 10, 1))):
        _, t = extractor_rows(name, w, h, nf, n, kind)
        print("| %s | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %g | %d | %d | %d | %.1f |" % (name, t["frames"], t["kp_ref"], t["kp_gpu"], t["count_mismatch"], t["x"], t["y"], t["size"], t["angle"],
                                                                                            t["response"], t["octave"], t["desc_bytes_diff"], t["angle_max_abs"], t["angle_diff_kp"], t["sincos_sensitive"],
                                                                                            t["sincos_sensitive_desc_diff"], t["ref_ms_per_frame_1core"]))
    print("\nThe rotation of `computeOrbDescriptor` is glibc's `sincosf` (FMA variant) restated operation by operation on the device (`det_math.cuh`); it equals the live libm "
          "for every float (`tests/test_oracle_cvlite.py`, `tests/test_gpu_extractor.py::test_device_sincos_is_glibc_sincosf`), so sincos-sensitive keypoints agree by construction, "
          "and 120 000 adversarial angles are pushed through `ProcessDesp` in `test_descriptors_at_sincos_sensitive_angles`.")
    rc = random_config_rows(max(a.frames // 3, 5), kind)
    print("\n### Randomised configurations (size 96-1400 x 96-900, 150-3000 features, scale factor 1.1-2.0, 1-8 levels, FAST thresholds 10-45 / 3-12, one frame with a padded row stride)\n")
    print("| configurations | frames | keypoints (ref) | frames with different count | keypoint field diffs | descriptor byte diffs | skipped (geometry the reference cannot run) |\n|---|---|---|---|---|---|---|")
    print("| %d | %d | %d | %d | %d | %d | %d |%s" % (rc["configs"], rc["frames"], rc["keypoints"], rc["count_mismatch"], rc["diff_fields"], rc["diff_desc"], rc["skipped"],
                                                     (" first differing configuration: " + rc["worst"]) if rc["worst"] else ""))
    m = matcher_rows(max(a.frames // 5, 1), kind)
    print("\n## Matchers (a13-a17), C2-style pairs (frame B = frame A warped by (+7, -4) px, 2 deg)\n")
    print("| pairs | F1 queries | accepted matches (all calls) | SearchForInitialization: index/prev diffs, count diffs | SearchByProjection(Frame,Frame) x3 th: index, count | SearchByProjection(Frame,MapPoints) x3 th: index, count |")
    print("|---|---|---|---|---|---|")
    print("| %d | %d | %d | %d, %d | %d, %d | %d, %d |" % (m["pairs"], m["init_queries"], m["matches"], m["init_idx_diff"], m["init_count_diff"], m["projf_idx_diff"], m["projf_count_diff"],
                                                         m["projp_idx_diff"], m["projp_count_diff"]))
    s = stereo_rows(max(a.frames // 20, 1), kind)
    print("\n## ComputeStereoMatches (a18), C4 pairs\n")
    print("| pairs | left keypoints | matched (ref) | mvuRight diffs | mvDepth diffs |\n|---|---|---|---|---|")
    print("| %d | %d | %d | %d | %d |" % (s["pairs"], s["left_keypoints"], s["matched"], s["u_right_diff"], s["depth_diff"]))
    w = bow_rows(max(a.frames // 5, 1), kind)
    print("\n## Bag of words (8f rank 2): DBoW2 transform (levelsup 1 on a k = 10, L = 3 tree grown from real descriptors) and SearchByBoW, both forms x 3 (ratio, orientation) settings\n")
    print("| frames | features | word id diffs | node id diffs | BowVector entries | BowVector id diffs | BowVector weight diffs (double, exact) | frames with a different FeatureVector | SearchByBoW calls | matches (ref) | match index diffs | count diffs | ref transform ms / 1000 features at ORBvoc size (1 core) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    print("| %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %.2f |" % (w["frames"], w["features"], w["word_diff"], w["node_diff"], w["bow_entries"], w["bow_id_diff"], w["bow_val_diff"], w["fv_diff"],
                                                                                 w["sb_calls"], w["sb_matches"], w["sb_idx_diff"], w["sb_count_diff"], w["ref_transform_ms"]))
    r3 = rank3_rows(max(a.frames // 10, 1))
    print("\n## Remaining windowed matchers (8f rank 3) and batched descriptor selection (rank 4), vs the port oracle (pinned to the reference bodies by `tests/golden/ref_match_kf.npz`, `ref_bow.npz`)\n")
    print("| frame pairs | function | accepted matches / selections (oracle) | differences (count + index) |\n|---|---|---|---|")
    for key, name in (("kf", "SearchByProjection(Frame&, KeyFrame*, set&, th, ORBdist) x3 settings"), ("kfp", "SearchByProjection(KeyFrame*, Scw, points, matched, th) x2 th"),
                      ("sim3", "SearchBySim3 x2 th"), ("fuse", "Fuse search, pose + Sim3 forms x2 th"), ("tri", "SearchForTriangulation x3 settings"), ("dist", "ComputeDistinctiveDescriptors, 300 map points per pair")):
        print("| %d | %s | %d | %d |" % (r3["pairs"], name, r3[key][0], r3[key][1]))
    f = frame_rows(max(a.frames // 2, 4), kind)
    print("\n## Device-resident Frame (8f rank 1): UndistortKeyPoints / ComputeStereoFromRGBD / AssignFeaturesToGrid, cameras TUM1, TUM2, 4-coefficient, rectified\n")
    print("| frames | keypoints | mvKeysUn diffs | mvuRight diffs | mvDepth diffs | frames with a different mGrid | bounds diffs |\n|---|---|---|---|---|---|---|")
    print("| %d | %d | %d | %d | %d | %d | %d |" % (f["frames"], f["keypoints"], f["keys_un_diff"], f["u_right_diff"], f["depth_diff"], f["grid_diff"], f["bounds_diff"]))


if __name__ == "__main__":
    main()
