import importlib, os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import match_cases as mc
orbx = importlib.import_module("amos-slam_b200")
L, R = mc.stereo_pair()
EL, ER = orbx.ORBextractor(2000, 1.2, 8, 20, 7), orbx.ORBextractor(2000, 1.2, 8, 20, 7)
kl, dl = EL(L); kr, dr = ER(R)
M = orbx.ORBmatcher()
for _ in range(3):
    ur, dep = M.ComputeStereoMatches(EL, ER, kl, dl, kr, dr, 0.0, mc.BF_KITTI)
print("ok", (ur >= 0).sum())
