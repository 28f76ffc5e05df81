import importlib, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from tools.synth import synth_frame
orbx = importlib.import_module("amos-slam_b200")
E = orbx.ORBextractor(1000, 1.2, 8, 20, 7)
A = synth_frame(0, 640, 480)
for _ in range(4): E(A)
