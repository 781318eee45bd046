import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, common
from optimal_conv_b200 import hec, params as PR, synth
N = 1 << PR.LOGN
c = hec.Context(PR.LOGN, PR.Q_SET6[:3], PR.P_ALL[:1])
v = np.tile(np.array(common.ENCODE_EDGE), 3)
pts = c.EncodeCoeffsNTTMany(np.stack([np.resize(v, 5000), np.resize(v[::-1], 5000)] * 9), 2, 2.0 ** 30)
out = c.download_pt(pts[17])
for p in pts: p.free()
try:
    c.EncodeCoeffsNTT(np.array([np.nan]), 1, 2.0 ** 30)
except hec.HecError as e:
    print("refused:", e)
c.close(); print("ok", out[:, :3])
