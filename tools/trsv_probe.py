#!/usr/bin/env python
"""configs[3] (BASELINE wording): time of the level-scheduled triangular solve for several grid sizes (VBC_TRSV_LEVELS_AHEAD)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb
from vbc_b200 import synth
DT = np.float32 if os.environ.get("PROBE_F32") else np.float64
A, pi, phi = synth.config_c4_triangular(dtype=DT)
B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
nlev = vb.trsv_analyse(B.T)
b = torch.from_numpy(synth.vector(A.n, 11, dtype=DT)).cuda(); x = torch.empty_like(b)
for _ in range(2): vb.ldiv_lower_(x, B.T, b)
B.sync()
import threading, time
clk = []
def sample():
    try:
        import pynvml
        pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not stop:
            clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); time.sleep(0.002)
    except Exception as e:
        clk.append(repr(e))
stop = False
th = threading.Thread(target=sample); th.start()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); vb.ldiv_lower_(x, B.T, b); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
stop = True; th.join()
print(json.dumps({"block": os.environ.get("VBC_TRSV_BLOCK"), "sm_mhz_min_med_max": [min(clk), sorted(clk)[len(clk) // 2], max(clk)] if clk and isinstance(clk[0], int) else clk[:1], "ahead": os.environ.get("VBC_TRSV_LEVELS_AHEAD"), "levels": nlev, "ms_med": sorted(ts)[2], "us_per_level": 1e3 * sorted(ts)[2] / nlev}))
