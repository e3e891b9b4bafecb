import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from h264decode_b200 import capi
from tests.test_hd_logic import random_stream
ctx = capi.Context(0)
rng = np.random.default_rng(0)
T = 16384
sizes = [1, 3, 4, 5, 15, 16, 17, 31, 33, 1000, T - 1, T, T + 1, 2 * T - 3, 2 * T + 5, 5 * T + 7, 40 * T + 11]
for n in sizes:
    for p_zero, p_sc in [(0.5, 0.02), (0.2, 0.002), (0.9, 0.0005), (0.02, 0.0001)]:
        s = random_stream(rng, n, p_zero, p_sc, ext_types=True)
        try:
            summ, nals, ext, rbsp = ctx.annexb_scan(s)
        except Exception as e:
            print("FAIL n=%d p_zero=%g p_sc=%g: %s" % (n, p_zero, p_sc, e)); sys.exit(1)
        print("ok", n, p_zero, p_sc, summ["n_nals"], flush=True)
