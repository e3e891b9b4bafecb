import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess
from h264decode_b200 import build as b
OUT = os.path.join(os.path.dirname(b.__file__), "exp")
if "--build" in sys.argv:
    os.makedirs(OUT, exist_ok=True)
    for w in (1, 2, 4, 8):
        b.build(extra=["-DH264B_CABAC_WARPS=%d" % w], out=os.path.join(OUT, "libw_%d.so" % w)); print("built", w)
if "--run" in sys.argv:
    for w in (1, 2, 4, 8):
        env = dict(os.environ, H264B_LIB=os.path.join(OUT, "libw_%d.so" % w))
        out = subprocess.run([sys.executable, "bench.py", "--steps", "2", "--warmup", "2", "--no-cpu"], env=env, capture_output=True, text=True).stdout
        import json
        d = json.loads(out.strip().splitlines()[-1]); print("warps/CTA", w, d["stage_ms"], flush=True)
