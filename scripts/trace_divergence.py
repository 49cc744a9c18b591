"""Where does a simulation's walk leave the previous one's?  CPU diagnostic on the C oracle built with -DTZO_TRACE
(histogram over simulations of previous path length x shared levels).  Usage: trace_divergence.py [workload] [envs] [moves]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from oracle import build as OB  # noqa: E402
from oracle import c_oracle as CO  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
moves = int(sys.argv[3]) if len(sys.argv) > 3 else 6
out = OB.OUT_DIR / "libtz_oracle_trace.so"
OB.OUT_DIR.mkdir(exist_ok=True)
subprocess.run(["gcc", *OB.FLAGS, "-DTZO_TRACE", f"-I{ROOT}/include", f"-I{ROOT}/standin/include", f"{ROOT}/oracle/tz_oracle.c", "-o",
                str(out), "-lm"], check=True)
real_build = OB.build
OB.build = lambda force=False: out
CO._lib = None
lib = CO.lib()
name, _, S, N, weighted, discount, _ = bench.WORKLOADS[wl]
_, g, cg, cfg, t, episode, core, payload = bench.cpu_selfplay_setup(wl, B, 0, 1000)
rng = np.random.default_rng(5)
dn = rng.dirichlet(np.full(g.F, 0.3), size=(moves, B)).astype(np.float32)
rn = rng.random((moves, B, g.F), dtype=np.float32)
u = rng.random((moves, B), dtype=np.float32)
CO.selfplay(t, cfg, cg, S, moves, 1.0, True, 0, dn, 0.25, rn, u, core, payload, episode, bench.host_threads(CO))
H = np.ctypeslib.as_array((C.c_longlong * (256 * 256)).in_dll(lib, "tzo_trace_hist")).reshape(256, 256).copy()
tot = H.sum()
L = np.arange(256)
pl = H.sum(1)
print(f"{wl} ({name}) envs {B} sims {S} moves {moves}: {tot} walks")
print("previous path length: mean %.1f  p50 %d  p90 %d  p99 %d  max %d" % (
    (pl * L).sum() / tot, *(int(np.searchsorted(np.cumsum(pl), q * tot)) for q in (0.5, 0.9, 0.99)), int(L[pl > 0].max())))
sh = H.sum(0)
print("levels shared with the previous walk: mean %.1f  p50 %d  p90 %d  p99 %d" % (
    (sh * L).sum() / tot, *(int(np.searchsorted(np.cumsum(sh), q * tot)) for q in (0.5, 0.9, 0.99))))
for lo, hi in ((1, 8), (8, 16), (16, 32), (32, 64), (64, 128), (128, 256)):
    blk = H[lo:hi]
    n = blk.sum()
    if n == 0:
        continue
    s = blk.sum(0)
    cs = np.cumsum(s)
    print(f"  previous length {lo:3d}-{hi - 1:3d}: {100.0 * n / tot:5.1f} % of walks; shared levels mean {(s * L).sum() / n:5.1f}"
          f"  p50 {int(np.searchsorted(cs, 0.5 * n))}  p90 {int(np.searchsorted(cs, 0.9 * n))}; shared <= 4 levels: {100.0 * s[:5].sum() / n:4.1f} %,"
          f" <= 8: {100.0 * s[:9].sum() / n:4.1f} %")
