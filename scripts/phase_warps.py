"""Diagnostic: per-WARP globaltimer stamps of k_sim at every phase boundary (libtz_b200_prof.so, -DTZ_PROFILE):
which phase makes the slowest warps of a launch slow."""
import ctypes as C, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbozero_b200 import _abi
from standin import abi as _sabi
_orig = _abi._load
def _load(name, symbols, lib_dir=None):
    return _orig("libtz_b200_prof.so" if name == "libtz_b200.so" else name, symbols, lib_dir)
_abi._load = _load
_sabi._load = lambda name, symbols, lib_dir=None: _orig(name, symbols, lib_dir)
import turbozero_b200 as tz
from standin.synthetic import SyntheticGame, SyntheticSelfPlay, make_synthetic_evaluator

ORDER = [0, 8, 9, 10, 1, 2, 3, 4, 5, 6]
NAMES = ["entry>rt1 issued", "rt1 back+sel check", "q/n issued", "rows issued", "expand+values", "decisions", "stores", "walk", "epilogue"]

def run(name, B, S, N):
    game = SyntheticGame.named(name, 1234)
    ev = make_synthetic_evaluator(tz.MCTS, game, action_selector=tz.PUCTSelector(), max_nodes=N, num_iterations=S)
    sp = SyntheticSelfPlay(game, ev, B)
    sp.dir_noise.copy_(torch.distributions.Dirichlet(torch.full((B, game.F), 0.3)).sample().cuda())
    sp.uniform01.uniform_()
    lib = _abi.lib()
    lib.tz_debug_prof_gt.argtypes = [C.c_void_p, C.c_int]
    for _ in range(3): sp.move()
    torch.cuda.synchronize()
    ts = sp.tree.struct()
    sp.game.root_eval(sp.state, sp.dir_noise, sp.dir_eps, out=(sp.root_policy, sp.root_value))
    ptrs = (C.c_void_p * 2)(sp.state["core"].data_ptr(), SyntheticGame._pay(sp.state))
    st = torch.cuda.current_stream().cuda_stream
    lib.tz_set_root(C.byref(ts), sp.root_policy.data_ptr(), sp.root_value.data_ptr(), ptrs, st)
    lib.tz_select(C.byref(ts), C.byref(sp.cfg), C.byref(sp.work), st)
    nw = min(B, 4096)
    gbuf = (C.c_longlong * (16 * nw))()
    fn, user, _ = sp._cb
    leaf = _sabi.synth_lib().tz_synth_leaf_cb
    med_ph, slow_ph, slow_tot, med_tot = [], [], [], []
    for s in range(S - 1):
        leaf(user, s, C.byref(sp.work), st)
        lib.tz_expand_backprop_select(C.byref(ts), C.byref(sp.cfg), C.byref(sp.work), st)
        torch.cuda.synchronize()
        if s < S // 2: continue
        lib.tz_debug_prof_gt(gbuf, nw)
        rows = [[gbuf[16 * i + k] for k in ORDER] for i in range(nw)]
        ph = [[r[j + 1] - r[j] for j in range(len(ORDER) - 1)] for r in rows]
        tot = [r[-1] - r[0] for r in rows]
        order = sorted(range(nw), key=lambda i: tot[i])
        slow = order[-8:]
        med_ph.append([statistics.median(p[j] for p in ph) for j in range(len(NAMES))])
        slow_ph.append([statistics.mean(ph[i][j] for i in slow) for j in range(len(NAMES))])
        slow_tot.append(statistics.mean(tot[i] for i in slow)); med_tot.append(statistics.median(tot))
    print(f"{name} B={B}: per-warp phase durations (ns), median over launches: [median warp] vs [mean of the 8 slowest warps]")
    for j, nm in enumerate(NAMES):
        print(f"  {nm:22s} {statistics.median(m[j] for m in med_ph):8.0f} {statistics.median(m[j] for m in slow_ph):8.0f}")
    print(f"  {'total':22s} {statistics.median(med_tot):8.0f} {statistics.median(slow_tot):8.0f}")

if __name__ == "__main__":
    run("connect_four", 1024, 128, 256)
    run("go_9x9", 1024, 800, 1600)
