"""Diagnostic: per-phase clock64 stamps of k_sim for tree 0 (needs libtz_b200_prof.so built with -DTZ_PROFILE)."""
import ctypes as C, os, sys
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

def run(name, B, S, N, weighted=False):
    game = SyntheticGame.named(name, 1234)
    ev = make_synthetic_evaluator(tz.WeightedMCTS if weighted else tz.MCTS, game, action_selector=tz.PUCTSelector(), max_nodes=N, num_iterations=S)
    sp = SyntheticSelfPlay(game, ev, B)
    sp.dir_noise.copy_(torch.distributions.Dirichlet(torch.full((B, game.F), 0.3)).sample().cuda())
    sp.uniform01.uniform_()
    lib = _abi.lib()
    lib.tz_debug_prof.argtypes = [C.c_void_p]
    for _ in range(3): sp.move()
    torch.cuda.synchronize()
    # one more move, sim by sim, reading the stamps after every fused launch
    ts = sp.tree.struct()
    sp.game.root_eval(sp.state, sp.dir_noise, sp.dir_eps, out=(sp.root_policy, sp.root_value))
    ptrs = (C.c_void_p * 2)(sp.state["core"].data_ptr(), SyntheticGame._pay(sp.state))
    st = torch.cuda.current_stream().cuda_stream
    lib.tz_set_root(C.byref(ts), sp.root_policy.data_ptr(), sp.root_value.data_ptr(), ptrs, st)
    lib.tz_select(C.byref(ts), C.byref(sp.cfg), C.byref(sp.work), st)
    buf = (C.c_longlong * 64)()
    nw = min(B, 4096)
    wbuf = (C.c_longlong * (4 * nw))()
    lib.tz_debug_prof_warps.argtypes = [C.c_void_p, C.c_int]
    spans, worst = [], []
    rows = []
    fn, user, _ = sp._cb
    leaf = _sabi.synth_lib().tz_synth_leaf_cb
    for s in range(S - 1):
        leaf(user, s, C.byref(sp.work), st)
        lib.tz_expand_backprop_select(C.byref(ts), C.byref(sp.cfg), C.byref(sp.work), st)
        torch.cuda.synchronize()
        lib.tz_debug_prof(buf)
        v = list(buf)
        rows.append(v)
        lib.tz_debug_prof_warps(wbuf, nw)
        w = [wbuf[4 * i:4 * i + 4] for i in range(nw)]
        t0 = min(x[0] for x in w); t1 = max(x[1] for x in w)
        durs = sorted(((x[1] - x[0]), x[2], x[3], x[0] - t0) for x in w)
        spans.append((t1 - t0, durs[len(durs) // 2][0], durs[-1], max(x[0] for x in w) - t0, max(x[2] for x in w), max(x[3] for x in w)))
    import statistics
    def med(f): return statistics.median(f(v) for v in rows[S // 2:])
    print(f"{name} B={B}: kernel span (first warp entry -> last warp exit, ns) / median warp / slowest warp (ns, old L, new L, start offset) / last warp start offset / max old L / max new L")
    for sp_ in spans[S // 2::max(1, S // 16)]:
        print("   ", sp_)
    print(f"  median span {statistics.median(x[0] for x in spans[S // 2:])} ns, median of median-warp {statistics.median(x[1] for x in spans[S // 2:])} ns")
    print(f"{name} B={B}: medians over the second half of the search (cycles, tree 0)")
    print("  total                      ", med(lambda v: v[6] - v[0]))
    print("  entry -> trip 2 issued     ", med(lambda v: v[1] - v[0]), "  [trip 1 issued", med(lambda v: v[8] - v[0]),
          " trip 1 back + selector check", med(lambda v: v[9] - v[8]), " q/n issued", med(lambda v: v[10] - v[9]),
          " rows + table staging issued", med(lambda v: v[1] - v[10]), "]")
    print("  expand + per-level values  ", med(lambda v: v[2] - v[1]))
    print("  path decisions (all levels)", med(lambda v: v[3] - v[2]))
    print("  stores + emb store + sync  ", med(lambda v: v[4] - v[3]))
    print("  walk                       ", med(lambda v: v[5] - v[4]), " levels", med(lambda v: v[7]))
    print("  epilogue (emb gather)      ", med(lambda v: v[6] - v[5]))

if __name__ == "__main__":
    run("connect_four", 1024, 128, 256)
    run("go_9x9", 256, 400, 800)
    run("othello", 256, 100, 200, weighted=True)
