"""Diagnostic: does k_sim pay for a cold front end (instruction / constant caches) after another kernel ran?
Two identical self-plays A and B; per simulation either  leaf(B) sim(B) leaf(A) sim(A)   [A's k_sim follows a leaf kernel]
or  leaf(A) leaf(B) sim(B) sim(A)   [A's k_sim follows an identical k_sim launch].  Reports tree-0 phase cycles of A.
Needs libtz_b200_prof.so (-DTZ_PROFILE)."""
import ctypes as C, os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbozero_b200 import _abi
from standin import abi as _sabi
_orig = _abi._load
_abi._load = lambda name, symbols: _orig("libtz_b200_prof.so" if name == "libtz_b200.so" else name, symbols)
import turbozero_b200 as tz
from standin.synthetic import SyntheticGame, SyntheticSelfPlay, make_synthetic_evaluator

def make(name, B, S, N):
    game = SyntheticGame.named(name, 1234)
    ev = make_synthetic_evaluator(tz.MCTS, game, action_selector=tz.PUCTSelector(), max_nodes=N, num_iterations=S)
    sp = SyntheticSelfPlay(game, ev, B)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    sp.dir_noise.copy_(torch.distributions.Dirichlet(torch.full((B, game.F), 0.3)).sample().cuda() * 0 + 1.0 / game.F)
    sp.uniform01.fill_(0.37)
    return sp

def run(order, name="connect_four", B=256, S=128, N=256):
    A, Bp = make(name, B, S, N), make(name, B, S, N)
    lib = _abi.lib(); lib.tz_debug_prof.argtypes = [C.c_void_p]
    for _ in range(3):
        A.move(); Bp.move()
    torch.cuda.synchronize()
    st = torch.cuda.current_stream().cuda_stream
    leaf = _sabi.synth_lib().tz_synth_leaf_cb
    def begin(sp):
        ts = sp.tree.struct()
        sp.game.root_eval(sp.state, sp.dir_noise, sp.dir_eps, out=(sp.root_policy, sp.root_value))
        ptrs = (C.c_void_p * 2)(sp.state["core"].data_ptr(), SyntheticGame._pay(sp.state))
        lib.tz_set_root(C.byref(ts), sp.root_policy.data_ptr(), sp.root_value.data_ptr(), ptrs, st)
        lib.tz_select(C.byref(ts), C.byref(sp.cfg), C.byref(sp.work), st)
        return ts
    tA, tB = begin(A), begin(Bp)
    buf = (C.c_longlong * 64)(); rows = []
    def L(sp): leaf(sp._cb[1], 0, C.byref(sp.work), st)
    def K(sp, ts): lib.tz_expand_backprop_select(C.byref(ts), C.byref(sp.cfg), C.byref(sp.work), st)
    for s in range(S - 1):
        if order == "after_leaf":
            L(Bp); K(Bp, tB); L(A); K(A, tA)
        else:
            L(A); L(Bp); K(Bp, tB); K(A, tA)
        torch.cuda.synchronize()
        lib.tz_debug_prof(buf); rows.append(list(buf))
    med = lambda f: statistics.median(f(v) for v in rows[S // 2:])
    print("   prologue split: to loads", med(lambda v: v[20]-v[0]), " loads issued", med(lambda v: v[21]-v[20]), " sel check", med(lambda v: v[22]-v[21]),
          " ring test", med(lambda v: v[23]-v[22]), " q/n issue", med(lambda v: v[24]-v[23]), " rows issue", med(lambda v: v[1]-v[24]))
    print(f"{order:12s}: total {med(lambda v: v[6]-v[0])}  entry->trip2 {med(lambda v: v[1]-v[0])}  expand {med(lambda v: v[2]-v[1])}  decisions {med(lambda v: v[3]-v[2])}  stores {med(lambda v: v[4]-v[3])}  walk {med(lambda v: v[5]-v[4])} ({med(lambda v: v[7])} lv)  epilogue {med(lambda v: v[6]-v[5])}")

if __name__ == "__main__":
    for o in ("after_leaf", "after_ksim"):
        run(o)
