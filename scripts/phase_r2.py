"""Diagnostic (libtz_b200_prof.so, -DTZ_PROFILE): per-tree %globaltimer stamps of k_sim_wide and k_reroot_bulk at their phase
boundaries: which phase bounds the slowest CTAs of a launch.   python scripts/phase_r2.py [wide|reroot] [game] [B] [S] [N] [warps] [weighted]"""
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


def setup(name, B, S, N, warps, weighted):
    game = SyntheticGame.named(name, 1234)
    base = tz.WeightedMCTS if weighted else tz.MCTS
    ev = make_synthetic_evaluator(base, game, action_selector=tz.PUCTSelector(), max_nodes=N, num_iterations=S)
    ev.sim_warps = warps
    sp = SyntheticSelfPlay(game, ev, B)
    sp.dir_noise.copy_(torch.distributions.Dirichlet(torch.full((B, game.F), 0.3)).sample().cuda())
    sp.uniform01.uniform_()
    lib = _abi.lib()
    lib.tz_debug_prof_gt.argtypes = [C.c_void_p, C.c_int]
    for _ in range(3):
        sp.move()
    torch.cuda.synchronize()
    return game, ev, sp, lib


def report(title, names, rows):
    """rows: per tree list of stamps (ns); prints median-CTA and slowest-8 phase durations"""
    ph = [[r[j + 1] - r[j] for j in range(len(names))] for r in rows]
    tot = [r[len(names)] - r[0] for r in rows]
    order = sorted(range(len(rows)), key=lambda i: tot[i])
    slow = order[-8:]
    t0 = min(r[0] for r in rows)
    print(title)
    for j, nm in enumerate(names):
        print(f"  {nm:34s} median {statistics.median(p[j] for p in ph):9.0f}   slowest-8 mean {statistics.mean(ph[i][j] for i in slow):9.0f}")
    print(f"  {'total per CTA':34s} median {statistics.median(tot):9.0f}   slowest-8 mean {statistics.mean(tot[i] for i in slow):9.0f}")
    print(f"  first CTA in -> last CTA out: {max(r[len(names)] for r in rows) - t0} ns; CTA start spread {max(r[0] for r in rows) - t0} ns")


def wide(name, B, S, N, warps, weighted):
    game, ev, sp, lib = setup(name, B, S, N, warps, weighted)
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
    names = ["entry -> record checked (RT1)", "RT2 + backup", "decisions (+ expansion writes)", "vote", "walk", "embedding rows"]
    acc = []
    for s in range(S - 1):
        leaf(user, s, C.byref(sp.work), st)
        lib.tz_expand_backprop_select(C.byref(ts), C.byref(sp.cfg), C.byref(sp.work), st)
        torch.cuda.synchronize()
        if s < S // 2 or s % 8:
            continue
        lib.tz_debug_prof_gt(gbuf, nw)
        acc.append([[gbuf[16 * i + k] for k in range(9)] for i in range(nw)])
    # median over sampled launches of per-launch summaries
    rows = acc[len(acc) // 2]
    report(f"k_sim_wide {name} B={B} S={S} warps={warps} weighted={weighted}: one launch (of {len(acc)} sampled), ns", names,
           [r[:7] for r in rows])
    # per-level cost of the decisions phase: (stamp3 - stamp2) against the path length, all sampled launches pooled
    import collections
    byL = collections.defaultdict(list)
    chain = collections.defaultdict(list)
    for rr in acc:
        for r in rr:
            if r[8] > 0:
                byL[min(int(r[8]), 64) // 4 * 4].append(r[3] - r[2])
                if weighted and r[7] > r[2]:
                    chain[min(int(r[8]), 64) // 4 * 4].append(r[7] - r[2])
    print("  decisions phase by path length L (bucketed by 4): L, trees, median ns" + (", median chain ns" if weighted else ""))
    for Lb in sorted(byL):
        extra = f" {statistics.median(chain[Lb]):8.0f}" if (weighted and chain[Lb]) else ""
        print(f"    {Lb:3d}+ {len(byL[Lb]):7d} {statistics.median(byL[Lb]):8.0f}{extra}")
    firsts = [max(r[6] for r in rr) - min(r[0] for r in rr) for rr in acc]
    print(f"  first-in -> last-out over the sampled launches: median {statistics.median(firsts)} ns, max {max(firsts)} ns")


def reroot(name, B, S, N, warps, weighted):
    game, ev, sp, lib = setup(name, B, S, N, warps, weighted)
    nw = min(B, 4096)
    gbuf = (C.c_longlong * (16 * nw))()
    names = ["entry -> root child known", "parents + pointer jumping", "scan", "chunk 0 gather issued", "chunk 0 landed", "chunk 0 scattered",
             "remaining chunks", "tail fill"]
    for rep in range(3):
        sp.move()
        torch.cuda.synchronize()
        lib.tz_debug_prof_gt(gbuf, nw)
        rows = [[gbuf[16 * i + k] for k in range(15)] for i in range(nw)]
        live = [r for r in rows if r[10] > 0 and all(r[k] > 0 for k in range(9))]  # re-rooted trees with at least one chunk
        if live:
            nfi = statistics.mean(r[9] for r in live)
            cnt = statistics.mean(r[10] for r in live)
            report(f"k_reroot_bulk {name} B={B}: move {rep}, {len(live)} re-rooted trees, rows before {nfi:.0f} kept {cnt:.0f}, ns", names,
                   [r[:9] for r in live])
            ch = [r[14] for r in live]
            print(f"  chunks per tree: mean {statistics.mean(ch):.1f} max {max(ch)}; per chunk (thread 0, mean over trees): issue "
                  f"{statistics.mean(r[11] / max(r[14], 1) for r in live):.0f} ns, wait {statistics.mean(r[12] / max(r[14], 1) for r in live):.0f} ns, "
                  f"scatter {statistics.mean(r[13] / max(r[14], 1) for r in live):.0f} ns")


if __name__ == "__main__":
    a = sys.argv[1:]
    which = a[0] if a else "wide"
    name = a[1] if len(a) > 1 else "go_9x9"
    B, S, N = (int(a[2]), int(a[3]), int(a[4])) if len(a) > 4 else (1024, 800, 1600)
    warps = int(a[5]) if len(a) > 5 else 4
    weighted = len(a) > 6 and a[6] == "weighted"
    (wide if which == "wide" else reroot)(name, B, S, N, warps, weighted)
