"""Diagnostic: where one simulation's wall time goes when a whole move replays as a CUDA graph -- per-LAUNCH globaltimer stamps
of the search kernel (k_sim) and of the leaf stand-in (k_leaf): first warp in, last warp past griddepcontrol.wait, last warp
out.  Needs the diagnostic builds (scripts/build_prof.sh: libtz_b200_prof.so, libtz_synth_prof.so; -DTZ_PROFILE)."""
import ctypes as C, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turbozero_b200 import _abi
from standin import abi as _sabi
_orig = _abi._load
def _load(name, symbols, lib_dir=None):
    return _orig({"libtz_b200.so": "libtz_b200_prof.so", "libtz_synth.so": "libtz_synth_prof.so"}.get(name, name), symbols, lib_dir)
_abi._load = _load
_sabi._load = _load
import turbozero_b200 as tz
from standin.synthetic import SyntheticGame, SyntheticSelfPlay, make_synthetic_evaluator


def read(fn, reset):
    buf = (C.c_ulonglong * (4 * 1024))()
    fn.argtypes = [C.c_void_p, C.c_int]
    assert fn(buf, reset) == 0
    rows = [tuple(buf[4 * i:4 * i + 3]) for i in range(1024)]
    return sorted(r for r in rows if r[0] != 0xFFFFFFFFFFFFFFFF and r[2] != 0)


def run(name, B, S, N, programmatic):
    game = SyntheticGame.named(name, 1234)
    ev = make_synthetic_evaluator(tz.MCTS, game, action_selector=tz.PUCTSelector(), max_nodes=N, num_iterations=S,
                                  programmatic=programmatic)
    _sabi.synth_lib().tz_synth_set_programmatic(1 if programmatic else 0)
    sp = SyntheticSelfPlay(game, ev, B)
    sp.dir_noise.copy_(torch.distributions.Dirichlet(torch.full((B, game.F), 0.3)).sample().cuda())
    sp.uniform01.uniform_()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        sp.move()
        with torch.cuda.graph(cg, stream=side):
            sp.move()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(4): cg.replay()
    torch.cuda.synchronize()
    lib, slib = _abi.lib(), _sabi.synth_lib()
    read(lib.tz_debug_timeline, 1); read(slib.tz_synth_debug_timeline, 1)
    cg.replay(); torch.cuda.synchronize()
    sims, leaves = read(lib.tz_debug_timeline, 0), read(slib.tz_synth_debug_timeline, 0)
    sims = sims[1:]  # drop the select-only launch; sims[i] now follows leaves[i]
    n = min(len(sims), len(leaves)) - 1
    half = range(n // 2, n)
    med = lambda f: statistics.median(f(i) for i in half)
    print(f"{name} B={B} S={S} programmatic={programmatic}: medians over the second half of one move's simulations (ns)")
    print(f"  per simulation (leaf first-in -> next leaf first-in)   {med(lambda i: leaves[i + 1][0] - leaves[i][0]):8.0f}")
    print(f"  leaf:   first warp in -> last warp out                 {med(lambda i: leaves[i][2] - leaves[i][0]):8.0f}")
    print(f"  leaf:   last warp past its wait -> last warp out       {med(lambda i: leaves[i][2] - leaves[i][1]):8.0f}")
    print(f"  k_sim:  first warp in, relative to the leaf's last out {med(lambda i: sims[i][0] - leaves[i][2]):8.0f}   (negative = resident early)")
    print(f"  k_sim:  leaf last out -> last warp has its leaf loads  {med(lambda i: sims[i][1] - leaves[i][2]):8.0f}")
    print(f"  k_sim:  leaf loads issued (last warp) -> last warp out {med(lambda i: sims[i][2] - sims[i][1]):8.0f}")
    print(f"  k_sim:  first warp in -> last warp out                 {med(lambda i: sims[i][2] - sims[i][0]):8.0f}")
    print(f"  next leaf: first warp in, relative to k_sim's last out {med(lambda i: leaves[i + 1][0] - sims[i][2]):8.0f}   (negative = resident early)")
    print(f"  next leaf: k_sim last out -> last warp past its wait   {med(lambda i: leaves[i + 1][1] - sims[i][2]):8.0f}")


if __name__ == "__main__":
    for prog in (False, True):
        run("connect_four", 1024, 128, 256, prog)
