"""Scratch timing of the CUDA self-play loop (not the driver's bench; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import turbozero_b200 as tz
from standin.synthetic import SyntheticGame, SyntheticSelfPlay

def run(name, B, S, N, weighted=False, moves=8, warm=3, graph=True, use_path=True, pipelines=1, pdl=True):
    game = SyntheticGame.named(name, 1234)
    base = tz.WeightedMCTS if weighted else tz.MCTS
    kw = dict(eval_fn=None, action_selector=tz.PUCTSelector(), branching_factor=game.F, max_nodes=N, num_iterations=S)
    ev = base(**kw)
    ev.programmatic_launch = pdl
    from standin import abi as _sabi
    _sabi.synth_lib().tz_synth_set_programmatic(1 if pdl else 0)
    sp = SyntheticSelfPlay(game, ev, B, dirichlet=True, use_path=use_path, pipelines=pipelines)
    sp.dir_noise.copy_(torch.distributions.Dirichlet(torch.full((B, game.F), 0.3)).sample().cuda())
    sp.uniform01.uniform_()
    if graph:
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            sp.move()  # warm-up outside capture
            with torch.cuda.graph(cg, stream=side):
                sp.move()
        torch.cuda.current_stream().wait_stream(side)
        step = cg.replay
    else:
        step = sp.move
    for _ in range(warm): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st0 = sp.tree.stats.sum(0).clone()
    e0.record()
    for _ in range(moves): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / moves
    st = (sp.tree.stats.sum(0) - st0).tolist()
    print(f"{name} B={B} S={S} N={N} weighted={weighted} graph={graph} K={pipelines} pdl={pdl}: {ms:.3f} ms/move, "
          f"{B*S/ms*1e3:.3e} sims/s, {ms*1e3/S:.2f} us/sim, levels/sim={st[0]/max(st[1],1):.2f}, nfi_mean={st[2]/moves/B:.1f} kept={st[3]/moves/B:.1f}", flush=True)

if __name__ == "__main__":
    import sys as _s
    if len(_s.argv) > 1 and _s.argv[1] == "pipe":
        for K in (1, 2, 4, 8, 16, 32):
            run("connect_four", 1024, 128, 256, moves=8, pipelines=K)
        for K in (1, 4, 16):
            run("othello", 512, 200, 400, weighted=True, moves=4, pipelines=K)
        for K in (1, 4, 16):
            run("go_9x9", 1024, 800, 1600, moves=2, warm=1, pipelines=K)
    elif len(_s.argv) > 1 and _s.argv[1] == "pdl":
        for v in (0, 1, 3):
            run("connect_four", 1024, 128, 256, pdl=v)
            run("go_9x9", 1024, 800, 1600, moves=2, warm=1, pdl=v)
            run("othello", 512, 200, 400, weighted=True, moves=4, pdl=v)
            run("connect_four", 16384, 128, 256, moves=2, warm=1, pdl=v)
    elif len(_s.argv) > 1 and _s.argv[1] == "sweep":
        for B in (128, 1024, 4096, 16384):
            run("connect_four", B, 128, 256, moves=4)
        run("tic_tac_toe", 32, 64, 128)
    else:
        run("connect_four", 1024, 128, 256)
        run("connect_four", 1024, 128, 256, pdl=False)
        run("connect_four", 1024, 128, 256, graph=False)
        run("connect_four", 1024, 128, 256, graph=False, pdl=False)
        run("othello", 512, 200, 400, weighted=True, moves=4, pdl=False)
        run("go_9x9", 1024, 800, 1600, moves=2, warm=1, pdl=False)
        run("tic_tac_toe", 32, 64, 128)
        run("othello", 512, 200, 400, weighted=True, moves=4)
        run("go_9x9", 1024, 800, 1600, moves=2, warm=1)
        run("2048", 2048, 100, 200)
        run("2048", 2048, 100, 200, pdl=False)
        run("connect_four", 65536, 128, 256, moves=2, warm=1, pdl=False)
        run("connect_four", 65536, 128, 256, moves=2, warm=1)
