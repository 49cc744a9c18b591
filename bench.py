#!/usr/bin/env python
"""bench.py -- MCTS simulations/sec of the batched search hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--named cfg3,cfg4] [--impl reference]

A STEP is one self-play move of every env on this rank: root evaluation + set_root, S simulations
(select -> leaf -> expand+backprop), root-action sampling, env step, and subtree re-rooting (or reset where the
episode ended) -- i.e. `step_env_and_evaluator` of the reference (core/common.py:32-103) with the user's pgx env and
network replaced by the synthetic stand-in (pgx / JAX are not installable here).

Headline workload = BASELINE.json configs[1]: connect_four shape, 1024 envs x 128 simulations, max_nodes 256, subtree
persistence on, one B200.  With --gpus N every rank owns its own 1024 envs (independent tree batches, no collective
in the search) => weak scaling; `value` is the whole-job aggregate.

The configs BASELINE.json names for N > 1 GPUs are measured in the same run and reported under `config.named`
(each with its own value / e2e / ms_per_step / roofline):
  N = 2, 4, 8   configs[2]  othello 8x8, 4096 envs (strong-scaled: 4096 / N per GPU) x 200 sims, WeightedMCTS backup
  N = 8         configs[3]  go_9x9, 8192 envs (1024 per GPU) x 800 sims
  N = 8         configs[4]  2048, 16384 envs (2048 per GPU) x 100 sims, discount +1, with the replay-buffer update in the
                            step, ONE cross-rank replay sample (NCCL all-gather / all-reduce) and the gradient-mean
                            all-reduce of a parameter-sized buffer per step (core/training/train.py:393-437)
`--named cfgX,...` forces named legs at any N (single-GPU shares of the configs; used for development runs).

Legs of every workload (all in one JSON line):
  value        device-resident inputs, one CUDA-graph replay per step, per-step CUDA events, L2 flushed between steps
  e2e          the public Python API (step_env_and_evaluator = MCTS.evaluate + MCTS.step, captured by the user in a CUDA
               graph) with per-step H2D copies of the step's random inputs from pinned memory and a D2H read of
               actions + policy weights
  roofline     the per-simulation kernel (expand+backprop+select) timed INSIDE the step it describes: the same graph
               instantiation (programmatic launches included) replayed with the launch timeline on (TzWork.timeline:
               %globaltimer at first warp in / last warp out of every launch); achieved = algorithmic bytes / that
               duration vs the measured HBM peak.  launches x avg duration <= ms_per_step is asserted.
  cpu_baseline the CPU oracle (C restatement, OpenMP over trees) on a bounded sample of the same workload (rank 0,
               N = 1), plus the literal per-tree NumPy restatement (the masked-dataflow oracle) on a smaller sample

`--impl reference`: the reference's algorithm on the host CPU only (the oracle port; the real reference needs JAX,
which is absent), all host threads, same config / metric.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "mcts_simulations_per_sec"
UNIT = "simulations/s"

# name -> (synthetic game shape, envs per GPU at the config's GPU count, simulations, max_nodes, weighted, discount, description)
# Launch mode per shape, measured (profiles/r2ak_modes.log: ordinary / search launched programmatically / + the leaf stand-in
# cooperating / + the search signalling its dependents early):
#   connect_four 106.3 / 102.8 / 111.4 / 123.6 M   othello 28.0 / 27.8 / 26.4 / 26.5 M
#   go_9x9       36.3 / 36.3 / 36.8 / 36.8 M        2048    174.3 / 170.6 / 182.0 / 168.0 M simulations/s
# -> (TzSearchCfg.programmatic bits, leaf stand-in cooperates) per workload; None = ordinary stream-ordered launches, the API default.
PDL_MODE = {"cfg1": (3, 1), "cfg2": (3, 1), "cfg3": None, "cfg4": (1, 1), "cfg5": (1, 1)}
NO_PDL_BY_DEFAULT = {w for w, m in PDL_MODE.items() if m is None}
REPLAY_CAPACITY = 256  # slots per env of the episode replay buffer in the cfg5 step
TRAIN_BATCH = 1024     # rows of the cross-rank replay sample per step (cfg5, N > 1)
GRAD_FLOATS = 2 << 20  # parameter-sized buffer of the gradient-mean all-reduce (cfg5, N > 1): 2 Mi fp32 = 8 MiB
WORKLOADS = {
    "cfg1": ("tic_tac_toe", 32, 64, 128, False, -1.0, "tic_tac_toe 32 envs x 64 sims (configs[0])"),
    "cfg2": ("connect_four", 1024, 128, 256, False, -1.0, "connect_four 1024 envs x 128 sims, persist_tree (configs[1])"),
    "cfg3": ("othello", 512, 200, 400, True, -1.0, "othello 4096 envs x 200 sims, WeightedMCTS (configs[2])"),
    "cfg4": ("go_9x9", 1024, 800, 1600, False, -1.0, "go_9x9 8192 envs / 8 GPUs x 800 sims (configs[3])"),
    "cfg5": ("2048", 2048, 100, 200, False, 1.0, "2048 16384 envs / 8 GPUs x 100 sims, discount +1, replay memory (configs[4])"),
}
CFG3_TOTAL_ENVS = 4096


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--named", default=None, help="comma-separated workloads to add under config.named (default: by --gpus)")
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU of the headline workload")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-pdl", action="store_true", help="ordinary stream-ordered launches (TzSearchCfg.programmatic = 0)")
    ap.add_argument("--pdl", action="store_true", help="force programmatic dependent launches on")
    ap.add_argument("--pdl-bits", type=int, default=0, help="TzSearchCfg.programmatic when programmatic launches are on (1: launch "
                    "programmatically; 3: also signal dependents early; 0: the workload's measured best, PDL_MODE)")
    ap.add_argument("--leaf-pdl", type=int, default=-1, help="the stand-in leaf kernel cooperates (waits, then signals) when "
                    "programmatic launches are on; 0: it is launched ordinarily; -1: the workload's measured best")
    ap.add_argument("--sim-warps", type=int, default=0, help="TzSearchCfg.sim_warps (0 = library's choice)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-roofline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_inputs(rng, steps, B, F, alpha=0.3):
    dn = rng.dirichlet([alpha] * F, size=(steps, B)).astype(np.float32)
    rn = (rng.random((steps, B, F), dtype=np.float32) * np.float32(1e-8)).astype(np.float32)
    u = rng.random((steps, B), dtype=np.float32)
    return dn, rn, u


def algorithmic_bytes(levels, sims, F, E, weighted):
    """per simulation L*(16F+28) + 8F + 4E + 25 bytes (weighted: + L*(12F+4))."""
    per_level = 16 * F + 28 + ((12 * F + 4) if weighted else 0)
    return levels * per_level + sims * (8 * F + 4 * E + 25)


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm on the host CPU (oracle port; JAX is not installable)
# ------------------------------------------------------------------------------------------------------------------
def cpu_selfplay_setup(wl, B, env_offset, seed):
    from oracle import c_oracle as CO
    from oracle import synth_numpy as SN

    name, _, S, N, weighted, discount, _ = WORKLOADS[wl]
    g = SN.make_game(name, seed)
    cg = CO.make_game(g.F, g.payload_bytes, g.rho256, g.tau1024, g.max_depth, g.seed)
    cfg = CO.make_cfg(discount=discount, weighted=weighted)
    t = CO.HostTrees(B, N, g.F, g.emb_row_bytes, weighted=weighted)
    episode = np.zeros((B,), np.int32)
    core = np.zeros((B, 4), np.int32)
    payload = np.zeros((B, g.payload_bytes), np.uint8) if g.payload_bytes > 0 else None
    CO.synth_init_states(cg, B, env_offset, episode, core, payload)
    return CO, g, cg, cfg, t, episode, core, payload


def host_threads(CO):
    """All the host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    silently make the CPU arm single-threaded: the thread count is passed to the oracle explicitly instead."""
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except AttributeError:
        return max(os.cpu_count() or 1, CO.num_threads())


def cpu_moves(CO, cg, cfg, t, S, moves, dn, rn, u, core, payload, episode, nthreads):
    t0 = time.perf_counter()
    CO.selfplay(t, cfg, cg, S, moves, 1.0, True, 0, dn, 0.25, rn, u, core, payload, episode, nthreads)
    return time.perf_counter() - t0


def numpy_oracle_baseline(wl, seed, seconds=6.0):
    """The literal per-tree NumPy restatement (oracle/mcts_numpy.py: the reference's gathers / scatters / where-selects /
    N-1 label rounds, one env at a time as vmap's per-example program) on a small sample of the same workload: BASELINE.md
    section 3's `oracle-batched` stand-in for the reference's vmapped JAX-CPU path.  One thread."""
    from helpers import Schedule, run_numpy
    from oracle import synth_numpy as SN

    name, _, S, N, weighted, discount, _ = WORKLOADS[wl]
    g = SN.make_game(name, seed)
    envs, spent, sims = 1, 0.0, 0
    while True:
        s = Schedule(game=g, B=envs, N=N, S=S, moves=1, temperature=1.0, weighted=weighted, discount=discount, seed=seed)
        t0 = time.perf_counter()
        run_numpy(s)
        dt = time.perf_counter() - t0
        spent += dt
        sims += envs * S
        last, last_envs = envs * S / dt, envs
        if spent + 2.2 * dt > seconds or envs >= 64:  # the next (doubled) sample would overrun the budget
            break
        envs *= 2
    return {"value": last, "unit": UNIT, "cores": 1, "kind": "port-numpy",
            "sample": f"one move of {last_envs} envs x {S} sims ({dt:.1f} s; {sims} simulations in {spent:.1f} s in all), "
                      "oracle/mcts_numpy.py per tree, single thread"}


def run_reference(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    name, B0, S, N, weighted, discount, desc = WORKLOADS[wl]
    B = args.envs or B0
    CO, g, cg, cfg, t, episode, core, payload = cpu_selfplay_setup(wl, B, 0, 1000)
    threads = host_threads(CO)
    rng = np.random.default_rng(1)
    total = max(args.warmup, 1) + args.steps
    dn, rn, u = host_inputs(rng, total, B, g.F)
    warm, i = 0.0, 0
    while i < args.warmup or warm < 2.0:  # at least W moves and 2 s: shared host cores ramp up slowly
        j = min(i, args.warmup - 1) if args.warmup > 0 else 0
        warm += cpu_moves(CO, cg, cfg, t, S, 1, dn[j:j + 1], rn[j:j + 1], u[j:j + 1], core, payload, episode, threads)
        i += 1
    t0 = time.perf_counter()
    for i in range(total - args.steps, total):
        cpu_moves(CO, cg, cfg, t, S, 1, dn[i:i + 1], rn[i:i + 1], u[i:i + 1], core, payload, episode, threads)
    dt = time.perf_counter() - t0
    value = B * S * args.steps / dt
    sample = f"{args.steps} moves of {B} envs x {S} sims (full per-GPU workload), tree-major C oracle"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": desc, "envs_per_gpu": B, "envs_total": B, "simulations": S, "max_nodes": N, "branching_factor": g.F,
                   "embedding_bytes": g.payload_bytes + 16, "weighted": weighted, "discount": discount,
                   "note": "reference algorithm on host CPU via the oracle port (oracle/tz_oracle.c, OpenMP over trees); the "
                           "reference itself needs JAX, which is not installable in this image.  One host runs ONE per-GPU "
                           "workload whatever --gpus says (rank 0 only)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    out.append(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def measure_workload(cx, wl, B, K, W, *, use_pdl, legs, strong_total=None):
    """All legs of one workload on this rank's share of `B` envs.  Returns a dict (identical on every rank where it
    matters: times are max-reduced over ranks)."""
    import torch
    import torch.distributed as dist

    tz, _abi, lib, slib, dev, rank, world, args = cx.tz, cx._abi, cx.lib, cx.slib, cx.dev, cx.rank, cx.world, cx.args
    from turbozero_b200.common import step_env_and_evaluator
    from standin.synthetic import SyntheticEnv, SyntheticGame, SyntheticSelfPlay, make_synthetic_evaluator

    name, _, S, N, weighted, discount, desc = WORKLOADS[wl]
    seed = 1000 + rank
    game = SyntheticGame.named(name, seed)
    F, E = game.F, game.emb_bytes
    base = tz.WeightedMCTS if weighted else tz.MCTS
    with_replay = wl == "cfg5"  # "full self-play + replay memory" (BASELINE.json configs[4])
    with_nccl = with_replay and world > 1

    def barrier():
        if world > 1:
            dist.barrier()

    mode = PDL_MODE.get(wl) or (3, 1)  # (forced on for a workload whose default is ordinary launches: both bits)
    pdl_bits = args.pdl_bits if args.pdl_bits > 0 else mode[0]
    leaf_pdl = args.leaf_pdl if args.leaf_pdl >= 0 else mode[1]

    def new_eval(programmatic):
        ev = make_synthetic_evaluator(base, game, action_selector=tz.PUCTSelector(), max_nodes=N, num_iterations=S,
                                      discount=discount, temperature=1.0, programmatic=programmatic)
        ev.sim_warps = args.sim_warps
        return ev

    def launches():
        return lib.tz_launch_count() + slib.tz_synth_launch_count()

    rng = np.random.default_rng(seed)
    total = W + K
    dn_h, rn_h, u_h = host_inputs(rng, total, B, F)

    def flush():
        if cx.flush_buf is not None:
            cx.flush_buf.fill_(1)

    dn_d, rn_d, u_d = (torch.from_numpy(x).to(dev) for x in (dn_h, rn_h, u_h))

    def make_replay(get_core, get_pw, get_done):
        """configs[4] names "full self-play + replay memory": the collection step's buffer update (Trainer.collect,
        core/training/train.py:300-340 -> tz_replay_collect) runs inside the step, fed from static buffers; with more than one
        rank also the training step's two exchange steps, once per move, OUTSIDE the search (north star: "NCCL only for the
        existing gradient mean and replay-memory gather"): one sample of TRAIN_BATCH rows over the buffers of all ranks
        (replay_memory.py:137-183 / train.py:435-437: all-reduce of the valid count, all-gather of every rank's candidates,
        all-reduce of the owners' rows) and the gradient mean (train.py:393,397) of a parameter-sized buffer, enqueued eagerly
        on a second stream behind the step's buffer update (OverlappedStep below).  Returns (collect(move_fn), after_step or None, extra config)."""
        rb = tz.EpisodeReplayBuffer(capacity=REPLAY_CAPACITY)
        obs0 = get_core().to(torch.float32)
        rstate = rb.init(B, tz.BaseExperience(reward=torch.zeros((1,)), policy_weights=torch.zeros((F,)),
                                              policy_mask=torch.zeros((F,), dtype=torch.bool), observation_nn=obs0[0].cpu(),
                                              cur_player_id=torch.zeros((), dtype=torch.int32)), device=dev)
        x_obs, x_mask = torch.empty_like(obs0), torch.ones((B, F), dtype=torch.bool, device=dev)
        x_rew0, x_rew = torch.zeros((B, 1), device=dev), torch.empty((B, 1), device=dev)
        x_player = torch.zeros((B,), dtype=torch.int32, device=dev)
        x_trunc = torch.zeros((B,), dtype=torch.uint8, device=dev)

        def search_half(move_fn):
            x_obs.copy_(get_core())  # the position the search runs on (int32 -> float32 features)
            move_fn()

        def buffer_half():
            x_rew.copy_(get_done().unsqueeze(1))  # stand-in reward: 1 where the episode ended
            exp = tz.BaseExperience(reward=x_rew0, policy_weights=get_pw(), policy_mask=x_mask, observation_nn=x_obs,
                                    cur_player_id=x_player)
            rb.collect_update(rstate, [exp], x_rew, get_done(), x_trunc)

        def collect(move_fn):
            search_half(move_fn)
            buffer_half()

        collect.search_half, collect.buffer_half = search_half, buffer_half
        if not with_nccl:
            return collect, None, {}
        grads = torch.zeros((GRAD_FLOATS,), dtype=torch.float32, device=dev)

        def after():
            rb.sample(rstate, 17, TRAIN_BATCH)
            dist.all_reduce(grads, op=dist.ReduceOp.AVG)

        return collect, after, {"nccl": {"replay_sample_rows": TRAIN_BATCH, "grad_allreduce_bytes": GRAD_FLOATS * 4,
                                         "overlap": "the exchange legs of step i run on a second stream behind its buffer update "
                                                    "and overlap the search of step i+1, whose buffer update waits for them"}}

    class OverlappedStep:
        """configs[4] at more than one GPU: [search of the move] [wait for the previous step's exchange legs] [replay-buffer
        update] on the main stream, then [cross-rank sample + gradient all-reduce] on a second stream.  The sample of step i
        sees exactly the buffer after update i (the next update waits for it), and its NCCL traffic overlaps search i+1."""

        def __init__(self, search, update, after):
            self.search, self.update, self.after = search, update, after
            self.comm = torch.cuda.Stream()
            self.done = None

        def __call__(self):
            main = torch.cuda.current_stream()
            self.search()
            if self.done is not None:
                main.wait_event(self.done)
            self.update()
            ready = torch.cuda.Event()
            ready.record(main)
            self.comm.wait_event(ready)
            with torch.cuda.stream(self.comm):
                self.after()
                self.done = torch.cuda.Event()
                self.done.record(self.comm)

        def join(self):
            torch.cuda.current_stream().wait_stream(self.comm)

    def capture(fn):
        if args.no_graph:
            return fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        return g.replay

    # ---------------- leg 1: device-resident inputs, whole move in the C-ABI (tz_search + leaf callback) ----------
    def timed_moves(programmatic, with_clocks, keep):
        """W warm-up + K timed self-play moves (one CUDA-graph replay each, per-step events, L2 flushed in between)."""
        ev_ = new_eval(pdl_bits if programmatic else False)
        slib.tz_synth_set_programmatic(leaf_pdl if programmatic else 0)
        sp_ = SyntheticSelfPlay(game, ev_, B, env_offset=rank * B, dirichlet=True, device=dev, stats=True)

        def load_inputs_(i):
            sp_.dir_noise.copy_(dn_d[i], non_blocking=True)
            sp_.root_noise.copy_(rn_d[i], non_blocking=True)
            sp_.uniform01.copy_(u_d[i], non_blocking=True)

        one_move = sp_.move
        after_step = None
        extra = {}
        if with_replay:
            collect, after_step, extra = make_replay(lambda: sp_.state["core"], lambda: sp_.policy_weights, lambda: sp_.reset_flag)

            def one_move():
                collect(sp_.move)

        l0 = launches()
        one_move()  # un-captured first move: loads modules, sizes caches
        if after_step:
            after_step()
        torch.cuda.synchronize()
        per_move = launches() - l0
        overlapped = None
        if after_step:  # two graphs, so that the buffer update alone waits for the previous step's exchange legs
            g_search = capture(lambda: collect.search_half(sp_.move))
            g_update = capture(collect.buffer_half)
            overlapped = OverlappedStep(g_search, g_update, after_step)
            step = overlapped

            def replay():
                g_search()
                g_update()
        else:
            replay = capture(one_move)
            step = replay

        for i in range(W):
            load_inputs_(i)
            step()
        torch.cuda.synchronize()
        st0 = sp_.tree.stats.sum(0).cpu().numpy().astype(np.int64)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(cx.local) if (rank == 0 and with_clocks) else None
        t_wall0 = time.perf_counter()
        for i in range(K):
            load_inputs_(W + i)
            flush()
            evs[i][0].record()
            step()
            evs[i][1].record()
        tail_ms = 0.0
        if overlapped is not None:  # the last step's exchange legs have no next search to hide behind: they count in full
            tail = torch.cuda.Event(enable_timing=True)
            overlapped.join()
            tail.record()
            torch.cuda.synchronize()
            tail_ms = evs[K - 1][1].elapsed_time(tail)
        torch.cuda.synchronize()
        ms_wall = time.perf_counter() - t_wall0
        barrier()
        if sampler is not None:
            # the timed region may be shorter than a few sampling periods: keep the SAME load running (untimed replays of the
            # same step) until ~0.6 s of it has been sampled, so that the clocks line describes the GPU under this load
            # (rank 0 only: the graph replay alone, never the step's collectives -- the other ranks are not in this loop)
            t_end = time.perf_counter() + max(0.0, 0.6 - ms_wall)
            while time.perf_counter() < t_end:
                replay()
                torch.cuda.synchronize()
        clocks_ = sampler.stop() if sampler else None
        per_step = [a_.elapsed_time(b_) for a_, b_ in evs]
        d_st = sp_.tree.stats.sum(0).cpu().numpy().astype(np.int64) - st0
        t_ms = torch.tensor([sum(per_step) + tail_ms, -min(per_step), max(per_step), statistics.median(per_step)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        res = {"ms_total": float(t_ms[0]), "ms_min": -float(t_ms[1]), "ms_max": float(t_ms[2]), "ms_median": float(t_ms[3]),
               "levels_per_sim": float(d_st[0]) / max(float(d_st[1]), 1.0), "launches_per_move": int(per_move), "clocks": clocks_,
               "extra": extra}
        if keep:
            res["sp"], res["load_inputs"], res["one_move"] = sp_, load_inputs_, one_move
        return res

    total_envs = strong_total if strong_total else B * world
    other = None
    if use_pdl and "ordinary" in legs:  # the same moves with ordinary launches: what a user's framework kernels get by default
        other = timed_moves(False, False, False)
    main = timed_moves(use_pdl, True, True)
    sp, load_inputs, one_move = main.pop("sp"), main.pop("load_inputs"), main.pop("one_move")
    ms_per_step = main["ms_total"] / K
    value = total_envs * S * K / (main["ms_total"] * 1e-3)

    # ---------------- leg 2: roofline of the per-simulation kernel, timed inside the step it describes ---------------
    roofline = None
    if "roofline" in legs:
        peak, peak_src = measured_peaks()
        moves_r = 3
        # the SAME instantiation the headline runs (programmatic or not, inside a graph), captured once more with the launch
        # timeline on: every search launch and every leaf launch records first-warp-in / inputs-ready / last-warp-out
        sp.timeline_begin()
        mark = sp.timeline_mark()
        if args.no_graph:
            replay_tl = one_move
        else:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            cg_tl = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(cg_tl, stream=side):
                    one_move()
            torch.cuda.current_stream().wait_stream(side)
            replay_tl = cg_tl.replay
        n_search, n_leaf = S + 1, S  # select, S-1 fused launches, expand-only; S leaf launches
        resident, after_inputs, leaf_exec, gap_ls, gap_sl, period = [], [], [], [], [], []
        st0 = sp.tree.stats.sum(0).cpu().numpy().astype(np.int64)
        for m in range(moves_r):
            load_inputs(W + m)
            sp.timeline_clear()
            if args.no_graph:
                mark = sp.timeline_mark()
            flush()
            replay_tl()
            torch.cuda.synchronize()
            ts, tl = sp.timeline_read(mark, n_search, n_leaf)
            # fused launch j (1 <= j <= S-1) sits between leaf j-1 and leaf j
            for j in range(1, S):
                s_in, s_rdy, s_out = (int(x) for x in ts[j])
                l_in, l_rdy, l_out = (int(x) for x in tl[j - 1])
                n_in, n_rdy, n_out = (int(x) for x in tl[j])
                if min(s_in, l_in, n_in) < 0 or s_out == 0 or l_out == 0 or n_out == 0:
                    continue  # a row that was not written (cannot happen; guards the arithmetic)
                resident.append(s_out - s_in)
                after_inputs.append(s_out - max(s_rdy, l_out) if s_rdy else s_out - max(s_in, l_out))
                leaf_exec.append(l_out - l_rdy)
                gap_ls.append(max(s_rdy, s_in) - l_out if s_rdy else s_in - l_out)
                gap_sl.append(n_rdy - s_out)
                period.append(n_rdy - l_rdy)
        sp.timeline_end()
        st1 = sp.tree.stats.sum(0).cpu().numpy().astype(np.int64)
        us = lambda xs: (sum(xs) / max(len(xs), 1)) * 1e-3
        dur_us = us(resident)
        dl, ds = int(st1[0] - st0[0]), int(st1[1] - st0[1])
        bytes_total = algorithmic_bytes(dl, ds, F, E, weighted)
        bytes_per_launch = bytes_total / max(ds // B, 1)
        achieved = bytes_per_launch / (dur_us * 1e-6) / 1e9 if dur_us > 0 else 0.0
        fits = (S - 1) * dur_us * 1e-3 <= ms_per_step * 1.02
        assert fits or args.no_graph, (f"roofline leg inconsistent: {S - 1} launches x {dur_us:.2f} us = {(S - 1) * dur_us * 1e-3:.3f} ms "
                                       f"exceeds ms_per_step {ms_per_step:.3f}")
        traffic = None
        if wl == "cfg2" and B == WORKLOADS["cfg2"][1]:  # per-launch DRAM bytes of this kernel on this shape, from the committed ncu capture
            try:
                with open(os.path.join(ROOT, "profiles", "ksim_traffic.json")) as f:
                    traffic = float(json.load(f)["traffic_bytes_per_launch"])
            except Exception:
                traffic = None
        wide = args.sim_warps > 1 or (args.sim_warps == 0 and F > 32)
        kname = (f"k_sim_wide (a CTA of {args.sim_warps or (2 if weighted else 4)} warps per tree)" if wide else "k_sim (one warp per tree)")
        roofline = {"bound": "hbm", "kernel": kname + ": expand+backprop of simulation i fused with select of i+1, the instantiation "
                    "this step runs, timed inside a replay of the step (first warp in -> last warp out, %globaltimer)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": "profiles/ksim_traffic.json (ncu --set full, cold L2)" if traffic else None,
                    "peak_source": peak_src, "avg_launch_us": dur_us, "median_launch_us": statistics.median(resident) * 1e-3 if resident else None,
                    "launches_timed": len(resident), "launches_per_step": S - 1,
                    "launches_x_avg_ms": (S - 1) * dur_us * 1e-3, "fits_in_ms_per_step": bool(fits),
                    "algorithmic_bytes_per_launch": bytes_per_launch, "levels_per_sim": dl / max(ds, 1),
                    "per_simulation_us": {"period": us(period), "leaf_stand_in_exec": us(leaf_exec),
                                          "leaf_out_to_search_inputs_ready": us(gap_ls), "search_after_inputs_ready": us(after_inputs),
                                          "search_out_to_next_leaf_running": us(gap_sl)},
                    "note": "a launch moves a few MB for ~1 K trees: it is bound by the dependent chain of one warp (or CTA) per "
                            "tree and by the two grid-completion -> dependent-start hand-overs per simulation, not by bandwidth; "
                            "see DESIGN.md section 3"}
        # the re-root launch of three more moves, timed by events around the launch (eager moves: the events bracket one kernel)
        rr_ms, rr_st0 = [], sp.tree.stats.sum(0).cpu().numpy().astype(np.int64)
        blocker = cx.flush_buf if cx.flush_buf is not None else torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        for m in range(moves_r):
            load_inputs(W + m)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sp.reroot_events = (e0, e1)
            for _ in range(4):  # queued fills: the move is enqueued before the GPU reaches it (no host gaps inside)
                blocker.fill_(m)
            sp.move()
            torch.cuda.synchronize()
            rr_ms.append(e0.elapsed_time(e1))
        sp.reroot_events = None
        rr_st = sp.tree.stats.sum(0).cpu().numpy().astype(np.int64) - rr_st0
        # bytes_reroot = 4 nfi + 2 K R + (nfi - K) R per tree, R = 8F + 13 + E (+4 weighted)
        R_row = 8 * F + 13 + E + (4 if weighted else 0)
        rr_bytes = (4 * int(rr_st[2]) + R_row * (int(rr_st[3]) + int(rr_st[2]))) / moves_r
        rr_avg_ms = sum(rr_ms) / len(rr_ms)
        rr_achieved = rr_bytes / (rr_avg_ms * 1e-3) / 1e9
        roofline["reroot"] = {"bound": "hbm", "kernel": "k_reroot_bulk (get_subtree / reset of every tree, one launch per move)",
                              "achieved": rr_achieved, "peak": peak, "unit": "GB/s", "frac": rr_achieved / peak,
                              "avg_launch_us": rr_avg_ms * 1e3, "algorithmic_bytes_per_launch": rr_bytes,
                              "rows_before": int(rr_st[2]) / moves_r / B, "rows_kept": int(rr_st[3]) / moves_r / B}
    slib.tz_synth_set_programmatic(leaf_pdl if use_pdl else 0)
    del sp, one_move, load_inputs

    # ---------------- leg 3: e2e through the public Python API with host buffers -----------------------------------
    e2e = None
    if "e2e" in legs:
        ev2 = new_eval(pdl_bits if use_pdl else False)
        slib.tz_synth_set_programmatic(leaf_pdl if use_pdl else 0)
        env = SyntheticEnv(game, B, env_offset=rank * B, device=dev)
        tree2 = ev2.init_batched(B, game.template_embedding(), device=dev)
        # the step's three random inputs live in ONE device buffer (views below), so the step costs one H2D copy
        s_in = torch.empty((2 * B * F + B,), dtype=torch.float32, device=dev)
        s_dn, s_rn, s_u = s_in[:B * F].view(B, F), s_in[B * F:2 * B * F].view(B, F), s_in[2 * B * F:]
        out_box = {}

        def api_move():
            out, _, _, _, _, _ = step_env_and_evaluator(
                key=None, env_state=env.state, env_state_metadata=env.metadata(), eval_state=tree2, params=None,
                evaluator=ev2, env_step_fn=env.env_step_fn, env_init_fn=None, max_steps=1 << 30,
                leaf_fn=game.leaf_fn, root_noise=s_rn, uniform01=s_u, dirichlet_noise=s_dn)
            out_box["action"], out_box["pw"] = out.action, out.policy_weights

        user_step, e2e_after = api_move, None
        if with_replay:  # the same step as `value`: replay-buffer update inside, the NCCL legs behind it
            out_box["pw"] = torch.zeros((B, F), dtype=torch.float32, device=dev)
            pw_static = out_box["pw"]

            def api_move_static():
                api_move()
                pw_static.copy_(out_box["pw"])
                out_box["pw"] = pw_static

            collect2, e2e_after, _ = make_replay(lambda: env.state["core"], lambda: pw_static, lambda: env.reset_flag)

            def user_step():
                collect2(api_move_static)

        in_p = torch.from_numpy(np.concatenate([dn_h.reshape(total, -1), rn_h.reshape(total, -1), u_h.reshape(total, -1)],
                                               axis=1)).pin_memory()  # [steps, 2BF + B], pinned

        def h2d(i):
            s_in.copy_(in_p[i], non_blocking=True)

        h2d(0)
        l0 = launches()
        user_step()
        torch.cuda.synchronize()
        api_launches = launches() - l0
        if e2e_after:  # (configs[4], more than one GPU) the exchange legs overlap the next step's search, as in `value`
            api_step = OverlappedStep(capture(lambda: collect2.search_half(api_move_static)), capture(collect2.buffer_half), e2e_after)
        else:
            api_step = capture(user_step)

        act_slots = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(2)]
        pw_slots = [torch.empty((B, F), dtype=torch.float32).pin_memory() for _ in range(2)]
        sink = [0.0]

        def enqueue(i, slot):
            """One step: H2D of its inputs, the captured API step (+ the NCCL legs of configs[4], on their own stream), D2H of
            its results.  (The timed region ends with a device-wide synchronize, so the last step's exchange legs count.)"""
            h2d(i)
            api_step()
            act_slots[slot].copy_(out_box["action"], non_blocking=True)
            pw_slots[slot].copy_(out_box["pw"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            return ev

        def consume(ev, slot):
            ev.synchronize()  # the host reads the step's actions and policy weights
            sink[0] += float(act_slots[slot][0]) + float(pw_slots[slot][0, 0])

        def run_pipelined(first, count):
            """The actor loop of a device-resident env: step i+1 is enqueued before the host consumes step i's results (the
            reference's jitted collect loop dispatches asynchronously too), so the launch latency of a step overlaps the GPU's
            work on the previous one; every step's results are read by the host, one step late."""
            prev = None
            for j in range(count):
                ev = enqueue(first + j, j & 1)
                if prev is not None:
                    consume(prev, (j - 1) & 1)
                prev = ev
            consume(prev, (count - 1) & 1)

        def run_lockstep(first, count):
            for j in range(count):
                consume(enqueue(first + j, j & 1), j & 1)  # the host waits for every step before it enqueues the next

        def timed(run):
            run(0, W)
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(W, K)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
            return float(t_e.item())

        dt_lock = timed(run_lockstep)
        dt_pipe = timed(run_pipelined)
        e2e = {"value": total_envs * S * K / dt_pipe, "unit": UNIT,
               "h2d_bytes_per_step": int(2 * B * F * 4 + B * 4), "d2h_bytes_per_step": int(B * 4 + B * F * 4),
               "ms_per_step": dt_pipe / K * 1e3,
               "lockstep": {"value": total_envs * S * K / dt_lock, "ms_per_step": dt_lock / K * 1e3,
                            "note": "the host waits for each step's results before it enqueues the next one: adds the host's "
                                    "graph-launch latency (40-100 us, box dependent) to every step"},
               "api": "step_env_and_evaluator(MCTS.evaluate + MCTS.step) captured in a CUDA graph by the caller; host wall clock "
                      "incl. pinned H2D of every step's noise inputs and D2H of its actions + policy weights, which the host "
                      "reads one step late (step i+1 is enqueued before step i's results are consumed)",
               "launches_per_step": int(api_launches)}
        del tree2, ev2, env

    res = {
        "value": value, "ms_per_step": ms_per_step,
        "per_step_ms": {"min": main["ms_min"], "median": main["ms_median"], "max": main["ms_max"]},
        "levels_per_sim": main["levels_per_sim"], "launches_per_move": main["launches_per_move"], "clocks": main["clocks"],
        "config": {"workload": desc, "envs_per_gpu": B, "envs_total": total_envs, "simulations": S, "max_nodes": N,
                   "branching_factor": F, "embedding_bytes": E, "weighted": weighted, "discount": discount,
                   "programmatic_dependent_launch": use_pdl, "programmatic_bits": pdl_bits if use_pdl else 0,
                   "leaf_stand_in_cooperates": bool(leaf_pdl) if use_pdl else False, "levels_per_sim": main["levels_per_sim"],
                   "sim_warps": args.sim_warps,
                   "step": "one self-play move of all envs: root eval, set_root, S x (select, leaf, expand+backprop), "
                           "root action, env step, re-root" + (", replay-buffer update (tz_replay_collect, capacity "
                                                                f"{REPLAY_CAPACITY})" if with_replay else "")
                           + (f", one cross-rank replay sample of {TRAIN_BATCH} rows and the gradient-mean all-reduce of "
                              f"{GRAD_FLOATS * 4 >> 20} MiB over NCCL" if with_nccl else "")},
        "ordinary_launches": (None if other is None else
                              {"value": total_envs * S * K / (other["ms_total"] * 1e-3), "ms_per_step": other["ms_total"] / K,
                               "note": "same moves with TzSearchCfg.programmatic = 0, the API default (the stand-in leaf kernel "
                                       "launched ordinarily too)"}),
        "roofline": roofline, "e2e": e2e,
    }
    res["config"].update(main["extra"])
    return res


def run_native(args, out):
    import torch
    import torch.distributed as dist

    import turbozero_b200 as tz
    from standin import abi as _sabi
    from turbozero_b200 import _abi

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (native arm) needs a CUDA device; there is no CPU fallback for the search kernels")
    cx = Ctx()
    cx.args, cx.tz, cx._abi = args, tz, _abi
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(cx.local)
    cx.dev = torch.device("cuda", cx.local)
    if cx.world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=cx.dev, timeout=datetime.timedelta(seconds=180))
    cx.lib, cx.slib = _abi.lib(), _sabi.synth_lib()
    cx.flush_buf = None if args.no_flush else torch.empty(512 << 20, dtype=torch.uint8, device=cx.dev)
    world, rank = cx.world, cx.rank

    wl = args.workload
    name, B0, S, N, weighted, discount, desc = WORKLOADS[wl]
    B = args.envs or B0
    K, W = args.steps, max(args.warmup, 3)  # timing rule: never fewer than 3 warm-up steps

    def pdl_for(w):
        return args.pdl or (not args.no_pdl and w not in NO_PDL_BY_DEFAULT)

    legs = {"ordinary"}
    if not args.skip_roofline:
        legs.add("roofline")
    if not args.skip_e2e:
        legs.add("e2e")
    head = measure_workload(cx, wl, B, K, W, use_pdl=pdl_for(wl), legs=legs)

    # ---------------- the configs BASELINE.json names for this GPU count --------------------------------------------
    if args.named is not None:
        named_wls = [w for w in args.named.split(",") if w]
    else:
        named_wls = (["cfg3"] if world in (2, 4, 8) else []) + (["cfg4", "cfg5"] if world == 8 else [])
    named = []
    for nw in named_wls:
        nB = WORKLOADS[nw][1]
        strong = None
        if nw == "cfg3" and world > 1 and CFG3_TOTAL_ENVS % world == 0:
            nB, strong = CFG3_TOTAL_ENVS // world, CFG3_TOTAL_ENVS  # configs[2] is ONE 4096-env job sharded over 2 / 4 / 8 GPUs
        nK = max(3, min(K, 4 if nw == "cfg4" else 8))
        r = measure_workload(cx, nw, nB, nK, 3, use_pdl=pdl_for(nw), legs=legs - {"ordinary"}, strong_total=strong)
        entry = {"name": nw, "metric": METRIC, "unit": UNIT, "value": r["value"], "ms_per_step": r["ms_per_step"], "steps": nK,
                 "warmup": 3, "per_step_ms": r["per_step_ms"], "scaling": "strong" if strong else "weak",
                 "env_steps_per_sec": r["value"] / WORKLOADS[nw][2], "config": r["config"], "e2e": r["e2e"],
                 "roofline": r["roofline"], "launches_per_step": r["launches_per_move"]}
        named.append(entry)

    # ---------------- leg 4: CPU baseline (oracle port) on a bounded sample ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        seed = 1000 + rank
        CO, g, cgm, cfg, t, episode, core, payload = cpu_selfplay_setup(wl, B, 0, seed)
        threads = host_threads(CO)
        dnc, rnc, uc = host_inputs(np.random.default_rng(2), 1, B, g.F)
        warm = 0.0
        while warm < 2.0:  # host cores of a shared box take a second or two to ramp up / schedule all threads
            warm += cpu_moves(CO, cgm, cfg, t, S, 1, dnc, rnc, uc, core, payload, episode, threads)
        moves_done, spent = 0, 0.0
        while spent < args.cpu_seconds and moves_done < 4096:
            spent += cpu_moves(CO, cgm, cfg, t, S, 1, dnc, rnc, uc, core, payload, episode, threads)
            moves_done += 1
        cpu = {"value": B * S * moves_done / spent, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{moves_done} moves of {B} envs x {S} sims ({spent:.1f} s), tree-major C oracle, OpenMP over trees; "
                         f"host has {os.cpu_count()} logical CPUs",
               "numpy_port": numpy_oracle_baseline(wl, seed, seconds=min(6.0, args.cpu_seconds)),
               "note": "the reference itself (JAX on the CPU backend) cannot run in this image; `value` is the multi-threaded C "
                       "restatement (the strongest CPU number), `numpy_port` the literal per-tree NumPy restatement of the "
                       "reference's array program (closest stand-in for its per-example dataflow, one thread)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": head["ms_per_step"], "per_step_ms": head["per_step_ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
            "config": dict(head["config"],
                           l2="not flushed" if args.no_flush else "flushed between steps (512 MiB fill, outside the per-step events)",
                           graph=not args.no_graph, ordinary_launches=head["ordinary_launches"], named=named),
            "ordinary_launches": head["ordinary_launches"],
            "clocks": head["clocks"],
            "env_steps_per_sec": head["value"] / S,  # the metric's second half: self-play env steps of the whole job per second
            "gpu_launches": int(head["launches_per_move"] * K),
            "launches_per_step": int(head["launches_per_move"]),
        }
        if head["roofline"]:
            line["roofline"] = head["roofline"]
        if head["e2e"]:
            line["e2e"] = head["e2e"]
        if cpu:
            line["cpu_baseline"] = cpu
        out.append(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner under torchrun) are
    # sent to stderr for the duration of the run, and the real stdout is restored for the final print
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = []
    try:
        if args.impl == "reference":
            run_reference(args, out)
        else:
            run_native(args, out)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in out:
        print(line, flush=True)


if __name__ == "__main__":
    main()
