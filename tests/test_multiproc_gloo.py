"""world_size-2 `gloo` test of the multi-GPU path's host logic: the env batch is split in contiguous
shards (core/common.py:12-29), every rank searches ITS shard with no data-path collective, and the per-rank results
gathered in rank order equal the unsharded run.  The search itself runs on the CPU oracle here (no GPU in CI); the
sharding arithmetic, env offsets, input slicing and the max-over-ranks timing reduction are the code under test."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import Schedule, run_c_treemajor
    from oracle import synth_numpy as SN
    from turbozero_b200.common import shard_slice

    g = SN.SynthGame(F=7, payload_bytes=8, rho256=230, tau1024=12, max_depth=42, seed=33)
    full = Schedule(game=g, B=B, N=48, S=24, moves=3)  # identical on every rank (same seed): the global inputs
    sl = shard_slice(B, rank, world)
    shard = Schedule(game=g, B=sl.stop - sl.start, N=48, S=24, moves=3, env_offset=sl.start)
    shard.dir_noise, shard.root_noise, shard.uniform01 = (np.ascontiguousarray(x[:, sl]) for x in
                                                          (full.dir_noise, full.root_noise, full.uniform01))
    res = run_c_treemajor(shard, nthreads=2)
    # gather actions (moves, B/D) in rank order, like the reference gathers pmap outputs
    mine = torch.from_numpy(res.actions.copy())
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    elapsed = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing rule
    if rank == 0:
        np.save(os.path.join(out_dir, "actions.npy"), torch.cat(parts, dim=1).numpy())
        np.save(os.path.join(out_dir, "elapsed.npy"), elapsed.numpy())
    np.save(os.path.join(out_dir, f"nfi{rank}.npy"), res.arrays["next_free_idx"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_search_equals_unsharded(tmp_path):
    from helpers import Schedule, run_c_treemajor
    from oracle import synth_numpy as SN

    B, world = 8, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    g = SN.SynthGame(F=7, payload_bytes=8, rho256=230, tau1024=12, max_depth=42, seed=33)
    ref = run_c_treemajor(Schedule(game=g, B=B, N=48, S=24, moves=3))
    assert np.array_equal(np.load(tmp_path / "actions.npy"), ref.actions)
    nfi = np.concatenate([np.load(tmp_path / f"nfi{r}.npy") for r in range(world)])
    assert np.array_equal(nfi, ref.arrays["next_free_idx"])
    assert np.load(tmp_path / "elapsed.npy")[0] == pytest.approx(0.002)


def _merge_worker(rank, world, port, n_per_rank, k, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from turbozero_b200.common import merge_topk

    scores = torch.from_numpy(np.load(os.path.join(out_dir, "scores.npy"))[rank])
    order = torch.sort(scores, stable=True).indices[:k]  # this rank's k best, like EpisodeReplayBuffer.sample
    owner, local = merge_topk(scores[order], order + rank * n_per_rank, k, n_per_rank)
    if rank == 1:
        np.save(os.path.join(out_dir, "owner.npy"), owner.numpy())
        np.save(os.path.join(out_dir, "local.npy"), local.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_cross_rank_replay_sample_merge_equals_global_argsort(tmp_path):
    """EpisodeReplayBuffer.sample across ranks (core/memory/replay_memory.py:157-169 samples over the device axis): the
    winners of the per-rank candidate merge are the first k of one stable argsort over the concatenated blocks."""
    world, n, k = 2, 40, 12
    rng = np.random.default_rng(3)
    scores = rng.standard_normal((world, n)).astype(np.float32)
    scores[0, 5:20] = np.inf          # unsampleable slots
    scores[1, 3] = scores[0, 2]       # a tie across ranks: the lower global index wins
    scores[1, :30] = np.inf           # rank 1 has fewer than k candidates
    scores[1, 3] = scores[0, 2]
    np.save(tmp_path / "scores.npy", scores)
    port = _free_port()
    mp.spawn(_merge_worker, args=(world, port, n, k, str(tmp_path)), nprocs=world, join=True)
    ref = np.argsort(scores.reshape(-1), kind="stable")[:k]
    assert np.array_equal(np.load(tmp_path / "owner.npy") * n + np.load(tmp_path / "local.npy"), ref)
