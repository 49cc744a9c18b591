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
