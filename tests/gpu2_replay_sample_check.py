"""Two-GPU check (run under torchrun, NCCL; not collected by pytest -- the `-m gpu` suite is single-GPU):
EpisodeReplayBuffer.sample across ranks equals the NumPy oracle's sample over the concatenation of the ranks' buffers
(core/memory/replay_memory.py:137-183 samples over the device axis of the (D, B/D, cap) state).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu2_replay_sample_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import turbozero_b200 as tz  # noqa: E402
from oracle import replay_numpy as RN  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, cap, P, F, steps, k = 96, 16, 2, 5, 30, 200
    tmpl = {"reward": np.zeros((P,), np.float32), "policy_weights": np.zeros((F,), np.float32), "policy_mask": np.zeros((F,), bool),
            "observation_nn": np.zeros((3,), np.float32), "cur_player_id": np.zeros((), np.int32)}
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    states = []
    for r in range(world):  # every rank builds EVERY rank's oracle state (same seeds), and its own buffer on its GPU
        rng = np.random.default_rng(100 + r)
        s = RN.init(B, cap, tmpl)
        if r == rank:
            buf = tz.EpisodeReplayBuffer(capacity=cap)
            st = buf.init(B, tz.BaseExperience(**{kk: torch.from_numpy(v) for kk, v in tmpl.items()}))
        for _ in range(steps):
            e = {"observation_nn": rng.standard_normal((B, 3)).astype(np.float32), "policy_mask": rng.random((B, F)) < 0.7,
                 "policy_weights": rng.random((B, F)).astype(np.float32), "reward": np.zeros((B, P), np.float32),
                 "cur_player_id": rng.integers(0, 2, (B,)).astype(np.int32)}
            rew, term, trunc = rng.standard_normal((B, P)).astype(np.float32), rng.random(B) < 0.15, rng.random(B) < 0.05
            RN.collect_update(s, [e], rew, term, trunc, cap)
            if r == rank:
                buf.collect_update(st, [tz.BaseExperience(**{kk: dev(v) for kk, v in e.items()})], dev(rew), dev(term), dev(trunc))
        states.append(s)
    gumbels = [np.random.default_rng(7 + r).gumbel(size=(B * cap,)).astype(np.float32) for r in range(world)]
    got = buf.sample(st, None, k, gumbel=dev(gumbels[rank]))
    # oracle: one state holding all ranks' envs, rank-major (the reference's (D, B/D, cap) flattened)
    big = RN.init(B * world, cap, tmpl)
    big.populated = np.concatenate([s.populated for s in states])
    big.has_reward = np.concatenate([s.has_reward for s in states])
    big.buffer = {kk: np.concatenate([s.buffer[kk] for s in states]) for kk in tmpl}
    ref = RN.sample(big, np.concatenate(gumbels), k)
    for kk in tmpl:
        a = getattr(got, kk).cpu().numpy()
        assert a.shape == ref[kk].shape and np.array_equal(a, ref[kk]), f"rank {rank}: leaf {kk} differs from the oracle"
    ok = torch.ones(1, device="cuda")
    dist.all_reduce(ok)
    if rank == 0:
        print(f"cross-rank replay sample over {world} ranks: {k} rows x {len(tmpl)} leaves bit-exact vs oracle/replay_numpy.py")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
