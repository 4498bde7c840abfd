"""world_size-2 CPU test (gloo) of the multi-GPU host logic: contiguous sharding
with global env ids, and the statistics all-reduce.  The env arithmetic is done
by the oracle here (no GPU in this container); what is under test is
guardx_b200.distributed and the rank-invariance of the keying scheme."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from guardx_b200 import distributed as gxd

TOTAL, T, POOL = 200, 120, 4096
CFG = {'num_steps': 25, '_seed': 4}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import OracleEngine
    n, off = gxd.shard(TOTAL, rank, world)
    env = OracleEngine(dict(CFG, env_num=n), env_id_offset=off, pool_size=POOL)
    obs0 = env.reset()
    ro = env.rollout(T)
    stats = torch.from_numpy(env.stats().copy())
    gxd.allreduce_stats(stats)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), obs0=obs0, obs=ro["obs"], done=ro["done"],
             reward=ro["reward"], stats=stats.numpy(), n=n, off=off)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_everything():
    for total, world in ((4096, 8), (65536, 4), (10, 3), (7, 7)):
        parts = [gxd.shard(total, r, world) for r in range(world)]
        assert sum(n for n, _ in parts) == total
        assert parts[0][1] == 0
        for (n0, o0), (_, o1) in zip(parts, parts[1:]):
            assert o1 == o0 + n0


@pytest.mark.timeout(300)
def test_two_ranks_reproduce_one_batch(tmp_path):
    from oracle.oracle import OracleEngine
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    whole = OracleEngine(dict(CFG, env_num=TOTAL), pool_size=POOL)
    obs0 = whole.reset()
    ro = whole.rollout(T)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    assert [int(p["off"]) for p in parts] == [0, 100]
    np.testing.assert_array_equal(np.concatenate([p["obs0"] for p in parts]), obs0)
    for k in ("obs", "done", "reward"):
        np.testing.assert_array_equal(np.concatenate([p[k] for p in parts], axis=1), ro[k])
    # the all-reduced statistics equal the single-process ones on every rank
    ref = whole.stats()
    assert ref[gxd.N_EPISODES] > TOTAL
    for p in parts:
        np.testing.assert_allclose(p["stats"], ref, rtol=1e-12)
    summ = gxd.episode_summary(parts[0]["stats"])
    assert summ['TotalEnvInteracts'] == TOTAL * T and summ['EpLen'] > 1 and summ['CostRate'] >= 0
