"""CPU, world_size 2 over gloo: the sharding / gather logic of layoutdm_b200.parallel.  The per-shard compute is
stood in for by the oracle (allowed in tests only); the GPU tests cover the kernels' global-index noise keying."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from layoutdm_b200.parallel import all_gather_ids, sample_sharded, shard_bounds, shard_cond


def test_shard_bounds_cover_everything():
    for total in (1, 7, 8, 1024, 8192, 1000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_shard_cond_slices_rows_and_broadcasts():
    cond = dict(seq=torch.arange(40).view(8, 5), mask=torch.ones(8, 5, dtype=torch.bool), type="c", refine_table=torch.zeros(3, 3))
    c = shard_cond(cond, 2, 5)
    assert torch.equal(c["seq"], cond["seq"][2:5]) and c["type"] == "c" and c["refine_table"].shape == (3, 3)
    one = dict(seq=torch.arange(5).view(1, 5), type="c")
    assert shard_cond(one, 3, 6)["seq"].shape == (1, 5)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from oracle import layoutdm_oracle as O
        vo, spec = O.RICO25, O.ModelSpec(layers=1)
        sd = O.make_weights(vo, spec, seed=2)
        orc = O.Oracle(vo, spec, sd)
        cfg = O.SamplingCfg(name="random", num_timesteps=2)

        def sample_fn(batch_size, cond, b_global0, seed):
            return orc.sample(batch_size, cfg, seed=seed, cond=cond, b_global0=b_global0)

        ids = sample_sharded(sample_fn, total, seed=4)
        # ragged gather primitive
        lo, hi = shard_bounds(total, world, rank)
        g = all_gather_ids(torch.arange(lo, hi).view(-1, 1).repeat(1, 3), total)
        assert torch.equal(g[:, 0], torch.arange(total))
        if rank == 0:
            q.put(ids)
    finally:
        dist.destroy_process_group()


def test_sharded_sampling_is_invariant_to_world_size():
    from oracle import layoutdm_oracle as O
    total = 5          # ragged over 2 ranks: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    ids2 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vo, spec = O.RICO25, O.ModelSpec(layers=1)
    orc = O.Oracle(vo, spec, O.make_weights(vo, spec, seed=2))
    ids1 = orc.sample(total, O.SamplingCfg(name="random", num_timesteps=2), seed=4)
    assert torch.equal(ids1, ids2)
