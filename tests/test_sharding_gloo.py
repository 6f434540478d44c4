"""The N>1 host logic on CPU: interval sharding + the single gather of call records, world_size 2 over gloo (no GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pisces_b200 import _native, sharding


def test_shard_loci_balances_and_cuts_on_blocks():
    rng = np.random.default_rng(0)
    w = rng.poisson(500, 10_500)
    w[3000:4000] *= 4
    for world in (1, 2, 4, 8):
        sh = sharding.shard_loci(w, world, first_position=1)
        assert sh[0][0] == 0 and sh[-1][1] == len(w)
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))                      # contiguous, no overlap
        assert all(lo % 1000 == 0 for lo, _ in sh[1:])                            # cuts at 1000-bp block starts
        loads = [w[lo:hi].sum() for lo, hi in sh]
        assert max(loads) <= 1.35 * sum(loads) / world + 4 * 500 * 1000
    # positions that do not start on a block boundary: first_position 501 -> first cut after 500 loci
    sh = sharding.shard_loci(np.ones(3000), 2, first_position=501)
    assert sh[1][0] % 1000 == 500
    assert sharding.shard_loci([], 2) == [(0, 0), (0, 0)]


def test_shard_plan_cuts_on_blocks_with_halo():
    """pb2_shard_plan (host logic, no GPU): block-aligned cuts balanced by read starts, halo = two blocks + the read span, read windows that hold every
    read able to touch the staged range."""
    rng = np.random.default_rng(1)
    pos0 = np.sort(np.concatenate([rng.integers(0, 40_000, 30_000), rng.integers(10_000, 14_000, 30_000)])).astype(np.int32)   # a deep stretch
    span = 150
    for n in (1, 2, 4, 8):
        plan = sharding.shard_plan(pos0, 1, 40_150, span, n)
        assert plan[0]["own_lo"] == 1 and plan[-1]["own_hi"] == 40_150
        assert all(a["own_hi"] + 1 == b["own_lo"] for a, b in zip(plan, plan[1:]))
        assert all(s["own_hi"] % 1000 == 0 for s in plan[:-1])                         # cuts at block boundaries
        for s in plan:
            assert s["stage_lo"] == max(1, s["own_lo"] - 3000) and s["stage_hi"] == s["own_hi"] + 3000
            touching = np.nonzero((pos0 + 1 + span >= s["stage_lo"]) & (pos0 + 1 <= s["stage_hi"]))[0]
            if len(touching):
                assert s["read_first"] <= touching[0] and s["read_end"] >= touching[-1] + 1
        loads = [int(((pos0 + 1 >= s["own_lo"]) & (pos0 + 1 <= s["own_hi"])).sum()) for s in plan]
        assert max(loads) <= len(pos0) / n + 0.2 * len(pos0)                          # balanced by reads, up to one block of the deep stretch
    by_pos = sharding.shard_plan(None, 1, 10_000, 100, 4)
    assert [s["own_hi"] for s in by_pos] == [3000, 5000, 8000, 10_000] or [s["own_hi"] for s in by_pos] == [2000, 5000, 8000, 10_000] or all(s["own_hi"] % 1000 == 0 for s in by_pos)


def _fake_records(lo, hi, step):
    pos = np.arange(lo, hi, step, dtype=np.int32)
    r = np.zeros(len(pos), dtype=_native.RECORD_DTYPE)
    r["position"] = pos
    r["variant_qscore"] = pos % 101
    r["total_coverage"] = 500
    return r


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards = sharding.shard_loci(np.full(5000, 300), world, first_position=1)
    lo, hi = shards[rank]
    mine = _fake_records(lo + 1, hi + 1, 7 if rank == 0 else 3)      # ragged: ranks contribute different record counts
    got = sharding.gather_call_records(mine)
    q.put((rank, got["position"].tolist(), got["variant_qscore"].tolist()))
    dist.destroy_process_group()


def test_gather_call_records_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = sharding.shard_loci(np.full(5000, 300), 2, first_position=1)
    exp = np.concatenate([_fake_records(shards[0][0] + 1, shards[0][1] + 1, 7), _fake_records(shards[1][0] + 1, shards[1][1] + 1, 3)])
    for rank, pos, vq in res:
        assert pos == exp["position"].tolist() and vq == exp["variant_qscore"].tolist()
        assert pos == sorted(pos)                                     # rank order == genome order
