"""world_size-2 gloo test of the N>1 host logic (sequence sharding, max-over-ranks timing, result gather).
The per-sequence work is done by the CPU oracle here (test infrastructure standing in for the GPU, which
this container does not have); the sharded result must equal the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nav24_b200 import shard

N_SEQ, H, W, NF = 5, 260, 340, 300


def _run_sequence(seq):
    from nav24_b200.synth import sequence
    from oracle import orb_oracle as oo
    fr = sequence(H, W, 1000 * seq + 3, 2, step=(3, 1))
    o = oo.OrbOracle(NF)
    (_, k1, d1), (_, k2, d2) = o.detect(fr[0]), o.detect(fr[1])
    ud1 = np.stack([k1["x"], k1["y"]], 1); ud2 = np.stack([k2["x"], k2["y"]], 1)
    m = oo.match_window(k1, ud1, d1, k2, ud2, d2, oo.grid_for(W, H))
    return (2, len(k1) + len(k2), int((m >= 0).sum()), shard.digest(k1, d1, k2, d2, m))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.sequences_for_rank(N_SEQ, rank, world)
        summary = {s: _run_sequence(s) for s in mine}
        dist.barrier()
        ms = shard.max_over_ranks(10.0 + rank)
        merged = shard.gather_summaries(summary)
        if rank == 0:
            out.put((ms, merged))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_round_robin_ownership():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            own = shard.sequences_for_rank(64, r, world)
            assert all(shard.owner_of(s, world) == r for s in own)
            seen += own
        assert sorted(seen) == list(range(64))
    assert len(shard.sequences_for_rank(64, 3, 8)) == 8
    with pytest.raises(ValueError):
        shard.sequences_for_rank(4, 2, 2)


@pytest.mark.timeout(300)
def test_two_rank_gloo_equals_unsharded():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ms, merged = out.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ms == 11.0                                   # MAX over ranks
    assert sorted(merged) == list(range(N_SEQ))
    for s in range(N_SEQ):
        assert merged[s] == _run_sequence(s)
    assert sum(v[2] for v in merged.values()) > 20
