"""N > 1 path on CPU: two gloo ranks shard independent streams, decode their shards (the CPU oracle stands in for the
GPU here — test infrastructure only) and exchange nothing but per-stream checksums, unit counts and times. The union
must equal the single-process result, and the aggregate must be sum(units) / max(time)."""
import os
import socket

import numpy as np
import pytest

from parseoggvorbis_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 1000):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_shard_by_cost_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    costs = rng.integers(50, 5000, size=101)
    for w in (2, 4, 8):
        parts = sharding.shard_by_cost(costs, w)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        loads = [int(costs[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= costs.max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from parseoggvorbis_b200 import abi, workloads
        from tests import oracle_binding as ob
        setup, batch = workloads.config2(P=60, streams=5, distinct=5, seed=3)     # the same corpus on every rank
        lo, hi = sharding.shard_range(len(batch.streams), world, rank)
        sums, units = [], 0
        for s in range(lo, hi):
            st = batch.streams[s:s + 1].copy()
            p0 = int(st["first_packet"][0]); p1 = p0 + int(st["n_packets"][0])
            pk = batch.packets[p0:p1].copy()
            pk["stream"] = 0
            st["first_packet"] = 0; st["pcm_base"] = 0; st["setup_id"] = 0
            sub = abi.Batch(st, pk, batch.ys, batch.payload, int(st["pcm_frames"][0]) * setup.channels)
            pcm, status = ob.synth_batch([setup], sub, imdct="fast")
            assert not status.any()
            sums.append(float(np.abs(pcm.astype(np.float64)).sum()))
            units += pcm.size
        all_sums = sharding.merge_checksums(np.asarray(sums), dist)
        tot_units, ms = sharding.aggregate_throughput(units, 10.0 * (rank + 1), dist)
        q.put((rank, all_sums.tolist(), tot_units, ms, units))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_equal_one_process():
    import torch.multiprocessing as mp
    from parseoggvorbis_b200 import workloads
    from tests import oracle_binding as ob
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    setup, batch = workloads.config2(P=60, streams=5, distinct=5, seed=3)
    ref, status = ob.synth_batch([setup], batch, imdct="fast")
    assert not status.any()
    C = setup.channels
    expect = []
    for s in range(len(batch.streams)):
        b0 = int(batch.streams["pcm_base"][s]); n = int(batch.streams["pcm_frames"][s]) * C
        expect.append(float(np.abs(ref[b0:b0 + n].astype(np.float64)).sum()))
    for rank, sums, tot_units, ms, units in res:
        assert np.allclose(sums, expect, rtol=0, atol=0)         # same code, same inputs: bit-identical sums
        assert tot_units == ref.size                               # every stream decoded exactly once
        assert ms == 20.0                                          # max over ranks
    assert sum(r[4] for r in res) == ref.size
