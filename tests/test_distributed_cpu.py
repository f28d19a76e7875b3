"""CPU, world_size 2 (gloo): the N>1 host logic of bench.py — shard split of one seeded global batch and the
metrics gather (the only communication on this path; problems are independent)."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from rrt_mpc_b200.synthetic import make_batch
    total = 601                                                          # a global batch that does not divide evenly
    start, count = bench.shard_range(rank, world, total)
    x0, ref, up = make_batch(total, 20, seed=5, start=start, count=count)
    local = {"ms_total": 10.0 + rank, "solves": float(count), "checksum": float(x0.sum() + ref.sum() + up.sum())}
    allm = bench.gather_metrics(local, world)
    q.put((rank, start, count, allm))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from rrt_mpc_b200.synthetic import make_batch
    full = make_batch(601, 20, seed=5)
    (r0, s0, c0, m0), (r1, s1, c1, m1) = got
    assert (s0, c0, s1, c1) == (0, 301, 301, 300)                       # disjoint, covering shards of the fixed global batch (strong scaling)
    assert m0 == m1 and len(m0) == 2                                    # every rank sees every rank's metrics
    chk = [float(full[0][a:b].sum() + full[1][a:b].sum() + full[2][a:b].sum()) for a, b in ((0, 301), (301, 601))]
    assert np.allclose([m0[0]["checksum"], m0[1]["checksum"]], chk, rtol=1e-12)
    # whole-job value = all solves / max-over-ranks time
    t = max(m["ms_total"] for m in m0)
    assert t == 11.0 and sum(m["solves"] for m in m0) == 601.0
    import bench
    assert [bench.shard_range(r, 8, 1 << 20) for r in (0, 7)] == [(0, 131072), (917504, 131072)]       # configs[4] over 8 GPUs


def test_flop_model_matches_baseline_table():
    import bench
    for N, fi, fc in ((15, 9704, 20570), (20, 12894, 27225), (50, 32026, 67155)):
        fm = bench.flop_model(N)
        assert fm["f_iter"] == fi and fm["f_chol"] == fc
    assert bench.io_bytes(50, 1) == (1680, 2456) and bench.io_bytes(20, 1) == (720, 1016)
