"""CPU suite, part 3: the N>1 path of bench.py (regions sharded over ranks, no data-path collective,
results gathered on rank 0) exercised with world_size 2 over gloo.  The per-rank compute is the
oracle here (no GPU in this container); what is under test is the sharding + gather logic."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from __graft_entry__ import load_package
    from _oracle import load_oracle
    import bench
    pkg = load_package()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    whole = pkg.synth.random_small(77, n_regions=9, general_gaps=False)
    mine, off = pkg.shard_regions(whole, rank, world)
    local = load_oracle().batch(mine)["log10"] if mine.n_pairs else np.zeros(0)
    gathered = bench.gather_results(local, off, whole.n_pairs, rank, world)       # the bench's own gather
    tmax = bench.max_over_ranks(float(rank + 1), world)
    if rank == 0:
        q.put((gathered.tolist(), tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_equals_single_process():
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from __graft_entry__ import load_package
    from _oracle import load_oracle
    pkg = load_package()
    whole = pkg.synth.random_small(77, n_regions=9, general_gaps=False)
    want = load_oracle().batch(whole)["log10"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    got, tmax = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert np.array_equal(np.array(got).view(np.uint64), want.view(np.uint64))
    assert tmax == 2.0            # the timing reduction is a MAX over ranks
