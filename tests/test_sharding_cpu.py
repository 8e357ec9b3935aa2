"""Multi-GPU host logic on CPU: world_size-2 `gloo` run of the sharding arithmetic bench.py uses
(step-major, rank-minor global sim ids; one all-reduce of integer counters).  The per-rank work is done
by the oracle here (no GPU in this container); the property under test is that the union of shards is
bit-identical to one contiguous run and independent of the rank count."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, per, steps, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mcpower_b200 as M
    from mcpower_b200 import spec as S
    from oracle import orc
    spec = S.sweep_spec()
    grid = [40, 80]
    total = None
    for i in range(steps):
        begin = (i * world + rank) * per           # same formula as bench.py
        r = orc.run_batch(spec, grid, per, 77, sim_begin=begin)
        c = M.counters_from_bytes(r["uncorrected"], r["corrected"], r["converged"], r["overall"], r["factor_excluded"])
        total = c if total is None else total + c
    t = torch.from_numpy(total.astype(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(Path(out_dir) / "counters.npy", t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_run(tmp_path):
    sys.path.insert(0, str(ROOT))
    import mcpower_b200 as M
    from mcpower_b200 import spec as S
    from oracle import orc
    per, steps, world = 48, 2, 2
    mp.spawn(_worker, args=(world, 29500 + os.getpid() % 1000, per, steps, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "counters.npy")
    r = orc.run_batch(S.sweep_spec(), [40, 80], per * steps * world, 77)
    want = M.counters_from_bytes(r["uncorrected"], r["corrected"], r["converged"], r["overall"], r["factor_excluded"])
    assert np.array_equal(got, want.astype(np.int64))
