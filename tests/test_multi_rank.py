"""N > 1 path on CPU: two gloo ranks each advance their shard of a Monte-Carlo batch (hostsim
test double) and rank 0 gathers the waveforms; the result must equal the single-process batch."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parity_util import GOLDEN, HOSTSIM, ngt, pkg, first_pattern

NS = 4


def _run_shard(lib, lo, hi, max_points=1024):
    flat = ngt.read(f"{GOLDEN}/ro17k.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz")
    wave = ngt.read(f"{GOLDEN}/ro17k.wave.ngt")
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=first_pattern(trace))
    dv = pkg.mc.draw_delvto(NS, 34, seed=11)[lo:hi]
    b = pkg.Batch(circ, hi - lo)
    b.put("b4.inst", pkg.mc.bsim4_inst_with_delvto(lib, flat, dv))
    res = b.tran(max_points, wave["save_eq"][:1])
    t, v = res.waves()
    return res, t, v


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = pkg.Library(HOSTSIM)
    lo, hi = pkg.parallel.shard(NS, rank, world)
    res, t, v = _run_shard(lib, lo, hi)
    allv = pkg.parallel.gather_results(dist, torch.from_numpy(v[:, :, 0].copy()))
    allt = pkg.parallel.gather_results(dist, torch.from_numpy(t.copy()))
    acc = pkg.parallel.gather_results(dist, torch.from_numpy(res.accepted.astype(np.int64)))
    if rank == 0:
        np.savez(out, v=allv.numpy(), t=allt.numpy(), acc=acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_single_process(hostsim_lib, tmp_path):
    out = str(tmp_path / "gathered.npz")
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    g = np.load(out)
    res, t, v = _run_shard(hostsim_lib, 0, NS)
    assert np.array_equal(g["acc"], res.accepted.astype(np.int64))
    assert np.array_equal(g["t"], t) and np.array_equal(g["v"], v[:, :, 0])


def test_shard_covers_all_samples():
    for n, w in ((4096, 8), (10, 4), (3, 8)):
        seen = []
        for r in range(w):
            lo, hi = pkg.parallel.shard(n, r, w)
            seen += list(range(lo, hi))
        assert seen == list(range(n))
