"""Multi-process host logic (time / latitude-band sharding + slab gather) on CPU with gloo, world_size 2 and 3.
The per-rank compute stand-in is the CPU oracle; on the GPU box the same plans drive libcdfgpu (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cdftools_b200 import shard, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    m = synth.make_mesh("SMALL")
    ib = oracle.basin_masks(*synth.basin_mask_inputs(m))
    e3m = oracle.mask_e3v(m.e3v_0, m.vmask.astype(np.float32))
    nrec = 5
    if mode == "time":
        mine = shard.time_plan(nrec, world, rank)
        loc = np.stack([oracle.cdfmoc_record(m.e1v, e3m, ib, synth.make_v_record(m, jt)[:-1]) for jt in mine]) \
            if mine else np.zeros((0, m.nz, m.ny, 5))
        full = shard.gather_time_slabs(torch.from_numpy(loc), nrec)
    elif mode == "band_moc":
        j0, j1 = shard.band_plan(m.ny, world)[rank]
        v = synth.make_v_record(m, 1)[:-1]
        loc = oracle.cdfmoc_record(shard.band_slice(m.e1v, j0, j1, 0), shard.band_slice(e3m, j0, j1, 1),
                                   shard.band_slice(ib, j0, j1, 0), shard.band_slice(v, j0, j1, 1))
        full = shard.gather_band_slabs(torch.from_numpy(loc), m.ny, row_dim=1)
    else:  # band_mocsig: the oracle skips the first/last row of what it is given, so bands carry a one-row halo
        j0, j1 = shard.band_plan(m.ny, world)[rank]
        h0, h1 = max(j0 - 1, 0), min(j1 + 1, m.ny)
        v = synth.make_v_record(m, 1)[:-1]
        t, s = (x[:-1] for x in synth.make_ts_record(m, 1))
        loc, _ = oracle.cdfmocsig_record(shard.band_slice(m.e1v, h0, h1, 0), shard.band_slice(m.e3v_0, h0, h1, 1),
                                         shard.band_slice(ib, h0, h1, 0), shard.band_slice(v, h0, h1, 1),
                                         shard.band_slice(t, h0, h1, 1), shard.band_slice(s, h0, h1, 1),
                                         0.0, 0.0, 0.0, 2000.0, 0, 30.0, 0.05, 158)
        loc = loc[j0 - h0: loc.shape[0] - (h1 - j1)]
        full = shard.gather_band_slabs(torch.from_numpy(np.ascontiguousarray(loc)), m.ny, row_dim=0)
    if rank == 0:
        q.put(full.numpy())
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def _run(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_plans():
    assert shard.time_plan(7, 3, 0) == [0, 3, 6] and shard.time_plan(7, 3, 2) == [2, 5]
    assert sum(len(shard.time_plan(365, 8, r)) for r in range(8)) == 365
    b = shard.band_plan(3059, 8)
    assert b[0][0] == 0 and b[-1][1] == 3059 and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert max(j1 - j0 for j0, j1 in b) - min(j1 - j0 for j0, j1 in b) <= 1


@pytest.mark.parametrize("world", [2, 3])
def test_time_sharding_matches_single_process(world, oracle_mod):
    got = _run(world, "time")
    m = synth.make_mesh("SMALL")
    ib = oracle_mod.basin_masks(*synth.basin_mask_inputs(m))
    e3m = oracle_mod.mask_e3v(m.e3v_0, m.vmask.astype(np.float32))
    ref = np.stack([oracle_mod.cdfmoc_record(m.e1v, e3m, ib, synth.make_v_record(m, jt)[:-1]) for jt in range(5)])
    assert np.array_equal(got, ref)


def test_band_sharding_cdfmoc(oracle_mod):
    got = _run(2, "band_moc")
    m = synth.make_mesh("SMALL")
    ib = oracle_mod.basin_masks(*synth.basin_mask_inputs(m))
    e3m = oracle_mod.mask_e3v(m.e3v_0, m.vmask.astype(np.float32))
    assert np.array_equal(got, oracle_mod.cdfmoc_record(m.e1v, e3m, ib, synth.make_v_record(m, 1)[:-1]))


def test_band_sharding_cdfmocsig(oracle_mod):
    got = _run(2, "band_mocsig")
    m = synth.make_mesh("SMALL")
    ib = oracle_mod.basin_masks(*synth.basin_mask_inputs(m))
    v = synth.make_v_record(m, 1)[:-1]
    t, s = (x[:-1] for x in synth.make_ts_record(m, 1))
    ref, _ = oracle_mod.cdfmocsig_record(m.e1v, m.e3v_0, ib, v, t, s, 0.0, 0.0, 0.0, 2000.0, 0, 30.0, 0.05, 158)
    assert np.array_equal(got, ref)
