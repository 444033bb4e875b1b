"""Multi-rank path on CPU (gloo, world_size 2): contiguous column shards solved
independently and gathered give exactly the single-rank result; the synthetic
generator does not depend on the sharding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hostcheck_lib
import oracle_lib
from spartacus_surface_b200 import config_type, canopy_flux_type, boundary_conds_out_type
from spartacus_surface_b200.sharding import shard_columns, gather_columns
from spartacus_surface_b200.synthetic import make_synthetic

NCOL, NLAY = 96, 6


def _config():
    cfg = config_type(n_vegetation_region_urban=2, n_vegetation_region_forest=2, n_stream_sw_urban=2,
                      n_stream_lw_urban=2, n_stream_sw_forest=2, n_stream_lw_forest=2)
    return cfg.consolidate(oracle_lib.legendre_gauss_init)


def _solve(cfg, ncol, col_offset):
    cp, sw, lw = make_synthetic(cfg, ncol, NLAY, col_offset=col_offset)
    bc = boundary_conds_out_type().allocate(ncol, 1, 1)
    fl = [canopy_flux_type().allocate(cfg, ncol, cp.ntotlay, 1, use_direct=d, do_save_flux_profile=False)
          for d in (True, True, False, False)]
    rc = hostcheck_lib.make_solver(fast=True)(cfg, cp, sw, lw, bc, None, None, *fl)
    assert rc == 0
    return bc, fl


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = _config()
    c0, c1 = shard_columns(np.full(NCOL, NLAY), world)[rank]
    bc, fl = _solve(cfg, c1 - c0, c0)
    gathered = {
        "top_net": gather_columns(fl[0].top_net, dist),
        "veg_abs": gather_columns(fl[0].veg_abs.reshape(c1 - c0, NLAY), dist),
        "lw_emission": gather_columns(bc.lw_emission, dist),
    }
    dist.barrier()
    if rank == 0:
        np.savez(out_path, **{k: v.numpy() for k, v in gathered.items()})
    dist.destroy_process_group()


def test_shard_columns_balances_layers():
    nlay = np.array([0, 5, 5, 0, 10, 1, 1, 1, 7, 3])
    shards = shard_columns(nlay, 3)
    assert shards[0][0] == 0 and shards[-1][1] == nlay.size
    assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
    loads = [int(np.maximum(nlay[a:b], 1).sum()) for a, b in shards]
    assert max(loads) - min(loads) <= 10
    assert shard_columns(np.full(8, 16), 8) == [(i, i + 1) for i in range(8)]


def test_two_rank_gloo_matches_single_rank(tmp_path):
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, 29531 + os.getpid() % 500, out), nprocs=2, join=True)
    got = np.load(out)
    cfg = _config()
    bc, fl = _solve(cfg, NCOL, 0)
    assert np.array_equal(got["top_net"], fl[0].top_net)
    assert np.array_equal(got["veg_abs"], fl[0].veg_abs.reshape(NCOL, NLAY))
    assert np.array_equal(got["lw_emission"], bc.lw_emission)
