"""world_size-2 gloo tests of the multi-rank drivers (host logic + collectives), CPU only."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cpu_standin_backend import OracleBackend
        from bixverse_b200 import distributed as D
        from bixverse_b200.synthetic import synthetic_bulk, synthetic_sparse_cells
        from oracle import oracle as o

        be = OracleBackend()
        n, p = 40, 300
        x = synthetic_bulk(n, p, seed=5)
        xt = torch.from_numpy(np.ascontiguousarray(x.T))  # p x n rows == column-major n x p
        # --- bulk correlation: contiguous, disjoint, complete packed ranges
        sl, (o0, o1) = D.sharded_cor_upper(be, xt, n, p, True, 1, 6.0)
        full = o.soft_threshold(o.cor_upper_triangle(x, True, True), 6.0)
        assert np.abs(sl.numpy() - full[o0:o1]).max() < 1e-14  # gene-sharded ranks + operand all-gather
        rng = [None] * world
        dist.all_gather_object(rng, (o0, o1))
        assert rng[0][0] == 0 and rng[-1][1] == full.shape[0] and all(a[1] == b[0] for a, b in zip(rng[:-1], rng[1:]))
        # --- host-to-host driver: only the rank's gene slab is uploaded
        r0, r1 = D.pair_balanced_rows(p, 1, world, rank)
        out_host = torch.empty(D.packed_offset(r1, p, 1) - D.packed_offset(r0, p, 1), dtype=torch.float64)
        h0, h1 = D.host_sharded_cor_upper(be, xt, n, p, True, out_host, 1, 6.0)
        assert (h0, h1) == (o0, o1) and np.abs(out_host.numpy() - full[o0:o1]).max() < 1e-14
        # --- diffcor
        xb = synthetic_bulk(n + 7, p, seed=6)
        res, (d0, d1) = D.sharded_diffcor(be, xt, n, torch.from_numpy(np.ascontiguousarray(xb.T)), n + 7, p, False)
        ref = o.differential_cor(x, xb, False)
        for k in ref:
            assert np.array_equal(res[k].numpy(), ref[k][d0:d1], equal_nan=True)
        # --- TOM: slabs exchanged, k all-reduced
        for signed in (True, False):
            t, (t0, t1) = D.sharded_tom_from_exp(be, xt, n, p, False, signed, "v1", beta=2.0)
            tref = o.dense_to_upper_triangle(o.tom_from_exp(x, signed, "v1", False, 2.0), True)
            assert np.abs(t.numpy() - tref[t0:t1]).max() < 1e-12
            # split-float kinds exchange the operand planes of each rank's slab instead of the f64 matrix
            t2, (u0, u1) = D.sharded_tom_from_exp(be, xt, n, p, False, signed, "v1", beta=2.0, precision=D._lib.PREC_F16X3)
            assert (u0, u1) == (t0, t1) and np.abs(t2.numpy() - tref[t0:t1]).max() < 1e-12
        # --- sparse: cell shards, Gram all-reduce
        cells, genes = 600, 30
        ip, ix, d = synthetic_sparse_cells(cells, genes, 3, density=0.2)
        c0, c1 = cells * rank // world, cells * (rank + 1) // world
        lp = torch.from_numpy(ip[c0 : c1 + 1] - ip[c0])
        li = torch.from_numpy(ix[ip[c0] : ip[c1]])
        ld = torch.from_numpy(d[ip[c0] : ip[c1]])
        out = D.sharded_sparse_gram_cor(be, lp, li, ld, c1 - c0, genes)
        sref = o.sparse_gram_cor(ip, ix, d, cells, genes)
        ok = ~np.isnan(sref)
        assert np.abs(out.numpy()[ok] - sref[ok]).max() < 1e-12
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world2_gloo(tmp_path, lib_built):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_even_slab_and_offsets(lib_built):
    from bixverse_b200 import distributed as D
    from oracle import oracle as o

    for p in (300, 20000):
        for world in (1, 2, 8):
            slabs = [D.even_slab(p, world, r) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == p
            assert all(a[1] == b[0] and a[0] % 128 == 0 for a, b in zip(slabs[:-1], slabs[1:]))
    assert D.packed_offset(17, 300, 1) == int(o.packed_offset(17, 300, 1))
    assert D.packed_offset(17, 300, 0) == int(o.packed_offset(17, 300, 0))
