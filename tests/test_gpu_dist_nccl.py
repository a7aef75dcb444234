"""world_size-2 NCCL run of the five multi-process drivers on real GPUs (skipped with fewer than 2 devices):
the same checks as tests/test_dist_gloo.py, but through libbixcor's device entry points and NCCL collectives."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from bixverse_b200 import _lib, api
        from bixverse_b200 import distributed as D
        from bixverse_b200.synthetic import synthetic_bulk, synthetic_sparse_cells
        from oracle import oracle as o

        api.init(1, rank)
        be = D.DeviceBackend(dev)
        n, p = 120, 1300
        x = synthetic_bulk(n, p, seed=5, tie_heavy=True)
        xt = torch.from_numpy(np.ascontiguousarray(x.T))
        xd = xt.to(dev)
        full = o.soft_threshold(o.cor_upper_triangle(x, True, True), 6.0)
        for prec, tol in ((_lib.PREC_F64, 1e-10), (_lib.PREC_I8EXACT, 1e-12), (_lib.PREC_TF32X3, 1e-5), (_lib.PREC_F16X3, 1e-5)):
            sl, (o0, o1) = D.sharded_cor_upper(be, xd, n, p, True, 1, 6.0, prec)
            assert np.abs(sl.cpu().numpy() - full[o0:o1]).max() < tol
            out_host = torch.empty(o1 - o0, dtype=torch.float64).pin_memory()
            assert D.host_sharded_cor_upper(be, xt.pin_memory(), n, p, True, out_host, 1, 6.0, prec) == (o0, o1)
            assert np.abs(out_host.numpy() - full[o0:o1]).max() < tol
        # fused operand exchange over peer memory (bixcor_xchg_*): the rank kernel stores its records into every rank's buffer,
        # a flag barrier replaces the all-gather; several steps on changing data exercise the two buffers and the epochs
        x2 = synthetic_bulk(n, p, seed=77)
        x2[:, 3] = 1.5       # zero variance -> NaN row / column
        x2[4, 900] = np.nan  # NaN column (a gene of the second slab)
        x2d = torch.from_numpy(np.ascontiguousarray(x2.T)).to(dev)
        full2 = o.soft_threshold(o.cor_upper_triangle(x2, True, True), 6.0)
        ok2 = ~np.isnan(full2)
        for prec, tol in ((_lib.PREC_I8EXACT, 1e-12), (_lib.PREC_AUTO, 1e-12), (_lib.PREC_F64, 1e-10), (_lib.PREC_TF32X3, 1e-5), (_lib.PREC_F16X3, 1e-5)):
            D.xchg_setup(be, n, p, prec)
            for step in range(5):
                if step % 2 == 0:
                    sl, (o0, o1) = D.sharded_cor_upper_xchg(be, xd, n, p, True, 1, 6.0)
                    assert np.abs(sl.cpu().numpy() - full[o0:o1]).max() < tol, (prec, step)
                else:
                    sl, (o0, o1) = D.sharded_cor_upper_xchg(be, x2d, n, p, True, 1, 6.0)
                    got, ref2 = sl.cpu().numpy(), full2[o0:o1]
                    assert np.isnan(got[~ok2[o0:o1]]).all() and np.abs(got[ok2[o0:o1]] - ref2[ok2[o0:o1]]).max() < tol, (prec, step)
            out_host = torch.empty(o1 - o0, dtype=torch.float64)  # pageable
            assert D.host_sharded_cor_upper_xchg(be, xt, n, p, True, out_host, 1, 6.0) == (o0, o1)
            assert np.abs(out_host.numpy() - full[o0:o1]).max() < tol
        # more than 1024 samples: the CTA rank kernel has no fused peer stores, its slab goes out through the peer-copy kernel
        n3, p3 = 1100, 300
        x3 = synthetic_bulk(n3, p3, seed=8, tie_heavy=True)
        x3d = torch.from_numpy(np.ascontiguousarray(x3.T)).to(dev)
        full3 = o.cor_upper_triangle(x3, True, True)
        D.xchg_setup(be, n3, p3, _lib.PREC_AUTO)
        for step in range(2):
            sl, (o0, o1) = D.sharded_cor_upper_xchg(be, x3d, n3, p3, True, 1, 1.0)
            assert np.abs(sl.cpu().numpy() - full3[o0:o1]).max() < 1e-12
        D.xchg_teardown(be)
        xb = synthetic_bulk(n + 9, p, seed=6)
        xbd = torch.from_numpy(np.ascontiguousarray(xb.T)).to(dev)
        res, (d0, d1) = D.sharded_diffcor(be, xd, n, xbd, n + 9, p, False)
        ref = o.differential_cor(x, xb, False)
        assert np.abs(res["r_a"].cpu().numpy() - ref["r_a"][d0:d1]).max() < 1e-10
        assert np.abs(res["z_score"].cpu().numpy() - ref["z_score"][d0:d1]).max() < 1e-9
        for prec, tol in ((_lib.PREC_F64, 1e-10), (_lib.PREC_OZAKI, 1e-10), (_lib.PREC_TF32X3, 1e-5), (_lib.PREC_F16X3, 1e-5)):
            for signed in (True, False):
                t, (t0, t1) = D.sharded_tom_from_exp(be, xd, n, p, False, signed, "v1", beta=2.0, precision=prec)
                tref = o.dense_to_upper_triangle(o.tom_from_exp(x, signed, "v1", False, 2.0), True)
                assert np.abs(t.cpu().numpy() - tref[t0:t1]).max() < tol
        cells, genes = 20000, 200
        ip, ix, d = synthetic_sparse_cells(cells, genes, 3, density=0.1)
        c0, c1 = cells * rank // world, cells * (rank + 1) // world
        lp = torch.from_numpy(ip[c0 : c1 + 1] - ip[c0]).to(dev)
        li = torch.from_numpy(ix[ip[c0] : ip[c1]]).to(dev)
        ld = torch.from_numpy(d[ip[c0] : ip[c1]]).to(dev)
        sref = o.sparse_gram_cor(ip, ix, d, cells, genes)
        ok = ~np.isnan(sref)
        for prec, tol in ((_lib.PREC_F64, 1e-9), (_lib.PREC_TF32X3, 1e-5), (_lib.PREC_F16X3, 1e-5)):
            out = D.sharded_sparse_gram_cor(be, lp, li, ld, c1 - c0, genes, prec)
            assert np.abs(out.cpu().numpy()[ok] - sref[ok]).max() < tol
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world2_nccl(tmp_path, lib_built):
    import torch
    import torch.multiprocessing as mp

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, 29700 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
