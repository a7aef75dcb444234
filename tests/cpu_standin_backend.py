"""CPU stand-in for bixverse_b200.distributed.DeviceBackend built on the oracle.
Lives in tests/ on purpose: it exists only to exercise the multi-rank plumbing
(partitioning + collectives) under gloo without a GPU."""
import numpy as np
import torch

from oracle import oracle as o


class OracleBackend:
    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float64)

    @staticmethod
    def _slice(packed, p, shift, r0, r1):
        return packed[int(o.packed_offset(r0, p, shift)) : int(o.packed_offset(r1, p, shift))]

    def cor_upper_rows(self, x, n, p, spearman, shift, beta, precision, r0, r1):
        full = o.soft_threshold(o.cor_upper_triangle(x.numpy().T, spearman, bool(shift)), beta)
        return torch.from_numpy(np.ascontiguousarray(self._slice(full, p, shift, r0, r1)))

    # operand exchange: the stand-in operand record of a gene is its standardised column (n doubles)
    def operand_row_bytes(self, n, precision):
        return 8 * n

    def empty_bytes(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    def build_operand(self, x, n, p, spearman, precision, g0, g1, operand, inv_norm):
        xs = x.numpy()[g0:g1].T  # n x (g1-g0)
        z = o._standardise(o.rank_matrix_col(xs) if spearman else xs)
        view = operand.numpy().view(np.float64).reshape(-1, n)
        view[g0:g1] = z.T
        inv_norm[g0:g1] = 1.0

    def operand_inv_norm(self, operand, n, p, precision, inv_norm):
        z = operand.numpy().view(np.float64).reshape(-1, n)[:p]
        inv_norm[:p] = torch.from_numpy((np.abs(z).sum(axis=1) > 0).astype(np.float64))  # 1 where a record arrived

    def cor_upper_rows_op(self, operand, inv_norm, n, p, precision, shift, beta, r0, r1):
        z = operand.numpy().view(np.float64).reshape(-1, n)[:p].T
        assert float(inv_norm[:p].sum()) == p  # every slab arrived
        full = o.soft_threshold(o.dense_to_upper_triangle(z.T @ z, bool(shift)), beta)
        return torch.from_numpy(np.ascontiguousarray(self._slice(full, p, shift, r0, r1)))

    def to_device(self, host_tensor):
        return host_tensor.clone()

    def build_operand_slab(self, x_slab, n, spearman, precision, g0, g1, operand, inv_norm):
        xs = x_slab.numpy().T  # n x (g1-g0)
        z = o._standardise(o.rank_matrix_col(xs) if spearman else xs)
        operand.numpy().view(np.float64).reshape(-1, n)[g0:g1] = z.T
        inv_norm[g0:g1] = 1.0

    def cor_upper_rows_op_to_host(self, operand, inv_norm, n, p, precision, shift, beta, r0, r1, out_host):
        out_host.copy_(self.cor_upper_rows_op(operand, inv_norm, n, p, precision, shift, beta, r0, r1))

    def diffcor_rows(self, xa, na, xb, nb, p, spearman, precision, r0, r1):
        res = o.differential_cor(xa.numpy().T, xb.numpy().T, spearman)
        return {k: torch.from_numpy(np.ascontiguousarray(self._slice(v, p, 1, r0, r1))) for k, v in res.items()}

    def affinity_cols(self, x, n, p, spearman, signed, beta, precision, c0, c1, a_full):
        r = o.column_pairwise_cor(x.numpy().T, spearman)
        a = o.soft_threshold(r, beta, signed=bool(signed)) if beta != 1.0 else r
        np.fill_diagonal(a, 1.0)
        a_full[c0:c1] = torch.from_numpy(a[c0:c1])

    def tom_connectivity(self, a_full, p, signed, c0, c1, k):
        k[c0:c1] = (a_full[c0:c1].abs() if signed else a_full[c0:c1]).sum(dim=1)

    def tom_rows(self, a_full, k, p, signed, version, precision, r0, r1):
        a = a_full.numpy()[:p]  # multi-rank: equal padded slabs, the rows beyond p are never written
        assert np.allclose(k.numpy(), (np.abs(a) if signed else a).sum(axis=1))  # the all-reduced k is the full connectivity
        t = o.calc_tom(a, bool(signed), "v1" if version == 1 else "v2")
        return torch.from_numpy(np.ascontiguousarray(self._slice(o.dense_to_upper_triangle(t, True), p, 1, r0, r1)))

    # split-float kinds: the stand-in's "planes" are the f64 rows themselves (8 bytes per entry), which exercises the
    # slab split / plane all-gather / k all-reduce plumbing of sharded_tom_from_exp
    def tom_plane_row_bytes(self, p, precision):
        return 8 * p

    def tom_split_planes(self, a_full, p, precision, c0, c1, planes):
        planes.numpy().view(np.float64).reshape(-1, p)[c0:c1] = a_full.numpy()[c0:c1]

    def tom_rows_planes(self, planes, k, p, signed, version, precision, r0, r1):
        a = torch.from_numpy(planes.numpy().view(np.float64).reshape(-1, p)[:p].copy())
        return self.tom_rows(a, k, p, signed, version, precision, r0, r1)

    def sparse_accumulate(self, indptr, indices, data, cells, genes, precision, gram, colsum):
        d = o._csr_to_dense(indptr.numpy(), indices.numpy(), data.numpy(), cells, genes)
        gram += torch.from_numpy(d.T @ d)
        colsum += torch.from_numpy(d.sum(axis=0))

    def sparse_finalise(self, gram, colsum, total_cells, genes):
        g, s = gram.numpy(), colsum.numpy()
        cov = g - np.outer(s, s) / total_cells
        var = np.diag(cov).copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            r = cov / np.sqrt(np.outer(var, var))
        return torch.from_numpy(o.dense_to_upper_triangle(r, True))
