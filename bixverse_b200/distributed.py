"""One-process-per-GPU drivers for the sharded path.

Bulk correlation / differential correlation shard the packed triangle into
contiguous block-row ranges balanced by pair count: no data-path collective.
TOM has one real exchange step: every rank builds the affinity columns of its
gene slab, the slabs are exchanged over NCCL (NVLink/NVSwitch) so that each
rank holds the full A operand, and the connectivity vector k is all-reduced.
The sparse single-cell Gram is sharded by cells and the partial Gram matrices
are all-reduced.

The compute calls go through a *backend* object; the product backend is
``DeviceBackend`` (device pointers into libbixcor, tensors in HBM).  The
world_size-2 gloo tests inject a CPU stand-in to exercise the partitioning and
collective plumbing without a GPU; that stand-in lives in ``tests/``, never here.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from ._lib import PREC_F64, TOM_V1, TOM_V2, check

TILE = 128


def packed_offset(i: int, p: int, shift: int) -> int:
    return i * (2 * p - i - 1) // 2 if shift else i * (2 * p - i + 1) // 2


def pair_balanced_rows(p: int, shift: int, world: int, rank: int) -> tuple[int, int]:
    """Same partition as bixcor_shard_rows (include/bixcor.h)."""
    import ctypes as C

    r0, r1 = C.c_int64(), C.c_int64()
    check(_lib.load().bixcor_shard_rows(p, int(shift), world, rank, C.byref(r0), C.byref(r1)))
    return r0.value, r1.value


def even_slab(p: int, world: int, rank: int) -> tuple[int, int]:
    """Gene slab [c0,c1) for full-width strips: equal tile counts, 128-aligned."""
    nb = (p + TILE - 1) // TILE
    b0 = nb * rank // world
    b1 = nb * (rank + 1) // world
    return min(p, b0 * TILE), min(p, b1 * TILE)


class DeviceBackend:
    """libbixcor on the current CUDA device; tensors stay resident in HBM."""

    def __init__(self, device: torch.device):
        self.device = device
        self.lib = _lib.load()

    def _sync_in(self):
        torch.cuda.current_stream(self.device).synchronize()

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float64, device=self.device)

    def cor_upper_rows(self, x, n, p, spearman, shift, beta, precision, r0, r1):
        ln = packed_offset(r1, p, shift) - packed_offset(r0, p, shift)
        out = self.empty(max(ln, 1))[:ln]
        self._sync_in()
        check(self.lib.bixcor_dev_cor_upper_rows(x.data_ptr(), n, p, int(spearman), int(shift), float(beta), precision,
                                                 r0, r1, out.data_ptr()))
        return out

    def operand_row_bytes(self, n, precision):
        return int(self.lib.bixcor_dev_operand_row_bytes(n, precision))

    def empty_bytes(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def build_operand(self, x, n, p, spearman, precision, g0, g1, operand, inv_norm):
        self._sync_in()
        check(self.lib.bixcor_dev_build_operand(x.data_ptr(), n, p, int(spearman), precision, g0, g1,
                                                operand.data_ptr(), inv_norm.data_ptr()))

    def operand_inv_norm(self, operand, n, p, precision, inv_norm):
        """Per-gene norms from the (all-gathered) operand records: saves the second collective."""
        self._sync_in()
        check(self.lib.bixcor_dev_operand_inv_norm(operand.data_ptr(), n, p, precision, inv_norm.data_ptr()))

    def cor_upper_rows_op(self, operand, inv_norm, n, p, precision, shift, beta, r0, r1):
        ln = packed_offset(r1, p, shift) - packed_offset(r0, p, shift)
        out = self.empty(max(ln, 1))[:ln]
        self._sync_in()
        check(self.lib.bixcor_dev_cor_upper_rows_op(operand.data_ptr(), inv_norm.data_ptr(), n, p, precision, int(shift),
                                                    float(beta), r0, r1, out.data_ptr()))
        return out

    def to_device(self, host_tensor):
        return host_tensor.to(self.device, non_blocking=True)

    def build_operand_slab(self, x_slab, n, spearman, precision, g0, g1, operand, inv_norm):
        """x_slab holds only genes [g0,g1) (rows of the gene-major matrix); their records land at [g0,g1)."""
        rb = self.operand_row_bytes(n, precision)
        self._sync_in()
        check(self.lib.bixcor_dev_build_operand(x_slab.data_ptr(), n, g1 - g0, int(spearman), precision, 0, g1 - g0,
                                                operand.data_ptr() + g0 * rb, inv_norm.data_ptr() + 8 * g0))

    def cor_upper_rows_op_to_host(self, operand, inv_norm, n, p, precision, shift, beta, r0, r1, out_host):
        """Gram rows [r0,r1) straight into host memory; finished bands are copied while later ones compute."""
        self._sync_in()
        check(self.lib.bixcor_dev_cor_upper_rows_op_to_host(operand.data_ptr(), inv_norm.data_ptr(), n, p, precision,
                                                            int(shift), float(beta), r0, r1, out_host.data_ptr()))

    def diffcor_rows(self, xa, na, xb, nb, p, spearman, precision, r0, r1):
        ln = packed_offset(r1, p, 1) - packed_offset(r0, p, 1)
        outs = [self.empty(max(ln, 1))[:ln] for _ in range(4)]
        self._sync_in()
        check(self.lib.bixcor_dev_diffcor_rows(xa.data_ptr(), na, xb.data_ptr(), nb, p, int(spearman), precision, r0, r1,
                                               *[o.data_ptr() for o in outs]))
        return dict(zip(("r_a", "r_b", "z_score", "p_val"), outs))

    def affinity_cols(self, x, n, p, spearman, signed, beta, precision, c0, c1, a_full):
        self._sync_in()
        check(self.lib.bixcor_dev_affinity_cols(x.data_ptr(), n, p, int(spearman), int(signed), float(beta), precision,
                                                c0, c1, a_full.data_ptr()))

    def tom_connectivity(self, a_full, p, signed, c0, c1, k):
        self._sync_in()
        check(self.lib.bixcor_dev_tom_connectivity(a_full.data_ptr(), p, int(signed), c0, c1, k.data_ptr()))

    def tom_plane_row_bytes(self, p, precision):
        return int(self.lib.bixcor_dev_tom_plane_row_bytes(p, precision))

    def tom_split_planes(self, a_full, p, precision, c0, c1, planes):
        """hi / lo operand planes of the affinity rows [c0,c1) into the full plane buffer (split-float kinds)."""
        self._sync_in()
        check(self.lib.bixcor_dev_tom_split_planes(a_full.data_ptr(), p, precision, c0, c1, planes.data_ptr()))

    def tom_rows_planes(self, planes, k, p, signed, version, precision, r0, r1):
        ln = packed_offset(r1, p, 1) - packed_offset(r0, p, 1)
        out = self.empty(max(ln, 1))[:ln]
        self._sync_in()
        check(self.lib.bixcor_dev_tom_rows_planes(planes.data_ptr(), k.data_ptr(), p, int(signed), version, precision, 1,
                                                  r0, r1, out.data_ptr()))
        return out

    def tom_rows(self, a_full, k, p, signed, version, precision, r0, r1):
        ln = packed_offset(r1, p, 1) - packed_offset(r0, p, 1)
        out = self.empty(max(ln, 1))[:ln]
        self._sync_in()
        check(self.lib.bixcor_dev_tom_rows(a_full.data_ptr(), k.data_ptr(), p, int(signed), version, precision, 1,
                                           r0, r1, out.data_ptr()))
        return out

    def sparse_accumulate(self, indptr, indices, data, cells, genes, precision, gram, colsum):
        self._sync_in()
        check(self.lib.bixcor_dev_sparse_gram_accumulate(indptr.data_ptr(), indices.data_ptr(), data.data_ptr(), cells,
                                                         genes, precision, gram.data_ptr(), colsum.data_ptr()))

    def sparse_finalise(self, gram, colsum, total_cells, genes):
        out = self.empty(max(genes * (genes - 1) // 2, 1))[: genes * (genes - 1) // 2]
        self._sync_in()
        check(self.lib.bixcor_dev_sparse_gram_finalise(gram.data_ptr(), colsum.data_ptr(), total_cells, genes,
                                                       out.data_ptr()))
        return out


def _world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gene_slab(p: int, world: int) -> int:
    """Equal 128-aligned gene slab width of the operand exchange (the last slab may be ragged)."""
    return ((p + world - 1) // world + TILE - 1) // TILE * TILE


def exchange_operand(backend, x, n, p, spearman, precision, group=None):
    """Every rank rank-transforms / standardises only its gene slab; the gene-major operand records
    (and the per-gene norms) are all-gathered, so the column kernel scales with the number of GPUs."""
    rank, world = _world(group)
    slab = gene_slab(p, world)
    rb = backend.operand_row_bytes(n, precision)
    operand = backend.empty_bytes(world * slab * rb)
    inv = backend.zeros(world * slab)
    g0, g1 = min(p, rank * slab), min(p, (rank + 1) * slab)
    backend.build_operand(x, n, p, spearman, precision, g0, g1, operand, inv)
    if world > 1:
        dist.all_gather_into_tensor(operand, operand[rank * slab * rb : (rank + 1) * slab * rb], group=group)
        backend.operand_inv_norm(operand, n, p, precision, inv)  # instead of a second all-gather
    return operand, inv


def xchg_setup(backend, n, p, precision, group=None):
    """Collective: create this rank's exchange block, all-gather the cudaIpc handles and map the peers' blocks
    (include/bixcor.h, bixcor_xchg_*).  Afterwards bixcor_dev_cor_upper_rows_xchg needs no collective library: the column
    kernel stores its records into every peer's operand buffer over NVLink and a flag barrier in peer memory replaces the
    all-gather.  One box, world <= 8."""
    import ctypes as C

    rank, world = _world(group)
    lib = backend.lib
    check(lib.bixcor_xchg_release_peers())  # (a previous exchange: nobody maps a block any more when its owner frees it)
    if world > 1:
        dist.barrier(group=group)
    def agree(rc):  # a failure on one rank fails the set-up on every rank (nobody waits in a collective for a rank that raised)
        if world > 1:
            worst = torch.tensor([rc], dtype=torch.int32, device=backend.device)
            dist.all_reduce(worst, op=dist.ReduceOp.MIN, group=group)
            if rc == _lib.OK and int(worst.item()) != _lib.OK:
                raise _lib.BixcorError(int(worst.item()), "operand exchange set-up failed on another rank")
        check(rc)

    h = (C.c_ubyte * _lib.XCHG_HANDLE_BYTES)()
    agree(lib.bixcor_xchg_create(n, p, precision, world, rank, h))
    mine = torch.tensor(list(bytes(h)), dtype=torch.uint8, device=backend.device)
    every = torch.empty(world * _lib.XCHG_HANDLE_BYTES, dtype=torch.uint8, device=backend.device)
    if world > 1:
        dist.all_gather_into_tensor(every, mine, group=group)
    else:
        every.copy_(mine)
    blob = bytes(every.cpu().numpy().tobytes())
    agree(lib.bixcor_xchg_open(blob))  # (also the barrier: every rank has mapped every block before anybody stores into one)


def xchg_teardown(backend, group=None):
    """Collective: unmap the peers' blocks, synchronise, free this rank's block."""
    _rank, world = _world(group)
    check(backend.lib.bixcor_xchg_release_peers())
    if world > 1:
        dist.barrier(group=group)
    check(backend.lib.bixcor_xchg_destroy())


def sharded_cor_upper_xchg(backend, x, n, p, spearman, shift=1, beta=1.0, group=None):
    """sharded_cor_upper through the fused operand exchange (xchg_setup first): one library call per rank and step."""
    rank, world = _world(group)
    r0, r1 = pair_balanced_rows(p, shift, world, rank)
    ln = packed_offset(r1, p, shift) - packed_offset(r0, p, shift)
    out = backend.empty(max(ln, 1))[:ln]
    backend._sync_in()
    check(backend.lib.bixcor_dev_cor_upper_rows_xchg(x.data_ptr(), n, p, int(spearman), int(shift), float(beta), r0, r1, out.data_ptr()))
    return out, (packed_offset(r0, p, shift), packed_offset(r1, p, shift))


def host_sharded_cor_upper_xchg(backend, x_host, n, p, spearman, out_host, shift=1, beta=1.0, group=None):
    """host_sharded_cor_upper through the fused operand exchange: H2D of this rank's gene slab, rank kernel with peer stores,
    flag barrier, Gram rows with the band-pipelined D2H into out_host."""
    rank, world = _world(group)
    r0, r1 = pair_balanced_rows(p, shift, world, rank)
    slab = gene_slab(p, world)
    g0, g1 = min(p, rank * slab), min(p, (rank + 1) * slab)
    x_slab = backend.to_device(x_host[g0:g1] if g1 > g0 else x_host[0:1])  # (an empty slab still takes part in the barrier; its pointer is not read)
    backend._sync_in()
    check(backend.lib.bixcor_dev_cor_upper_rows_xchg_to_host(x_slab.data_ptr(), n, p, int(spearman), int(shift), float(beta), r0, r1,
                                                             out_host.data_ptr()))
    return packed_offset(r0, p, shift), packed_offset(r1, p, shift)


def sharded_cor_upper(backend, x, n, p, spearman, shift=1, beta=1.0, precision=PREC_F64, group=None):
    """Rank-local slice of the packed correlation vector and its [o0,o1) range.

    world == 1: one fused call.  world > 1: gene-sharded column kernel + operand all-gather over
    NCCL, then the pair-balanced block rows of the Gram (no further collective)."""
    rank, world = _world(group)
    r0, r1 = pair_balanced_rows(p, shift, world, rank)
    if world == 1:
        out = backend.cor_upper_rows(x, n, p, spearman, shift, beta, precision, r0, r1)
    else:
        operand, inv = exchange_operand(backend, x, n, p, spearman, precision, group)
        out = backend.cor_upper_rows_op(operand, inv, n, p, precision, shift, beta, r0, r1)
    return out, (packed_offset(r0, p, shift), packed_offset(r1, p, shift))


def host_sharded_cor_upper(backend, x_host, n, p, spearman, out_host, shift=1, beta=1.0, precision=PREC_F64,
                           group=None, workspace=None):
    """End-to-end packed correlation of one rank, host memory to host memory.

    x_host: the p x n gene-major host tensor (== R's column-major n x p matrix), ideally pinned; only
    this rank's gene slab crosses PCIe.  The slab is rank-transformed / standardised on the device, the
    operand records are all-gathered over NCCL (NVLink), and the rank's pair-balanced block rows are
    written into out_host (its packed slice, length o1 - o0) with the device-to-host copy of finished
    bands overlapped with the Gram of later bands.  `workspace` (dict) caches the device buffers."""
    rank, world = _world(group)
    r0, r1 = pair_balanced_rows(p, shift, world, rank)
    slab = gene_slab(p, world)
    rb = backend.operand_row_bytes(n, precision)
    ws = workspace if workspace is not None else {}
    key = (n, p, precision, world)
    if ws.get("key") != key:
        ws.clear()
        ws.update(key=key, operand=backend.empty_bytes(world * slab * rb), inv=backend.zeros(world * slab))
    operand, inv = ws["operand"], ws["inv"]
    g0, g1 = min(p, rank * slab), min(p, (rank + 1) * slab)
    if g1 > g0:
        x_slab = backend.to_device(x_host[g0:g1])
        backend.build_operand_slab(x_slab, n, spearman, precision, g0, g1, operand, inv)
    if world > 1:
        dist.all_gather_into_tensor(operand, operand[rank * slab * rb : (rank + 1) * slab * rb], group=group)
        backend.operand_inv_norm(operand, n, p, precision, inv)  # instead of a second all-gather
    backend.cor_upper_rows_op_to_host(operand, inv, n, p, precision, shift, beta, r0, r1, out_host)
    return packed_offset(r0, p, shift), packed_offset(r1, p, shift)


def sharded_diffcor(backend, xa, na, xb, nb, p, spearman, precision=PREC_F64, group=None):
    rank, world = _world(group)
    r0, r1 = pair_balanced_rows(p, 1, world, rank)
    res = backend.diffcor_rows(xa, na, xb, nb, p, spearman, precision, r0, r1)
    return res, (packed_offset(r0, p, 1), packed_offset(r1, p, 1))


def sharded_tom_from_exp(backend, x, n, p, spearman, signed, version, beta=1.0, precision=PREC_F64, group=None):
    """calculate_tom_from_exp, packed (shift = 1) output slice of this rank.

    phase 1  affinity columns of this rank's gene slab (full-width strip)
    phase 2  connectivity of the slab
    exchange ONE collective for the operand + all-reduce(sum) of k.  Split-float kinds (3xTF32 / 3xF16): every rank splits
             its own slab into the hi / lo operand planes and the PLANES are all-gathered (f16: half the bytes of the f64
             matrix; no rank converts the whole matrix) - phase 3 rebuilds a_ij from hi + lo.  FP64 / Ozaki: the f64 slabs
             are all-gathered.  Slabs are equal (128-aligned, the last one ragged) so that a single all-gather serves.
    phase 3  TOM block rows of the pair-balanced range, upper tiles only
    """
    rank, world = _world(group)
    ver = TOM_V1 if version in ("v1", TOM_V1) else TOM_V2
    cor_prec = precision if precision in (PREC_F64, _lib.PREC_TF32X3, _lib.PREC_F16X3, _lib.PREC_OZAKI, _lib.PREC_OZAKI4) else PREC_F64
    r0, r1 = pair_balanced_rows(p, 1, world, rank)
    rng = (packed_offset(r0, p, 1), packed_offset(r1, p, 1))
    k = backend.zeros(p)
    if world == 1:
        a_full = backend.empty(p, p)  # row i of this tensor == column i of the column-major matrix
        backend.affinity_cols(x, n, p, spearman, signed, beta, cor_prec, 0, p, a_full)
        backend.tom_connectivity(a_full, p, signed, 0, p, k)
        return backend.tom_rows(a_full, k, p, signed, ver, precision, r0, r1), rng
    slab = gene_slab(p, world)
    c0, c1 = min(p, rank * slab), min(p, (rank + 1) * slab)
    a_full = backend.empty(world * slab, p)
    backend.affinity_cols(x, n, p, spearman, signed, beta, cor_prec, c0, c1, a_full)
    backend.tom_connectivity(a_full, p, signed, c0, c1, k)
    if precision in (_lib.PREC_TF32X3, _lib.PREC_F16X3):
        rb = backend.tom_plane_row_bytes(p, precision)
        planes = backend.empty_bytes(world * slab * rb)
        backend.tom_split_planes(a_full, p, precision, c0, c1, planes)
        dist.all_gather_into_tensor(planes, planes[rank * slab * rb : (rank + 1) * slab * rb], group=group)
        dist.all_reduce(k, op=dist.ReduceOp.SUM, group=group)
        return backend.tom_rows_planes(planes, k, p, signed, ver, precision, r0, r1), rng
    dist.all_gather_into_tensor(a_full, a_full[rank * slab : (rank + 1) * slab], group=group)
    dist.all_reduce(k, op=dist.ReduceOp.SUM, group=group)
    return backend.tom_rows(a_full, k, p, signed, ver, precision, r0, r1), rng


def sharded_sparse_gram_cor(backend, indptr, indices, data, local_cells, genes, precision=PREC_F64, group=None):
    """Cell-sharded all-pairs Pearson: partial Gram + column sums, all-reduced, finalised on every rank."""
    rank, world = _world(group)
    gram = backend.zeros(genes, genes)
    colsum = backend.zeros(genes)
    backend.sparse_accumulate(indptr, indices, data, local_cells, genes, precision, gram, colsum)
    total = torch.tensor([float(local_cells)], dtype=torch.float64, device=gram.device)
    if world > 1:
        dist.all_reduce(gram, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(colsum, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return backend.sparse_finalise(gram, colsum, int(total.item()), genes)
