"""Synthetic inputs of the BASELINE.json configs (numpy only, no device work).

``synthetic_bulk`` restates the shape of bixverse's bulk RNA-seq generator
(defaults of ``params_synthetic_bulk_rnaseq``, R/param_wrappers.R:763-784:
negative-binomial counts, gene mean ~ Gamma(5, 10), dispersion
1/(0.2 + 0.3*mean)... planted co-expression modules on Normal(0, 0.5) latent
factors with LogNormal(0, 0.7) loadings) followed by log2-CPM as in
inst/tinytest/test_cor_modules.R:124-126, returned samples x genes (the
orientation BulkCoExp hands to rs_cor_upper_triangle).  The generator itself
lives in the un-vendored crate, so this is a restatement of its documented
parameters, not a bit-level copy; it only has to produce realistic inputs.
"""
from __future__ import annotations

import numpy as np


def synthetic_bulk(n_samples: int, n_genes: int, seed: int, n_modules: int | None = None,
                   tie_heavy: bool = False, structure_seed: int | None = None, rewire_frac: float = 0.0) -> np.ndarray:
    """`structure_seed` (optional) fixes the gene means and the planted module structure independently of the sample
    draw (`seed`), so two matrices can share their co-expression structure; `rewire_frac` of the modules (the first
    ones) then get a freshly drawn gene set instead - the rewired modules of a differential-correlation pair."""
    rng = np.random.default_rng(seed)
    srng = rng if structure_seed is None else np.random.default_rng(structure_seed)
    if n_modules is None:
        n_modules = max(3, min(20, n_genes // 400))
    mean = srng.gamma(shape=5.0, scale=10.0, size=n_genes)
    log_mu = np.log(mean)[None, :] + rng.normal(0.0, 0.1, size=(n_samples, n_genes))
    # planted modules: disjoint gene sets sharing a latent factor
    sizes = srng.integers(100, 201, size=n_modules)
    sizes = np.minimum(sizes, max(2, n_genes // (n_modules + 1)))
    perm = srng.permutation(n_genes)
    n_rewired = int(round(rewire_frac * n_modules)) if structure_seed is not None else 0
    wrng = np.random.default_rng((structure_seed or 0) + 7919)
    start = 0
    for m in range(n_modules):
        genes = perm[start : start + sizes[m]]
        start += sizes[m]
        if m < n_rewired:  # this module is made of other genes in the rewired condition
            genes = wrng.choice(n_genes, size=genes.shape[0], replace=False)
        factor = rng.normal(0.0, 0.5, size=n_samples)
        loading = srng.lognormal(0.0, 0.7, size=genes.shape[0]) * srng.choice([-1.0, 1.0], size=genes.shape[0], p=[0.2, 0.8])
        log_mu[:, genes] += factor[:, None] * loading[None, :]
    mu = np.exp(log_mu)
    disp = 1.0 / (0.2 + 0.3 * mean)  # NB size parameter r = 1/disp
    r = 1.0 / disp
    lam = rng.gamma(shape=np.broadcast_to(r, mu.shape), scale=mu / r)
    counts = rng.poisson(lam).astype(np.float64)
    lib = counts.sum(axis=1, keepdims=True)
    x = np.log2(counts / lib * 1e6 + 1.0)
    if tie_heavy:
        x = np.round(x, 1)
    return np.asfortranarray(x)


def synthetic_pair(n_a: int, n_b: int, n_genes: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    """Config 3: two conditions with the same gene means and module structure; in the second, 25 % of the
    modules are rewired (made of a different gene set), which is what the differential correlation should find."""
    return (synthetic_bulk(n_a, n_genes, seed, structure_seed=seed),
            synthetic_bulk(n_b, n_genes, seed + 1, structure_seed=seed, rewire_frac=0.25))


def synthetic_sparse_cells(n_cells: int, n_genes: int, seed: int, density: float = 0.05):
    """Config 5: cells x genes CSR, ~density non-zeros per row, log1p-normalised float32 values."""
    rng = np.random.default_rng(seed)
    nnz_per_cell = np.clip(rng.poisson(density * n_genes, size=n_cells), 1, n_genes).astype(np.int64)
    indptr = np.zeros(n_cells + 1, dtype=np.int64)
    np.cumsum(nnz_per_cell, out=indptr[1:])
    nnz = int(indptr[-1])
    # gene popularity skewed so that some genes are dense and most are rare
    w = rng.gamma(0.7, 1.0, size=n_genes)
    w /= w.sum()
    indices = np.empty(nnz, dtype=np.int32)
    blk = 4096
    for c0 in range(0, n_cells, blk):
        c1 = min(n_cells, c0 + blk)
        kmax = int(nnz_per_cell[c0:c1].max())
        # Gumbel top-k: a weighted sample without replacement per cell
        g = np.log(w)[None, :] + rng.gumbel(size=(c1 - c0, n_genes))
        top = np.argpartition(-g, kmax - 1, axis=1)[:, :kmax]
        for r in range(c1 - c0):
            k = int(nnz_per_cell[c0 + r])
            sel = np.sort(top[r, :k])
            indices[indptr[c0 + r] : indptr[c0 + r + 1]] = sel
    counts = 1 + rng.negative_binomial(2, 0.5, size=nnz)
    data = np.log1p(counts.astype(np.float32) * np.float32(1e4 / 2500.0)).astype(np.float32)
    return indptr, indices, data


def synthetic_sparse_cells_fast(n_cells: int, n_genes: int, seed: int, nnz_per_cell: int = 250):
    """Config-5-scale generator (vectorised): every cell expresses exactly `nnz_per_cell` genes, one drawn
    uniformly from each of `nnz_per_cell` equal strata of the gene axis, so the column indices of a row are
    unique and sorted by construction.  Values are log1p-normalised counts as float32."""
    rng = np.random.default_rng(seed)
    width = n_genes // nnz_per_cell
    assert width >= 1
    indptr = np.arange(n_cells + 1, dtype=np.int64) * nnz_per_cell
    base = (np.arange(nnz_per_cell, dtype=np.int32) * width)[None, :]
    indices = np.empty((n_cells, nnz_per_cell), dtype=np.int32)
    data = np.empty((n_cells, nnz_per_cell), dtype=np.float32)
    blk = 1 << 16
    for c0 in range(0, n_cells, blk):
        c1 = min(n_cells, c0 + blk)
        indices[c0:c1] = base + rng.integers(0, width, size=(c1 - c0, nnz_per_cell), dtype=np.int32)
        counts = 1 + rng.negative_binomial(2, 0.5, size=(c1 - c0, nnz_per_cell))
        data[c0:c1] = np.log1p(counts.astype(np.float32) * np.float32(4.0))
    return indptr, indices.reshape(-1), data.reshape(-1)
