/*
 * bixcor.h - C ABI of libbixcor, the B200-native drop-in for the co-expression
 * correlation hot path of GregorLueg/bixverse.
 *
 * The library replaces the four calls that bixverse's extendr binding layer makes
 * into the CPU crate `bixverse-rs` 0.4.4 on this path; everything above the
 * `#[extendr] fn` bodies (R wrappers, result classes) is unchanged.  Reference
 * interfaces replaced (paths relative to the bixverse repository root):
 *
 *   column_pairwise_cor(&MatRef<f64>, bool) -> Mat<f64>
 *       src/rust/src/base/r_cors_similarity.rs:70-77      -> bixcor_cor_dense
 *   column_pairwise_cor + upper_triangle_indices + scalar gather
 *       src/rust/src/base/r_cors_similarity.rs:251-268    -> bixcor_cor_upper
 *   calculate_diff_correlation::<f64>(&cor_a,&cor_b,n_a,n_b,spearman)
 *       src/rust/src/methods/r_diffcor.rs:45-74           -> bixcor_diffcor
 *   calc_tom(MatRef<f64>, signed, TomType) (+ parse_tom_types)
 *       src/rust/src/methods/r_coremo.rs:46-59            -> bixcor_tom
 *   rs_cor -> abs -> rs_tom chain of calculate_tom_from_exp
 *       R/functions_stats.R:103-121                       -> bixcor_tom_from_exp
 *   rank_matrix_col (average ties)
 *       src/rust/src/enrichment/r_singscore.rs:89-95      -> bixcor_rank_columns
 *   packed triangle index order
 *       src/rust/src/base/r_helpers.rs:75-134             -> bixcor_packed_offset,
 *                                                            bixcor_upper_to_dense,
 *                                                            bixcor_dense_to_upper
 *   rs_pairwise_gene_cors_mc semantic anchor (pair list over sparse columns)
 *       src/rust/src/meta_cell/r_mc_processing.rs:258-283 -> bixcor_sparse_gram_cor
 *
 * Conventions
 *   - All matrices are column-major `double` exactly as R stores them: an n x p
 *     expression matrix has the n samples of gene j contiguous at x + j*n.
 *   - Pointers are HOST pointers unless the function name contains `_dev_`; host
 *     buffers may be pageable or pinned (pinned buffers are DMA'd directly).
 *   - Outputs are caller-allocated; the library never keeps a caller pointer
 *     past return and never frees caller memory.
 *   - Every function returns BIXCOR_OK (0) or a negative error code; the message
 *     is available from bixcor_last_error().  No C++ exception crosses the ABI.
 *   - There is NO CPU fallback: without a usable sm_100 GPU every compute entry
 *     point fails with BIXCOR_ERR_NO_GPU.
 *   - Calls are blocking and not re-entrant (R calls on one thread).
 *   - Packed upper-triangle order (0-based, i < j, shift = 1):
 *       idx(i,j) = i*(2p - i - 1)/2 + (j - i - 1);  shift = 0 keeps the diagonal:
 *       idx(i,j) = i*(2p - i + 1)/2 + (j - i).
 */
#ifndef BIXCOR_H
#define BIXCOR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BIXCOR_OK 0
#define BIXCOR_ERR_NO_GPU (-1)
#define BIXCOR_ERR_INVALID (-2)
#define BIXCOR_ERR_CUDA (-3)
#define BIXCOR_ERR_OOM (-4)
#define BIXCOR_ERR_UNSUPPORTED (-5)
#define BIXCOR_ERR_INTERNAL (-6) /* a C++ exception other than std::bad_alloc was stopped at the ABI */

/* `precision` argument */
#define BIXCOR_PREC_AUTO (-1) /* what the drop-in patch passes: the fastest path that provably meets the
                               * reference's FP64 contract (<= 1e-10) for this call: the exact-integer tcgen05 Gram for Spearman
                               * with n <= 1860 samples (exact by construction), FP64 DMMA otherwise.  The environment variable
                               * BIXCOR_PRECISION = auto | f64 | tf32 | f16 | i8 | ozaki | ozaki4 replaces the AUTO rule
                               * process-wide (an R user opting into the <= 1e-5 fast paths without an API change); a value
                               * that is not defined for the call (i8 for Pearson or n > 1860) falls back to the AUTO rule. */
#define BIXCOR_PREC_F64 0    /* FP64 tensor-core (DMMA) Gram: exact path, <=1e-10 */
#define BIXCOR_PREC_TF32X3 1 /* tcgen05 3xTF32 split, FP32 accumulate in TMEM: fast path, <=1e-5 */
#define BIXCOR_PREC_I8EXACT 2 /* Spearman only: exact integer Gram of centred ranks on tcgen05 kind::i8 */
#define BIXCOR_PREC_OZAKI 3   /* FP64-class Gram on the int8 tensor cores: 6 balanced base-128 digit slices per operand
                               * (Ozaki splitting), 24 exact kind::i8 MMAs per product, FP64 recombination; <=1e-10 */
#define BIXCOR_PREC_OZAKI4 4  /* same with the four leading slices (three passes, 12 MMAs per product): the dropped digit
                               * products grow with sqrt(K): <=1e-7 at K = 1000, 5e-6 on the 20k TOM (the 3xTF32 class,
                               * 10 % faster there and free of the chunked FP32 promotion) */
#define BIXCOR_PREC_F16X3 5   /* tcgen05 kind::f16 on binary16 hi / lo planes (3 MMAs per product like 3xTF32, 22 significant
                               * bits, half the operand bytes, twice the MMA rate): fast path, <=1e-5; dense Grams and TOM */

/* `version` argument of the TOM entry points ("v1"/"v2" of parse_tom_types) */
#define BIXCOR_TOM_V1 1
#define BIXCOR_TOM_V2 2

/* ---- lifecycle --------------------------------------------------------- */
/* Bind the calling process to `n_gpus` visible devices starting at `first_device`
 * (bixcor_init(n) == bixcor_init_devices(0, n)).  One process per GPU (torchrun)
 * passes first_device = LOCAL_RANK, n_gpus = 1.  Idempotent; lazily called with
 * (0, 1) by the compute entry points. */
int bixcor_init(int n_gpus);
int bixcor_init_devices(int first_device, int n_gpus);
/* With n_gpus > 1 the library opens libnccl.so.2 at run time and creates one communicator per device
 * (ncclCommInitAll): bixcor_cor_upper / bixcor_diffcor / bixcor_cor_dense / bixcor_coremo_stability shard their rows /
 * slabs / samples over the devices; bixcor_tom* exchange affinity slabs by grouped ncclBroadcast and all-reduce the
 * connectivity vector k; bixcor_sparse_gram_cor shards the cells and all-reduces the partial Gram + column sums.
 * Returns NCCL's version code, or 0 when the collectives are not active (one device, BIXCOR_NCCL=0 or no libnccl:
 * slabs then move by peer copies and k through the host). */
int bixcor_nccl_version(void);
void bixcor_shutdown(void);
const char* bixcor_last_error(void);
/* number of kernels launched by this library since init (bench `gpu_launches`) */
int64_t bixcor_launch_count(void);
/* Bytes the host entry points moved over PCIe since the last reset (bench `e2e.h2d_bytes_per_step` / `d2h_bytes_per_step`).
 * The exact Spearman Gram (BIXCOR_PREC_I8EXACT / AUTO with beta = 1 or 6) travels as int32 - 4 bytes per pair instead of 8 -
 * and the library's host staging threads finish r = S / (|d_i| |d_j|) while they move it into the caller's f64 buffer
 * (bit-identical to the device epilogue; env BIXCOR_S32_TRANSPORT=0 sends the f64 values instead).
 * Results of the <= 1e-5-class precisions (BIXCOR_PREC_TF32X3 / BIXCOR_PREC_F16X3: packed correlation, TOM) carry ~22
 * significant bits; their finished bands are rounded to binary32 on the device, travel as 4 bytes per value and are widened
 * into the caller's f64 buffer by the same threads (relative 6e-8, far inside the class; env BIXCOR_F32_TRANSPORT=0 sends
 * f64).  The FP64-class precisions (F64, OZAKI) and every differential-correlation output always travel as f64. */
int bixcor_transfer_counters(int64_t* h2d_bytes, int64_t* d2h_bytes, int reset);
/* Run-time form of BIXCOR_S32_TRANSPORT / BIXCOR_F32_TRANSPORT (1 = on, 0 = off: f64 on the wire, every arithmetic operation of
 * the result on the GPU; negative = leave as it is).  bench.py reports the end-to-end time both ways. */
int bixcor_set_transport(int int32_exact_spearman, int binary32_split_float);
/* The host halves of the two transports as plain functions (no GPU involved; what the staging threads do to one piece):
 * dst[t] = (double)src[t] * inv_norm[i] * inv_norm[j] (and, with pow6, sq = v v, sq sq sq) for the packed elements
 * [g0, g0 + count) of a p-gene triangle - the device epilogue's IEEE operations in the device epilogue's order - and the
 * float -> double widening.  Exposed for the CPU-only tests (tests/test_abi.py). */
int bixcor_host_expand_s32(const int32_t* src, int64_t count, int64_t g0, const double* inv_norm, int64_t p, int shift, int pow6,
                           double* dst);
int bixcor_host_widen_f32(const float* src, int64_t count, double* dst);
/* Fisher-z variance inflation for Spearman (default 1.06, Fieller); Pearson uses 1. */
int bixcor_set_fieller(double c);
/* Test hook of the exception barrier: the NEXT API call throws from inside its body
 * (1 = std::bad_alloc -> BIXCOR_ERR_OOM, 2 = std::runtime_error -> BIXCOR_ERR_INTERNAL). */
int bixcor_debug_inject(int what);

/* ---- packed triangle helpers (pure index logic, run on the host) ------- */
int64_t bixcor_packed_len(int64_t p, int shift);
int64_t bixcor_packed_offset(int64_t i, int64_t p, int shift);
int bixcor_upper_to_dense(const double* packed, int shift, int64_t p, double* out_pp);
int bixcor_dense_to_upper(const double* dense_pp, int shift, int64_t p, double* out_packed);
/* rs_upper_triangle_to_sparse (src/rust/src/data/r_sparse.rs:43-55; crate CompressedSparseData2::from_upper_triangle_sym)
 * behind upper_triangular_sym_mat$get_sparse_matrix() (R/classes_triangular_mat.R:125-140): the symmetric matrix of a
 * packed vector as compressed sparse columns (== rows, it is symmetric): exact zeros dropped, diagonal := 1 when shift,
 * row indices ascending inside a column, 0-based (inst/tinytest/test_sparse.R:5-50 compares with Matrix's @i/@p/@x).
 * Two calls: the count, then the fill into caller-allocated indptr[p+1] / indices[nnz] / data[nnz]. */
int64_t bixcor_upper_to_sparse_nnz(const double* packed, int shift, int64_t p);
int bixcor_upper_to_sparse(const double* packed, int shift, int64_t p, int64_t* indptr, int32_t* indices, double* data);
/* Block-row range [row_begin,row_end) of rank `rank` of `world`: equal pair-count
 * quantiles rounded to 128 rows, so every rank owns one contiguous packed range. */
int bixcor_shard_rows(int64_t p, int shift, int world, int rank, int64_t* row_begin, int64_t* row_end);

/* ---- host-pointer entry points (the drop-in boundary) ------------------ */
int bixcor_rank_columns(const double* x, int64_t n, int64_t p, double* out_np);

int bixcor_cor_dense(const double* x, int64_t n, int64_t p, int spearman, int precision, double* out_pp);

/* beta == 1: plain r.  beta != 1: soft-threshold extension |r|^beta. */
int bixcor_cor_upper(const double* x, int64_t n, int64_t p, int spearman, int shift, double beta,
                     int precision, double* out_packed);
/* Same, restricted to rows [row_begin,row_end): out_slice receives the packed range
 * starting at bixcor_packed_offset(row_begin,p,shift).  Row ranges are block-row ranges of the 128 x 128 tiling:
 * row_begin must be a multiple of 128 and row_end a multiple of 128 or p (bixcor_shard_rows produces such ranges);
 * anything else is BIXCOR_ERR_INVALID.  An empty range (row_begin == row_end) is a no-op.  This holds for every
 * *_rows entry point below, host or device. */
int bixcor_cor_upper_rows(const double* x, int64_t n, int64_t p, int spearman, int shift, double beta,
                          int precision, int64_t row_begin, int64_t row_end, double* out_slice);

int bixcor_diffcor(const double* xa, int64_t na, const double* xb, int64_t nb, int64_t p, int spearman,
                   int precision, double* r_a, double* r_b, double* z, double* pval);
int bixcor_diffcor_rows(const double* xa, int64_t na, const double* xb, int64_t nb, int64_t p, int spearman,
                        int precision, int64_t row_begin, int64_t row_end, double* r_a, double* r_b,
                        double* z, double* pval);

/* rs_tom: the affinity matrix is used AS PASSED.  signed_: k_i = sum_j |a_ij| and |a_ij| in the denominators;
 * unsigned: k_i = sum_j a_ij and a_ij itself (R/functions_stats.R:15-37).  abs() for unsigned networks is the R
 * wrappers' job (calculate_tom / calculate_tom_from_exp, R/functions_stats.R:52-54,114-116) and stays there;
 * cor_module_tom passes the raw correlation (R/methods_coexp_cor.R:151), so bixcor_tom_packed does not take abs either.
 * bixcor_tom_from_exp replaces the whole of calculate_tom_from_exp and therefore does (|r|^beta when unsigned). */
int bixcor_tom(const double* a_pp, int64_t p, int signed_, int version, int precision, double* out_pp);

/* cor_module_tom (R/methods_coexp_cor.R:118-164) in one call: the stored packed correlation (shift as stored) in,
 * the packed shift = 1 TOM out; replaces rs_upper_triangle_to_dense + rs_tom + rs_dense_to_upper_triangle. */
int bixcor_tom_packed(const double* packed_in, int shift_in, int64_t p, int signed_, int version, int precision,
                      double* out_packed);
/* packed != 0: out is the shift=1 packed vector (cor_module_tom layout), else dense p x p. */
int bixcor_tom_from_exp(const double* x, int64_t n, int64_t p, int spearman, int signed_, int version,
                        double beta, int precision, int packed, double* out);

/* Sibling Gram kernels sharing the same GEMM (FP64 path):
 *   rs_covariance  src/rust/src/base/r_cors_similarity.rs:49-55   == base R cov()
 *   rs_cos         src/rust/src/base/r_cors_similarity.rs:90-97   column cosine similarity
 *   rs_cor2        src/rust/src/base/r_cors_similarity.rs:114-122 == base R cor(x, y): p x q, column-major
 *   rs_cov2cor     src/rust/src/base/r_cors_similarity.rs:135-142 */
int bixcor_covariance(const double* x, int64_t n, int64_t p, double* out_pp);
int bixcor_cos(const double* x, int64_t n, int64_t p, double* out_pp);
int bixcor_cor2(const double* x, int64_t n, int64_t p, const double* y, int64_t q, int spearman, double* out_pq);
/* rs_rbh_cor (src/rust/src/methods/r_rbh.rs:187-272; crate calculate_rbh_cor) for ONE (origin, target) pair:
 * x (n x p) and y (n x q) hold the two module matrices restricted to their n shared features (rows aligned).
 * similarity = |cor(x_i, y_j)|; (i, j) is a hit when it is among the k_best largest of row i AND of column j and
 * exceeds min_similarity.  Hits are written origin-major with 0-based indices; *n_hits is the number found (an
 * error is returned when it exceeds capacity; p * q always suffices). */
int bixcor_rbh_cor(const double* x, int64_t n, int64_t p, const double* y, int64_t q, int k_best, int spearman,
                   double min_similarity, int32_t* out_origin, int32_t* out_target, double* out_sim, int64_t capacity,
                   int64_t* n_hits);
int bixcor_cov2cor(const double* cov_pp, int64_t p, double* out_pp);

/* RBF affinities.  `rbf_type` mirrors parse_rbf_types ("gaussian", "bump", "inverse_quadratic"):
 *   rs_rbf_function / rs_rbf_function_mat  src/rust/src/base/r_rbf.rs:36-83  -> bixcor_rbf (a matrix is its flat storage)
 *   rs_rbf_iterate_epsilons                src/rust/src/base/r_rbf.rs:118-131 -> bixcor_rbf_iterate_epsilons:
 *     out is p x n_eps column-major, column e = column sums of the symmetric affinity matrix rebuilt from the
 *     packed distances with epsilons[e] (diagonal := 1 when shift, as rs_upper_triangle_to_dense)
 *   bixcor_cor_upper_rbf: fused correlation -> d = 1 - |r| -> phi(d) (or 1 - phi(d) when complement) in the Gram epilogue,
 *     packed output; the R side computes this as rs_cor_upper_triangle + rs_rbf_function (R/methods_coexp_cor.R:357-370). */
#define BIXCOR_RBF_GAUSSIAN 1
#define BIXCOR_RBF_BUMP 2
#define BIXCOR_RBF_INVERSE_QUADRATIC 3
int bixcor_rbf(const double* x, int64_t len, double epsilon, int rbf_type, double* out);
int bixcor_rbf_iterate_epsilons(const double* dist_packed, int64_t p, int shift, const double* epsilons, int n_eps,
                                int rbf_type, double* out_p_x_neps);
int bixcor_cor_upper_rbf(const double* x, int64_t n, int64_t p, int spearman, int shift, double epsilon, int rbf_type,
                         int complement, int precision, double* out_packed);
/* rs_coremo_stability (src/rust/src/methods/r_coremo.rs:191-235): for every 1-based sample index
 * the packed (shift = 1) vector 1 - phi(1 - |r|) of the correlation without that sample; outs[k] receives
 * p(p-1)/2 doubles for indices_1based[k].  Each leave-one-out pass re-ranks / re-standardises on the device and
 * runs the Gram with the transform fused; the device-to-host copy of pass k overlaps the kernels of pass k + 1. */
int bixcor_coremo_stability(const double* x, int64_t n, int64_t p, const int32_t* indices_1based, int n_idx,
                            double epsilon, int rbf_type, int spearman, int precision, double* const* outs);

/* cells x genes CSR (int64 indptr[cells+1], int32 indices, float data) -> packed shift=1 Pearson */
int bixcor_sparse_gram_cor(const int64_t* indptr, const int32_t* indices, const float* data, int64_t cells,
                           int64_t genes, int precision, double* out_packed);

/* Pair-list gene-gene correlations over a sparse cells x genes matrix:
 *   rs_pairwise_gene_cors_mc  src/rust/src/meta_cell/r_mc_processing.rs:258-283
 * The sparse list {indptr, indices (0-based), data} is CSR with rows = cells (is_csr != 0) or CSC with
 * columns = genes (is_csr == 0).  gene_idx1/2 are 0-based (as the reference requires); out[k] is the Pearson
 * (or Spearman: average-tie ranks over ALL cells, zeros tie) correlation of the two genes of pair k over all
 * cells including the implicit zeros.  Only the genes named in the pair lists are densified on the device.
 * Spearman over more than 16384 cells ranks the non-zero entries of a gene only (the zeros' shared rank follows
 * analytically); genes with more than 16384 non-zero cells are ranked by the chunked dense rank kernels. */
int bixcor_sparse_pairwise_cor(const int64_t* indptr, const int32_t* indices, const float* data, int64_t cells,
                               int64_t genes, int is_csr, const int32_t* gene_idx1, const int32_t* gene_idx2,
                               int64_t n_pairs, int spearman, double* out);
/* rs_pairwise_gene_cors (src/rust/src/single_cell/r_sc_processing.rs:474-499): the same over the n_keep cells listed
 * in cells_to_keep (0-based, unique; the Rust side keeps reading counts_genes.bin and hands over the gene columns). */
int bixcor_sparse_pairwise_cor_cells(const int64_t* indptr, const int32_t* indices, const float* data, int64_t cells,
                                     int64_t genes, int is_csr, const int32_t* cells_to_keep, int64_t n_keep,
                                     const int32_t* gene_idx1, const int32_t* gene_idx2, int64_t n_pairs, int spearman,
                                     double* out);

/* ---- device-pointer entry points (inputs/outputs resident in HBM) ------ */
/* All pointers are device pointers on the library's current device; work is
 * enqueued on the library's compute stream and the call returns after
 * synchronising it.  Used by bench.py for the HBM-resident `value` and by the
 * one-process-per-GPU multi-GPU drivers, which run collectives between phases. */
int bixcor_dev_rank_columns(const double* d_x, int64_t n, int64_t p, double* d_out_np);
int bixcor_dev_cor_upper_rows(const double* d_x, int64_t n, int64_t p, int spearman, int shift, double beta,
                              int precision, int64_t row_begin, int64_t row_end, double* d_out_slice);
/* Operand exchange for multi-process drivers.  The Gram operand (standardised columns, TF32 hi/lo or
 * int8 digit planes) is gene-major: bixcor_dev_operand_row_bytes(n, precision) contiguous bytes per gene,
 * so a gene slab is one contiguous range that ranks can build independently and all-gather over NVLink.
 * d_operand / d_inv_norm (p doubles, written for BIXCOR_PREC_I8EXACT) are base pointers of the full buffers. */
int64_t bixcor_dev_operand_row_bytes(int64_t n, int precision);
int bixcor_dev_build_operand(const double* d_x, int64_t n, int64_t p, int spearman, int precision,
                             int64_t gene_begin, int64_t gene_end, void* d_operand, double* d_inv_norm);
/* Rebuilds the p per-gene norms of a BIXCOR_PREC_I8EXACT operand from its digit records (exact integer sums), so a
 * multi-process driver only all-gathers the operand; a no-op for the other precisions. */
int bixcor_dev_operand_inv_norm(const void* d_operand, int64_t n, int64_t p, int precision, double* d_inv_norm);
int bixcor_dev_cor_upper_rows_op(const void* d_operand, const double* d_inv_norm, int64_t n, int64_t p,
                                 int precision, int shift, double beta, int64_t row_begin, int64_t row_end,
                                 double* d_out_slice);
/* Same Gram, but the packed rows land in HOST memory: every finished band of block rows is copied
 * device-to-host while later bands compute (the multi-process end-to-end path: H2D of one gene slab,
 * operand all-gather over NVLink, then this call). h_out_slice may be pinned or pageable. */
int bixcor_dev_cor_upper_rows_op_to_host(const void* d_operand, const double* d_inv_norm, int64_t n, int64_t p,
                                         int precision, int shift, double beta, int64_t row_begin,
                                         int64_t row_end, double* h_out_slice);
/* Fused operand exchange over peer memory: the multi-process (one process per GPU, one box) form of the three calls above
 * without a collective library in the data path.  Every rank allocates one exchange block (two operand buffers, two norm
 * buffers, flags) and exports it with cudaIpc (bixcor_xchg_create: `handle_out` receives BIXCOR_XCHG_HANDLE_BYTES bytes); the
 * caller all-gathers the handles in rank order by whatever means it has (torch.distributed, MPI, a file) and every rank maps
 * its peers' blocks (bixcor_xchg_open).  bixcor_dev_cor_upper_rows_xchg then (1) runs the column kernel on this rank's
 * gene slab (equal 128-aligned slabs in rank order) and stores every finished record straight into the operand buffer of
 * EVERY rank over NVLink / NVSwitch - for the exact Spearman kind with n <= 1024 from inside the rank kernel, so the
 * transfer overlaps the sort; other kinds copy the slab with a peer-store kernel -, (2) synchronises the ranks with a
 * one-CTA flag barrier in the same peer-mapped block, (3) computes the packed rows [row_begin,row_end) from the local
 * buffer.  It replaces the all-gather collective + bixcor_dev_operand_inv_norm of the three-call form (no separate
 * collective kernel, no norm kernel).  A collective call: every rank of the exchange must make it, with its own row range
 * (an empty one is fine).  precision: as for the other calls, BIXCOR_PREC_AUTO resolved for Spearman; FP64, 3xTF32, 3xF16 and
 * exact int8 have per-gene records (Ozaki does not).  world <= 8.  d_x: the full n x p matrix in HBM; the _to_host form
 * takes only this rank's gene slab (genes [rank*slab, ...)) and returns the rows through the band-pipelined D2H. */
#define BIXCOR_XCHG_HANDLE_BYTES 64
int bixcor_xchg_create(int64_t n, int64_t p, int precision, int world, int rank, void* handle_out);
int bixcor_xchg_open(const void* handles_in_rank_order);
/* Teardown / re-creation is two-phase because a block must outlive its remote mappings: every rank unmaps its peers' blocks
 * (bixcor_xchg_release_peers), the caller synchronises the ranks, then bixcor_xchg_destroy / the next bixcor_xchg_create
 * frees the rank's own block. */
int bixcor_xchg_release_peers(void);
int bixcor_xchg_destroy(void);
int bixcor_dev_cor_upper_rows_xchg(const double* d_x, int64_t n, int64_t p, int spearman, int shift, double beta,
                                   int64_t row_begin, int64_t row_end, double* d_out_slice);
int bixcor_dev_cor_upper_rows_xchg_to_host(const double* d_x_slab, int64_t n, int64_t p, int spearman, int shift, double beta,
                                           int64_t row_begin, int64_t row_end, double* h_out_slice);
int bixcor_dev_cor_dense(const double* d_x, int64_t n, int64_t p, int spearman, int precision, double* d_out_pp);
int bixcor_dev_diffcor_rows(const double* d_xa, int64_t na, const double* d_xb, int64_t nb, int64_t p,
                            int spearman, int precision, int64_t row_begin, int64_t row_end, double* d_r_a,
                            double* d_r_b, double* d_z, double* d_pval);
/* TOM in three phases so that a multi-process driver can all-gather A and
 * all-reduce k between them:
 *   phase 1: affinity block rows [row_begin,row_end) of the dense p x p matrix
 *            d_a (full columns of those genes: d_a + row_begin*p) from expression;
 *   phase 2: connectivity k_i = sum_j |a_ij| (signed) / sum_j a_ij (unsigned) for columns [row_begin,row_end);
 *   phase 3: TOM for block rows [row_begin,row_end) from the full A and full k. */
int bixcor_dev_affinity_cols(const double* d_x, int64_t n, int64_t p, int spearman, int signed_, double beta,
                             int precision, int64_t row_begin, int64_t row_end, double* d_a_pp);
int bixcor_dev_tom_connectivity(const double* d_a_pp, int64_t p, int signed_, int64_t row_begin, int64_t row_end,
                                double* d_k);
int bixcor_dev_tom_rows(const double* d_a_pp, const double* d_k, int64_t p, int signed_, int version,
                        int precision, int packed, int64_t row_begin, int64_t row_end, double* d_out);
/* Split-float kinds (BIXCOR_PREC_TF32X3 / BIXCOR_PREC_F16X3) with several ranks: the A * A operand is the pair of hi / lo
 * planes of A, bixcor_dev_tom_plane_row_bytes(p, precision) contiguous bytes per gene.  Every rank splits only its own
 * slab (rows [row_begin,row_end) of d_a_pp) into the full plane buffer, the planes are all-gathered (half the bytes of
 * the f64 matrix for the f16 kind, and no rank converts the whole matrix), and phase 3 runs from the planes alone: the
 * epilogue rebuilds a_ij and the diagonal from hi + lo (22 significant bits, inside the 1e-5 class of these kinds). */
int64_t bixcor_dev_tom_plane_row_bytes(int64_t p, int precision);
int bixcor_dev_tom_split_planes(const double* d_a_pp, int64_t p, int precision, int64_t row_begin, int64_t row_end,
                                void* d_planes);
int bixcor_dev_tom_rows_planes(const void* d_planes, const double* d_k, int64_t p, int signed_, int version,
                               int precision, int packed, int64_t row_begin, int64_t row_end, double* d_out);
/* sparse: accumulate the partial Gram (dense genes x genes, upper tiles valid) and
 * column sums of this rank's cell shard; finalise after the all-reduce. */
int bixcor_dev_sparse_gram_accumulate(const int64_t* d_indptr, const int32_t* d_indices, const float* d_data,
                                      int64_t cells, int64_t genes, int precision, double* d_gram_pp,
                                      double* d_colsum);
int bixcor_dev_sparse_gram_finalise(const double* d_gram_pp, const double* d_colsum, int64_t total_cells,
                                    int64_t genes, double* d_out_packed);

/* Run the device-pointer entry points on a caller-owned CUDA stream (external != 0;
 * e.g. torch's current stream, a NULL handle then means the legacy default stream) or
 * back on the library's own stream (external == 0), and optionally let them return
 * without synchronising so that consecutive calls queue back to back. */
int bixcor_dev_set_stream(void* cuda_stream, int external);
int bixcor_dev_set_async(int on);
int bixcor_dev_synchronize(void);

/* Per-kernel CUDA-event timing on the compute stream.  Enable, run any number of
 * calls, then read the accumulated milliseconds: what[0]=rank/standardise kernels,
 * [1]=Gram/TOM GEMM kernels, [2]=other kernels, [3]=number of GEMM launches.
 * bixcor_timing_enable(x) also resets the accumulators. */
int bixcor_timing_enable(int on);
int bixcor_last_timing(double* what4);

#ifdef __cplusplus
}
#endif
#endif /* BIXCOR_H */
