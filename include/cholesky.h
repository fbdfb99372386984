/*
 * cholesky.h -- C ABI of the B200-native sparse Cholesky numeric-factorization engine.
 *
 * Drop-in boundary for the hot path of syamajala/cholesky.  In the reference this file exports
 * one symbol, `void register_mappers()` (cholesky.h:19-27, cholesky.cc:88-91), and the numeric
 * path sits behind Terra FFI calls into CBLAS/LAPACKE issued by the fused leaf tasks
 * (blas.rg:292-504) that the level loop launches (mmat.rg:1227-1355).  The entry points below
 * replace exactly that: `chol_factor` is the level loop, `chol_fused_*` are the per-level fused
 * leaf tasks, and the loaders/writers keep the reference's file formats (mmio.c, mnd.c,
 * mmat.rg:102-147).  Plain pointers and sizes only; int return 0 = ok, <0 = error
 * (chol_last_error gives the text).  Host pointers are borrowed; device memory is owned by the
 * handle.  One host thread drives a handle.  A handle drives one GPU, or (chol_create with ngpu = 2, 4
 * or 8) a group of GPUs of one node from one process: the same calls then run subtree-partitioned over
 * the devices listed, which exchange data through NVLink peer memory.
 */
#ifndef __CHOLESKY_H__
#define __CHOLESKY_H__

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct chol chol_t;

/* reference `fspace Filled` (blas.rg:55-61): Legion int1d/int2d/rect2d are 64-bit => 9 x int64.
 * filled == 0 means structurally FILLED (mmat.rg:615, 1205-1206). */
typedef struct {
  int64_t filled, sep_x, sep_y, interval, cluster, lo_x, lo_y, hi_x, hi_y;
} chol_filled_t;

typedef struct {
  double seconds_best, seconds_median, seconds_last; /* level loop only (mmat.rg:1226-1355 region) */
  double assemble_seconds;                           /* fill_block equivalent (mmat.rg:1216-1224) */
  double flops;                                      /* algorithmic flops of the reference BLAS call list */
  int64_t kernel_launches;                           /* kernels launched per iteration in the timed region */
  int info;                                          /* 0, or 1-based permuted column of a non-positive pivot */
} chol_stats_t;

/* ---- lifecycle.  replaces: regentlib.start(main, register_mappers) (mmat.rg:1498) and the processor
 * selection of its command line (-ll:gpu / -ll:cpu).  devices[0 .. ngpu): CUDA device ordinals, ngpu = 1, 2, 4
 * or 8 (anything else is refused); a device may be listed more than once (several ranks then share it: a
 * way to exercise the partitioned path on one GPU). */
int chol_create(const int *devices, int ngpu, chol_t **out);
int chol_num_ranks(chol_t *);                /* 1, or ngpu of a group handle */
chol_t *chol_rank_handle(chol_t *, int r);   /* borrowed handle of rank r of a group, for the per-rank inspection calls
                                                (chol_partition_stats, chol_get_launch, chol_launch_times, chol_solve_stats) */
void chol_destroy(chol_t *);
const char *chol_last_error(chol_t *);
void register_mappers(void); /* kept so that anything linking the old symbol still resolves; no-op */

/* ---- inputs.  replaces: read_matrix_banner/read_separators/read_clusters/read_matrix
 * (mmat.rg:76-100, 851-883 -> mmio.c:96-217, mnd.c:22-199) */
int chol_load(chol_t *, const char *matrix_mtx, const char *separators_txt, const char *clusters_txt);
/* same data from memory: lower-triangle coordinate entries (0-based, row >= col); separator dof
 * lists concatenated in ascending separator id (sep_ptr has nsep+1 entries); cluster interval lists
 * concatenated per separator per interval: iv_ptr has (total intervals + 1) entries, sep_iv_ptr has
 * nsep+1 entries indexing iv_ptr. */
int chol_load_arrays(chol_t *, int n, int64_t nz, const int32_t *I, const int32_t *J, const double *V, int levels,
                     int nsep, const int64_t *sep_ptr, const int32_t *sep_dofs, const int64_t *sep_iv_ptr,
                     const int64_t *iv_ptr, const int32_t *iv_vals);
/* synthetic inputs of BASELINE.json: nx*ny*nz grid Laplacian, stencil 5 (2-D), 7 or 27 (3-D), dof
 * index x + nx*(y + ny*z), geometric nested dissection to `levels` levels (0 = the utils.py:6-7
 * rule ceil(log2(n/64))+1) with the cluster interval hierarchy of the reference's clust files. */
int chol_generate(chol_t *, int nx, int ny, int nz, int stencil, int levels);
/* dump the loaded/generated problem in the reference's three text formats (any path may be NULL) */
int chol_write_inputs(chol_t *, const char *matrix_mtx, const char *separators_txt, const char *clusters_txt);

/* ---- host symbolic analysis.  replaces: build_separator_tree, partition_matrix,
 * find_index_space_2d/3d, fill_block's flags, compute_filled_clusters (mmat.rg:299-1028).
 * keep_records != 0 keeps every Filled record for chol_get_filled (counts and checksums are
 * always kept). */
int chol_analyze(chol_t *, int keep_records);
/* the result of chol_analyze as a file, so that the ranks of a node (one process per GPU) analyse once: one rank
 * calls chol_analyze + chol_save_analysis, the others chol_load_analysis instead of chol_analyze (after loading the
 * same problem and setting their partition; the file is refused if it belongs to a different problem) */
int chol_save_analysis(chol_t *, const char *path);
int chol_load_analysis(chol_t *, const char *path);
int chol_n(chol_t *);
int64_t chol_nz(chol_t *);
int chol_levels(chol_t *);
int chol_num_separators(chol_t *);
int chol_max_int_size(chol_t *);
int64_t chol_num_blocks(chol_t *);
int64_t chol_num_clusters0(chol_t *);
int chol_get_perm(chol_t *, int32_t *perm);                 /* permuted row -> original dof */
int chol_get_sep_sizes(chol_t *, int32_t *sizes_by_label);  /* nsep entries */
int64_t chol_get_block_bounds(chol_t *, int64_t *out6);     /* (row_sep, col_sep, lo_x, lo_y, hi_x, hi_y) */
int64_t chol_num_filled(chol_t *, int interval_lbl);
int64_t chol_get_filled(chol_t *, int interval_lbl, chol_filled_t *out); /* sorted (sep_x, sep_y, cluster) */
uint64_t chol_filled_checksum(chol_t *, int interval_lbl);  /* order independent, same hash as the oracle */
double chol_flops(chol_t *);
int chol_flops_by_level(chol_t *, double *potrf, double *trsm, double *syrk, double *gemm);
int chol_call_counts(chol_t *, int64_t *c4);                /* reference BLAS calls: potrf, trsm, syrk, gemm */
int64_t chol_factor_doubles(chol_t *);                      /* device doubles of factor storage */
/* algorithmic HBM bytes of one tree level: [0] its panels (factored in place), [1] distinct operands of its
 * Schur updates, [2] their destination clusters (read-modify-written) -- the numerators of the HBM roofline */
int chol_level_bytes(chol_t *, int lvl, double *out3);

/* ---- numeric factorization on the GPU.  replaces: the level loop mmat.rg:1211-1358 and the
 * fused leaf tasks blas.rg:292-504.  chol_factor = `iterations` x (assemble; level loop), timing
 * the level loop with CUDA events on the launching stream (warmup iterations run first, untimed). */
int chol_assemble(chol_t *);
int chol_factor(chol_t *, int iterations, int warmup, chol_stats_t *stats);
/* piecewise: one tree level, one phase -- mirror fused_dpotrf / fused_dtrsm / fused_dsyrk+dgemm */
int chol_fused_dpotrf(chol_t *, int lvl);
int chol_fused_dtrsm(chol_t *, int lvl);
int chol_fused_update(chol_t *, int lvl);
/* end to end with HOST buffers: upload A's values (in the order given at load time), assemble,
 * factor, and read back diag(L) in permuted order.  values may be NULL (reuse the loaded ones). */
int chol_factor_host(chol_t *, const double *values, int64_t nz, double *diag_out, chol_stats_t *stats);
int chol_synchronize(chol_t *);
/* the compiled launch list: kind 0 panel_kernel (diagonal block of a block column + the rows below it) /
 * 2 gemm_grouped / 3 peer_sync / 4 reduce_rects / 5 (no kernel: a cross-stream dependency) / 6 push_rects, tree
 * level, phase (1 fused_dpotrf, 2 fused_dtrsm, 4 fused_dsyrk+dgemm), CTAs or rectangles, executed flops,
 * cfg = GEMM kernel (0: 64x64 tiles, 3: warp tiles) + 16 * stream (0 update, 1 chain, 2 background pushes)
 * + 256 * widest block column of a panel launch */
int64_t chol_num_launches(chol_t *);
int chol_get_launch(chol_t *, int64_t i, int *kind, int *level, int *phase, int64_t *ctas, double *flops, int *cfg);
/* per-kernel accounting of one instrumented factorization (device time by kernel class, ms): panel_kernel
 * (fused_dpotrf + fused_dtrsm of the block columns), the multi-GPU exchange kernels (0 on one GPU), the grouped
 * GEMM (fused_dsyrk + fused_dgemm and the in-panel trailing updates) and the flops the latter executed */
int chol_kernel_times(chol_t *, double *panel_ms, double *exchange_ms, double *gemm_ms, double *gemm_flops);
/* per-launch device time (ms) of that instrumented pass, in launch-list order; returns the count */
int64_t chol_launch_times(chol_t *, float *ms, int64_t cap);

/* ---- multi-GPU, one process and one handle per GPU (world = 1, 2, 4 or 8; the form torchrun / MPI
 * launchers need -- inside one process use a group handle instead).  Rank r owns the subtree under heap
 * index world + r; the rows of the top log2(world) levels' panels are dealt to the ranks below each
 * separator in blocks of 256 and every rank ends up holding the complete factored top panels.  Call
 * chol_set_partition before chol_analyze, then exchange the 128-byte IPC blobs of all ranks (any
 * host-side all-gather) and hand the concatenation to chol_ipc_import; the factorization then exchanges
 * data through NVLink peer memory from inside its own kernels.  Every rank must make the same sequence of
 * chol_factor / chol_factor_host / chol_kernel_times calls.  Result accessors report the panels the rank
 * owns (rank 0 also the top panels). */
int chol_set_partition(chol_t *, int rank, int world);
int chol_ipc_export(chol_t *, void *handles128);
int chol_ipc_import(chol_t *, const void *all_handles, int world);
/* what this rank's schedule covers: [0] matrix entries it assembles, [1] GEMM flops it executes,
 * [2] peer-store launches (rows pushed to other ranks), [3] doubles of the top panels, [4] diagonal tiles it factors,
 * [5] 64-row slabs it solves,
 * [6] bytes it pulls from peers for the partial-sum reduction, [7] bytes it pushes to peers, per factorization */
int chol_partition_stats(chol_t *, double *out8);
int chol_rank(chol_t *);
int chol_world(chol_t *);
/* verification: largest |difference| between the ranks' copies of the factored top panels, compared on the GPUs
 * through peer memory (each rank against the next one; a group handle reports the largest); 0 = bit-identical */
int chol_top_copies_diff(chol_t *, double *maxdiff);
/* verification: largest |difference| between the ranks' copies of the factored top panels, compared on the GPUs
 * through peer memory (each rank against the next one; a group handle reports the largest); 0 = bit-identical */
int chol_top_copies_diff(chol_t *, double *maxdiff);

/* ---- results.  replaces: write_matrix (mmat.rg:102-147) */
int64_t chol_factor_nnz(chol_t *);                                           /* entries != 0 */
int64_t chol_get_factor_coo(chol_t *, int32_t *I, int32_t *J, double *V);   /* 0-based permuted */
int chol_get_factor_dense(chol_t *, double *out_row_major_nxn);             /* small n only */
int chol_write_factor(chol_t *, const char *path, int full_precision);     /* "%0.8g" or "%.17g" */
/* the factor at scale (row f-4): binary block dump (header + one dense record per filled cluster, layout
 * in csrc/factor_file.cc) and its streaming conversion to write_matrix's text format, host only,
 * one record in memory at a time.  Converter returns 0, or < 0: -1 cannot read, -2 not a dump,
 * -3 cannot write, -4 truncated or corrupt, -5 entry count differs from the header. */
int chol_write_factor_binary(chol_t *, const char *path);
int chol_factor_binary_to_mtx(const char *bin_path, const char *mtx_path, int full_precision);
/* relative residual estimate ||(A - L L^T) W||_F / ||A W||_F, W = k <= 4 Rademacher columns (seeded).  Z = L (L^T W)
 * is evaluated on the GPU from the factor where it sits (two coalesced passes over it, nothing but n x 4 doubles
 * comes back), A W on the host from the loaded entries.  Partitioned handles: every rank calls
 * chol_residual_partial (its panels' share of Z, n x 4 doubles, permuted row order), the caller sums the shares
 * over the ranks and hands the sum to chol_residual_finish on any rank. */
int chol_residual(chol_t *, int k, uint64_t seed, double *rel);
int chol_residual_partial(chol_t *, int k, uint64_t seed, double *z_out_nx4);
int chol_residual_finish(chol_t *, int k, uint64_t seed, const double *z_sum_nx4, double *rel);

/* ---- debug trace (next row f-3).  replaces: the `-d <dir>` path of mmat.rg (1086-1090): the log lines
 * printed with debug = true (mmat.rg:331, 352, 396, 432, 1010; blas.rg:308, 340, 405, 422, 490) and the
 * per-task snapshots of write_blocks (mmat.rg:149-218), i.e. the inputs of verify.debug_factor
 * (verify.py:216-275).  Analyze with keep_records first.  The log is host-only (path NULL: stdout).
 * chol_factor_debug assembles, then runs the level loop one fused task group at a time on the GPU and
 * writes <dir>/{potrf,trsm,gemm}_lvl<L>_a..[_b..][_c..].mtx after each (and the "%0.2f" block dump
 * .txt when with_txt != 0); it leaves the complete factor in the handle. */
int chol_write_debug_log(chol_t *, const char *log_path);
int chol_factor_debug(chol_t *, const char *dir, int full_precision, int with_txt);

/* ---- solve (next row f-1).  replaces: mmat.rg:1364-1495, blas.rg:217-290, mnd.c:201-229 */
int chol_solve(chol_t *, const double *b, double *x); /* original dof order in and out; single-GPU or group handle */
/* the same sweeps on a partitioned handle (after chol_set_partition / chol_ipc_import / chol_factor): every
 * rank calls chol_solve_forward with the whole b and gets its contribution to the shared top rows
 * (chol_solve_top_size doubles); the caller sums those over the ranks (any host-side all-reduce) and hands
 * the sum to chol_solve_backward, which returns the entries of x the rank owns (original dof order, zeros
 * elsewhere: the sum over the ranks is x).  On a single-GPU handle the pair equals chol_solve. */
int64_t chol_solve_top_size(chol_t *);
int chol_solve_forward(chol_t *, const double *b, double *top_partial);
int chol_solve_backward(chol_t *, const double *top_sum, double *x_owned);
int chol_solve_stats(chol_t *, double *out12); /* tiles of the rank's solve schedule: [0..5] subtree, [6..11] shared top */
int chol_matvec(chol_t *, const double *x, double *y); /* host check helper: y = A x, original dof order */
int chol_read_vector(const char *path, int n, double *out);
int chol_write_solution(const char *path, int n, const double *x);

#ifdef __cplusplus
}
#endif

#endif /* __CHOLESKY_H__ */
