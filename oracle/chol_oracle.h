/*
 * chol_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the reference's symbolic analysis + numeric level loop
 * (reference: mmat.rg:299-1028, 1211-1358; blas.rg:63-504; mnd.c:22-229) over host
 * BLAS/LAPACK (dlopen'd OpenBLAS).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (cholesky_b200/) never links, imports or calls it.
 *
 * Parity pin: the reference ships no golden factors and cannot be built here (Regent /
 * Legion absent).  The oracle is pinned against the reference's own acceptance check,
 * verify.py (check_matrix / permute_matrix / check_solution, imported unmodified in the
 * build container) on its four fixtures, and the resulting golden vectors are committed
 * under tests/golden/ (see tests/golden/make_golden.py).
 */
#ifndef CHOL_ORACLE_H
#define CHOL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc orc_t;

/* One reference `Filled` record (blas.rg:55-61): 9 x int64. filled==0 means FILLED. */
typedef struct {
  int64_t filled, sep_x, sep_y, interval, cluster, lo_x, lo_y, hi_x, hi_y;
} orc_filled_t;

/* dlopen the host BLAS (scipy-bundled OpenBLAS, `scipy_` symbol prefix, or a plain one). */
int orc_set_blas(const char *libpath);
const char *orc_blas_config(void);

orc_t *orc_create(void);
void orc_destroy(orc_t *);
const char *orc_last_error(orc_t *);

/* readers: mnd.c:22-229 / mmat.rg:76-100 semantics */
int orc_load(orc_t *, const char *mtx, const char *ord, const char *clust);
/* tree, permutation, block bounds, assembly flags, symbolic fill (mmat.rg:299-1028).
 * literal_assembly=1 uses the reference's hash-table probe per dense entry (fill_block);
 * 0 scatters the nonzeros directly (same result, O(nz)); -1 picks by problem size. */
int orc_analyze(orc_t *, int literal_assembly);

int orc_n(orc_t *);
int orc_nz(orc_t *);
int orc_levels(orc_t *);
int orc_num_separators(orc_t *);
int orc_max_int_size(orc_t *);
int64_t orc_num_blocks(orc_t *);
int64_t orc_num_clusters0(orc_t *); /* allocated interval-0 clusters */
int orc_get_perm(orc_t *, int32_t *perm_out);            /* permuted row -> original dof */
int orc_get_sep_sizes(orc_t *, int32_t *sizes_by_label); /* nsep entries, label 1..nsep */
/* block bounds in reference order-independent form: for every allocated block
 * (row_sep, col_sep, lo_x, lo_y, hi_x, hi_y); returns count; out may be NULL */
int64_t orc_get_block_bounds(orc_t *, int64_t *out6);
int64_t orc_num_filled(orc_t *, int interval_lbl);
int64_t orc_get_filled(orc_t *, int interval_lbl, orc_filled_t *out); /* sorted (sep_x,sep_y,cluster) */
uint64_t orc_filled_checksum(orc_t *, int interval_lbl);              /* order independent */
int64_t orc_factor_nnz_alloc(orc_t *); /* doubles stored (filled-cluster storage) */
double orc_flops(orc_t *);             /* algorithmic flops of the reference BLAS call list */
int orc_flops_by_level(orc_t *, double *potrf, double *trsm, double *syrk, double *gemm); /* levels entries each */
int orc_call_counts(orc_t *, int64_t *c4); /* potrf, trsm, syrk, gemm calls */

/* numeric: (re)assemble A then run the level loop (mmat.rg:1211-1358).
 * threads: worker threads across the tasks of one level (1 BLAS thread each, mmat.rg:1057);
 * phases with fewer tasks than threads run serially with `threads` BLAS threads.
 * first_level/last_level bound the loop (levels-1 .. 0 for everything). */
int orc_assemble(orc_t *);
int orc_factor(orc_t *, int threads, double *seconds);
int orc_factor_levels(orc_t *, int threads, int from_level, int to_level, double *seconds);
/* piecewise entry points mirroring the fused tasks, one tree level each */
int orc_fused_dpotrf(orc_t *, int lvl);
int orc_fused_dtrsm(orc_t *, int lvl);
int orc_fused_update(orc_t *, int lvl); /* fused_dsyrk + fused_dgemm */

/* results */
int64_t orc_factor_nnz(orc_t *);                        /* entries != 0, as write_matrix counts */
int64_t orc_get_factor_coo(orc_t *, int32_t *I, int32_t *J, double *V); /* 0-based permuted */
int orc_get_factor_dense(orc_t *, double *out_row_major_nxn);           /* small n only */
int orc_write_factor(orc_t *, const char *path, int full_precision);   /* mmat.rg:102-147 */
/* solve (mmat.rg:1364-1495): b in original dof order -> x in original dof order */
int orc_solve(orc_t *, const double *b, double *x);
int orc_read_vector(const char *path, int n, double *out);               /* mnd.c:201-229 */
int orc_write_solution(const char *path, int n, const double *x);      /* mmat.rg:785-798 */

/* the `-d` debug path: log (mmat.rg:331,352,396,432,1010; blas.rg:308,340,405,422,490) and one
 * snapshot <dir>/{potrf,trsm,gemm}_lvl*.mtx per fused task (write_blocks, mmat.rg:149-218), in
 * program order; what verify.debug_factor (verify.py:216-275) replays */
int orc_debug_trace(orc_t *, const char *dir, const char *log_path, int full_precision);

uint64_t orc_hash_sax(uint64_t key); /* uthash.h:602-610 over the 8 key bytes */

#ifdef __cplusplus
}
#endif
#endif
