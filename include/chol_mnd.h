/*
 * chol_mnd.h -- readers of the reference's nested-dissection text files, plain-array outputs.
 * The reference's versions (mnd.h:28-66, mnd.c:22-271) write into Legion accessors
 * (legion_physical_region_t / legion_field_id_t arguments); Legion does not exist in this build,
 * so each reader keeps the reference's name behind an `mnd_` prefix, the same file semantics
 * token for token, and returns its data through caller-provided arrays.
 *   mnd_read_separators  <- read_separators (mnd.c:22-69)
 *   mnd_read_clusters    <- read_clusters   (mnd.c:71-150)
 *   mnd_read_matrix      <- read_matrix     (mnd.c:152-199; the hash-table insert is dropped, the
 *                           engine looks entries up by permuted position instead)
 *   mnd_read_vector      <- read_vector     (mnd.c:201-229)
 *   mnd_hash_sax         <- hash_sax        (mnd.c:252-257, uthash.h:602-610)
 */
#ifndef CHOL_MND_H
#define CHOL_MND_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mnd_SepInfo { /* SepInfo, mnd.h:23-26 */
  int levels;
  int num_separators;
} mnd_SepInfo;

/* returns the number of dofs read (>= 0) or < 0; dofs[p] / sep_of_row[p] for p in file order */
int mnd_read_separators(const char *file, int n, mnd_SepInfo *info, int32_t *dofs, int32_t *sep_of_row);
/* (idx, interval, separator label) triples in file order; call with idx == NULL to count.
 * *max_int_size is what the reference's read_clusters returns. */
int64_t mnd_read_clusters(const char *file, int64_t cap, int32_t *idx, int32_t *interval, int32_t *sep,
                          int *max_int_size);
int mnd_read_matrix(const char *file, int64_t nz, int32_t *I, int32_t *J, double *V); /* 0-based out */
int mnd_read_vector(const char *file, int n, double *out);
uint64_t mnd_hash_sax(uint64_t key);

#ifdef __cplusplus
}
#endif
#endif
