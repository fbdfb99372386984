/*
 * chol_mmio.h -- the subset of the NIST Matrix Market I/O API the reference uses, same names and
 * semantics: mm_read_banner (reference mmio.c:96-179), mm_read_mtx_crd_size (mmio.c:189-217),
 * mm_write_banner (mmio.c:386-397), mm_write_mtx_crd_size (mmio.c:181-187), mm_typecode_to_str
 * (mmio.c:448-511), and the MM_typecode accessors (mmio.h:18-75).  Callers: mmat.rg:76-100,128-129.
 */
#ifndef CHOL_MMIO_H
#define CHOL_MMIO_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MM_MAX_LINE_LENGTH 1025
#define MatrixMarketBanner "%%MatrixMarket"
#define MM_MAX_TOKEN_LENGTH 64

typedef char MM_typecode[4];

int mm_read_banner(FILE *f, MM_typecode *matcode);
int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz);
int mm_write_banner(FILE *f, MM_typecode matcode);
int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz);
char *mm_typecode_to_str(MM_typecode matcode); /* malloc'd; caller frees */

#define mm_is_matrix(t) ((t)[0] == 'M')
#define mm_is_sparse(t) ((t)[1] == 'C')
#define mm_is_coordinate(t) ((t)[1] == 'C')
#define mm_is_dense(t) ((t)[1] == 'A')
#define mm_is_array(t) ((t)[1] == 'A')
#define mm_is_complex(t) ((t)[2] == 'C')
#define mm_is_real(t) ((t)[2] == 'R')
#define mm_is_pattern(t) ((t)[2] == 'P')
#define mm_is_integer(t) ((t)[2] == 'I')
#define mm_is_symmetric(t) ((t)[3] == 'S')
#define mm_is_general(t) ((t)[3] == 'G')
#define mm_is_skew(t) ((t)[3] == 'K')
#define mm_is_hermitian(t) ((t)[3] == 'H')

#define mm_set_matrix(t) ((*t)[0] = 'M')
#define mm_set_coordinate(t) ((*t)[1] = 'C')
#define mm_set_array(t) ((*t)[1] = 'A')
#define mm_set_dense(t) mm_set_array(t)
#define mm_set_sparse(t) mm_set_coordinate(t)
#define mm_set_complex(t) ((*t)[2] = 'C')
#define mm_set_real(t) ((*t)[2] = 'R')
#define mm_set_pattern(t) ((*t)[2] = 'P')
#define mm_set_integer(t) ((*t)[2] = 'I')
#define mm_set_symmetric(t) ((*t)[3] = 'S')
#define mm_set_general(t) ((*t)[3] = 'G')
#define mm_set_skew(t) ((*t)[3] = 'K')
#define mm_set_hermitian(t) ((*t)[3] = 'H')
#define mm_clear_typecode(t) ((*t)[0] = (*t)[1] = (*t)[2] = ' ', (*t)[3] = 'G')
#define mm_initialize_typecode(t) mm_clear_typecode(t)

#define MM_COULD_NOT_READ_FILE 11
#define MM_PREMATURE_EOF 12
#define MM_NOT_MTX 13
#define MM_NO_HEADER 14
#define MM_UNSUPPORTED_TYPE 15
#define MM_LINE_TOO_LONG 16
#define MM_COULD_NOT_WRITE_FILE 17

#define MM_MTX_STR "matrix"
#define MM_ARRAY_STR "array"
#define MM_DENSE_STR "array"
#define MM_COORDINATE_STR "coordinate"
#define MM_SPARSE_STR "coordinate"
#define MM_COMPLEX_STR "complex"
#define MM_REAL_STR "real"
#define MM_INT_STR "integer"
#define MM_GENERAL_STR "general"
#define MM_SYMM_STR "symmetric"
#define MM_HERM_STR "hermitian"
#define MM_SKEW_STR "skew-symmetric"
#define MM_PATTERN_STR "pattern"

#ifdef __cplusplus
}
#endif
#endif
