// File formats of the reference, kept as the drop-in surface:
//   matrix      Matrix Market coordinate, lower triangle (mmio.c:96-217; entries as mnd.c:152-199)
//   separators  "levels nsep" then "id;d0,d1,...,"   (mnd.c:22-69)
//   clusters    header then "id;iv0;iv1;...;"         (mnd.c:71-150)
//   vector      Matrix Market array, 3 header lines   (mnd.c:201-229)
// Outputs are plain arrays (no Legion accessors).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/chol_mmio.h"
#include "../../include/chol_mnd.h"
#include "chol_internal.h"

// ---------------------------------------------------------------------------- mmio subset
extern "C" {

int mm_read_banner(FILE *f, MM_typecode *matcode) {
  char line[MM_MAX_LINE_LENGTH];
  char banner[MM_MAX_TOKEN_LENGTH], mtx[MM_MAX_TOKEN_LENGTH], crd[MM_MAX_TOKEN_LENGTH];
  char data_type[MM_MAX_TOKEN_LENGTH], storage[MM_MAX_TOKEN_LENGTH];
  mm_clear_typecode(matcode);
  if (fgets(line, MM_MAX_LINE_LENGTH, f) == NULL) return MM_PREMATURE_EOF;
  if (sscanf(line, "%63s %63s %63s %63s %63s", banner, mtx, crd, data_type, storage) != 5) return MM_PREMATURE_EOF;
  char *fields[4] = {mtx, crd, data_type, storage};
  for (char *s : fields)
    for (char *p = s; *p; p++) *p = (char)tolower((unsigned char)*p);
  if (strncmp(banner, MatrixMarketBanner, strlen(MatrixMarketBanner)) != 0) return MM_NO_HEADER;
  if (strcmp(mtx, MM_MTX_STR) != 0) return MM_UNSUPPORTED_TYPE;
  mm_set_matrix(matcode);
  if (!strcmp(crd, MM_SPARSE_STR)) mm_set_sparse(matcode);
  else if (!strcmp(crd, MM_DENSE_STR)) mm_set_dense(matcode);
  else return MM_UNSUPPORTED_TYPE;
  if (!strcmp(data_type, MM_REAL_STR)) mm_set_real(matcode);
  else if (!strcmp(data_type, MM_COMPLEX_STR)) mm_set_complex(matcode);
  else if (!strcmp(data_type, MM_PATTERN_STR)) mm_set_pattern(matcode);
  else if (!strcmp(data_type, MM_INT_STR)) mm_set_integer(matcode);
  else return MM_UNSUPPORTED_TYPE;
  if (!strcmp(storage, MM_GENERAL_STR)) mm_set_general(matcode);
  else if (!strcmp(storage, MM_SYMM_STR)) mm_set_symmetric(matcode);
  else if (!strcmp(storage, MM_HERM_STR)) mm_set_hermitian(matcode);
  else if (!strcmp(storage, MM_SKEW_STR)) mm_set_skew(matcode);
  else return MM_UNSUPPORTED_TYPE;
  return 0;
}

int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz) {
  char line[MM_MAX_LINE_LENGTH];
  *M = *N = *nz = 0;
  do {
    if (fgets(line, MM_MAX_LINE_LENGTH, f) == NULL) return MM_PREMATURE_EOF;
  } while (line[0] == '%');
  if (sscanf(line, "%d %d %d", M, N, nz) == 3) return 0;
  int got;
  do {
    got = fscanf(f, "%d %d %d", M, N, nz);
    if (got == EOF) return MM_PREMATURE_EOF;
  } while (got != 3);
  return 0;
}

char *mm_typecode_to_str(MM_typecode matcode) {
  const char *t0, *t1, *t2, *t3;
  if (!mm_is_matrix(matcode)) return NULL;
  t0 = MM_MTX_STR;
  if (mm_is_sparse(matcode)) t1 = MM_SPARSE_STR;
  else if (mm_is_dense(matcode)) t1 = MM_DENSE_STR;
  else return NULL;
  if (mm_is_real(matcode)) t2 = MM_REAL_STR;
  else if (mm_is_complex(matcode)) t2 = MM_COMPLEX_STR;
  else if (mm_is_pattern(matcode)) t2 = MM_PATTERN_STR;
  else if (mm_is_integer(matcode)) t2 = MM_INT_STR;
  else return NULL;
  if (mm_is_general(matcode)) t3 = MM_GENERAL_STR;
  else if (mm_is_symmetric(matcode)) t3 = MM_SYMM_STR;
  else if (mm_is_hermitian(matcode)) t3 = MM_HERM_STR;
  else if (mm_is_skew(matcode)) t3 = MM_SKEW_STR;
  else return NULL;
  char buf[MM_MAX_LINE_LENGTH];
  snprintf(buf, sizeof buf, "%s %s %s %s", t0, t1, t2, t3);
  char *out = (char *)malloc(strlen(buf) + 1); /* caller frees, as mmio.c:448-453 */
  strcpy(out, buf);
  return out;
}

int mm_write_banner(FILE *f, MM_typecode matcode) {
  char *s = mm_typecode_to_str(matcode);
  if (!s) return MM_UNSUPPORTED_TYPE;
  int r = fprintf(f, "%s %s\n", MatrixMarketBanner, s);
  free(s);
  return r < 0 ? MM_COULD_NOT_WRITE_FILE : 0;
}

int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz) {
  return fprintf(f, "%d %d %d\n", M, N, nz) < 0 ? MM_COULD_NOT_WRITE_FILE : 0;
}

// ---------------------------------------------------------------------------- mnd subset
uint64_t mnd_hash_sax(uint64_t key) { /* uthash.h:602-610 over the 8 key bytes; mnd.c:252-257 */
  const unsigned char *k = (const unsigned char *)&key;
  uint64_t h = 0;
  for (unsigned i = 0; i < sizeof(uint64_t); i++) h ^= (h << 5) + (h >> 2) + k[i];
  return h;
}

static char *slurp(const char *path, size_t *len) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)sz + 2);
  size_t got = fread(buf, 1, (size_t)sz, f);
  fclose(f);
  buf[got] = 0;
  *len = got;
  return buf;
}

/* mnd.c:152-199: skip exactly two lines, then nz triples "i j val" (1-based). */
int mnd_read_matrix(const char *file, int64_t nz, int32_t *I, int32_t *J, double *V) {
  size_t len;
  char *buf = slurp(file, &len);
  if (!buf) return -1;
  char *p = buf;
  for (int l = 0; l < 2; l++) {
    char *nl = strchr(p, '\n');
    if (!nl) {
      free(buf);
      return -2;
    }
    p = nl + 1;
  }
  for (int64_t e = 0; e < nz; e++) {
    char *q;
    long i = strtol(p, &q, 10);
    if (q == p) {
      free(buf);
      return -3;
    }
    p = q;
    long j = strtol(p, &q, 10);
    if (q == p) {
      free(buf);
      return -3;
    }
    p = q;
    double v = strtod(p, &q);
    if (q == p) {
      free(buf);
      return -3;
    }
    p = q;
    I[e] = (int32_t)(i - 1), J[e] = (int32_t)(j - 1), V[e] = v;
  }
  free(buf);
  return 0;
}

/* mnd.c:201-229 */
int mnd_read_vector(const char *file, int n, double *out) {
  FILE *f = fopen(file, "r");
  if (!f) return -1;
  char buff[1024];
  for (int i = 0; i < 3; i++)
    if (!fgets(buff, sizeof buff, f)) {
      fclose(f);
      return -2;
    }
  for (int i = 0; i < n; i++) {
    double v = 0.0;
    if (fscanf(f, "%lg\n", &v) != 1) {
      fclose(f);
      return -3;
    }
    out[i] = v;
  }
  fclose(f);
  return 0;
}

/* mnd.c:22-69.  sep_of_row/dofs get one entry per listed dof, in file order. */
int mnd_read_separators(const char *file, int n, mnd_SepInfo *info, int32_t *dofs, int32_t *sep_of_row) {
  FILE *fp = fopen(file, "r");
  if (!fp) return -1;
  char *line = NULL;
  size_t cap = 0;
  int i = 0, pos = 0, rc = 0;
  info->levels = info->num_separators = 0;
  while (getline(&line, &cap, fp) != -1) {
    if (i == 0) {
      info->levels = atoi(&line[0]);
      info->num_separators = atoi(&line[2]); /* as mnd.c:45 */
      i++;
      continue;
    }
    char *save = NULL;
    char *rows = strtok_r(line, ";", &save);
    if (!rows) break;
    int separator = atoi(rows) + 1;
    rows = strtok_r(NULL, ",", &save);
    while (rows != NULL) {
      if (isspace((unsigned char)*rows)) break;
      if (pos >= n) {
        rc = -4;
        break;
      }
      dofs[pos] = atoi(rows);
      sep_of_row[pos] = separator;
      pos++;
      rows = strtok_r(NULL, ",", &save);
    }
    if (rc) break;
    i++;
  }
  free(line);
  fclose(fp);
  if (rc) return rc;
  return pos;
}

/* mnd.c:71-150, token for token.  Emits (idx, interval, separator) triples like the reference's
 * ClusterIndex region; returns the count, *max_int_size as the reference computes it. */
int64_t mnd_read_clusters(const char *file, int64_t cap_out, int32_t *idx, int32_t *interval_out, int32_t *sep_out,
                          int *max_int_size) {
  FILE *fp = fopen(file, "r");
  if (!fp) return -1;
  char *line = NULL;
  size_t cap = 0;
  int i = 0;
  int64_t k = 0;
  int mx = -1;
  while (getline(&line, &cap, fp) != -1) {
    if (i == 0) {
      i++;
      continue;
    }
    char *save = NULL;
    char *rows = strtok_r(line, "; ", &save);
    if (!rows) break;
    int separator = atoi(rows) + 1;
    int interval = 0, dofs = 0;
    rows = strtok_r(NULL, ",; ", &save);
    while (rows != NULL) {
      int row = atoi(rows);
      dofs++;
      rows = strtok_r(NULL, ",; ", &save);
      if (rows == NULL) {
        if (dofs > mx) mx = dofs;
      } else {
        if (idx) {
          if (k >= cap_out) {
            free(line), fclose(fp);
            return -2;
          }
          idx[k] = row, interval_out[k] = interval, sep_out[k] = separator;
        }
        k++;
        if (strcmp("0", rows) == 0) {
          if (dofs > mx) mx = dofs;
          interval++;
          dofs = 0;
        }
      }
    }
    i++;
  }
  free(line);
  fclose(fp);
  if (max_int_size) *max_int_size = mx;
  return k;
}

}  // extern "C"

// ---------------------------------------------------------------------------- Problem
namespace chb {

int finish_problem(Problem &P, std::string &err) {
  if (P.levels < 1 || P.levels > 24) return err = "bad level count", -1;
  if (P.N != (1 << P.levels) - 1) return err = "num_separators != 2^levels - 1", -1;
  P.start.assign(P.N + 2, 0);
  int acc = 0;
  for (int label = 1; label <= P.N; label++) {
    int h = P.heap_of(label);
    P.start[h] = acc;
    acc += P.sz[h];
  }
  if (acc != P.n) return err = "separator lists cover " + std::to_string(acc) + " dofs, matrix has " + std::to_string(P.n), -1;
  std::vector<char> seen(P.n, 0);
  for (int p = 0; p < P.n; p++) {
    int d = P.perm[p];
    if (d < 0 || d >= P.n || seen[d]) return err = "separator lists are not a permutation (dof " + std::to_string(d) + ")", -1;
    seen[d] = 1;
  }
  for (int h = 1; h <= P.N; h++) {
    if (P.iv[h].empty()) return err = "cluster file: separator id " + std::to_string(P.label_of(h) - 1) + " missing", -1;
    for (auto &l : P.iv[h])
      if (l.size() < 2 && P.sz[h] > 0) return err = "cluster interval too short", -1;
  }
  return 0;
}

int read_problem(Problem &P, const char *mtx, const char *ord, const char *clust, std::string &err) {
  FILE *f = fopen(mtx, "r");
  if (!f) return err = std::string("cannot open ") + mtx, -1;
  MM_typecode tc;
  if (mm_read_banner(f, &tc) != 0) {
    fclose(f);
    return err = "Unable to read banner.", -1; /* mmat.rg:83-86 */
  }
  int M, N, nz;
  if (mm_read_mtx_crd_size(f, &M, &N, &nz) != 0) {
    fclose(f);
    return err = "Unable to read matrix size.", -1; /* mmat.rg:93-96 */
  }
  fclose(f);
  memcpy(P.typecode, tc, 4);
  P.n = M, P.ncols = N, P.nz = nz;
  P.ei.resize(nz), P.ej.resize(nz), P.ev.resize(nz);
  if (mnd_read_matrix(mtx, nz, P.ei.data(), P.ej.data(), P.ev.data()) != 0) return err = "bad matrix entries", -1;

  mnd_SepInfo info;
  std::vector<int32_t> dofs(P.n), sepof(P.n);
  int got = mnd_read_separators(ord, P.n, &info, dofs.data(), sepof.data());
  if (got < 0) return err = std::string("cannot read separators from ") + ord, -1;
  P.levels = info.levels, P.N = info.num_separators;
  if (P.N < 1 || P.N != (1 << P.levels) - 1) return err = "num_separators != 2^levels - 1", -1;
  if (got != P.n) return err = "separator file lists " + std::to_string(got) + " dofs, matrix has " + std::to_string(P.n), -1;
  P.perm.assign(dofs.begin(), dofs.end());
  P.sz.assign(P.N + 2, 0);
  int prev = 0;
  for (int p = 0; p < P.n; p++) {
    int label = sepof[p];
    if (label < 1 || label > P.N || label < prev) return err = "separator ids must be ascending (fill_block indexes dofs by permuted row, mmat.rg:578)", -1;
    prev = label;
    P.sz[P.heap_of(label)]++;
  }

  int mx = -1;
  int64_t cnt = mnd_read_clusters(clust, 0, NULL, NULL, NULL, &mx);
  if (cnt < 0) return err = std::string("cannot read clusters from ") + clust, -1;
  std::vector<int32_t> ci(cnt), cint(cnt), csep(cnt);
  mnd_read_clusters(clust, cnt, ci.data(), cint.data(), csep.data(), &mx);
  P.max_int_size = mx;
  P.iv.assign(P.N + 2, {});
  for (int64_t k = 0; k < cnt; k++) {
    int label = csep[k];
    if (label < 1 || label > P.N) return err = "cluster file: bad separator id", -1;
    auto &lists = P.iv[P.heap_of(label)];
    if ((int)lists.size() <= cint[k]) lists.resize(cint[k] + 1);
    lists[cint[k]].push_back(ci[k]);
  }
  return finish_problem(P, err);
}

int write_problem(const Problem &P, const char *mtx, const char *ord, const char *clust, std::string &err) {
  if (mtx) {
    FILE *f = fopen(mtx, "w");
    if (!f) return err = std::string("cannot write ") + mtx, -1;
    MM_typecode tc;
    memcpy(tc, P.typecode, 4);
    mm_write_banner(f, tc);
    mm_write_mtx_crd_size(f, P.n, P.ncols, (int)P.nz);
    for (int64_t e = 0; e < P.nz; e++) {
      double v = P.ev[e];
      if (v == std::floor(v) && std::fabs(v) < 1e15) fprintf(f, "%d %d %.1f\n", P.ei[e] + 1, P.ej[e] + 1, v);
      else fprintf(f, "%d %d %.17g\n", P.ei[e] + 1, P.ej[e] + 1, v);
    }
    fclose(f);
  }
  if (ord) {
    FILE *f = fopen(ord, "w");
    if (!f) return err = std::string("cannot write ") + ord, -1;
    fprintf(f, "%d %d\n", P.levels, P.N);
    for (int label = 1; label <= P.N; label++) {
      int h = P.heap_of(label);
      fprintf(f, "%d;", label - 1);
      for (int i = 0; i < P.sz[h]; i++) fprintf(f, "%d,", P.perm[P.start[h] + i]);
      fprintf(f, "\n");
    }
    fclose(f);
  }
  if (clust) {
    FILE *f = fopen(clust, "w");
    if (!f) return err = std::string("cannot write ") + clust, -1;
    fprintf(f, "%d %d\n", P.levels, P.N);
    for (int label = 1; label <= P.N; label++) {
      int h = P.heap_of(label);
      fprintf(f, "%d;", label - 1);
      for (auto &l : P.iv[h]) {
        for (int v : l) fprintf(f, "%d,", v);
        fprintf(f, ";");
      }
      fprintf(f, "\n");
    }
    fclose(f);
  }
  return 0;
}

}  // namespace chb
