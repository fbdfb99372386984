/*
 * chol_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see chol_oracle.h).
 *
 * A plain-C restatement of the reference's algorithm for the numeric-factorization path:
 *   readers            mnd.c:22-229, mmat.rg:76-100
 *   separator tree     mmat.rg:834-849
 *   block bounds       mmat.rg:299-362
 *   cluster bounds     mmat.rg:364-451 (interval composition 405-410, 417-422)
 *   allocated sets     mmat.rg:697-767
 *   assembly           mmat.rg:501-633 (hash probe `search` + fill_block)
 *   symbolic fill      mmat.rg:896-1028, coarsening 635-695
 *   numeric level loop mmat.rg:1211-1358, leaf tasks blas.rg:63-504
 *   solve              mmat.rg:1364-1495
 *   writers            mmat.rg:102-147, 785-798
 * The arithmetic itself lives in a third-party library the reference resolves by name at
 * JIT time ("libcblas.so"/"liblapacke.so", blas.rg:18-22; OpenBLAS per mmat.rg:1057), not
 * vendored and not version pinned.  Here it is the scipy-bundled OpenBLAS, dlopen'd.
 *
 * One deliberate difference from the reference, stated once: the reference gives every
 * (ancestor, descendant) block a dense Legion instance (ld = rows of the row separator,
 * cholesky.cc:65-73).  That is O(n * sum of ancestor sizes) memory and cannot reach the
 * BASELINE sizes, so the oracle stores, per block, only the rows that belong to a cluster
 * that is structurally filled when the block's column separator is eliminated (a superset
 * of every earlier filled cluster, because coarsening only merges).  All other entries
 * are exact zeros in the reference too.  BLAS receives (pointer, ld) of that storage with
 * the same m/n/k, so results are the reference's results.
 */
#define _GNU_SOURCE
#include "chol_oracle.h"
#include <ctype.h>
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ host BLAS (dlopen) */
enum { ColMajor = 102, NoTrans = 111, Trans = 112, Lower = 122, NonUnit = 131, Right = 142 };
typedef int (*potrf_fn)(int, char, int, double *, int);
typedef void (*trsm_fn)(int, int, int, int, int, int, int, double, const double *, int, double *, int);
typedef void (*syrk_fn)(int, int, int, int, int, double, const double *, int, double, double *, int);
typedef void (*gemm_fn)(int, int, int, int, int, int, double, const double *, int, const double *, int, double,
                        double *, int);
typedef void (*trsv_fn)(int, int, int, int, int, const double *, int, double *, int);
typedef void (*gemv_fn)(int, int, int, int, double, const double *, int, const double *, int, double, double *, int);
typedef void (*setthr_fn)(int);
typedef char *(*config_fn)(void);
static struct {
  void *h;
  potrf_fn potrf;
  trsm_fn trsm;
  syrk_fn syrk;
  gemm_fn gemm;
  trsv_fn trsv;
  gemv_fn gemv;
  setthr_fn set_threads;
  config_fn config;
} B;

static void *sym2(void *h, const char *name) {
  char buf[128];
  snprintf(buf, sizeof buf, "scipy_%s", name);
  void *p = dlsym(h, buf);
  if (!p) p = dlsym(h, name);
  return p;
}

int orc_set_blas(const char *libpath) {
  void *h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    fprintf(stderr, "orc_set_blas: %s\n", dlerror());
    return -1;
  }
  B.h = h;
  B.potrf = (potrf_fn)sym2(h, "LAPACKE_dpotrf");
  B.trsm = (trsm_fn)sym2(h, "cblas_dtrsm");
  B.syrk = (syrk_fn)sym2(h, "cblas_dsyrk");
  B.gemm = (gemm_fn)sym2(h, "cblas_dgemm");
  B.trsv = (trsv_fn)sym2(h, "cblas_dtrsv");
  B.gemv = (gemv_fn)sym2(h, "cblas_dgemv");
  B.set_threads = (setthr_fn)sym2(h, "openblas_set_num_threads");
  B.config = (config_fn)sym2(h, "openblas_get_config");
  if (!B.potrf || !B.trsm || !B.syrk || !B.gemm || !B.trsv || !B.gemv) return -2;
  return 0;
}
const char *orc_blas_config(void) { return (B.config ? B.config() : "unknown"); }

/* ------------------------------------------------------------------ hash (uthash.h:602-610) */
uint64_t orc_hash_sax(uint64_t key) {
  const unsigned char *k = (const unsigned char *)&key;
  uint64_t h = 0;
  for (unsigned i = 0; i < sizeof(uint64_t); i++) h ^= (h << 5) + (h >> 2) + k[i];
  return h;
}

/* ------------------------------------------------------------------ data */
typedef struct {
  int32_t b, z, lox, loy, hix, hiy;
} rec_t;

struct orc {
  char err[512];
  /* matrix (mnd.c:152-199): entries + the reference's open-addressing table */
  int n, ncols, nz;
  char typecode[4];
  int64_t *ei, *ej; /* 0-based as read */
  double *ev;
  uint64_t hsize; /* ceil(nz/0.75) */
  int64_t *hi, *hj;
  double *hv;
  /* separators (mnd.c:22-69): dofs in file order == permuted order */
  int levels, N;
  int *sepdof;   /* [n] permuted row -> original dof */
  int *seplabel; /* [n] permuted row -> separator label */
  int *sz, *start; /* by heap index 1..N */
  int max_int_size;
  /* clusters (mnd.c:71-150): raw interval lists by heap index */
  int *niv;   /* number of intervals */
  int ***iv;  /* iv[h][k] raw list */
  int **ivn;  /* ivn[h][k] length */
  int ***cb;  /* composed boundaries cb[h][k][j] in local dof positions */
  /* blocks */
  int64_t nblocks;
  int64_t *boff; /* by heap index of the column separator; block = boff[hc] + distance */
  uint8_t **flag; /* per block, nc0(r)*nc0(c) bytes, 0 = FILLED, 1 = empty (mmat.rg:615) */
  int *cbiv;      /* interval at which the block's cluster bounds were last computed */
  int analyzed;
  /* snapshots F[t] (mmat.rg:1000-1016) */
  int64_t *nrec;   /* [levels] */
  rec_t **rec;     /* [levels][nrec] sorted by (b,z) */
  int64_t **rptr;  /* [levels][nblocks+1] */
  /* storage */
  int64_t *doff;  /* per block offset into data */
  int *ld;        /* per block stored rows */
  int *nseg;      /* per block */
  int64_t *sptr;  /* per block start in seg arrays */
  int *seg_lo, *seg_len, *seg_off;
  int64_t ndata;
  double *data;
  /* flop accounting */
  double *f_potrf, *f_trsm, *f_syrk, *f_gemm;
  int64_t calls[4];
  int literal;
};

static int fail(orc_t *o, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(o->err, sizeof o->err, fmt, ap);
  va_end(ap);
  return -1;
}
const char *orc_last_error(orc_t *o) { return o->err; }
orc_t *orc_create(void) { return (orc_t *)calloc(1, sizeof(orc_t)); }

static int ilog2(int x) {
  int l = 0;
  while (x > 1) {
    x >>= 1;
    l++;
  }
  return l;
}
#define LEVEL_OF(h) ilog2(h)
#define LABEL_OF(o, h) ((o)->N + 1 - (h))
#define HEAP_OF(o, label) ((o)->N + 1 - (label))
static inline int nc_of(orc_t *o, int h, int k) { return o->ivn[h][k] - 1; }
static inline int64_t blk(orc_t *o, int hr, int hc) { return o->boff[hc] + (LEVEL_OF(hc) - LEVEL_OF(hr)); }

void orc_destroy(orc_t *o) {
  if (!o) return;
  free(o->ei), free(o->ej), free(o->ev), free(o->hi), free(o->hj), free(o->hv);
  free(o->sepdof), free(o->seplabel), free(o->sz), free(o->start);
  if (o->iv)
    for (int h = 1; h <= o->N; h++) {
      for (int k = 0; k < o->niv[h]; k++) {
        free(o->iv[h][k]);
        if (o->cb && o->cb[h]) free(o->cb[h][k]);
      }
      free(o->iv[h]), free(o->ivn[h]);
      if (o->cb) free(o->cb[h]);
    }
  free(o->iv), free(o->ivn), free(o->cb), free(o->niv);
  if (o->flag)
    for (int64_t b = 0; b < o->nblocks; b++) free(o->flag[b]);
  free(o->flag), free(o->cbiv), free(o->boff);
  if (o->rec)
    for (int t = 0; t < o->levels; t++) free(o->rec[t]), free(o->rptr[t]);
  free(o->rec), free(o->rptr), free(o->nrec);
  free(o->doff), free(o->ld), free(o->nseg), free(o->sptr), free(o->seg_lo), free(o->seg_len), free(o->seg_off);
  free(o->data);
  free(o->f_potrf), free(o->f_trsm), free(o->f_syrk), free(o->f_gemm);
  free(o);
}

/* ------------------------------------------------------------------ readers */
/* mmat.rg:76-100 -> mmio.c:96-179 (banner) and 189-217 (sizes). */
static int read_banner(orc_t *o, const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) return fail(o, "cannot open %s", path);
  char line[1025], banner[64], mtx[64], crd[64], dt[64], ss[64];
  if (!fgets(line, sizeof line, f) || sscanf(line, "%63s %63s %63s %63s %63s", banner, mtx, crd, dt, ss) != 5) {
    fclose(f);
    return fail(o, "Unable to read banner.");
  }
  for (char *p = mtx; *p; p++) *p = tolower(*p);
  for (char *p = crd; *p; p++) *p = tolower(*p);
  for (char *p = dt; *p; p++) *p = tolower(*p);
  for (char *p = ss; *p; p++) *p = tolower(*p);
  if (strncmp(banner, "%%MatrixMarket", 14) || strcmp(mtx, "matrix")) {
    fclose(f);
    return fail(o, "Unable to read banner.");
  }
  o->typecode[0] = 'M';
  o->typecode[1] = !strcmp(crd, "coordinate") ? 'C' : 'A';
  o->typecode[2] = !strcmp(dt, "real") ? 'R' : !strcmp(dt, "integer") ? 'I' : !strcmp(dt, "complex") ? 'C' : 'P';
  o->typecode[3] = !strcmp(ss, "general") ? 'G' : !strcmp(ss, "symmetric") ? 'S' : !strcmp(ss, "hermitian") ? 'H' : 'K';
  do {
    if (!fgets(line, sizeof line, f)) {
      fclose(f);
      return fail(o, "Unable to read matrix size.");
    }
  } while (line[0] == '%');
  if (sscanf(line, "%d %d %d", &o->n, &o->ncols, &o->nz) != 3) {
    fclose(f);
    return fail(o, "Unable to read matrix size.");
  }
  fclose(f);
  return 0;
}

/* mnd.c:152-199: skip exactly two lines, read nz "i j val" triples, insert into an
 * open-addressing table of size ceil(nz/0.75) at hash_sax(i*cols+j), linear probing,
 * "empty" meaning val == 0. */
static int read_matrix(orc_t *o, const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) return fail(o, "cannot open %s", path);
  char buff[1024];
  if (!fgets(buff, sizeof buff, f) || !fgets(buff, sizeof buff, f)) {
    fclose(f);
    return fail(o, "short matrix file");
  }
  o->ei = malloc(sizeof(int64_t) * (size_t)o->nz);
  o->ej = malloc(sizeof(int64_t) * (size_t)o->nz);
  o->ev = malloc(sizeof(double) * (size_t)o->nz);
  o->hsize = (uint64_t)ceil(o->nz / 0.75);
  o->hi = malloc(sizeof(int64_t) * o->hsize);
  o->hj = malloc(sizeof(int64_t) * o->hsize);
  o->hv = calloc(o->hsize, sizeof(double));
  for (uint64_t p = 0; p < o->hsize; p++) o->hi[p] = o->hj[p] = -1;
  for (int e = 0; e < o->nz; e++) {
    unsigned long i = 0, j = 0;
    double v = 0.0;
    if (fscanf(f, "%lu %lu %lg\n", &i, &j, &v) != 3) {
      fclose(f);
      return fail(o, "bad matrix entry %d", e);
    }
    i -= 1, j -= 1;
    o->ei[e] = (int64_t)i, o->ej[e] = (int64_t)j, o->ev[e] = v;
    uint64_t p = orc_hash_sax((uint64_t)i * (uint64_t)o->ncols + j) % o->hsize;
    while (o->hv[p] != 0) p = (p + 1) % o->hsize;
    o->hi[p] = (int64_t)i, o->hj[p] = (int64_t)j, o->hv[p] = v;
  }
  fclose(f);
  return 0;
}

/* mmat.rg:501-527 `search`: probe for (row, col), row >= col. */
static double search(orc_t *o, int64_t r, int64_t c) {
  uint64_t k = orc_hash_sax((uint64_t)r * (uint64_t)o->ncols + (uint64_t)c) % o->hsize;
  if (o->hi[k] == r && o->hj[k] == c) return o->hv[k];
  while (o->hv[k] != 0) {
    k = (k + 1) % o->hsize;
    if (o->hi[k] == r && o->hj[k] == c) return o->hv[k];
  }
  return 0.0;
}

/* mnd.c:22-69. First line "levels nsep" parsed with atoi(&line[0]), atoi(&line[2]).
 * Every further line "id;d0,d1,...,dk,": label = id+1, dofs appended in file order. */
static int read_separators(orc_t *o, const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) return fail(o, "cannot open %s", path);
  char *line = NULL;
  size_t cap = 0;
  ssize_t rd;
  int i = 0, pos = 0;
  o->sepdof = malloc(sizeof(int) * (size_t)o->n);
  o->seplabel = malloc(sizeof(int) * (size_t)o->n);
  int prev_label = 0;
  while ((rd = getline(&line, &cap, f)) != -1) {
    if (i == 0) {
      o->levels = atoi(&line[0]);
      o->N = atoi(&line[2]);
      o->sz = calloc((size_t)o->N + 2, sizeof(int));
      o->start = calloc((size_t)o->N + 2, sizeof(int));
      i++;
      continue;
    }
    char *save = NULL;
    char *rows = strtok_r(line, ";", &save);
    if (!rows) break;
    int label = atoi(rows) + 1;
    if (label < 1 || label > o->N || label < prev_label) {
      free(line), fclose(f);
      return fail(o, "separator file: ids must be ascending in 0..%d (got %d)", o->N - 1, label - 1);
    }
    prev_label = label;
    rows = strtok_r(NULL, ",", &save);
    while (rows != NULL) {
      if (isspace((unsigned char)*rows)) break; /* the trailing "\n" token */
      if (pos >= o->n) {
        free(line), fclose(f);
        return fail(o, "separator file lists more than %d dofs", o->n);
      }
      o->sepdof[pos] = atoi(rows);
      o->seplabel[pos] = label;
      o->sz[HEAP_OF(o, label)]++;
      pos++;
      rows = strtok_r(NULL, ",", &save);
    }
    i++;
  }
  free(line);
  fclose(f);
  if (o->N != (1 << o->levels) - 1) return fail(o, "num_separators %d != 2^%d-1", o->N, o->levels);
  if (pos != o->n) return fail(o, "separator file lists %d dofs, matrix has %d", pos, o->n);
  /* permuted offsets: ascending label (mmat.rg:315-339 allocates from the bottom-right
   * corner backwards in heap order, which is the same thing). */
  int acc = 0;
  for (int label = 1; label <= o->N; label++) {
    o->start[HEAP_OF(o, label)] = acc;
    acc += o->sz[HEAP_OF(o, label)];
  }
  return 0;
}

/* mnd.c:71-150, token for token: delimiters ",; " ; a token equal to "0" after the first
 * opens a new interval; the last token of a line (the "\n") is counted but never stored. */
static int read_clusters(orc_t *o, const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) return fail(o, "cannot open %s", path);
  char *line = NULL;
  size_t cap = 0;
  ssize_t rd;
  int i = 0;
  o->max_int_size = -1;
  o->niv = calloc((size_t)o->N + 2, sizeof(int));
  o->iv = calloc((size_t)o->N + 2, sizeof(int **));
  o->ivn = calloc((size_t)o->N + 2, sizeof(int *));
  while ((rd = getline(&line, &cap, f)) != -1) {
    if (i == 0) {
      i++;
      continue;
    }
    char *save = NULL;
    char *rows = strtok_r(line, "; ", &save);
    if (!rows) break;
    int label = atoi(rows) + 1;
    if (label < 1 || label > o->N) {
      free(line), fclose(f);
      return fail(o, "cluster file: bad id %d", label - 1);
    }
    int h = HEAP_OF(o, label);
    int maxiv = o->levels + 1;
    o->iv[h] = calloc((size_t)maxiv, sizeof(int *));
    o->ivn[h] = calloc((size_t)maxiv, sizeof(int));
    int interval = 0, dofs = 0;
    size_t ccap = 16;
    int *cur = malloc(sizeof(int) * ccap);
    int ncur = 0;
    rows = strtok_r(NULL, ",; ", &save);
    while (rows != NULL) {
      int row = atoi(rows);
      dofs++;
      rows = strtok_r(NULL, ",; ", &save);
      if (rows == NULL) {
        if (dofs > o->max_int_size) o->max_int_size = dofs;
      } else {
        if ((size_t)ncur == ccap) cur = realloc(cur, sizeof(int) * (ccap *= 2));
        cur[ncur++] = row;
        if (strcmp("0", rows) == 0) {
          if (dofs > o->max_int_size) o->max_int_size = dofs;
          if (interval >= maxiv) {
            free(line), fclose(f);
            return fail(o, "cluster file: too many intervals for id %d", label - 1);
          }
          o->iv[h][interval] = cur, o->ivn[h][interval] = ncur;
          interval++;
          dofs = 0;
          ccap = 16, cur = malloc(sizeof(int) * ccap), ncur = 0;
        }
      }
    }
    if (ncur > 0) {
      o->iv[h][interval] = cur, o->ivn[h][interval] = ncur;
      interval++;
    } else
      free(cur);
    o->niv[h] = interval;
    i++;
  }
  free(line);
  fclose(f);
  for (int h = 1; h <= o->N; h++)
    if (o->niv[h] < 1) return fail(o, "cluster file: separator id %d missing", LABEL_OF(o, h) - 1);
  return 0;
}

int orc_load(orc_t *o, const char *mtx, const char *ord, const char *clust) {
  if (read_banner(o, mtx)) return -1;
  if (read_separators(o, ord)) return -1;
  if (read_clusters(o, clust)) return -1;
  if (read_matrix(o, mtx)) return -1;
  return 0;
}

/* mnd.c:201-229: skip three lines, then n values. */
int orc_read_vector(const char *path, int n, double *out) {
  FILE *f = fopen(path, "r");
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
/* mmat.rg:785-798 */
int orc_write_solution(const char *path, int n, const double *x) {
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  for (int i = 0; i < n; i++) fprintf(f, "%0.8g\n", x[i]);
  fclose(f);
  return 0;
}

/* ------------------------------------------------------------------ symbolic */
/* Composed cluster boundaries (mmat.rg:400-422): interval k lists index interval k-1. */
static int compose_clusters(orc_t *o) {
  o->cb = calloc((size_t)o->N + 2, sizeof(int **));
  for (int h = 1; h <= o->N; h++) {
    o->cb[h] = calloc((size_t)o->niv[h], sizeof(int *));
    for (int k = 0; k < o->niv[h]; k++) {
      int len = o->ivn[h][k];
      o->cb[h][k] = malloc(sizeof(int) * (size_t)len);
      for (int j = 0; j < len; j++) {
        int v = o->iv[h][k][j];
        for (int i = k - 1; i >= 0; i--) {
          if (v < 0 || v >= o->ivn[h][i]) return fail(o, "cluster interval %d of id %d indexes out of range", k, LABEL_OF(o, h) - 1);
          v = o->iv[h][i][v];
        }
        o->cb[h][k][j] = v;
      }
      if (len < 1 || o->cb[h][k][0] != 0 || o->cb[h][k][len - 1] != o->sz[h])
        if (!(o->sz[h] == 0))
          return fail(o, "cluster interval %d of id %d does not span the separator (size %d)", k, LABEL_OF(o, h) - 1, o->sz[h]);
    }
  }
  return 0;
}

/* interval index used while eliminating tree level lvl (mmat.rg:1350-1354, 1018-1026) */
static inline int interval_of_level(orc_t *o, int lvl) {
  int k = o->levels - 2 - lvl;
  return k < 0 ? 0 : k;
}
static inline int has_interval(orc_t *o, int h, int k) { return k < o->niv[h]; }

static void cluster_rect(orc_t *o, int hr, int hc, int k, int z, rec_t *r) {
  int ncc = nc_of(o, hc, k);
  int rc = z / ncc, cc = z % ncc;
  r->lox = o->start[hr] + o->cb[hr][k][rc];
  r->hix = o->start[hr] + o->cb[hr][k][rc + 1] - 1;
  r->loy = o->start[hc] + o->cb[hc][k][cc];
  r->hiy = o->start[hc] + o->cb[hc][k][cc + 1] - 1;
}

static int cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return x < y ? -1 : x > y;
}

/* which interval-0 cluster of separator h holds local position p */
static int find_cluster0(orc_t *o, int h, int p) {
  int *b = o->cb[h][0];
  int lo = 0, hi = nc_of(o, h, 0); /* b[lo] <= p < b[hi] */
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2;
    if (b[mid] <= p) lo = mid;
    else hi = mid;
  }
  return lo;
}

static int is_ancestor_or_self(int hr, int hc) {
  int d = LEVEL_OF(hc) - LEVEL_OF(hr);
  return d >= 0 && (hc >> d) == hr;
}

/* Assembly flags (fill_block, mmat.rg:561-627): interval-0 cluster z of a block is FILLED
 * iff it received a nonzero.  Two routes to the same flags. */
static int assembly_flags(orc_t *o) {
  int *iperm = malloc(sizeof(int) * (size_t)o->n);
  for (int p = 0; p < o->n; p++) iperm[o->sepdof[p]] = p;
  if (o->literal) {
    /* literally: every dense entry of every allocated block probes the hash table */
    for (int hc = 1; hc <= o->N; hc++)
      for (int hr = hc; hr >= 1; hr >>= 1) {
        int64_t b = blk(o, hr, hc);
        int ncr = nc_of(o, hr, 0), ncc = nc_of(o, hc, 0);
        for (int col = 0; col < ncc; col++)
          for (int row = 0; row < ncr; row++) {
            int nnz = 0;
            for (int i = o->cb[hr][0][row]; i < o->cb[hr][0][row + 1]; i++)
              for (int j = o->cb[hc][0][col]; j < o->cb[hc][0][col + 1]; j++) {
                int64_t idxi = o->sepdof[o->start[hr] + i], idxj = o->sepdof[o->start[hc] + j];
                if (idxj > idxi) {
                  int64_t t = idxi;
                  idxi = idxj, idxj = t;
                }
                double val = search(o, idxi, idxj);
                int gx = o->start[hr] + i, gy = o->start[hc] + j;
                if (hr == hc ? (gy <= gx && val != 0.0) : (val != 0.0)) nnz++;
              }
            if (nnz > 0) o->flag[b][row * ncc + col] = 0;
          }
      }
  } else {
    for (int e = 0; e < o->nz; e++) {
      if (o->ev[e] == 0.0) continue;
      int pi = iperm[o->ei[e]], pj = iperm[o->ej[e]];
      if (pi < pj) {
        int t = pi;
        pi = pj, pj = t;
      }
      int hr = HEAP_OF(o, o->seplabel[pi]), hc = HEAP_OF(o, o->seplabel[pj]);
      if (!is_ancestor_or_self(hr, hc)) continue; /* no block: silently dropped (mmat.rg:1191) */
      int rc = find_cluster0(o, hr, pi - o->start[hr]), cc = find_cluster0(o, hc, pj - o->start[hc]);
      o->flag[blk(o, hr, hc)][rc * nc_of(o, hc, 0) + cc] = 0;
    }
  }
  free(iperm);
  return 0;
}

/* merge_filled_clusters, mmat.rg:635-695 */
static void merge_filled(orc_t *o, int k) {
  for (int hc = 1; hc <= o->N; hc++)
    for (int hr = hc; hr >= 1; hr >>= 1) {
      int64_t b = blk(o, hr, hc);
      size_t tot = (size_t)nc_of(o, hr, 0) * (size_t)nc_of(o, hc, 0);
      if (!has_interval(o, hr, k) || !has_interval(o, hc, k)) {
        memset(o->flag[b], 1, tot);
        continue;
      }
      int pr = nc_of(o, hr, k - 1), pc = nc_of(o, hc, k - 1);
      int nr = nc_of(o, hr, k), ncn = nc_of(o, hc, k);
      uint8_t *old = malloc((size_t)pr * pc);
      memcpy(old, o->flag[b], (size_t)pr * pc);
      memset(o->flag[b], 1, tot);
      for (int row = 0; row < nr; row++) {
        int top = o->iv[hr][k][row], bottom = o->iv[hr][k][row + 1];
        for (int col = 0; col < ncn; col++) {
          int left = o->iv[hc][k][col], right = o->iv[hc][k][col + 1];
          for (int i = top; i < bottom; i++)
            for (int j = left; j < right; j++)
              if (old[(size_t)i * pc + j] == 0) o->flag[b][(size_t)row * ncn + col] = 0;
        }
      }
      free(old);
    }
}

static int cmp_rec(const void *a, const void *b) {
  const rec_t *x = a, *y = b;
  if (x->b != y->b) return x->b < y->b ? -1 : 1;
  return x->z < y->z ? -1 : x->z > y->z;
}

/* compute_filled_clusters, mmat.rg:896-1028 */
static int symbolic(orc_t *o) {
  int L = o->levels;
  o->nrec = calloc((size_t)L, sizeof(int64_t));
  o->rec = calloc((size_t)L, sizeof(rec_t *));
  o->rptr = calloc((size_t)L, sizeof(int64_t *));
  int interval = 0, interval_lbl = 0;
  for (int lvl = L - 1; lvl >= 0; lvl--) {
    /* partition_separators(depth = lvl): recompute cluster bounds of every block whose two
     * separators are on levels <= lvl (mmat.rg:453-499) */
    for (int hc = 1; hc < (1 << (lvl + 1)); hc++)
      for (int hr = hc; hr >= 1; hr >>= 1) {
        if (!has_interval(o, hr, interval) || !has_interval(o, hc, interval))
          return fail(o, "separator lacks interval %d needed at level %d", interval, lvl);
        o->cbiv[blk(o, hr, hc)] = interval;
      }
    /* fill propagation (mmat.rg:926-998) */
    for (int hs = (1 << lvl); hs < (1 << (lvl + 1)); hs++) {
      if (nc_of(o, hs, interval) != 1 && o->sz[hs] > 0)
        return fail(o, "separator id %d has %d clusters when eliminated (must be 1)", LABEL_OF(o, hs) - 1, nc_of(o, hs, interval));
      for (int hp = hs >> 1; hp >= 1; hp >>= 1)
        for (int hg = hp; hg >= 1; hg >>= 1) {
          int col_cluster_size = nc_of(o, hp, interval);
          int A_clusters = nc_of(o, hg, interval) * nc_of(o, hs, interval);
          int B_clusters = nc_of(o, hp, interval) * nc_of(o, hs, interval);
          uint8_t *fa = o->flag[blk(o, hg, hs)], *fb = o->flag[blk(o, hp, hs)], *fc = o->flag[blk(o, hg, hp)];
          for (int i = 0; i < A_clusters; i++) {
            if (fa[i] != 0) continue;
            for (int j = 0; j < B_clusters; j++) {
              if (fb[j] != 0) continue;
              if (hg == hp && j > i) continue;
              fc[(size_t)i * col_cluster_size + j] = 0;
            }
          }
        }
    }
    /* snapshot (mmat.rg:1000-1016) */
    int64_t cnt = 0;
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 1) {
        o->rec[interval_lbl] = malloc(sizeof(rec_t) * (size_t)(cnt ? cnt : 1));
        o->nrec[interval_lbl] = cnt;
        cnt = 0;
      }
      for (int hc = 1; hc <= o->N; hc++)
        for (int hr = hc; hr >= 1; hr >>= 1) {
          int64_t b = blk(o, hr, hc);
          int k = o->cbiv[b];
          if (k < 0) continue;
          int64_t tot = (int64_t)nc_of(o, hr, k) * nc_of(o, hc, k);
          for (int64_t z = 0; z < tot; z++)
            if (o->flag[b][z] == 0) {
              if (pass == 1) {
                rec_t *r = &o->rec[interval_lbl][cnt];
                r->b = (int32_t)b, r->z = (int32_t)z;
                cluster_rect(o, hr, hc, k, (int)z, r);
              }
              cnt++;
            }
        }
    }
    qsort(o->rec[interval_lbl], (size_t)cnt, sizeof(rec_t), cmp_rec);
    int64_t *rp = calloc((size_t)o->nblocks + 1, sizeof(int64_t));
    for (int64_t i = 0; i < cnt; i++) rp[o->rec[interval_lbl][i].b + 1]++;
    for (int64_t b = 0; b < o->nblocks; b++) rp[b + 1] += rp[b];
    o->rptr[interval_lbl] = rp;
    interval_lbl++;
    if (lvl <= L - 2) {
      interval++;
      if (interval < L) {
        /* blocks that lose their partition also lose their (stale) bounds' relevance */
        merge_filled(o, interval);
      }
    }
  }
  return 0;
}

/* storage: rows of clusters filled when the block's column separator is eliminated */
static int allocate_storage(orc_t *o) {
  int L = o->levels;
  o->doff = calloc((size_t)o->nblocks, sizeof(int64_t));
  o->ld = calloc((size_t)o->nblocks, sizeof(int));
  o->nseg = calloc((size_t)o->nblocks, sizeof(int));
  o->sptr = calloc((size_t)o->nblocks + 1, sizeof(int64_t));
  int64_t nseg_total = 0;
  for (int hc = 1; hc <= o->N; hc++) {
    int t = L - 1 - LEVEL_OF(hc);
    for (int hr = hc; hr >= 1; hr >>= 1) {
      int64_t b = blk(o, hr, hc);
      o->nseg[b] = (hr == hc) ? 1 : (int)(o->rptr[t][b + 1] - o->rptr[t][b]);
      nseg_total += o->nseg[b];
    }
  }
  o->seg_lo = malloc(sizeof(int) * (size_t)(nseg_total + 1));
  o->seg_len = malloc(sizeof(int) * (size_t)(nseg_total + 1));
  o->seg_off = malloc(sizeof(int) * (size_t)(nseg_total + 1));
  int64_t sp = 0, dp = 0;
  for (int64_t b = 0; b < o->nblocks; b++) {
    o->sptr[b] = sp;
    sp += o->nseg[b];
  }
  o->sptr[o->nblocks] = sp;
  for (int hc = 1; hc <= o->N; hc++) {
    int t = L - 1 - LEVEL_OF(hc);
    for (int hr = hc; hr >= 1; hr >>= 1) {
      int64_t b = blk(o, hr, hc);
      int64_t s0 = o->sptr[b];
      int rows = 0;
      if (hr == hc) {
        o->seg_lo[s0] = o->start[hr], o->seg_len[s0] = o->sz[hr], o->seg_off[s0] = 0;
        rows = o->sz[hr];
      } else {
        for (int64_t i = o->rptr[t][b], s = s0; i < o->rptr[t][b + 1]; i++, s++) {
          rec_t *r = &o->rec[t][i];
          o->seg_lo[s] = r->lox, o->seg_len[s] = r->hix - r->lox + 1, o->seg_off[s] = rows;
          rows += o->seg_len[s];
        }
      }
      o->ld[b] = rows > 0 ? rows : 1;
      o->doff[b] = dp;
      dp += (int64_t)rows * o->sz[hc];
    }
  }
  o->ndata = dp;
  o->data = calloc((size_t)(dp ? dp : 1), sizeof(double));
  if (!o->data) return fail(o, "out of memory for %lld doubles", (long long)dp);
  return 0;
}

/* pointer to global permuted (row, col) inside block b = (hr, hc); NULL when the row is not stored */
static inline double *at(orc_t *o, int64_t b, int hc, int grow, int gcol) {
  int64_t s0 = o->sptr[b];
  int lo = 0, hi = o->nseg[b];
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (o->seg_lo[s0 + mid] + o->seg_len[s0 + mid] <= grow) lo = mid + 1;
    else hi = mid;
  }
  if (lo >= o->nseg[b] || grow < o->seg_lo[s0 + lo]) return NULL;
  return o->data + o->doff[b] + (o->seg_off[s0 + lo] + (grow - o->seg_lo[s0 + lo])) + (int64_t)(gcol - o->start[hc]) * o->ld[b];
}

static int dry_run(orc_t *o);

int orc_analyze(orc_t *o, int literal_assembly) {
  if (compose_clusters(o)) return -1;
  /* allocated blocks (find_index_space_2d, mmat.rg:740-767) */
  o->boff = calloc((size_t)o->N + 2, sizeof(int64_t));
  int64_t nb = 0;
  for (int h = 1; h <= o->N; h++) {
    o->boff[h] = nb;
    nb += LEVEL_OF(h) + 1;
  }
  o->nblocks = nb;
  /* allocated interval-0 clusters (find_index_space_3d, mmat.rg:697-738), all flags "empty" */
  o->flag = calloc((size_t)nb, sizeof(uint8_t *));
  o->cbiv = malloc(sizeof(int) * (size_t)nb);
  double area = 0;
  for (int hc = 1; hc <= o->N; hc++)
    for (int hr = hc; hr >= 1; hr >>= 1) {
      int64_t b = blk(o, hr, hc);
      size_t tot = (size_t)nc_of(o, hr, 0) * (size_t)nc_of(o, hc, 0);
      o->flag[b] = malloc(tot ? tot : 1);
      memset(o->flag[b], 1, tot ? tot : 1);
      o->cbiv[b] = -1;
      area += (double)o->sz[hr] * o->sz[hc];
    }
  o->literal = literal_assembly < 0 ? (area <= 5e7) : literal_assembly;
  if (assembly_flags(o)) return -1;
  if (symbolic(o)) return -1;
  if (allocate_storage(o)) return -1;
  o->f_potrf = calloc((size_t)o->levels, sizeof(double));
  o->f_trsm = calloc((size_t)o->levels, sizeof(double));
  o->f_syrk = calloc((size_t)o->levels, sizeof(double));
  o->f_gemm = calloc((size_t)o->levels, sizeof(double));
  o->analyzed = 1;
  return dry_run(o);
}

/* (re)assembly of A's values into the blocks (fill(block,0) + fill_block, mmat.rg:1216-1224) */
int orc_assemble(orc_t *o) {
  if (!o->analyzed) return fail(o, "analyze first");
  memset(o->data, 0, sizeof(double) * (size_t)o->ndata);
  int *iperm = malloc(sizeof(int) * (size_t)o->n);
  for (int p = 0; p < o->n; p++) iperm[o->sepdof[p]] = p;
  if (o->literal) {
    for (int hc = 1; hc <= o->N; hc++)
      for (int hr = hc; hr >= 1; hr >>= 1) {
        int64_t b = blk(o, hr, hc);
        for (int64_t s = o->sptr[b]; s < o->sptr[b + 1]; s++)
          for (int gi = o->seg_lo[s]; gi < o->seg_lo[s] + o->seg_len[s]; gi++)
            for (int gj = o->start[hc]; gj < o->start[hc] + o->sz[hc]; gj++) {
              int64_t idxi = o->sepdof[gi], idxj = o->sepdof[gj];
              if (idxj > idxi) {
                int64_t t = idxi;
                idxi = idxj, idxj = t;
              }
              double val = search(o, idxi, idxj);
              if (val != 0.0 && (hr != hc || gj <= gi)) *at(o, b, hc, gi, gj) = val;
            }
      }
  } else {
    for (int e = 0; e < o->nz; e++) {
      if (o->ev[e] == 0.0) continue;
      int pi = iperm[o->ei[e]], pj = iperm[o->ej[e]];
      if (pi < pj) {
        int t = pi;
        pi = pj, pj = t;
      }
      int hr = HEAP_OF(o, o->seplabel[pi]), hc = HEAP_OF(o, o->seplabel[pj]);
      if (!is_ancestor_or_self(hr, hc)) continue;
      double *p = at(o, blk(o, hr, hc), hc, pi, pj);
      if (!p) {
        free(iperm);
        return fail(o, "internal: nonzero (%d,%d) outside the filled pattern", pi, pj);
      }
      *p = o->ev[e];
    }
  }
  free(iperm);
  return 0;
}

/* ------------------------------------------------------------------ numeric leaf tasks */
typedef struct {
  orc_t *o;
  int dry, lvl, t;
  double fl[4];
  int64_t calls[4];
  FILE *log; /* debug trace (`-d`): the fused tasks print one line per BLAS call, blas.rg:308,340,405,422,490 */
} ctx_t;

static void rec_to_filled(orc_t *o, int t, const rec_t *r, orc_filled_t *f);
/* "'X': (sx, sy, z), 'X_Lo': (..), 'X_Hi': (..), '<size>X': (..), " -- one operand of a debug line */
static void log_operand(ctx_t *c, const char *name, const char *size_key, const rec_t *r) {
  orc_filled_t f;
  rec_to_filled(c->o, c->t, r, &f);
  fprintf(c->log, "'%s': (%lld, %lld, %lld), '%s_Lo': (%lld, %lld), '%s_Hi': (%lld, %lld), '%s%s': (%lld, %lld), ", name,
          (long long)f.sep_x, (long long)f.sep_y, (long long)f.cluster, name, (long long)f.lo_x, (long long)f.lo_y, name,
          (long long)f.hi_x, (long long)f.hi_y, size_key, name, (long long)(f.hi_x - f.lo_x + 1), (long long)(f.hi_y - f.lo_y + 1));
}
static void log_tail(ctx_t *c, const rec_t *blockrec) {
  orc_filled_t f;
  rec_to_filled(c->o, c->t, blockrec, &f);
  fprintf(c->log, "'Block': (%lld, %lld), 'Level': %d, 'Interval': %d}\n", (long long)f.sep_x, (long long)f.sep_y, c->lvl, c->t);
}

static inline void rec_ptr(orc_t *o, const rec_t *r, int hc, double **p, int *ld) {
  *p = at(o, r->b, hc, r->lox, r->loy);
  *ld = o->ld[r->b];
}

/* fused_dpotrf, blas.rg:292-315 -> dpotrf_terra 63-76 */
static void fused_dpotrf(ctx_t *c, int hs) {
  orc_t *o = c->o;
  int64_t b = blk(o, hs, hs);
  for (int64_t i = o->rptr[c->t][b]; i < o->rptr[c->t][b + 1]; i++) {
    const rec_t *a = &o->rec[c->t][i];
    int m = a->hix - a->lox + 1;
    if (c->log) {
      fprintf(c->log, "POTRF: {");
      log_operand(c, "A", "Size", a);
      log_tail(c, a);
    }
    if (m == 0) continue;
    c->fl[0] += (double)m * m * m / 3.0 + (double)m * m / 2.0 + (double)m / 6.0;
    c->calls[0]++;
    if (c->dry) continue;
    double *A;
    int lda;
    rec_ptr(o, a, hs, &A, &lda);
    B.potrf(ColMajor, 'L', m, A, lda); /* return code ignored, as blas.rg:71 */
  }
}

/* fused_dtrsm, blas.rg:317-351 -> dtrsm_terra 88-104 */
static void fused_dtrsm(ctx_t *c, int hs, int hp) {
  orc_t *o = c->o;
  int64_t ba = blk(o, hs, hs), bb = blk(o, hp, hs);
  for (int64_t i = o->rptr[c->t][ba]; i < o->rptr[c->t][ba + 1]; i++) {
    const rec_t *a = &o->rec[c->t][i];
    for (int64_t j = o->rptr[c->t][bb]; j < o->rptr[c->t][bb + 1]; j++) {
      const rec_t *b = &o->rec[c->t][j];
      int m = b->hix - b->lox + 1, n = b->hiy - b->loy + 1;
      c->fl[1] += (double)m * n * n;
      c->calls[1]++;
      if (c->log) {
        fprintf(c->log, "TRSM: {");
        log_operand(c, "A", "Size", a);
        log_operand(c, "B", "Size", b);
        log_tail(c, b);
      }
      if (c->dry) continue;
      double *A, *Bp;
      int lda, ldb;
      rec_ptr(o, a, hs, &A, &lda);
      rec_ptr(o, b, hs, &Bp, &ldb);
      B.trsm(ColMajor, Right, Lower, Trans, NonUnit, m, n, 1.0, A, lda, Bp, ldb);
    }
  }
}

/* destination lookup: the reference scans filled_rC linearly for the colour (blas.rg:385-392) */
static const rec_t *find_rec(orc_t *o, int t, int64_t b, int z) {
  int64_t lo = o->rptr[t][b], hi = o->rptr[t][b + 1];
  while (lo < hi) {
    int64_t mid = (lo + hi) / 2;
    if (o->rec[t][mid].z < z) lo = mid + 1;
    else hi = mid;
  }
  if (lo < o->rptr[t][b + 1] && o->rec[t][lo].z == z) return &o->rec[t][lo];
  return NULL;
}

/* fused_dsyrk (blas.rg:353-436) when hg == hp, fused_dgemm (blas.rg:438-504) otherwise */
static void fused_update(ctx_t *c, int hs, int hp, int hg) {
  orc_t *o = c->o;
  int64_t bA = blk(o, hg, hs), bB = blk(o, hp, hs), bC = blk(o, hg, hp);
  int col_cluster_size = nc_of(o, hp, interval_of_level(o, c->lvl));
  for (int64_t i = o->rptr[c->t][bA]; i < o->rptr[c->t][bA + 1]; i++) {
    const rec_t *a = &o->rec[c->t][i];
    int row = a->z;
    int ax = a->hix - a->lox + 1, ay = a->hiy - a->loy + 1;
    for (int64_t j = o->rptr[c->t][bB]; j < o->rptr[c->t][bB + 1]; j++) {
      const rec_t *b = &o->rec[c->t][j];
      int col = b->z;
      int bx = b->hix - b->lox + 1;
      const rec_t *cc = find_rec(o, c->t, bC, row * col_cluster_size + col);
      if (!cc) continue;
      int cx = cc->hix - cc->lox + 1;
      if (c->log && (hg != hp || col <= row)) {
        fprintf(c->log, "GEMM: {");
        log_operand(c, "A", "size", a);
        log_operand(c, "B", "size", b);
        log_operand(c, "C", "size", cc);
        log_tail(c, cc);
      }
      double *A = NULL, *Bp = NULL, *C = NULL;
      int lda = 0, ldb = 0, ldc = 0;
      if (!c->dry) {
        rec_ptr(o, a, hs, &A, &lda);
        rec_ptr(o, b, hs, &Bp, &ldb);
        rec_ptr(o, cc, hp, &C, &ldc);
      }
      if (hg == hp) {
        if (col < row) {
          c->fl[3] += 2.0 * ax * bx * ay;
          c->calls[3]++;
          if (!c->dry) B.gemm(ColMajor, NoTrans, Trans, ax, bx, ay, -1.0, A, lda, Bp, ldb, 1.0, C, ldc);
        } else if (col == row) {
          c->fl[2] += (double)ay * cx * (cx + 1);
          c->calls[2]++;
          if (!c->dry) B.syrk(ColMajor, Lower, NoTrans, cx, ay, -1.0, A, lda, 1.0, C, ldc);
        }
      } else {
        c->fl[3] += 2.0 * ax * bx * ay;
        c->calls[3]++;
        if (!c->dry) B.gemm(ColMajor, NoTrans, Trans, ax, bx, ay, -1.0, A, lda, Bp, ldb, 1.0, C, ldc);
      }
    }
  }
}

/* ------------------------------------------------------------------ level loop + workers */
typedef struct {
  int kind; /* 0 potrf, 1 trsm, 2 update */
  int hs, hp, hg;
  int64_t key; /* tasks sharing a key touch the same written block: run in program order */
} task_t;

typedef struct {
  ctx_t base;
  task_t *tasks;
  int64_t *gstart; /* group boundaries */
  int64_t ngroups;
  volatile int64_t next;
  pthread_mutex_t mu;
} pool_t;

static void run_task(ctx_t *c, const task_t *t) {
  if (t->kind == 0) fused_dpotrf(c, t->hs);
  else if (t->kind == 1) fused_dtrsm(c, t->hs, t->hp);
  else fused_update(c, t->hs, t->hp, t->hg);
}

static void *worker(void *arg) {
  pool_t *p = arg;
  ctx_t c = p->base;
  for (;;) {
    int64_t g = __sync_fetch_and_add(&p->next, 1);
    if (g >= p->ngroups) break;
    for (int64_t i = p->gstart[g]; i < p->gstart[g + 1]; i++) run_task(&c, &p->tasks[i]);
  }
  return NULL;
}

static int cmp_task(const void *a, const void *b) {
  const task_t *x = a, *y = b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  /* program order inside a group: ascending (hs, hp desc-depth, hg) as enumerated */
  if (x->hs != y->hs) return x->hs < y->hs ? -1 : 1;
  if (x->hp != y->hp) return x->hp > y->hp ? -1 : 1;
  return x->hg > y->hg ? -1 : x->hg < y->hg;
}

static void run_phase(orc_t *o, ctx_t *c, task_t *tasks, int64_t nt, int threads) {
  if (nt == 0) return;
  if (c->dry || threads <= 1 || nt < 2 * (int64_t)threads) {
    /* few tasks (top of the tree): program order, BLAS gets all the threads */
    if (!c->dry && B.set_threads) B.set_threads(threads > 1 ? threads : 1);
    for (int64_t i = 0; i < nt; i++) run_task(c, &tasks[i]);
    return;
  }
  if (B.set_threads) B.set_threads(1); /* mmat.rg:1057 */
  qsort(tasks, (size_t)nt, sizeof(task_t), cmp_task);
  pool_t p;
  memset(&p, 0, sizeof p);
  p.base = *c;
  p.tasks = tasks;
  p.gstart = malloc(sizeof(int64_t) * (size_t)(nt + 1));
  p.ngroups = 0;
  for (int64_t i = 0; i < nt; i++)
    if (i == 0 || tasks[i].key != tasks[i - 1].key) p.gstart[p.ngroups++] = i;
  p.gstart[p.ngroups] = nt;
  p.next = 0;
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
  for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, worker, &p);
  for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
  free(th), free(p.gstart);
  (void)o;
}

/* one tree level, three phases in program order (mmat.rg:1240-1347) */
static void do_level(orc_t *o, int lvl, int dry, int threads, int phases) {
  ctx_t c;
  memset(&c, 0, sizeof c);
  c.o = o, c.dry = dry, c.lvl = lvl, c.t = o->levels - 1 - lvl;
  int first = 1 << lvl, last = (1 << (lvl + 1)) - 1;
  int64_t nsep = last - first + 1;
  if (phases & 1) {
    task_t *ts = malloc(sizeof(task_t) * (size_t)nsep);
    int64_t nt = 0;
    for (int hs = first; hs <= last; hs++) ts[nt++] = (task_t){0, hs, 0, 0, blk(o, hs, hs)};
    run_phase(o, &c, ts, nt, threads);
    free(ts);
  }
  if (phases & 2) {
    task_t *ts = malloc(sizeof(task_t) * (size_t)(nsep * (lvl + 1)));
    int64_t nt = 0;
    for (int hs = first; hs <= last; hs++)
      for (int hp = hs >> 1; hp >= 1; hp >>= 1) ts[nt++] = (task_t){1, hs, hp, 0, blk(o, hp, hs)};
    run_phase(o, &c, ts, nt, threads);
    free(ts);
  }
  if (phases & 4) {
    task_t *ts = malloc(sizeof(task_t) * (size_t)(nsep * (lvl + 1) * (lvl + 2) / 2 + 1));
    int64_t nt = 0;
    for (int hs = first; hs <= last; hs++)
      for (int hp = hs >> 1; hp >= 1; hp >>= 1)
        for (int hg = hp; hg >= 1; hg >>= 1) ts[nt++] = (task_t){2, hs, hp, hg, blk(o, hg, hp)};
    run_phase(o, &c, ts, nt, threads);
    free(ts);
  }
  if (dry) {
    o->f_potrf[lvl] += c.fl[0], o->f_trsm[lvl] += c.fl[1], o->f_syrk[lvl] += c.fl[2], o->f_gemm[lvl] += c.fl[3];
    for (int i = 0; i < 4; i++) o->calls[i] += c.calls[i];
  }
}

static int dry_run(orc_t *o) {
  for (int lvl = o->levels - 1; lvl >= 0; lvl--) do_level(o, lvl, 1, 1, 7);
  return 0;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int orc_factor_levels(orc_t *o, int threads, int from_level, int to_level, double *seconds) {
  if (!o->analyzed) return fail(o, "analyze first");
  if (!B.potrf) return fail(o, "no BLAS loaded (orc_set_blas)");
  double t0 = now_s();
  for (int lvl = from_level; lvl >= to_level; lvl--) do_level(o, lvl, 0, threads, 7);
  if (seconds) *seconds = now_s() - t0;
  return 0;
}
int orc_factor(orc_t *o, int threads, double *seconds) {
  if (orc_assemble(o)) return -1;
  return orc_factor_levels(o, threads, o->levels - 1, 0, seconds);
}
int orc_fused_dpotrf(orc_t *o, int lvl) {
  if (!B.potrf) return fail(o, "no BLAS loaded");
  do_level(o, lvl, 0, 1, 1);
  return 0;
}
int orc_fused_dtrsm(orc_t *o, int lvl) {
  if (!B.potrf) return fail(o, "no BLAS loaded");
  do_level(o, lvl, 0, 1, 2);
  return 0;
}
int orc_fused_update(orc_t *o, int lvl) {
  if (!B.potrf) return fail(o, "no BLAS loaded");
  do_level(o, lvl, 0, 1, 4);
  return 0;
}

/* ------------------------------------------------------------------ accessors */
int orc_n(orc_t *o) { return o->n; }
int orc_nz(orc_t *o) { return o->nz; }
int orc_levels(orc_t *o) { return o->levels; }
int orc_num_separators(orc_t *o) { return o->N; }
int orc_max_int_size(orc_t *o) { return o->max_int_size; }
int64_t orc_num_blocks(orc_t *o) { return o->nblocks; }
int64_t orc_num_clusters0(orc_t *o) {
  int64_t t = 0;
  for (int hc = 1; hc <= o->N; hc++)
    for (int hr = hc; hr >= 1; hr >>= 1) t += (int64_t)nc_of(o, hr, 0) * nc_of(o, hc, 0);
  return t;
}
int orc_get_perm(orc_t *o, int32_t *perm) {
  for (int p = 0; p < o->n; p++) perm[p] = o->sepdof[p];
  return 0;
}
int orc_get_sep_sizes(orc_t *o, int32_t *sizes) {
  for (int label = 1; label <= o->N; label++) sizes[label - 1] = o->sz[HEAP_OF(o, label)];
  return 0;
}
/* partition_matrix, mmat.rg:299-362 */
int64_t orc_get_block_bounds(orc_t *o, int64_t *out) {
  int64_t k = 0;
  for (int hc = 1; hc <= o->N; hc++)
    for (int hr = hc; hr >= 1; hr >>= 1) {
      if (out) {
        int64_t *r = out + 6 * k;
        r[0] = LABEL_OF(o, hr), r[1] = LABEL_OF(o, hc);
        r[2] = o->start[hr], r[3] = o->start[hc];
        r[4] = o->start[hr] + o->sz[hr] - 1, r[5] = o->start[hc] + o->sz[hc] - 1;
      }
      k++;
    }
  return k;
}
int64_t orc_num_filled(orc_t *o, int t) { return (t < 0 || t >= o->levels) ? -1 : o->nrec[t]; }

static void rec_to_filled(orc_t *o, int t, const rec_t *r, orc_filled_t *f) {
  /* recover (hr, hc) from the block id */
  int lo = 1, hi = o->N;
  while (lo < hi) {
    int mid = (lo + hi + 1) / 2;
    if (o->boff[mid] <= r->b) lo = mid;
    else hi = mid - 1;
  }
  int hc = lo, hr = hc >> (int)(r->b - o->boff[hc]);
  f->filled = 0;
  f->sep_x = LABEL_OF(o, hr), f->sep_y = LABEL_OF(o, hc);
  f->interval = t, f->cluster = r->z;
  f->lo_x = r->lox, f->lo_y = r->loy, f->hi_x = r->hix, f->hi_y = r->hiy;
}
static int cmp_filled(const void *a, const void *b) {
  const orc_filled_t *x = a, *y = b;
  if (x->sep_x != y->sep_x) return x->sep_x < y->sep_x ? -1 : 1;
  if (x->sep_y != y->sep_y) return x->sep_y < y->sep_y ? -1 : 1;
  return x->cluster < y->cluster ? -1 : x->cluster > y->cluster;
}
int64_t orc_get_filled(orc_t *o, int t, orc_filled_t *out) {
  if (t < 0 || t >= o->levels) return -1;
  for (int64_t i = 0; i < o->nrec[t]; i++) rec_to_filled(o, t, &o->rec[t][i], &out[i]);
  qsort(out, (size_t)o->nrec[t], sizeof(orc_filled_t), cmp_filled);
  return o->nrec[t];
}
static uint64_t mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
uint64_t orc_filled_checksum(orc_t *o, int t) {
  if (t < 0 || t >= o->levels) return 0;
  uint64_t sum = 0;
  for (int64_t i = 0; i < o->nrec[t]; i++) {
    orc_filled_t f;
    rec_to_filled(o, t, &o->rec[t][i], &f);
    const int64_t *w = (const int64_t *)&f;
    uint64_t h = 0;
    for (int k = 0; k < 9; k++) h = mix64(h ^ (uint64_t)w[k]);
    sum += h;
  }
  return sum;
}
int64_t orc_factor_nnz_alloc(orc_t *o) { return o->ndata; }
double orc_flops(orc_t *o) {
  double s = 0;
  for (int l = 0; l < o->levels; l++) s += o->f_potrf[l] + o->f_trsm[l] + o->f_syrk[l] + o->f_gemm[l];
  return s;
}
int orc_flops_by_level(orc_t *o, double *p, double *t, double *s, double *g) {
  for (int l = 0; l < o->levels; l++) p[l] = o->f_potrf[l], t[l] = o->f_trsm[l], s[l] = o->f_syrk[l], g[l] = o->f_gemm[l];
  return 0;
}
int orc_call_counts(orc_t *o, int64_t *c4) {
  for (int i = 0; i < 4; i++) c4[i] = o->calls[i];
  return 0;
}

/* visit stored entries block by block ((row_sep, col_sep) ascending labels), row-major inside
 * a block, as write_matrix does (mmat.rg:114-144) */
typedef void (*visit_fn)(void *u, int gi, int gj, double v);
static void visit(orc_t *o, visit_fn fn, void *u) {
  for (int lr = 1; lr <= o->N; lr++) {
    int hr = HEAP_OF(o, lr);
    /* column separators: hr itself and every descendant, ascending label */
    int lv = LEVEL_OF(hr);
    for (int lc = 1; lc <= lr; lc++) {
      int hc = HEAP_OF(o, lc);
      int d = LEVEL_OF(hc) - lv;
      if (d < 0 || (hc >> d) != hr) continue;
      int64_t b = blk(o, hr, hc);
      for (int64_t s = o->sptr[b]; s < o->sptr[b + 1]; s++)
        for (int r = 0; r < o->seg_len[s]; r++)
          for (int c = 0; c < o->sz[hc]; c++) {
            double v = o->data[o->doff[b] + o->seg_off[s] + r + (int64_t)c * o->ld[b]];
            if (v != 0) fn(u, o->seg_lo[s] + r, o->start[hc] + c, v);
          }
    }
  }
}
static void v_count(void *u, int i, int j, double v) {
  (void)i, (void)j, (void)v;
  (*(int64_t *)u)++;
}
int64_t orc_factor_nnz(orc_t *o) {
  int64_t c = 0;
  visit(o, v_count, &c);
  return c;
}
typedef struct {
  int32_t *I, *J;
  double *V;
  int64_t k;
} coo_t;
static void v_coo(void *u, int i, int j, double v) {
  coo_t *c = u;
  c->I[c->k] = i, c->J[c->k] = j, c->V[c->k] = v, c->k++;
}
int64_t orc_get_factor_coo(orc_t *o, int32_t *I, int32_t *J, double *V) {
  coo_t c = {I, J, V, 0};
  visit(o, v_coo, &c);
  return c.k;
}
typedef struct {
  double *d;
  int n;
} dense_t;
static void v_dense(void *u, int i, int j, double v) {
  dense_t *d = u;
  d->d[(size_t)i * d->n + j] = v;
}
int orc_get_factor_dense(orc_t *o, double *out) {
  memset(out, 0, sizeof(double) * (size_t)o->n * o->n);
  dense_t d = {out, o->n};
  visit(o, v_dense, &d);
  return 0;
}
typedef struct {
  FILE *f;
  int full;
} wr_t;
static void v_write(void *u, int i, int j, double v) {
  wr_t *w = u;
  fprintf(w->f, w->full ? "%d %d %.17g\n" : "%d %d %0.8g\n", i + 1, j + 1, v);
}
static const char *typecode_str(const char *tc, char *buf) {
  sprintf(buf, "%s %s %s %s", "matrix", tc[1] == 'C' ? "coordinate" : "array",
          tc[2] == 'R' ? "real" : tc[2] == 'I' ? "integer" : tc[2] == 'C' ? "complex" : "pattern",
          tc[3] == 'G' ? "general" : tc[3] == 'S' ? "symmetric" : tc[3] == 'H' ? "hermitian" : "skew-symmetric");
  return buf;
}
/* write_matrix, mmat.rg:102-147 (banner copied from the input typecode) */
int orc_write_factor(orc_t *o, const char *path, int full) {
  FILE *f = fopen(path, "w");
  if (!f) return fail(o, "cannot write %s", path);
  char buf[128];
  fprintf(f, "%s %s\n", "%%MatrixMarket", typecode_str(o->typecode, buf));
  fprintf(f, "%d %d %lld\n", o->n, o->ncols, (long long)orc_factor_nnz(o));
  wr_t w = {f, full};
  visit(o, v_write, &w);
  fclose(f);
  return 0;
}

/* ------------------------------------------------------------------ solve, mmat.rg:1364-1495 */
int orc_solve(orc_t *o, const double *b, double *x) {
  if (!B.trsv) return fail(o, "no BLAS loaded");
  if (B.set_threads) B.set_threads(1);
  int L = o->levels;
  double *Bv = malloc(sizeof(double) * (size_t)o->n);
  for (int p = 0; p < o->n; p++) Bv[p] = b[o->sepdof[p]]; /* fill_b, mmat.rg:769-783 */
  /* The reference passes whole dense blocks to dgemv; rows outside the stored pattern are
   * exact zeros there, so applying the stored row segments is the same product. */
  for (int lvl = L - 1; lvl >= 0; lvl--)
    for (int hs = 1 << lvl; hs < (1 << (lvl + 1)); hs++) {
      if (o->sz[hs] == 0) continue;
      int64_t bd = blk(o, hs, hs);
      B.trsv(ColMajor, Lower, NoTrans, NonUnit, o->sz[hs], o->data + o->doff[bd], o->ld[bd], Bv + o->start[hs], 1);
      for (int hp = hs >> 1; hp >= 1; hp >>= 1) {
        int64_t bb = blk(o, hp, hs);
        for (int64_t s = o->sptr[bb]; s < o->sptr[bb + 1]; s++)
          B.gemv(ColMajor, NoTrans, o->seg_len[s], o->sz[hs], -1.0, o->data + o->doff[bb] + o->seg_off[s], o->ld[bb],
                 Bv + o->start[hs], 1, 1.0, Bv + o->seg_lo[s], 1);
      }
    }
  for (int plvl = 0; plvl < L; plvl++)
    for (int hp = 1 << plvl; hp < (1 << (plvl + 1)); hp++) {
      if (o->sz[hp] == 0) continue;
      int64_t bd = blk(o, hp, hp);
      B.trsv(ColMajor, Lower, Trans, NonUnit, o->sz[hp], o->data + o->doff[bd], o->ld[bd], Bv + o->start[hp], 1);
      for (int lvl = plvl + 1; lvl < L; lvl++)
        for (int hs = hp << (lvl - plvl); hs < ((hp + 1) << (lvl - plvl)); hs++) {
          int64_t bb = blk(o, hp, hs);
          if (o->sz[hs] == 0) continue;
          for (int64_t s = o->sptr[bb]; s < o->sptr[bb + 1]; s++)
            B.gemv(ColMajor, Trans, o->seg_len[s], o->sz[hs], -1.0, o->data + o->doff[bb] + o->seg_off[s], o->ld[bb],
                   Bv + o->seg_lo[s], 1, 1.0, Bv + o->start[hs], 1);
        }
    }
  for (int p = 0; p < o->n; p++) x[o->sepdof[p]] = Bv[p]; /* mmat.rg:1483-1491 */
  free(Bv);
  return 0;
}

/* ------------------------------------------------------------------ debug trace, the `-d` path
 * Log grammar: partition_matrix mmat.rg:331,352; partition_separator 396,432; compute_filled_clusters
 * 1010; fused tasks blas.rg:308,340,405,422,490.  Snapshots: write_blocks (mmat.rg:174-218) after
 * every fused task of the level loop (mmat.rg:1245-1343), file names from gen_filename (149-172).
 * verify.debug_factor (verify.py:216-275) replays the log and compares each op's block with its
 * snapshot.  Serial, program order. */
static void dbg_clusters(orc_t *o, FILE *log, int hr, int hc, int k, int t) {
  int rows = has_interval(o, hr, k) ? nc_of(o, hr, k) : -1, cols = has_interval(o, hc, k) ? nc_of(o, hc, k) : -1;
  fprintf(log, "\t\tPartitioning (%d, %d) Cluster: %d Rows: %d Cols: %d\n", LABEL_OF(o, hr), LABEL_OF(o, hc), k, rows, cols);
  for (int row = 0; row < rows; row++) {
    for (int col = 0; col < cols; col++) {
      rec_t r;
      cluster_rect(o, hr, hc, k, row * cols + col, &r);
      long long sx = r.hix - r.lox + 1, sy = r.hiy - r.loy + 1;
      fprintf(log,
              "\t\tCluster: {'Block': (%d, %d), 'color': (%d, %d, %d), 'Lo': (%d, %d), 'Hi': (%d, %d), 'size': (%lld, %lld), "
              "'vol': %lld, 'Interval': %d}\n",
              LABEL_OF(o, hr), LABEL_OF(o, hc), LABEL_OF(o, hr), LABEL_OF(o, hc), row * cols + col, r.lox, r.loy, r.hix, r.hiy, sx,
              sy, (sx > 0 && sy > 0) ? sx * sy : 0LL, t);
    }
    fprintf(log, "\n");
  }
}

static int dbg_snapshot(orc_t *o, const char *dir, const char *name, int full) {
  char path[2048];
  snprintf(path, sizeof path, "%s/%s.mtx", dir, name);
  return orc_write_factor(o, path, full);
}

int orc_debug_trace(orc_t *o, const char *dir, const char *log_path, int full_precision) {
  if (!o->analyzed) return fail(o, "analyze first");
  if (!B.potrf) return fail(o, "no BLAS loaded (orc_set_blas)");
  FILE *log = fopen(log_path, "w");
  if (!log) return fail(o, "cannot write %s", log_path);
  const int L = o->levels;
  /* partition_matrix, mmat.rg:316-360 */
  for (int lvl = 0; lvl < L; lvl++)
    for (int h = 1 << lvl; h < (1 << (lvl + 1)); h++)
      for (int hr = h; hr >= 1; hr >>= 1)
        fprintf(log, "Block: {'Block': (%d, %d), 'Lo': (%d, %d), 'Hi': (%d, %d)}\n", LABEL_OF(o, hr), LABEL_OF(o, h), o->start[hr],
                o->start[h], o->start[hr] + o->sz[hr] - 1, o->start[h] + o->sz[h] - 1);
  /* compute_filled_clusters, mmat.rg:918-1026: cluster rectangles of every block down to the level
   * being eliminated, then the filled records of that interval label */
  for (int t = 0; t < L; t++) {
    const int lvl = L - 1 - t, k = interval_of_level(o, lvl);
    for (int l2 = 0; l2 <= lvl; l2++)
      for (int hr = 1 << l2; hr < (1 << (l2 + 1)); hr++)
        for (int cl = l2; cl <= lvl; cl++)
          for (int hc = hr << (cl - l2); hc < ((hr + 1) << (cl - l2)); hc++) dbg_clusters(o, log, hr, hc, k, t);
    orc_filled_t *f = malloc(sizeof(orc_filled_t) * (size_t)(o->nrec[t] + 1));
    orc_get_filled(o, t, f);
    for (int64_t i = 0; i < o->nrec[t]; i++)
      fprintf(log,
              "Fill: {'Level': %d, 'Interval': %d, 'Block': (%lld, %lld), 'Cluster': (%lld, %lld, %lld), 'Filled': 0, 'Lo': (%lld, "
              "%lld), 'Hi': (%lld, %lld), 'Size': (%lld, %lld)}\n",
              lvl, t, (long long)f[i].sep_x, (long long)f[i].sep_y, (long long)f[i].sep_x, (long long)f[i].sep_y,
              (long long)f[i].cluster, (long long)f[i].lo_x, (long long)f[i].lo_y, (long long)f[i].hi_x, (long long)f[i].hi_y,
              (long long)(f[i].hi_x - f[i].lo_x + 1), (long long)(f[i].hi_y - f[i].lo_y + 1));
    free(f);
  }
  if (orc_assemble(o)) {
    fclose(log);
    return -1;
  }
  if (B.set_threads) B.set_threads(1);
  char name[256];
  int rc = 0;
  for (int lvl = L - 1; lvl >= 0 && !rc; lvl--) {
    ctx_t c;
    memset(&c, 0, sizeof c);
    c.o = o, c.lvl = lvl, c.t = L - 1 - lvl, c.log = log;
    fprintf(log, "Factoring Level: %d Interval: %d Iteration: %d\n", lvl, interval_of_level(o, lvl), 0);
    const int first = 1 << lvl, last = (1 << (lvl + 1)) - 1;
    for (int hs = first; hs <= last && !rc; hs++) {
      fused_dpotrf(&c, hs);
      snprintf(name, sizeof name, "potrf_lvl%d_a%d%d", lvl, LABEL_OF(o, hs), LABEL_OF(o, hs));
      rc = dbg_snapshot(o, dir, name, full_precision);
    }
    for (int hs = first; hs <= last && !rc; hs++)
      for (int hp = hs >> 1; hp >= 1 && !rc; hp >>= 1) {
        fused_dtrsm(&c, hs, hp);
        snprintf(name, sizeof name, "trsm_lvl%d_a%d%d_b%d%d", lvl, LABEL_OF(o, hs), LABEL_OF(o, hs), LABEL_OF(o, hp), LABEL_OF(o, hs));
        rc = dbg_snapshot(o, dir, name, full_precision);
      }
    for (int hs = first; hs <= last && !rc; hs++)
      for (int hp = hs >> 1; hp >= 1 && !rc; hp >>= 1)
        for (int hg = hp; hg >= 1 && !rc; hg >>= 1) {
          fused_update(&c, hs, hp, hg);
          snprintf(name, sizeof name, "gemm_lvl%d_a%d%d_b%d%d_c%d%d", lvl, LABEL_OF(o, hg), LABEL_OF(o, hs), LABEL_OF(o, hp),
                   LABEL_OF(o, hs), LABEL_OF(o, hg), LABEL_OF(o, hp));
          rc = dbg_snapshot(o, dir, name, full_precision);
        }
  }
  fprintf(log, "Done factoring Iteration: %d.\n", 0);
  fclose(log);
  return rc;
}
