// C ABI of the engine (include/cholesky.h): handle, loaders, host symbolic analysis, GPU numeric
// factorization driven by the compiled level schedule, result access.
//
// A handle drives one GPU, or -- chol_create(devices, ngpu > 1) -- a group of GPUs from one process: the
// parent handle keeps the problem and the symbolic structure, one rank handle per device keeps that
// rank's schedule, buffers and streams, and every device-side call fans out over one host thread per
// rank (the ranks meet in flag words in peer memory, so their launch lists must be issued concurrently).
// The one-process-per-GPU form (chol_set_partition + chol_ipc_*) runs the same rank code with the peers'
// buffers mapped through CUDA IPC instead.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/chol_mmio.h"
#include "../../include/chol_mnd.h"
#include "../../include/cholesky.h"
#include "chol_internal.h"
#include "kernels.cuh"
#include "solve.h"
#include "solve_kernels.cuh"
#include "verify_kernels.cuh"

using namespace chb;

constexpr int kStreams = 4;  // update stream, chain stream (look-ahead), background pushes, rows stream (top panels of a partition)

struct SolveDev;
struct ResDev;
struct chol {
  std::string err;
  Problem own_P;
  Symbolic own_S;
  Problem &P;   // a rank handle of an in-process group refers to its parent's
  Symbolic &S;
  chol *parent = nullptr;
  std::vector<chol *> sub;  // in-process multi-GPU: one rank handle per device (empty otherwise)
  SolveDev *solve = nullptr;  // solve schedule and buffers, built on first use
  ResDev *res = nullptr;      // residual-check descriptors, built on first use
  int device = 0;
  Schedule D;
  bool loaded = false, analyzed = false, device_ready = false, assembled = false, factored = false;
  cudaStream_t streams[kStreams] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t &stream = streams[0];
  cudaStream_t cur = nullptr;  // stream of the launch being issued
  double *d_fac = nullptr;
  double *d_vals = nullptr;
  int64_t *d_aoff = nullptr;
  GemmProblem *d_probs = nullptr;
  GemmContrib *d_contribs = nullptr;
  TileRef *d_tiles = nullptr;
  PanelDesc *d_pdesc = nullptr;
  PanelSlab *d_pslabs = nullptr;
  TrsmDesc *d_trsm = nullptr;
  TileRef *d_trsm_tiles = nullptr;
  int *d_pflags = nullptr;  // four flag words per panel descriptor, zeroed at the start of every run of the level loop
  RectDesc *d_rects = nullptr;
  int *d_info = nullptr;  // [0] first non-positive pivot (1-based permuted column), [1] a peer wait timed out
  int64_t *d_diag_off = nullptr;
  double *d_diag = nullptr;
  double *h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  std::vector<double> h_fac;  // host copy of the factor, fetched lazily
  bool h_fac_valid = false;
  std::vector<cudaEvent_t> evs;  // cross-stream events of the launch list
  cudaEvent_t ev_fork = nullptr, ev_join[kStreams] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};  // timing events of a step
  std::vector<cudaEvent_t> kev;                      // two per launch, for the instrumented pass
  double k_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double k_gemm_flops = 0;
  std::vector<float> launch_ms;  // per launch, from the last instrumented pass
  // multi-GPU: flag words in peer memory, peers' buffers (IPC or in-process peer access)
  int rank = 0, world = 1;
  Peers peers = {};
  unsigned long long *d_flags = nullptr;
  unsigned *d_counters = nullptr;
  unsigned long long run_id = 0;  // whole-factorization runs so far: the high half of every flag value
  bool peers_ready = false;
  int num_sms = 148;
  // CUDA-graph replay of the level loop (the analogue of the reference's __demand(__trace), mmat.rg:1211): the
  // launch list of a single-GPU handle, both streams and their events, is captured once and replayed.  Pays
  // only where the step is launch-bound (512^2: 2.64 -> 2.50 ms; 128^3: 933 -> 948 ms, measured), so by
  // default (CHOL_GRAPH unset) it is used when the factorization has less than kGraphFlops flops.
  // graph_runs counts the eager runs since the schedule was uploaded: the first one stays eager (one-time
  // function attributes), the second is captured.
  int use_graph = -1, graph_runs = 0;
  cudaGraphExec_t graph_exec = nullptr;
  std::vector<void *> ipc_opened;
  chol() : P(own_P), S(own_S) {}
  explicit chol(chol *par) : P(par->P), S(par->S), parent(par) {}
};
constexpr double kGraphFlops = 2e10;

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
      return -100;                                                                                \
    }                                                                                             \
  } while (0)

static int fail(chol_t *c, const std::string &m) {
  c->err = m;
  return -1;
}

// fn(rank handle) on every rank of the handle, one host thread per rank for a group
template <typename F>
static int for_ranks(chol_t *c, F fn) {
  if (c->sub.empty()) {
    cudaSetDevice(c->device);
    return fn(c);
  }
  std::vector<int> rc(c->sub.size(), 0);
  std::vector<std::thread> th;
  for (size_t i = 0; i < c->sub.size(); i++)
    th.emplace_back([&, i] {
      cudaSetDevice(c->sub[i]->device);
      rc[i] = fn(c->sub[i]);
    });
  for (auto &t : th) t.join();
  for (size_t i = 0; i < rc.size(); i++)
    if (rc[i]) {
      c->err = "rank " + std::to_string(i) + ": " + c->sub[i]->err;
      return rc[i];
    }
  return 0;
}
static bool is_group(const chol_t *c) { return !c->sub.empty(); }

extern "C" {

void register_mappers(void) {}

int chol_create(const int *devices, int ngpu, chol_t **out) {
  if (!out) return -1;
  *out = nullptr;
  if (ngpu < 0 || ngpu > kMaxPeers || (ngpu > 1 && (ngpu & (ngpu - 1)))) return -1;  // 1, 2, 4 or 8 GPUs
  chol_t *c = new chol();
  c->device = (devices && ngpu > 0) ? devices[0] : 0;
  if (const char *e = getenv("CHOL_GRAPH")) c->use_graph = atoi(e);
  for (int r = 0; ngpu > 1 && r < ngpu; r++) {
    chol_t *s = new chol(c);
    s->device = devices[r];
    s->rank = r, s->world = ngpu;
    s->use_graph = 0;
    c->sub.push_back(s);
  }
  if (ngpu > 1) {
    c->world = ngpu;
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);  // ranks sharing a device: their streams must not share hardware queues
  }
  *out = c;
  return 0;
}
int chol_num_ranks(chol_t *c) { return is_group(c) ? (int)c->sub.size() : 1; }
chol_t *chol_rank_handle(chol_t *c, int r) {
  if (!is_group(c)) return r == 0 ? c : nullptr;
  return (r >= 0 && r < (int)c->sub.size()) ? c->sub[r] : nullptr;
}

static void free_solve(chol_t *c);
static void free_res(chol_t *c);
static void drop_graph(chol_t *c);
static void free_device(chol_t *c) {
  for (chol_t *s : c->sub) free_device(s);
  c->assembled = c->factored = c->h_fac_valid = false;
  if (!c->device_ready) return;
  cudaSetDevice(c->device);
  drop_graph(c);
  free_solve(c);
  free_res(c);
  cudaFree(c->d_fac), cudaFree(c->d_vals), cudaFree(c->d_aoff), cudaFree(c->d_probs), cudaFree(c->d_contribs);
  cudaFree(c->d_tiles), cudaFree(c->d_pdesc), cudaFree(c->d_pslabs), cudaFree(c->d_pflags), cudaFree(c->d_trsm), cudaFree(c->d_trsm_tiles), cudaFree(c->d_rects), cudaFree(c->d_info);
  cudaFree(c->d_diag_off), cudaFree(c->d_diag);
  for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
  c->ipc_opened.clear();
  cudaFree(c->d_flags), cudaFree(c->d_counters);
  c->d_flags = nullptr, c->d_counters = nullptr, c->peers_ready = false;
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  c->h_pinned = nullptr, c->h_pinned_bytes = 0;
  for (cudaEvent_t e : c->evs) cudaEventDestroy(e);
  c->evs.clear();
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  c->ev_fork = nullptr;
  for (cudaEvent_t &e : c->tev) {
    if (e) cudaEventDestroy(e);
    e = nullptr;
  }
  for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
  c->kev.clear();
  for (int i = 0; i < kStreams; i++) {
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    c->ev_join[i] = nullptr, c->streams[i] = nullptr;
  }
  c->device_ready = false;
}

void chol_destroy(chol_t *c) {
  if (!c) return;
  free_device(c);
  for (chol_t *s : c->sub) delete s;
  delete c;
}
const char *chol_last_error(chol_t *c) { return c ? c->err.c_str() : "null handle"; }

// a new problem or partition invalidates everything derived from the old one, on the host and on the device
static void invalidate(chol_t *c) {
  free_device(c);
  c->loaded = c->analyzed = false;
  for (chol_t *s : c->sub) s->loaded = s->analyzed = false;
}

int chol_load(chol_t *c, const char *mtx, const char *ord, const char *clust) {
  if (c->parent) return fail(c, "load through the group handle");
  invalidate(c);
  if (read_problem(c->P, mtx, ord, clust, c->err)) return -1;
  c->loaded = true;
  return 0;
}

int chol_load_arrays(chol_t *c, int n, int64_t nz, const int32_t *I, const int32_t *J, const double *V, int levels, int nsep,
                     const int64_t *sep_ptr, const int32_t *sep_dofs, const int64_t *sep_iv_ptr, const int64_t *iv_ptr,
                     const int32_t *iv_vals) {
  if (c->parent) return fail(c, "load through the group handle");
  invalidate(c);
  Problem &P = c->P;
  P = Problem();
  P.n = P.ncols = n, P.nz = nz;
  P.ei.assign(I, I + nz), P.ej.assign(J, J + nz), P.ev.assign(V, V + nz);
  P.levels = levels, P.N = nsep;
  if (nsep != (1 << levels) - 1) return fail(c, "num_separators != 2^levels - 1");
  P.perm.assign(sep_dofs, sep_dofs + sep_ptr[nsep]);
  if ((int)P.perm.size() != n) return fail(c, "separator lists do not cover the matrix");
  P.sz.assign(nsep + 2, 0);
  P.iv.assign(nsep + 2, {});
  int mx = -1;
  for (int id = 0; id < nsep; id++) {
    int h = P.heap_of(id + 1);
    P.sz[h] = (int)(sep_ptr[id + 1] - sep_ptr[id]);
    for (int64_t k = sep_iv_ptr[id]; k < sep_iv_ptr[id + 1]; k++) {
      P.iv[h].emplace_back(iv_vals + iv_ptr[k], iv_vals + iv_ptr[k + 1]);
      mx = std::max(mx, (int)(iv_ptr[k + 1] - iv_ptr[k]) + 1);
    }
  }
  P.max_int_size = mx;
  if (finish_problem(P, c->err)) return -1;
  c->loaded = true;
  return 0;
}

int chol_generate(chol_t *c, int nx, int ny, int nz, int stencil, int levels) {
  if (c->parent) return fail(c, "load through the group handle");
  invalidate(c);
  if (generate_problem(c->P, nx, ny, nz, stencil, levels, c->err)) return -1;
  c->loaded = true;
  return 0;
}

int chol_write_inputs(chol_t *c, const char *mtx, const char *ord, const char *clust) {
  if (!c->loaded) return fail(c, "nothing loaded");
  return write_problem(c->P, mtx, ord, clust, c->err);
}

int chol_analyze(chol_t *c, int keep_records) {
  if (c->parent) return fail(c, "analyze through the group handle");
  if (!c->loaded) return fail(c, "load a problem first");
  free_device(c);
  c->analyzed = false;
  if (analyze(c->P, c->S, keep_records != 0, c->err)) return -1;
  if (is_group(c)) {  // one symbolic analysis, one schedule per rank
    std::vector<int> rc(c->sub.size(), 0);
    std::vector<std::thread> th;
    for (size_t i = 0; i < c->sub.size(); i++)
      th.emplace_back([&, i] { rc[i] = build_schedule(c->P, c->S, c->sub[i]->D, (int)i, (int)c->sub.size(), false, c->sub[i]->err); });
    for (auto &t : th) t.join();
    for (size_t i = 0; i < rc.size(); i++) {
      if (rc[i]) return fail(c, c->sub[i]->err);
      c->sub[i]->loaded = c->sub[i]->analyzed = true;
    }
  } else if (build_schedule(c->P, c->S, c->D, c->rank, c->world, false, c->err))
    return -1;
  c->analyzed = true;
  return 0;
}

/* the symbolic analysis as a file: one rank of a node analyses and saves, the others load instead of repeating the
 * analysis (same result: the schedules built from it are identical; the file is checked against the loaded problem) */
int chol_save_analysis(chol_t *c, const char *path) {
  if (!c->analyzed) return fail(c, "analyze first");
  return save_symbolic(c->P, c->S, path, c->err) ? -1 : 0;
}
int chol_load_analysis(chol_t *c, const char *path) {
  if (c->parent) return fail(c, "analyze through the group handle");
  if (!c->loaded) return fail(c, "load a problem first");
  free_device(c);
  c->analyzed = false;
  if (load_symbolic(c->P, c->S, path, c->err)) return -1;
  if (is_group(c)) {
    for (size_t i = 0; i < c->sub.size(); i++) {
      if (build_schedule(c->P, c->S, c->sub[i]->D, (int)i, (int)c->sub.size(), false, c->sub[i]->err)) return fail(c, c->sub[i]->err);
      c->sub[i]->loaded = c->sub[i]->analyzed = true;
    }
  } else if (build_schedule(c->P, c->S, c->D, c->rank, c->world, false, c->err))
    return -1;
  c->analyzed = true;
  return 0;
}

int chol_n(chol_t *c) { return c->P.n; }
int64_t chol_nz(chol_t *c) { return c->P.nz; }
int chol_levels(chol_t *c) { return c->P.levels; }
int chol_num_separators(chol_t *c) { return c->P.N; }
int chol_max_int_size(chol_t *c) { return c->P.max_int_size; }
int64_t chol_num_blocks(chol_t *c) { return c->S.nblocks; }
int64_t chol_num_clusters0(chol_t *c) { return c->S.nclusters0; }
int chol_get_perm(chol_t *c, int32_t *perm) {
  for (int p = 0; p < c->P.n; p++) perm[p] = c->P.perm[p];
  return 0;
}
int chol_get_sep_sizes(chol_t *c, int32_t *sizes) {
  for (int label = 1; label <= c->P.N; label++) sizes[label - 1] = c->P.sz[c->P.heap_of(label)];
  return 0;
}
/* partition_matrix, mmat.rg:299-362 */
int64_t chol_get_block_bounds(chol_t *c, int64_t *out) {
  const Problem &P = c->P;
  int64_t k = 0;
  for (int hc = 1; hc <= P.N; hc++)
    for (int hr = hc; hr >= 1; hr >>= 1) {
      if (out) {
        int64_t *r = out + 6 * k;
        r[0] = P.label_of(hr), r[1] = P.label_of(hc);
        r[2] = P.start[hr], r[3] = P.start[hc];
        r[4] = P.start[hr] + P.sz[hr] - 1, r[5] = P.start[hc] + P.sz[hc] - 1;
      }
      k++;
    }
  return k;
}
int64_t chol_num_filled(chol_t *c, int t) { return (!c->analyzed || t < 0 || t >= c->P.levels) ? -1 : c->S.nfilled[t]; }
int64_t chol_get_filled(chol_t *c, int t, chol_filled_t *out) {
  if (!c->analyzed || t < 0 || t >= c->P.levels) return -1;
  if (c->S.records.empty()) return fail(c, "analyze with keep_records to read Filled records");
  const auto &r = c->S.records[t];
  static_assert(sizeof(chol_filled_t) == sizeof(FilledRec), "layout");
  memcpy(out, r.data(), r.size() * sizeof(FilledRec));
  return (int64_t)r.size();
}
uint64_t chol_filled_checksum(chol_t *c, int t) { return (!c->analyzed || t < 0 || t >= c->P.levels) ? 0 : c->S.checksum[t]; }
double chol_flops(chol_t *c) { return c->S.flops(); }
int chol_flops_by_level(chol_t *c, double *p, double *t, double *s, double *g) {
  for (int l = 0; l < c->P.levels; l++) p[l] = c->S.f_potrf[l], t[l] = c->S.f_trsm[l], s[l] = c->S.f_syrk[l], g[l] = c->S.f_gemm[l];
  return 0;
}
int chol_call_counts(chol_t *c, int64_t *c4) {
  for (int i = 0; i < 4; i++) c4[i] = c->S.calls[i];
  return 0;
}
int64_t chol_factor_doubles(chol_t *c) { return c->S.total_doubles; }

// ------------------------------------------------------------------------------ device side
}  // extern "C"
template <typename T>
static int upload(chol_t *c, T **dst, const std::vector<T> &src) {
  size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
  CK(cudaMalloc((void **)dst, bytes));
  if (!src.empty()) CK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

static void drop_graph(chol_t *c) {
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  c->graph_exec = nullptr, c->graph_runs = 0;
}
static int upload_schedule(chol_t *c) {
  drop_graph(c);  // the captured launches point into the descriptor arrays replaced below
  cudaFree(c->d_probs), cudaFree(c->d_contribs), cudaFree(c->d_tiles), cudaFree(c->d_pdesc), cudaFree(c->d_pslabs), cudaFree(c->d_pflags), cudaFree(c->d_trsm), cudaFree(c->d_trsm_tiles), cudaFree(c->d_rects);
  c->d_trsm = nullptr, c->d_trsm_tiles = nullptr;
  c->d_probs = nullptr, c->d_contribs = nullptr, c->d_tiles = nullptr, c->d_pdesc = nullptr, c->d_pslabs = nullptr, c->d_pflags = nullptr, c->d_rects = nullptr;
  if (upload(c, &c->d_probs, c->D.probs)) return -100;
  if (upload(c, &c->d_contribs, c->D.contribs)) return -100;
  if (upload(c, &c->d_tiles, c->D.tiles)) return -100;
  if (upload(c, &c->d_pdesc, c->D.pdesc)) return -100;
  if (upload(c, &c->d_pslabs, c->D.pslabs)) return -100;
  if (upload(c, &c->d_trsm, c->D.trsm)) return -100;
  if (upload(c, &c->d_trsm_tiles, c->D.trsm_tiles)) return -100;
  CK(cudaMalloc((void **)&c->d_pflags, std::max<size_t>(1, c->D.pdesc.size()) * 4 * sizeof(int)));
  if (upload(c, &c->d_rects, c->D.rects)) return -100;
  return 0;
}
// chol_factor runs the fused schedule (pivot block and off-diagonal rows advance together); the
// piecewise fused_dpotrf / fused_dtrsm entry points need the two phases as separate launch sequences
static int ensure_schedule(chol_t *c, bool split) {
  if (c->D.split_phases == split) return 0;
  CK(cudaStreamSynchronize(c->stream));
  if (build_schedule(c->P, c->S, c->D, c->rank, c->world, split, c->err)) return -1;
  return upload_schedule(c);
}

using GemmMain = GemmWsCfg<64, 64, 16, 32, 32, 4>;  // 64x64 CTA tiles, four-stage operand ring, three CTAs per SM
static int rank_device(chol_t *c) {  // one rank: streams, buffers, descriptor arrays
  if (!c->analyzed) return fail(c, "analyze first");
  if (c->device_ready) return 0;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(c, "no CUDA device: the numeric factorization runs on the GPU only (no CPU fallback)");
  if (c->device < 0 || c->device >= ndev) return fail(c, "no such CUDA device: " + std::to_string(c->device));
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device));
  {  // the chain stream gets the highest priority: its few CTAs must slip in between the CTAs of a
     // trailing update that fills the GPU, not queue behind them; the background pushes get the lowest
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->streams[0], cudaStreamNonBlocking, std::min(lo, hi + 1)));
    CK(cudaStreamCreateWithPriority(&c->streams[1], cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&c->streams[2], cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&c->streams[3], cudaStreamNonBlocking, hi));
  }
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < kStreams; i++) CK(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
  c->cur = c->stream;
  CK(cudaMalloc((void **)&c->d_fac, (size_t)c->S.total_doubles * sizeof(double)));
  c->device_ready = true;
  if (upload(c, &c->d_vals, c->P.ev)) return -100;
  if (upload(c, &c->d_aoff, c->D.a_off)) return -100;
  if (upload_schedule(c)) return -100;
  CK(cudaMalloc((void **)&c->d_info, 2 * sizeof(int)));
  std::vector<int64_t> doff(c->P.n);
  for (int h = 1; h <= c->P.N; h++)
    for (int i = 0; i < c->P.sz[h]; i++) doff[c->P.start[h] + i] = c->S.poff[h] + i + (int64_t)i * c->S.ld[h];
  if (upload(c, &c->d_diag_off, doff)) return -100;
  CK(cudaMalloc((void **)&c->d_diag, std::max(1, c->P.n) * sizeof(double)));
  CK(cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem_bytes(kRowBlock)));
  CK(cudaFuncSetAttribute(trsm_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmemBytes));
  CK(cudaFuncSetAttribute(gemm_grouped_ws<64, 64, 16, 32, 32, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmMain::kSmemBytes));
  {  // load every kernel now: with lazy module loading the first launch of a function synchronises the context, which
     // must not happen while another rank that shares this device is already spinning on a flag
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, panel_kernel));
    CK(cudaFuncGetAttributes(&fa, trsm_tile));
    CK(cudaFuncGetAttributes(&fa, gemm_small_warp));
    CK(cudaFuncGetAttributes(&fa, gemm_grouped_ws<64, 64, 16, 32, 32, 4, 3>));
    CK(cudaFuncGetAttributes(&fa, assemble_kernel));
    CK(cudaFuncGetAttributes(&fa, gather_diag_kernel));
    CK(cudaFuncGetAttributes(&fa, peer_sync));
    CK(cudaFuncGetAttributes(&fa, push_rects));
    CK(cudaFuncGetAttributes(&fa, reduce_rects));
    CK(cudaFuncGetAttributes(&fa, compare_rects));
  }
  for (cudaEvent_t &e : c->tev) CK(cudaEventCreate(&e));
  CK(cudaMalloc((void **)&c->d_flags, kFlagSlots * kMaxPeers * sizeof(unsigned long long)));
  CK(cudaMemset(c->d_flags, 0, kFlagSlots * kMaxPeers * sizeof(unsigned long long)));
  CK(cudaMalloc((void **)&c->d_counters, kStreams * sizeof(unsigned)));
  CK(cudaMemset(c->d_counters, 0, kStreams * sizeof(unsigned)));
  c->peers.n = 1, c->peers.rank = 0;
  c->peers.fac[0] = c->d_fac, c->peers.flags[0] = c->d_flags;
  c->peers_ready = (c->world == 1);
  c->run_id = 0;
  return 0;
}
// a group: every rank's buffers, then peer access between the devices and the peer tables
static int ensure_device(chol_t *c) {
  if (!is_group(c)) return rank_device(c);
  if (!c->analyzed) return fail(c, "analyze first");
  bool ready = true;
  for (chol_t *s : c->sub) ready = ready && s->device_ready && s->peers_ready;
  if (ready) return 0;
  for (chol_t *s : c->sub)
    if (rank_device(s)) return fail(c, s->err);
  const int n = (int)c->sub.size();
  for (int a = 0; a < n; a++)
    for (int b = 0; b < n; b++) {
      const int da = c->sub[a]->device, db = c->sub[b]->device;
      if (da == db) continue;
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, da, db));
      if (!can) return fail(c, "device " + std::to_string(da) + " cannot map the memory of device " + std::to_string(db));
      CK(cudaSetDevice(da));
      cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
      (void)cudaGetLastError();
    }
  for (int a = 0; a < n; a++) {
    chol_t *s = c->sub[a];
    s->peers.n = n, s->peers.rank = a;
    for (int b = 0; b < n; b++) s->peers.fac[b] = c->sub[b]->d_fac, s->peers.flags[b] = c->sub[b]->d_flags;
    s->peers_ready = true;
  }
  c->device_ready = true;
  return 0;
}

static int do_assemble(chol_t *c) {
  CK(cudaMemsetAsync(c->d_fac, 0, (size_t)c->S.total_doubles * sizeof(double), c->stream));
  const int info0[2] = {0x7fffffff, 0};
  CK(cudaMemcpyAsync(c->d_info, info0, sizeof info0, cudaMemcpyHostToDevice, c->stream));
  if (c->P.nz > 0) {
    int64_t nzv = c->P.nz;
    assemble_kernel<<<(unsigned)((nzv + 255) / 256), 256, 0, c->stream>>>(c->d_vals, c->d_aoff, nzv, c->d_fac);
  }
  CK(cudaGetLastError());
  c->assembled = true;
  c->factored = false;
  c->h_fac_valid = false;
  return 0;
}

// One kernel on the stream of the launch being issued.  (Programmatic dependent launch was tried here: every kernel
// letting its successor be scheduled early and waiting for its predecessor after reading its descriptors.  Measured
// on B200: 512^2 2.08 -> 2.03 ms, 64^3 23.2 -> 22.8 ms, 128^3 902.6 -> 913.1 ms -- the early-resident CTAs of the chain
// stream take SM slots from the trailing updates -- and with several ranks on one GPU they starve the rank whose
// flag they wait for.  Dropped.)
template <typename... KArgs, typename... Args>
static void launch(chol_t *c, void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, Args... args) {
  kern<<<grid, block, smem, c->cur>>>(static_cast<KArgs>(args)...);
}

static int run_launch(chol_t *c, const Launch &l) {
  const unsigned long long flag = (c->run_id << 32) | (unsigned long long)l.seq;
  switch (l.kind) {
    case K_PANEL:
      launch(c, panel_kernel, (unsigned)l.count, kPanelThreads, panel_smem_bytes(l.cfg), c->d_pdesc, c->d_pslabs + l.begin, c->d_fac, c->d_pflags,
             (l.cfg + kNB - 1) / kNB * kNB, c->d_info);
      break;
    case K_TRSM:
      launch(c, trsm_tile, (unsigned)l.count, kSlab, kTrsmSmemBytes, c->d_trsm, c->d_trsm_tiles + l.begin, c->d_fac);
      break;
    case K_GEMM:
      if (l.count <= 0) break;
      if (l.cfg == 3)
        launch(c, gemm_small_warp, (unsigned)((l.count + kSmallWarps - 1) / kSmallWarps), kSmallWarps * 32, 0, c->d_probs, c->d_contribs,
               c->d_tiles + l.begin, l.count, c->d_fac);
      else
        launch(c, gemm_grouped_ws<64, 64, 16, 32, 32, 4, 3>, (unsigned)l.count, GemmMain::kThreads, GemmMain::kSmemBytes, c->d_probs, c->d_contribs,
               c->d_tiles + l.begin, c->d_fac);
      break;
    case K_SYNC:
      launch(c, peer_sync, 1, 32, 0, c->peers, l.slot, flag, l.sig_mask, l.wait_mask, c->d_info + 1);
      break;
    case K_PUSH:
      if (l.count > 0 && l.mask) {  // enough CTAs per rectangle to spread a small push over about two waves of SMs
        int cg = 8;
        while (cg < 64 && l.count * cg < 2 * c->num_sms) cg *= 2;
        launch(c, push_rects, (unsigned)(l.count * cg), kPushThreads, 0, c->d_rects + l.begin, c->d_fac, c->peers, l.mask, cg, l.slot, flag, l.sig_mask,
               c->d_counters + l.stream);
      } else if (l.sig_mask)
        launch(c, peer_sync, 1, 32, 0, c->peers, l.slot, flag, l.sig_mask, 0u, c->d_info + 1);
      break;
    case K_REDUCE: {
      const int cg = std::max(1, std::min(64, (int)(4 * c->num_sms / std::max<int64_t>(1, l.count))));
      launch(c, reduce_rects, (unsigned)(l.count * cg), kPushThreads, 0, c->d_rects + l.begin, c->d_fac, c->peers, l.mask, cg);
      break;
    }
    default:
      break;
  }
  return 0;
}
static bool launches_kernel(const Launch &l) {
  if (l.kind == K_NOP) return false;
  if (l.kind == K_GEMM) return l.count > 0;
  if (l.kind == K_PUSH) return (l.count > 0 && l.mask) || l.sig_mask;
  return true;
}

extern "C" {
static int run_levels(chol_t *c, int lvl_from, int lvl_to, int phase_mask, bool per_kernel_timing) {
  if (c->world > 1 && !c->peers_ready) return fail(c, "multi-GPU handle: exchange IPC handles first (chol_ipc_export / chol_ipc_import)");
  std::vector<cudaEvent_t> ev;
  std::vector<int> kinds;
  std::vector<double> fl;
  if ((int)c->evs.size() < c->D.num_events || (per_kernel_timing && c->kev.size() < 2 * c->D.launches.size()))
    return fail(c, "internal: prepare_rank has to run before the level loop");
  size_t kev_next = 0;
  const bool whole = lvl_from >= c->P.levels - 1 && lvl_to <= 0 && phase_mask == 7;
  if (c->world > 1 && !whole) return fail(c, "partial runs of the level loop need a single-GPU handle");
  if (whole) c->run_id++;
  // graph replay / capture: whole level loop, single GPU, uninstrumented (see chol::use_graph)
  const bool want_graph = c->use_graph < 0 ? c->S.flops() < kGraphFlops : c->use_graph != 0;
  const bool graphable = want_graph && c->world == 1 && whole && !per_kernel_timing;
  if (graphable && c->graph_exec) {
    CK(cudaGraphLaunch(c->graph_exec, c->stream));
    return 0;
  }
  const bool capture = graphable && c->graph_runs >= 1;
  if (graphable) c->graph_runs++;
  if (capture) CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  CK(cudaMemsetAsync(c->d_pflags, 0, std::max<size_t>(1, c->D.pdesc.size()) * 4 * sizeof(int), c->stream));  // no diagonal tile is factored yet
  if (c->D.lookahead) {
    CK(cudaEventRecord(c->ev_fork, c->stream));
    for (int i = 1; i < kStreams; i++) CK(cudaStreamWaitEvent(c->streams[i], c->ev_fork, 0));
  }
  for (const Launch &l : c->D.launches) {
    if (l.level > lvl_from || l.level < lvl_to) continue;
    if (l.kind != K_NOP && !(l.phase & phase_mask)) continue;
    // the instrumented pass runs everything on one stream so that an event pair brackets one kernel alone
    c->cur = per_kernel_timing ? c->stream : c->streams[l.stream];
    if (l.wait_ev >= 0 && !per_kernel_timing) cudaStreamWaitEvent(c->cur, c->evs[l.wait_ev], 0);
    if (per_kernel_timing) {
      cudaEvent_t a = c->kev[kev_next++], b = c->kev[kev_next++];
      cudaEventRecord(a, c->cur);
      run_launch(c, l);
      cudaEventRecord(b, c->cur);
      ev.push_back(a), ev.push_back(b), kinds.push_back(l.kind), fl.push_back(l.flops);
    } else
      run_launch(c, l);
    if (l.rec_ev >= 0 && !per_kernel_timing) cudaEventRecord(c->evs[l.rec_ev], c->cur);
  }
  c->cur = c->stream;
  if (c->D.lookahead) {  // partial runs (piecewise calls) may leave work on the other streams: join them
    for (int i = 1; i < kStreams; i++) {
      CK(cudaEventRecord(c->ev_join[i], c->streams[i]));
      CK(cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
    }
  }
  if (whole) c->factored = true;
  if (capture) {
    cudaGraph_t g = nullptr;
    CK(cudaStreamEndCapture(c->stream, &g));
    cudaError_t e = cudaGraphInstantiate(&c->graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
      c->graph_exec = nullptr;
      c->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e);
      return -100;
    }
    CK(cudaGraphLaunch(c->graph_exec, c->stream));  // the captured launches have not run yet
    return 0;
  }
  CK(cudaGetLastError());
  if (per_kernel_timing) {
    CK(cudaStreamSynchronize(c->stream));
    for (double &m : c->k_ms) m = 0;
    c->k_gemm_flops = 0;
    c->launch_ms.assign(kinds.size(), 0.f);
    for (size_t i = 0; i < kinds.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
      c->k_ms[kinds[i]] += ms;
      c->launch_ms[i] = ms;
      if (kinds[i] == K_GEMM) c->k_gemm_flops += fl[i];
    }
  }
  return 0;
}

int chol_assemble(chol_t *c) {
  if (ensure_device(c)) return -1;
  return for_ranks(c, [](chol_t *r) {
    chol_t *c = r;
    if (do_assemble(c)) return -1;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
  });
}

// kernels one step launches
static int64_t count_kernels(chol_t *c) {
  int64_t k = 0;
  for (const Launch &l : c->D.launches) k += launches_kernel(l) ? 1 : 0;
  return k;
}

static int fetch_info(chol_t *c, int *info) {
  int v[2] = {0, 0};
  CK(cudaMemcpyAsync(v, c->d_info, sizeof v, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *info = (v[0] == 0x7fffffff) ? 0 : v[0];
  if (v[1]) return fail(c, "a wait on a peer GPU timed out (a rank of the partition is missing or failed)");
  return 0;
}
static int info_error(chol_t *c, int info) {
  c->factored = false;  // the panels past the bad pivot are meaningless: results and solves are refused
  return fail(c, "matrix is not positive definite: pivot at permuted column " + std::to_string(info));
}

// Everything a step needs that allocates or may synchronise the device (schedule upload, events, pinned staging
// buffer) is done here, for ALL ranks of a group, before any rank issues a launch: once a rank's kernels spin on
// flag words, a device-wide synchronisation by another rank that shares the GPU could never return.
static int prepare_rank(chol_t *c, bool pinned, bool instrumented) {
  if (ensure_schedule(c, false)) return -1;
  while ((int)c->evs.size() < c->D.num_events) {  // cross-stream events of the launch list (look-ahead)
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->evs.push_back(e);
  }
  while (instrumented && c->kev.size() < 2 * c->D.launches.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    c->kev.push_back(e);
  }
  const size_t need = std::max((size_t)c->P.nz, (size_t)c->P.n) * sizeof(double);
  if (pinned && c->h_pinned_bytes < need) {
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    CK(cudaMallocHost((void **)&c->h_pinned, need));
    c->h_pinned_bytes = need;
  }
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// one rank: `warmup + iterations` x (assemble; level loop), device time of every timed level loop
static int factor_rank(chol_t *c, int iterations, int warmup, std::vector<double> &secs, double &asm_s, int &info) {
  cudaEvent_t e0 = c->tev[0], e1 = c->tev[1], e2 = c->tev[2];
  for (int it = 0; it < warmup + iterations; it++) {
    CK(cudaEventRecord(e0, c->stream));
    if (do_assemble(c)) return -1;
    CK(cudaEventRecord(e1, c->stream));
    if (run_levels(c, c->P.levels - 1, 0, 7, false)) return -1;
    CK(cudaEventRecord(e2, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ma = 0, mf = 0;
    CK(cudaEventElapsedTime(&ma, e0, e1));
    CK(cudaEventElapsedTime(&mf, e1, e2));
    if (it >= warmup) secs.push_back(mf * 1e-3), asm_s = ma * 1e-3;
  }
  return fetch_info(c, &info);
}

int chol_factor(chol_t *c, int iterations, int warmup, chol_stats_t *st) {
  if (c->parent) return fail(c, "factor through the group handle");
  if (ensure_device(c)) return -1;
  if (iterations < 1) iterations = 1;
  const int nr = chol_num_ranks(c);
  std::vector<std::vector<double>> secs(nr);
  std::vector<double> asm_s(nr, 0.0);
  std::vector<int> info(nr, 0);
  if (for_ranks(c, [](chol_t *r) { return prepare_rank(r, false, false); })) return -1;
  if (for_ranks(c, [&](chol_t *r) { return factor_rank(r, iterations, warmup, secs[r->parent ? r->rank : 0], asm_s[r->parent ? r->rank : 0], info[r->parent ? r->rank : 0]); }))
    return -1;
  // a step of a group takes as long as its slowest rank
  std::vector<double> step(iterations, 0.0);
  double as = 0;
  int bad = 0;
  for (int r = 0; r < nr; r++) {
    for (int i = 0; i < iterations; i++) step[i] = std::max(step[i], secs[r][i]);
    as = std::max(as, asm_s[r]);
    if (info[r] && (!bad || info[r] < bad)) bad = info[r];
  }
  if (st) {
    std::vector<double> s = step;
    std::sort(s.begin(), s.end());
    st->seconds_best = s.front();
    st->seconds_median = s[s.size() / 2];
    st->seconds_last = step.back();
    st->assemble_seconds = as;
    st->flops = c->S.flops();
    st->kernel_launches = 0;
    for (int r = 0; r < nr; r++) st->kernel_launches += count_kernels(chol_rank_handle(c, r));
    st->info = bad;
  }
  c->assembled = c->factored = true;
  if (bad) {
    for (int r = 0; r < nr; r++) chol_rank_handle(c, r)->factored = false;
    return info_error(c, bad);
  }
  return 0;
}

static int piecewise(chol_t *c, int lvl, int phase) {
  if (is_group(c) || c->world > 1) return fail(c, "the piecewise fused tasks run on a single-GPU handle");
  if (ensure_device(c)) return -1;
  cudaSetDevice(c->device);
  if (!c->assembled) return fail(c, "assemble first");
  if (lvl < 0 || lvl >= c->P.levels) return fail(c, "bad level");
  if (ensure_schedule(c, phase != PH_UPDATE || c->D.split_phases)) return -1;
  while ((int)c->evs.size() < c->D.num_events) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->evs.push_back(e);
  }
  c->h_fac_valid = false;
  if (run_levels(c, lvl, lvl, phase, false)) return -1;
  CK(cudaStreamSynchronize(c->stream));
  if (lvl == 0 && phase == PH_POTRF) c->factored = true;  // the last fused task of the level loop (the root has no ancestors)
  return 0;
}
int chol_fused_dpotrf(chol_t *c, int lvl) { return piecewise(c, lvl, PH_POTRF); }
int chol_fused_dtrsm(chol_t *c, int lvl) { return piecewise(c, lvl, PH_TRSM); }
int chol_fused_update(chol_t *c, int lvl) { return piecewise(c, lvl, PH_UPDATE); }

// one rank of chol_factor_host: H2D of the values, assemble, level loop, D2H of diag(L)
static int factor_host_rank(chol_t *c, const double *values, double &secs, int &info) {
  cudaEvent_t e0 = c->tev[0], e1 = c->tev[1];
  CK(cudaEventRecord(e0, c->stream));
  const double *src = values ? values : c->P.ev.data();
  memcpy(c->h_pinned, src, (size_t)c->P.nz * sizeof(double));
  CK(cudaMemcpyAsync(c->d_vals, c->h_pinned, (size_t)c->P.nz * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (do_assemble(c)) return -1;
  if (run_levels(c, c->P.levels - 1, 0, 7, false)) return -1;
  gather_diag_kernel<<<(c->P.n + 255) / 256, 256, 0, c->stream>>>(c->d_diag_off, c->P.n, c->d_fac, c->d_diag);
  CK(cudaMemcpyAsync(c->h_pinned, c->d_diag, (size_t)c->P.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaEventRecord(e1, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  secs = ms * 1e-3;
  return fetch_info(c, &info);
}
// does this rank report separator h (its own subtree; rank 0 also the top panels, which every rank holds complete)
static bool reports(const chol_t *c, int h) {
  if (c->world == 1) return true;
  const int lv = c->P.level_of(h), own = lv < c->D.depth ? -1 : (h >> (lv - c->D.depth)) - (1 << c->D.depth);
  return own == c->rank || (own < 0 && c->rank == 0);
}

int chol_factor_host(chol_t *c, const double *values, int64_t nz, double *diag_out, chol_stats_t *st) {
  if (c->parent) return fail(c, "factor through the group handle");
  if (ensure_device(c)) return -1;
  if (values && nz != c->P.nz) return fail(c, "value count differs from the loaded pattern");
  const int nr = chol_num_ranks(c);
  std::vector<double> secs(nr, 0.0);
  std::vector<int> info(nr, 0);
  if (for_ranks(c, [](chol_t *r) { return prepare_rank(r, true, false); })) return -1;
  if (for_ranks(c, [&](chol_t *r) { return factor_host_rank(r, values, secs[r->parent ? r->rank : 0], info[r->parent ? r->rank : 0]); })) return -1;
  double worst = 0;
  int bad = 0;
  int64_t kernels = 0;
  for (int r = 0; r < nr; r++) {
    chol_t *rk = chol_rank_handle(c, r);
    worst = std::max(worst, secs[r]);
    if (info[r] && (!bad || info[r] < bad)) bad = info[r];
    kernels += count_kernels(rk) + 2;
    if (diag_out)  // every rank hands back the diagonal entries of the separators it reports (zeros elsewhere on a partitioned handle)
      for (int h = 1; h <= c->P.N; h++) {
        const bool mine = reports(rk, h);
        if (!mine && is_group(c)) continue;
        for (int i = 0; i < c->P.sz[h]; i++) diag_out[c->P.start[h] + i] = mine ? rk->h_pinned[c->P.start[h] + i] : 0.0;
      }
  }
  if (st) {
    st->seconds_best = st->seconds_median = st->seconds_last = worst;
    st->assemble_seconds = 0;
    st->flops = c->S.flops();
    st->kernel_launches = kernels;
    st->info = bad;
  }
  c->assembled = c->factored = true;
  if (bad) {
    for (int r = 0; r < nr; r++) chol_rank_handle(c, r)->factored = false;
    return info_error(c, bad);
  }
  return 0;
}

/* ---- multi-GPU plumbing, one process per GPU */
int chol_set_partition(chol_t *c, int rank, int world) {
  if (is_group(c) || c->parent) return fail(c, "a group handle is partitioned by chol_create");
  if (world < 1 || world > kMaxPeers || (world & (world - 1)) || rank < 0 || rank >= world) return fail(c, "world must be 1, 2, 4 or 8 and 0 <= rank < world");
  free_device(c);
  c->rank = rank, c->world = world;
  c->analyzed = false;
  return 0;
}
int chol_ipc_export(chol_t *c, void *handles128) {
  if (is_group(c) || c->parent) return fail(c, "the ranks of a group handle share one process: nothing to export");
  if (ensure_device(c)) return -1;
  cudaIpcMemHandle_t h[2];
  CK(cudaIpcGetMemHandle(&h[0], c->d_fac));
  CK(cudaIpcGetMemHandle(&h[1], c->d_flags));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(handles128, h, 128);
  return 0;
}
int chol_ipc_import(chol_t *c, const void *all_handles, int world) {
  if (is_group(c) || c->parent) return fail(c, "the ranks of a group handle share one process: nothing to import");
  if (ensure_device(c)) return -1;
  if (world != c->world) return fail(c, "world size mismatch");
  const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)all_handles;
  c->peers.n = world, c->peers.rank = c->rank;
  for (int p = 0; p < world; p++) {
    if (p == c->rank) {
      c->peers.fac[p] = c->d_fac, c->peers.flags[p] = c->d_flags;
      continue;
    }
    void *f = nullptr, *g = nullptr;
    CK(cudaIpcOpenMemHandle(&f, h[2 * p], cudaIpcMemLazyEnablePeerAccess));
    CK(cudaIpcOpenMemHandle(&g, h[2 * p + 1], cudaIpcMemLazyEnablePeerAccess));
    c->ipc_opened.push_back(f), c->ipc_opened.push_back(g);
    c->peers.fac[p] = (double *)f, c->peers.flags[p] = (unsigned long long *)g;
  }
  c->peers_ready = true;
  return 0;
}
static chol_t *first_rank(chol_t *c) { return is_group(c) ? c->sub[0] : c; }
/* what this rank's schedule covers: [0] matrix entries it assembles, [1] GEMM flops it executes,
 * [2] peer-store launches (rows pushed to other ranks), [3] doubles of the top panels (one copy per rank),
 * [4] diagonal tiles it factors, [5] 64-row slabs it solves, [6] bytes it pulls from peers for the partial-sum reduction, [7] bytes it pushes
 * to peers per factorization.  On a group handle: rank 0 (chol_rank_handle gives the others). */
int chol_partition_stats(chol_t *c, double *out8) {
  if (!c->analyzed) return fail(c, "analyze first");
  c = first_rank(c);
  double a = 0, f = 0, sh = 0, pt = 0, ts = 0, pull = 0, pushed = 0;
  for (int64_t o : c->D.a_off) a += (o >= 0);
  auto bits = [](unsigned m) { return (double)__builtin_popcount(m); };
  auto area = [](const RectDesc &d) {  // entries at or below the rectangle's diagonal offset
    double e = 0;
    for (int r = 0; r < d.rows; r++) e += std::max(0, std::min(d.cols, d.tri0 >= (1 << 29) ? d.cols : d.tri0 + r + 1));
    return e;
  };
  for (const Launch &l : c->D.launches) {
    if (l.kind == K_GEMM) f += l.flops;
    if (l.kind == K_PUSH) sh += (l.count > 0 && l.mask);
    for (int64_t i = l.begin; i < l.begin + l.count && l.kind == K_PANEL; i++) (c->D.pslabs[i].t >= 0 ? pt : ts) += 1.0;
    if (l.kind == K_TRSM) ts += 2.0 * (double)l.count;  // 128-row slabs
    for (int64_t i = l.begin; i < l.begin + l.count && (l.kind == K_PUSH || l.kind == K_REDUCE); i++) {
      const RectDesc &d = c->D.rects[i];
      if (l.kind == K_PUSH) pushed += 8.0 * area(d) * bits(l.mask);
      else pull += 8.0 * area(d) * bits(d.mask & l.mask & ~(1u << c->rank));
    }
  }
  out8[0] = a, out8[1] = f, out8[2] = sh, out8[3] = (double)c->D.top_doubles, out8[4] = pt, out8[5] = ts, out8[6] = pull, out8[7] = pushed;
  return 0;
}
/* Algorithmic HBM bytes of one tree level (8 B x distinct clusters read + written, read-modify-
 * written clusters counted twice): [0] panels of the level's separators, factored in place (pivot block
 * lower triangle + filled off-diagonal rows, 16 B per entry); [1] operands of the level's Schur updates
 * (every filled off-diagonal row cluster once, 8 B per entry); [2] their destination clusters (16 B per
 * entry, lower triangle only on diagonal clusters).  Single-GPU handles. */
int chol_level_bytes(chol_t *c, int lvl, double *out3) {
  if (!c->analyzed) return fail(c, "analyze first");
  if (lvl < 0 || lvl >= c->P.levels) return fail(c, "bad level");
  if (is_group(c) || c->world > 1) return fail(c, "chol_level_bytes: single-GPU handles");
  const Problem &P = c->P;
  const Symbolic &S = c->S;
  double panel = 0, oper = 0, dest = 0;
  for (int h = 1 << lvl; h < (1 << (lvl + 1)); h++) {
    const double n = P.sz[h];
    double off = 0;
    for (int64_t s = S.seg_ptr[h] + 1; s < S.seg_ptr[h + 1]; s++) off += S.segs[s].hi - S.segs[s].lo;
    panel += 16.0 * (n * (n + 1) / 2 + off * n);
    oper += 8.0 * off * n;
  }
  for (const Launch &l : c->D.launches) {
    if (l.level != lvl || l.kind != K_GEMM || l.phase != PH_UPDATE) continue;
    int last = -1;
    for (int64_t t = l.begin; t < l.begin + l.count; t++) {
      const int p = c->D.tiles[t].prob;
      if (p == last) continue;  // the tiles of one problem are consecutive
      last = p;
      const GemmProblem &g = c->D.probs[p];
      dest += 16.0 * ((double)g.M * g.N - ((g.tri & 1) ? 0.5 * g.N * (g.N - 1.0) : 0.0));
    }
  }
  out3[0] = panel, out3[1] = oper, out3[2] = dest;
  return 0;
}
/* Largest |difference| between this rank's copy of the factored top panels (lower triangle of the pivot blocks and
 * every stored off-diagonal row) and the copy of the next rank, compared on the GPU through peer memory; on a group
 * handle the largest over all ranks (so 0 means all copies are bit-identical).  0 on a single-GPU handle. */
static int top_diff_rank(chol_t *c, double *out) {
  *out = 0;
  if (c->world == 1) return 0;
  if (!c->device_ready || !c->factored || !c->peers_ready) return fail(c, "factor first");
  std::vector<RectDesc> rects;
  for (int h = 1; h < (1 << c->D.depth); h++)
    if (c->P.sz[h] > 0) rects.push_back(RectDesc{c->S.poff[h], c->S.ld[h], c->S.rows[h], c->P.sz[h], 0, 0u, 0});
  if (rects.empty()) return 0;
  RectDesc *d_r = nullptr;
  unsigned long long *d_w = nullptr, w = 0;
  if (upload(c, &d_r, rects)) return -100;
  CK(cudaMalloc((void **)&d_w, sizeof w));
  CK(cudaMemsetAsync(d_w, 0, sizeof w, c->stream));
  const int cg = std::max(1, 4 * c->num_sms / (int)rects.size());
  compare_rects<<<(unsigned)(rects.size() * cg), kPushThreads, 0, c->stream>>>(d_r, c->d_fac, c->peers, (c->rank + 1) % c->world, cg, d_w);
  CK(cudaMemcpyAsync(&w, d_w, sizeof w, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_r), cudaFree(d_w);
  memcpy(out, &w, sizeof w);
  return 0;
}
int chol_top_copies_diff(chol_t *c, double *maxdiff) {
  const int nr = chol_num_ranks(c);
  std::vector<double> d(nr, 0.0);
  if (for_ranks(c, [&](chol_t *r) { return top_diff_rank(r, &d[r->parent ? r->rank : 0]); })) return -1;
  *maxdiff = *std::max_element(d.begin(), d.end());
  return 0;
}
int chol_rank(chol_t *c) { return c->rank; }
int chol_world(chol_t *c) { return c->world; }

/* per-launch device time (ms) of the last chol_kernel_times pass, in launch-list order */
int64_t chol_launch_times(chol_t *c, float *ms, int64_t cap) {
  c = first_rank(c);
  int64_t n = std::min<int64_t>(cap, (int64_t)c->launch_ms.size());
  for (int64_t i = 0; i < n; i++) ms[i] = c->launch_ms[i];
  return (int64_t)c->launch_ms.size();
}
int64_t chol_num_launches(chol_t *c) { return c->analyzed ? (int64_t)first_rank(c)->D.launches.size() : -1; }
int chol_get_launch(chol_t *c, int64_t i, int *kind, int *level, int *phase, int64_t *ctas, double *flops, int *cfg) {
  if (!c->analyzed) return -1;
  c = first_rank(c);
  if (i < 0 || i >= (int64_t)c->D.launches.size()) return -1;
  const Launch &l = c->D.launches[i];
  *kind = l.kind, *level = l.level, *phase = l.phase, *ctas = l.count, *flops = l.flops, *cfg = (l.kind == K_GEMM ? l.cfg : 0) | (l.stream << 4) | (l.kind == K_PANEL ? l.cfg << 8 : 0);
  return 0;
}

int chol_synchronize(chol_t *c) {
  if (!c->device_ready) return 0;
  return for_ranks(c, [](chol_t *r) {
    chol_t *c = r;
    for (int i = 0; i < kStreams; i++) CK(cudaStreamSynchronize(c->streams[i]));
    return 0;
  });
}

/* one instrumented factorization: everything on one stream per rank, CUDA events around every launch */
int chol_kernel_times(chol_t *c, double *potrf_ms /* panel_ms */, double *trsm_ms /* exchange_ms */, double *gemm_ms, double *gemm_flops) {
  if (c->parent) return fail(c, "through the group handle");
  if (ensure_device(c)) return -1;
  if (for_ranks(c, [](chol_t *r) { return prepare_rank(r, false, true); })) return -1;
  if (for_ranks(c, [](chol_t *r) {
        if (do_assemble(r)) return -1;
        return run_levels(r, r->P.levels - 1, 0, 7, true);
      }))
    return -1;
  chol_t *r = first_rank(c);
  if (potrf_ms) *potrf_ms = r->k_ms[K_PANEL] + r->k_ms[K_TRSM];  // panel_kernel + trsm_tile: diagonal blocks and the rows below them
  if (trsm_ms) *trsm_ms = r->k_ms[K_SYNC] + r->k_ms[K_PUSH] + r->k_ms[K_REDUCE];  // multi-GPU exchange (0 on one GPU)
  if (gemm_ms) *gemm_ms = r->k_ms[K_GEMM];
  if (gemm_flops) *gemm_flops = r->k_gemm_flops;
  return 0;
}

// ------------------------------------------------------------------------------ results
static int fetch_factor(chol_t *c) {
  if (!c->device_ready || !c->assembled) return fail(c, "nothing factored yet");
  if (c->h_fac_valid) return 0;
  c->h_fac.resize((size_t)c->S.total_doubles);
  if (!is_group(c)) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(c->h_fac.data(), c->d_fac, (size_t)c->S.total_doubles * sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    // a group: the top panels from rank 0 (every rank holds them complete), every subtree from its owner; the
    // panels of one rank on one tree level are contiguous in the factor buffer
    const Problem &P = c->P;
    const Symbolic &S = c->S;
    const int world = (int)c->sub.size(), depth = c->sub[0]->D.depth;
    auto copy = [&](chol_t *r, int h0, int h1) -> int {
      const int64_t o0 = S.poff[h0], o1 = S.poff[h1];
      CK(cudaSetDevice(r->device));
      CK(cudaStreamSynchronize(r->stream));
      if (o1 > o0) CK(cudaMemcpy(c->h_fac.data() + o0, r->d_fac + o0, (size_t)(o1 - o0) * sizeof(double), cudaMemcpyDeviceToHost));
      return 0;
    };
    if (copy(c->sub[0], 1, 1 << depth)) return -100;
    for (int lvl = depth; lvl < P.levels; lvl++)
      for (int r = 0; r < world; r++) {
        const int h0 = ((1 << depth) + r) << (lvl - depth);
        if (copy(c->sub[r], h0, h0 + (1 << (lvl - depth)))) return -100;
      }
  }
  c->h_fac_valid = true;
  return 0;
}

// visit stored entries != 0 block by block ((row_sep, col_sep) ascending labels), row-major inside
// a block, as write_matrix does (mmat.rg:114-144)
}  // extern "C"
template <typename F>
static void visit(chol_t *c, F fn) {
  const Problem &P = c->P;
  const Symbolic &S = c->S;
  // blocks of row separator r: column separators r and its descendants
  for (int lr = 1; lr <= P.N; lr++) {
    int hr = P.heap_of(lr), lv = P.level_of(hr);
    std::vector<int> cols;
    for (int d = 0; lv + d < P.levels; d++)
      for (int hc = hr << d; hc < ((hr + 1) << d); hc++) cols.push_back(hc);
    std::sort(cols.begin(), cols.end(), [](int a, int b) { return a > b; });  // ascending label
    for (int hc : cols) {
      if (!is_group(c) && !reports(c, hc)) continue;  // a rank reports its own subtree; rank 0 also the top panels
      const double *pan = c->h_fac.data() + S.poff[hc];
      int ld = S.ld[hc];
      for (int64_t s = S.seg_ptr[hc]; s < S.seg_ptr[hc + 1]; s++) {
        const Seg &sg = S.segs[s];
        if (sg.anc != hr) continue;
        for (int r = 0; r < sg.hi - sg.lo; r++)
          for (int col = 0; col < P.sz[hc]; col++) {
            double v = pan[sg.off + r + (size_t)col * ld];
            if (v != 0) fn(P.start[hr] + sg.lo + r, P.start[hc] + col, v);
          }
      }
    }
  }
}

extern "C" {
int64_t chol_factor_nnz(chol_t *c) {
  if (fetch_factor(c)) return -1;
  int64_t k = 0;
  visit(c, [&](int, int, double) { k++; });
  return k;
}
int64_t chol_get_factor_coo(chol_t *c, int32_t *I, int32_t *J, double *V) {
  if (fetch_factor(c)) return -1;
  int64_t k = 0;
  visit(c, [&](int i, int j, double v) { I[k] = i, J[k] = j, V[k] = v, k++; });
  return k;
}
int chol_get_factor_dense(chol_t *c, double *out) {
  if (fetch_factor(c)) return -1;
  size_t n = (size_t)c->P.n;
  memset(out, 0, n * n * sizeof(double));
  visit(c, [&](int i, int j, double v) { out[(size_t)i * n + j] = v; });
  return 0;
}
static int write_factor_file(chol_t *c, const char *path, int full) {
  FILE *f = fopen(path, "w");
  if (!f) return fail(c, std::string("cannot write ") + path);
  int64_t nnz = 0;
  visit(c, [&](int, int, double) { nnz++; });
  MM_typecode tc;
  memcpy(tc, c->P.typecode, 4);
  mm_write_banner(f, tc);
  if (nnz <= 0x7fffffff) mm_write_mtx_crd_size(f, c->P.n, c->P.ncols, (int)nnz);
  else fprintf(f, "%d %d %lld\n", c->P.n, c->P.ncols, (long long)nnz);  // mmio's int count overflows (128^3: 3.4e9 entries)
  visit(c, [&](int i, int j, double v) { fprintf(f, full ? "%d %d %.17g\n" : "%d %d %0.8g\n", i + 1, j + 1, v); });
  const bool bad = ferror(f) != 0;
  if (fclose(f) != 0 || bad) return fail(c, std::string("write error on ") + path);
  return 0;
}
int chol_write_factor(chol_t *c, const char *path, int full) {
  if (fetch_factor(c)) return -1;
  return write_factor_file(c, path, full);
}

int chol_write_factor_binary(chol_t *c, const char *path) {
  if (fetch_factor(c)) return -1;
  if (is_group(c)) return write_factor_binary(c->P, c->S, c->h_fac.data(), 0, 1, 0, path, c->err) ? -1 : 0;
  return write_factor_binary(c->P, c->S, c->h_fac.data(), c->rank, c->world, c->D.depth, path, c->err) ? -1 : 0;
}

// ------------------------------------------------------------------------------ debug trace (`-d`)
int chol_write_debug_log(chol_t *c, const char *path) {
  if (!c->analyzed) return fail(c, "analyze first");
  FILE *f = path ? fopen(path, "w") : stdout;
  if (!f) return fail(c, std::string("cannot write ") + path);
  int rc = write_debug_log(c->P, c->S, f, c->err);
  if (path) fclose(f);
  else fflush(f);
  return rc;
}

// the human-readable companion of a snapshot (write_blocks, mmat.rg:185-217): every block, dense, "%0.2f"
static int write_blocks_txt(chol_t *c, const char *path, const DebugStep &st) {
  const Problem &P = c->P;
  const Symbolic &S = c->S;
  FILE *f = fopen(path, "w");
  if (!f) return fail(c, std::string("cannot write ") + path);
  const int ls = P.label_of(st.hs), lp = st.hp ? P.label_of(st.hp) : 0, lg = st.hg ? P.label_of(st.hg) : 0;
  if (st.op == 0) fprintf(f, "Level: %d POTRF A=(%d, %d)\n", st.lvl, ls, ls);
  else if (st.op == 1) fprintf(f, "Level: %d TRSM A=(%d, %d) B=(%d, %d)\n", st.lvl, ls, ls, lp, ls);
  else fprintf(f, "Level: %d GEMM A=(%d, %d) B=(%d, %d) C=(%d, %d)\n", st.lvl, lg, ls, lp, ls, lg, lp);
  std::vector<double> blk;
  for (int lr = 1; lr <= P.N; lr++)
    for (int lc = 1; lc <= lr; lc++) {
      const int hr = P.heap_of(lr), hc = P.heap_of(lc), d = P.level_of(hc) - P.level_of(hr);
      if (d < 0 || (hc >> d) != hr) continue;
      const int m = P.sz[hr], n = P.sz[hc];
      if (m == 0 || n == 0) continue;
      fprintf(f, "Color: %d %d size: %dx%d bounds.lo: %d %d bounds.hi: %d %d vol: %lld\n", lr, lc, m, n, P.start[hr], P.start[hc],
              P.start[hr] + m - 1, P.start[hc] + n - 1, (long long)m * n);
      blk.assign((size_t)m * n, 0.0);
      const double *pan = c->h_fac.data() + S.poff[hc];
      for (int64_t s = S.seg_ptr[hc]; s < S.seg_ptr[hc + 1]; s++) {
        const Seg &sg = S.segs[s];
        if (sg.anc != hr) continue;
        for (int r = 0; r < sg.hi - sg.lo; r++)
          for (int col = 0; col < n; col++) blk[(size_t)(sg.lo + r) * n + col] = pan[sg.off + r + (size_t)col * S.ld[hc]];
      }
      for (int i = 0; i < m; i++) {
        for (int j = 0; j < n; j++) {
          const double v = blk[(size_t)i * n + j];
          fprintf(f, v < 0 ? "%0.2f, " : " %0.2f, ", v);
        }
        fprintf(f, "\n");
      }
    }
  fclose(f);
  return 0;
}

/* The level loop one fused task group at a time (mmat.rg:1240-1343 with debug = true): the schedule
 * compiler is asked for the launches of ONE separator and ONE phase, the GPU runs them, and the factor
 * as it then stands is written under every file name the reference would write at that point.
 * fused_dtrsm of one separator runs for all its ancestors at once, and so do its Schur updates, so the
 * snapshots of one such group are identical; verify.debug_factor compares only the block an operation
 * wrote (verify.py:98-124), which is final within its group. */
int chol_factor_debug(chol_t *c, const char *dir, int full_precision, int with_txt) {
  if (!c->analyzed) return fail(c, "analyze first");
  if (c->world > 1) return fail(c, "the debug trace runs on a single-GPU handle");
  if (c->P.n > 20000) return fail(c, "the debug trace writes the whole factor after every fused task: small problems only (n <= 20000)");
  if (ensure_device(c)) return -1;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (do_assemble(c)) return -1;
  Schedule saved = std::move(c->D);
  int rc = 0;
  const std::vector<DebugStep> steps = debug_steps(c->P);
  for (size_t i = 0; i < steps.size() && !rc;) {
    size_t j = i;
    while (j < steps.size() && steps[j].op == steps[i].op && steps[j].hs == steps[i].hs) j++;
    const int phase = steps[i].op == 0 ? PH_POTRF : steps[i].op == 1 ? PH_TRSM : PH_UPDATE;
    if (build_schedule(c->P, c->S, c->D, 0, 1, true, c->err, steps[i].hs)) rc = -1;
    if (!rc) rc = upload_schedule(c);
    if (!rc) rc = run_levels(c, steps[i].lvl, steps[i].lvl, phase, false);
    c->h_fac_valid = false;
    if (!rc) rc = fetch_factor(c);
    for (size_t k = i; k < j && !rc; k++) {
      const std::string base = std::string(dir) + "/" + steps[k].name;
      rc = write_factor_file(c, (base + ".mtx").c_str(), full_precision);
      if (!rc && with_txt) rc = write_blocks_txt(c, (base + ".txt").c_str(), steps[k]);
    }
    i = j;
  }
  cudaStreamSynchronize(c->stream);
  c->D = std::move(saved);
  if (upload_schedule(c)) return -100;
  if (rc) return rc;
  int info = 0;
  if (fetch_info(c, &info)) return -1;
  if (info != 0) return info_error(c, info);
  c->factored = true;
  return 0;
}

}  // extern "C"
// ------------------------------------------------------------------------------ residual check on the GPU
struct ResDev {
  ResPanel *panels = nullptr;
  ResTile *t_ltw = nullptr, *t_ly = nullptr;
  int *rowmap = nullptr;
  double *y = nullptr, *z = nullptr;
  int64_t n_ltw = 0, n_ly = 0;
};
static void free_res(chol_t *c) {
  ResDev *v = c->res;
  if (!v) return;
  cudaFree(v->panels), cudaFree(v->t_ltw), cudaFree(v->t_ly), cudaFree(v->rowmap), cudaFree(v->y), cudaFree(v->z);
  delete v;
  c->res = nullptr;
}
static int ensure_res(chol_t *c) {
  if (c->res) return 0;
  const Problem &P = c->P;
  const Symbolic &S = c->S;
  std::vector<ResPanel> panels;
  std::vector<ResTile> ltw, ly;
  std::vector<int> rowmap;
  for (int h = 1; h <= P.N; h++) {
    if (!reports(c, h) || P.sz[h] == 0) continue;
    const int n = P.sz[h], r0 = (n + 1) / 2 * 2, rows = S.rows[h];
    ResPanel rp{S.poff[h], (int64_t)rowmap.size(), S.ld[h], n, rows, r0, P.start[h], 0};
    rowmap.resize(rowmap.size() + (size_t)std::max(0, rows - r0), -1);
    for (int64_t s = S.seg_ptr[h] + 1; s < S.seg_ptr[h + 1]; s++) {
      const Seg &sg = S.segs[s];
      for (int r = 0; r < sg.hi - sg.lo; r++) rowmap[rp.map_off + (sg.off - r0) + r] = P.start[sg.anc] + sg.lo + r;
    }
    const int pi = (int)panels.size();
    panels.push_back(rp);
    for (int c0 = 0; c0 < n; c0 += kResColG) ltw.push_back(ResTile{pi, c0, 0, 0});
    for (int sl = 0; sl * kResSlab < rows; sl++)
      for (int ch = 0; ch * kResChunk < n; ch++) {
        const int rlast = std::min(rows, (sl + 1) * kResSlab) - 1;
        if (rlast < n && rlast < ch * kResChunk) continue;  // wholly above the diagonal of the pivot block
        ly.push_back(ResTile{pi, sl, ch, 0});
      }
  }
  ResDev *v = c->res = new ResDev();
  v->n_ltw = (int64_t)ltw.size(), v->n_ly = (int64_t)ly.size();
  if (upload(c, &v->panels, panels) || upload(c, &v->t_ltw, ltw) || upload(c, &v->t_ly, ly) || upload(c, &v->rowmap, rowmap)) return -100;
  CK(cudaMalloc((void **)&v->y, std::max<size_t>(1, (size_t)P.n * kResK) * sizeof(double)));
  CK(cudaMalloc((void **)&v->z, std::max<size_t>(1, (size_t)P.n * kResK) * sizeof(double)));
  return 0;
}
// this rank's part of Z = L (L^T W): the panels it reports, read where they sit; z_out is n x kResK, permuted rows
static int residual_rank(chol_t *c, int k, uint64_t seed, double *z_out) {
  if (!c->device_ready || !c->factored) return fail(c, "factor first");
  if (ensure_res(c)) return -1;
  ResDev *v = c->res;
  const size_t bytes = (size_t)c->P.n * kResK * sizeof(double);
  CK(cudaMemsetAsync(v->y, 0, bytes, c->stream));
  CK(cudaMemsetAsync(v->z, 0, bytes, c->stream));
  if (v->n_ltw) res_ltw<<<(unsigned)v->n_ltw, kResColG * 32, 0, c->stream>>>(v->panels, v->t_ltw, v->rowmap, c->d_fac, k, seed, v->y);
  if (v->n_ly) res_ly<<<(unsigned)v->n_ly, kResSlab, 0, c->stream>>>(v->panels, v->t_ly, v->rowmap, c->d_fac, k, v->y, v->z);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(z_out, v->z, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" {
/* Z_partial = L_r (L_r^T W) over the panels this handle reports, W = k <= 4 Rademacher columns generated from
 * `seed`; z_out holds n x 4 doubles (row = permuted row, first k columns used).  The sum over the ranks of a
 * partition goes to chol_residual_finish. */
int chol_residual_partial(chol_t *c, int k, uint64_t seed, double *z_out) {
  if (k < 1 || k > kResK) return fail(c, "1 <= k <= 4 probe columns");
  if (!is_group(c)) {
    cudaSetDevice(c->device);
    return residual_rank(c, k, seed, z_out);
  }
  const size_t len = (size_t)c->P.n * kResK;
  std::vector<std::vector<double>> part(c->sub.size(), std::vector<double>(len));
  if (for_ranks(c, [&](chol_t *r) { return residual_rank(r, k, seed, part[r->rank].data()); })) return -1;
  for (size_t i = 0; i < len; i++) {
    double s = 0;
    for (auto &pz : part) s += pz[i];
    z_out[i] = s;
  }
  return 0;
}
/* rel = ||A W - Z||_F / ||A W||_F with the same W (host: A W from the loaded entries, O(nz k)) */
int chol_residual_finish(chol_t *c, int k, uint64_t seed, const double *z_sum, double *rel) {
  if (k < 1 || k > kResK) return fail(c, "1 <= k <= 4 probe columns");
  const Problem &P = c->P;
  const size_t n = (size_t)P.n;
  std::vector<double> AW(n * kResK, 0.0);
  std::vector<int> iperm(n);
  for (int p = 0; p < P.n; p++) iperm[P.perm[p]] = p;
  for (int64_t e = 0; e < P.nz; e++) {
    const int pi = iperm[P.ei[e]], pj = iperm[P.ej[e]];
    const double v = P.ev[e];
    for (int q = 0; q < k; q++) {
      AW[(size_t)pi * kResK + q] += v * res_w(seed, pj, q);
      if (pi != pj) AW[(size_t)pj * kResK + q] += v * res_w(seed, pi, q);
    }
  }
  double num = 0, den = 0;
  for (size_t i = 0; i < n; i++)
    for (int q = 0; q < k; q++) {
      const double a = AW[i * kResK + q], d = a - z_sum[i * kResK + q];
      num += d * d, den += a * a;
    }
  *rel = std::sqrt(num / (den > 0 ? den : 1));
  return 0;
}
/* relative residual estimate ||(A - L L^T) W||_F / ||A W||_F on a single-GPU or group handle */
int chol_residual(chol_t *c, int k, uint64_t seed, double *rel) {
  if (!is_group(c) && c->world > 1) return fail(c, "partitioned handle: sum chol_residual_partial over the ranks, then chol_residual_finish");
  k = std::max(1, std::min(k, kResK));
  std::vector<double> z((size_t)c->P.n * kResK);
  if (chol_residual_partial(c, k, seed ? seed : 1, z.data())) return -1;
  return chol_residual_finish(c, k, seed ? seed : 1, z.data(), rel);
}

// ------------------------------------------------------------------------------ solve on the GPU
}  // extern "C"
struct SolveDev {
  bool ready = false;
  SolveSchedule V;
  SolveTile *tiles = nullptr;
  SolveGemv *gemv = nullptr;
  TileRef *gemv_tiles = nullptr, *gather_tiles = nullptr;
  PullTile *pull_tiles = nullptr;
  PullSum *pull_sums = nullptr;
  double *scratch = nullptr;
  PullDest *pull = nullptr;
  PullContrib *pull_contrib = nullptr;
  GatherDesc *gather = nullptr;
  int *rowmap = nullptr, *perm = nullptr;
  double *x = nullptr, *io = nullptr;
};
static void free_solve(chol_t *c) {
  SolveDev *v = c->solve;
  if (!v) return;
  cudaFree(v->tiles), cudaFree(v->gemv), cudaFree(v->gemv_tiles), cudaFree(v->pull_tiles), cudaFree(v->gather_tiles);
  cudaFree(v->pull), cudaFree(v->pull_contrib), cudaFree(v->gather), cudaFree(v->rowmap), cudaFree(v->perm), cudaFree(v->x), cudaFree(v->io);
  cudaFree(v->pull_sums), cudaFree(v->scratch);
  delete v;
  c->solve = nullptr;
}
static int ensure_solve(chol_t *c) {
  if (c->solve && c->solve->ready) return 0;
  free_solve(c);
  SolveDev *v = c->solve = new SolveDev();
  if (build_solve(c->P, c->S, v->V, c->rank, c->world, c->err)) return -1;
  if (upload(c, &v->tiles, v->V.tiles) || upload(c, &v->gemv, v->V.gemv) || upload(c, &v->gemv_tiles, v->V.gemv_tiles) ||
      upload(c, &v->pull, v->V.pull) || upload(c, &v->pull_contrib, v->V.pull_contrib) || upload(c, &v->pull_tiles, v->V.pull_tiles) ||
      upload(c, &v->gather, v->V.gather) || upload(c, &v->gather_tiles, v->V.gather_tiles) || upload(c, &v->rowmap, v->V.rowmap) ||
      upload(c, &v->pull_sums, v->V.pull_sums) ||
      upload(c, &v->perm, c->P.perm))
    return -100;
  CK(cudaMalloc((void **)&v->x, std::max(1, c->P.n) * sizeof(double)));
  CK(cudaMalloc((void **)&v->io, std::max(1, c->P.n) * sizeof(double)));
  CK(cudaMalloc((void **)&v->scratch, std::max<size_t>(1, (size_t)v->V.pull_slots * kSolveSlab) * sizeof(double)));
  v->ready = true;
  return 0;
}
extern "C" {

// launches [from, to) of the solve schedule on the handle's stream (every step is a CUDA kernel on the
// factor as it sits in HBM, csrc/solve_kernels.cuh)
static void run_solve_launches(chol_t *c, size_t from, size_t to) {
  SolveDev *v = c->solve;
  cudaStream_t st = c->stream;
  // CHOL_SOLVE_TIMES=1: device time by kernel class and tree level of this sweep on stderr (a measurement aid)
  static const bool timing = getenv("CHOL_SOLVE_TIMES") && atoi(getenv("CHOL_SOLVE_TIMES"));
  std::vector<cudaEvent_t> ev;
  for (size_t i = from; i < to; i++) {
    const SolveLaunch &l = v->V.launches[i];
    const unsigned g = (unsigned)l.count;
    if (timing) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, st);
      ev.push_back(e);
    }
    switch (l.kind) {
      case SK_TILE_F:
        if (l.width <= kSolveNB) solve_tile<false><<<g, kSolveTileThreads, 0, st>>>(v->tiles + l.begin, c->d_fac, v->x);
        else solve_block<false><<<g, kSolveBlockThreads, 0, st>>>(v->tiles + l.begin, c->d_fac, v->x);
        break;
      case SK_TILE_B:
        if (l.width <= kSolveNB) solve_tile<true><<<g, kSolveTileThreads, 0, st>>>(v->tiles + l.begin, c->d_fac, v->x);
        else solve_block<true><<<g, kSolveBlockThreads, 0, st>>>(v->tiles + l.begin, c->d_fac, v->x);
        break;
      case SK_GEMV_F:
        solve_gemv_fwd<<<g, kSolveSlab, 0, st>>>(v->gemv, v->gemv_tiles + l.begin, c->d_fac, v->x);
        break;
      case SK_GEMV_B:
        solve_gemv_bwd<<<g, kSolveColG * 32, 0, st>>>(v->gemv, v->gemv_tiles + l.begin, c->d_fac, v->x);
        break;
      case SK_PULL:
        solve_pull<<<g, kSolveSlab, 0, st>>>(v->pull, v->pull_contrib, v->pull_tiles + l.begin, c->d_fac, v->x, v->scratch);
        break;
      case SK_PULL_SUM:
        solve_pull_sum<<<g, kSolveSlab, 0, st>>>(v->pull_sums + l.begin, v->scratch, v->x);
        break;
      case SK_GATHER:
        solve_gather<<<g, kSolveColG * 32, 0, st>>>(v->gather, v->gather_tiles + l.begin, v->rowmap, c->d_fac, v->x);
        break;
      case SK_EXCHANGE:
        break;
    }
  }
  if (timing && !ev.empty()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    cudaStreamSynchronize(st);
    double by[8][32] = {};
    int cnt[8] = {};
    for (size_t i = from; i < to; i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i - from], ev[i - from + 1]);
      const SolveLaunch &l = v->V.launches[i];
      by[l.kind][std::min(l.level, 31)] += ms, cnt[l.kind]++;
    }
    const char *names[8] = {"tile_fwd", "gemv_fwd", "pull", "gather", "tile_bwd", "gemv_bwd", "exchange", "pull_sum"};
    for (int k = 0; k < 8; k++) {
      double tot = 0;
      for (int l = 0; l < 32; l++) tot += by[k][l];
      if (!cnt[k]) continue;
      fprintf(stderr, "solve %-9s %5d launches %8.3f ms; by level:", names[k], cnt[k], tot);
      for (int l = 0; l < c->P.levels; l++) fprintf(stderr, " %.2f", by[k][l]);
      fprintf(stderr, "\n");
    }
    for (cudaEvent_t x : ev) cudaEventDestroy(x);
  }
}
static size_t solve_exchange_index(const SolveSchedule &V) {
  for (size_t i = 0; i < V.launches.size(); i++)
    if (V.launches[i].kind == SK_EXCHANGE) return i;
  return V.launches.size();
}

// one rank, forward: the rank's subtree; its pulls leave in the top rows only this rank's contributions (rank 0
// starts them from b, the others from zero), so the SUM of top_partial over the ranks is the right-hand side
// the top levels see
static int solve_forward_rank(chol_t *c, const double *b, double *top_partial) {
  if (!c->device_ready || !c->factored) return fail(c, "factor first");
  if (ensure_solve(c)) return -1;
  SolveDev *v = c->solve;
  const int n = c->P.n, t0 = v->V.top_row0;
  cudaStream_t st = c->stream;
  CK(cudaMemcpyAsync(v->io, b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
  permute_in_kernel<<<(n + 255) / 256, 256, 0, st>>>(v->io, v->perm, n, v->x);
  if (c->rank != 0 && t0 < n) CK(cudaMemsetAsync(v->x + t0, 0, (size_t)(n - t0) * sizeof(double), st));
  run_solve_launches(c, 0, solve_exchange_index(v->V));
  CK(cudaGetLastError());
  if (t0 < n) CK(cudaMemcpyAsync(top_partial, v->x + t0, (size_t)(n - t0) * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}
// top_sum: the sum of every rank's top_partial.  x_owned (original dof order): the entries of the separators
// the rank reports, zeros elsewhere, so that the sum over the ranks is the solution.
static int solve_backward_rank(chol_t *c, const double *top_sum, double *x_owned) {
  if (!c->device_ready || !c->factored) return fail(c, "factor first");
  if (!c->solve || !c->solve->ready) return fail(c, "chol_solve_forward first");
  SolveDev *v = c->solve;
  const Problem &P = c->P;
  const int n = P.n, t0 = v->V.top_row0;
  cudaStream_t st = c->stream;
  if (t0 < n) CK(cudaMemcpyAsync(v->x + t0, top_sum, (size_t)(n - t0) * sizeof(double), cudaMemcpyHostToDevice, st));
  run_solve_launches(c, solve_exchange_index(v->V), v->V.launches.size());
  permute_out_kernel<<<(n + 255) / 256, 256, 0, st>>>(v->x, v->perm, n, v->io);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(x_owned, v->io, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (c->world > 1)
    for (int h = 1; h <= P.N; h++) {
      if (reports(c, h)) continue;
      for (int i = 0; i < P.sz[h]; i++) x_owned[P.perm[P.start[h] + i]] = 0.0;
    }
  return 0;
}

/* mmat.rg:1364-1495: permute b, forward substitution leaves -> root, backward root -> leaves, un-permute.
 * On a group handle every rank sweeps its subtree, the top rows of the right-hand side are summed over the
 * ranks on the host (chol_solve_top_size doubles per rank), every rank sweeps the top levels and its subtree
 * backwards and the owned pieces of x are merged. */
int chol_solve(chol_t *c, const double *b, double *x) {
  if (!is_group(c)) {
    if (c->world > 1)
      return fail(c, "chol_solve runs on a single-GPU or group handle; on a partitioned handle use chol_solve_forward / "
                     "chol_solve_backward with a sum of the top part over the ranks in between");
    if (!c->device_ready || !c->factored) return fail(c, "factor first");
    cudaSetDevice(c->device);
    if (ensure_solve(c)) return -1;
    SolveDev *v = c->solve;
    const int n = c->P.n;
    cudaStream_t st = c->stream;
    CK(cudaMemcpyAsync(v->io, b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    permute_in_kernel<<<(n + 255) / 256, 256, 0, st>>>(v->io, v->perm, n, v->x);
    run_solve_launches(c, 0, v->V.launches.size());
    permute_out_kernel<<<(n + 255) / 256, 256, 0, st>>>(v->x, v->perm, n, v->io);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(x, v->io, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
  }
  if (!c->factored) return fail(c, "factor first");
  const int nr = (int)c->sub.size(), n = c->P.n;
  const int64_t nt = chol_solve_top_size(c->sub[0]);
  std::vector<std::vector<double>> top(nr, std::vector<double>((size_t)std::max<int64_t>(nt, 1))), xs(nr, std::vector<double>((size_t)n));
  if (for_ranks(c, [&](chol_t *r) { return solve_forward_rank(r, b, top[r->rank].data()); })) return -1;
  for (int r = 1; r < nr; r++)
    for (int64_t i = 0; i < nt; i++) top[0][i] += top[r][i];
  if (for_ranks(c, [&](chol_t *r) { return solve_backward_rank(r, top[0].data(), xs[r->rank].data()); })) return -1;
  for (int i = 0; i < n; i++) {
    double v = 0;
    for (int r = 0; r < nr; r++) v += xs[r][i];
    x[i] = v;
  }
  return 0;
}

/* The same sweeps on a partitioned handle (one process per GPU, section 6 of DESIGN.md). */
int64_t chol_solve_top_size(chol_t *c) {
  if (!c->analyzed) return -1;
  c = first_rank(c);
  if (c->world == 1) return 0;
  return (int64_t)c->P.n - c->P.start[(1 << c->D.depth) - 1];
}
int chol_solve_forward(chol_t *c, const double *b, double *top_partial) {
  if (is_group(c)) return fail(c, "chol_solve does this on a group handle");
  cudaSetDevice(c->device);
  return solve_forward_rank(c, b, top_partial);
}
int chol_solve_backward(chol_t *c, const double *top_sum, double *x_owned) {
  if (is_group(c)) return fail(c, "chol_solve does this on a group handle");
  cudaSetDevice(c->device);
  return solve_backward_rank(c, top_sum, x_owned);
}
/* what the rank's solve schedule covers, [0..5] on its subtree levels and [6..11] on the shared top levels:
 * forward tiles, forward gemv slabs, pull slabs, gather column groups, backward tiles, backward gemv groups */
int chol_solve_stats(chol_t *c, double *out12) {
  if (!c->analyzed) return fail(c, "analyze first");
  c = first_rank(c);
  SolveSchedule V;
  if (build_solve(c->P, c->S, V, c->rank, c->world, c->err)) return -1;
  for (int i = 0; i < 12; i++) out12[i] = 0;
  for (const SolveLaunch &l : V.launches) {
    if (l.kind == SK_EXCHANGE) continue;
    if (l.kind == SK_PULL_SUM) continue;
    const int slot = l.kind == SK_TILE_F ? 0 : l.kind == SK_GEMV_F ? 1 : l.kind == SK_PULL ? 2 : l.kind == SK_GATHER ? 3 : l.kind == SK_TILE_B ? 4 : 5;
    const bool top = c->world > 1 && l.level < V.depth;
    out12[slot + (top ? 6 : 0)] += (double)l.count;
  }
  return 0;
}

/* verification helper (host): y = A x with the loaded lower-triangle entries, original dof order */
int chol_matvec(chol_t *c, const double *x, double *y) {
  if (!c->loaded) return fail(c, "nothing loaded");
  const Problem &P = c->P;
  for (int i = 0; i < P.n; i++) y[i] = 0.0;
  for (int64_t e = 0; e < P.nz; e++) {
    const int i = P.ei[e], j = P.ej[e];
    y[i] += P.ev[e] * x[j];
    if (i != j) y[j] += P.ev[e] * x[i];
  }
  return 0;
}

int chol_read_vector(const char *path, int n, double *out) { return mnd_read_vector(path, n, out); }
int chol_write_solution(const char *path, int n, const double *x) { /* mmat.rg:785-798 */
  FILE *f = fopen(path, "w");
  if (!f) return -1;
  for (int i = 0; i < n; i++) fprintf(f, "%0.8g\n", x[i]);
  fclose(f);
  return 0;
}

}  // extern "C"
