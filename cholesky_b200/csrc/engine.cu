// C ABI of the engine (include/cholesky.h): handle, loaders, host symbolic analysis, GPU numeric
// factorization driven by the compiled level schedule, result access.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/chol_mmio.h"
#include "../../include/chol_mnd.h"
#include "../../include/cholesky.h"
#include "chol_internal.h"
#include "kernels.cuh"
#include "solve.h"
#include "solve_kernels.cuh"

using namespace chb;

constexpr int kMaxDevices = 16;


struct SolveDev;
struct chol {
  std::string err;
  SolveDev *solve = nullptr;  // solve schedule and buffers, built on first use
  int device = 0;
  Problem P;
  Symbolic S;
  Schedule D;
  bool loaded = false, analyzed = false, device_ready = false, assembled = false;
  cudaStream_t stream = nullptr;
  double *d_fac = nullptr;
  double *d_vals = nullptr;
  int64_t *d_aoff = nullptr;
  GemmProblem *d_probs = nullptr;
  GemmContrib *d_contribs = nullptr;
  TileRef *d_tiles = nullptr;
  PotrfDesc *d_potrf = nullptr;
  TrsmDesc *d_trsm = nullptr;
  TileRef *d_trsm_tiles = nullptr;
  int *d_info = nullptr;
  int64_t *d_diag_off = nullptr;
  double *d_diag = nullptr;
  double *h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  std::vector<double> h_fac;  // host copy of the factor, fetched lazily
  bool h_fac_valid = false;
  cudaStream_t stream1 = nullptr, cur = nullptr;  // chain stream (look-ahead); stream of the launch being issued
  std::vector<cudaEvent_t> evs;                   // cross-stream events of the launch list
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  double k_ms[6] = {0, 0, 0, 0, 0, 0};
  double k_gemm_flops = 0;
  std::vector<float> launch_ms;  // per launch, from the last instrumented pass
  // multi-GPU: one handle per rank; peers' factor buffers and flag words mapped through CUDA IPC
  int rank = 0, world = 1;
  Peers peers = {};
  unsigned long long *d_flags = nullptr;
  unsigned long long epoch = 0;
  bool peers_ready = false;
  // CHOL_GRAPH=1 (experimental, not measured yet): the whole level loop of a single-GPU handle is captured once
  // into a CUDA graph (both streams and their events) and replayed; for the launch-bound workloads (512^2: 400
  // launches in 3 ms).  graph_runs counts the eager runs since the schedule was uploaded: the first one stays
  // eager (one-time function attributes), the second is captured.
  int use_graph = 0, graph_runs = 0;
  cudaGraphExec_t graph_exec = nullptr;
  int gemm_stages = 3;      // CHOL_GEMM_STAGES: 4 = experimental four-stage operand ring for the 64x64 tiles (not measured yet)
  int trsm_batch = 0;       // CHOL_TRSM_BATCH: 1 = experimental trsm_tile<true> (all slab loads in flight at once)
  int potrf_r = 0;          // CHOL_POTRF_R: 1 / 2 / 3 = experimental register-resident right-looking pivot tile (potrf_tile_r / _r2)
  int potrf_w = 1;          // CHOL_POTRF_W: 1 = single-warp column steps (potrf_tile_w), 0 = 64-thread version
  int gemm_ws = 1;          // CHOL_GEMM_WS: 1 = warp-specialised TMA bulk-copy kernel (default, 8% faster on
                            // 128^3), 0 = the earlier cp.async kernel
  std::vector<void *> ipc_opened;
};

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

extern "C" {

void register_mappers(void) {}

int chol_create(const int *devices, int ngpu, chol_t **out) {
  if (!out) return -1;
  chol_t *c = new chol();
  c->device = (devices && ngpu > 0) ? devices[0] : 0;
  if (const char *e = getenv("CHOL_GEMM_WS")) c->gemm_ws = atoi(e);
  if (const char *e = getenv("CHOL_POTRF_W")) c->potrf_w = atoi(e);
  if (const char *e = getenv("CHOL_POTRF_R")) c->potrf_r = atoi(e);
  if (const char *e = getenv("CHOL_TRSM_BATCH")) c->trsm_batch = atoi(e);
  if (const char *e = getenv("CHOL_GEMM_STAGES")) c->gemm_stages = atoi(e);
  if (const char *e = getenv("CHOL_GRAPH")) c->use_graph = atoi(e);
  *out = c;
  return 0;
}

static void free_solve(chol_t *c);
static void drop_graph(chol_t *c);
static void free_device(chol_t *c) {
  if (!c->device_ready) return;
  drop_graph(c);
  free_solve(c);
  cudaSetDevice(c->device);
  cudaFree(c->d_fac), cudaFree(c->d_vals), cudaFree(c->d_aoff), cudaFree(c->d_probs), cudaFree(c->d_contribs);
  cudaFree(c->d_tiles), cudaFree(c->d_potrf), cudaFree(c->d_trsm), cudaFree(c->d_trsm_tiles), cudaFree(c->d_info);
  cudaFree(c->d_diag_off), cudaFree(c->d_diag);
  for (void *p : c->ipc_opened) cudaIpcCloseMemHandle(p);
  c->ipc_opened.clear();
  cudaFree(c->d_flags);
  c->d_flags = nullptr, c->peers_ready = false;
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  c->h_pinned = nullptr, c->h_pinned_bytes = 0;
  for (cudaEvent_t e : c->evs) cudaEventDestroy(e);
  c->evs.clear();
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  c->ev_fork = c->ev_join = nullptr;
  if (c->stream1) cudaStreamDestroy(c->stream1);
  c->stream1 = nullptr;
  if (c->stream) cudaStreamDestroy(c->stream);
  c->device_ready = false;
}

void chol_destroy(chol_t *c) {
  if (!c) return;
  free_device(c);
  delete c;
}
const char *chol_last_error(chol_t *c) { return c ? c->err.c_str() : "null handle"; }

int chol_load(chol_t *c, const char *mtx, const char *ord, const char *clust) {
  c->loaded = c->analyzed = false;
  if (read_problem(c->P, mtx, ord, clust, c->err)) return -1;
  c->loaded = true;
  return 0;
}

int chol_load_arrays(chol_t *c, int n, int64_t nz, const int32_t *I, const int32_t *J, const double *V, int levels, int nsep,
                     const int64_t *sep_ptr, const int32_t *sep_dofs, const int64_t *sep_iv_ptr, const int64_t *iv_ptr,
                     const int32_t *iv_vals) {
  c->loaded = c->analyzed = false;
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
  c->loaded = c->analyzed = false;
  if (generate_problem(c->P, nx, ny, nz, stencil, levels, c->err)) return -1;
  c->loaded = true;
  return 0;
}

int chol_write_inputs(chol_t *c, const char *mtx, const char *ord, const char *clust) {
  if (!c->loaded) return fail(c, "nothing loaded");
  return write_problem(c->P, mtx, ord, clust, c->err);
}

int chol_analyze(chol_t *c, int keep_records) {
  if (!c->loaded) return fail(c, "load a problem first");
  free_device(c);
  c->analyzed = false;
  if (analyze(c->P, c->S, keep_records != 0, c->err)) return -1;
  if (build_schedule(c->P, c->S, c->D, c->rank, c->world, false, c->err)) return -1;
  c->analyzed = true;
  c->assembled = false;
  c->h_fac_valid = false;
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

extern "C" {
static void drop_graph(chol_t *c) {
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  c->graph_exec = nullptr, c->graph_runs = 0;
}
static int upload_schedule(chol_t *c) {
  drop_graph(c);  // the captured launches point into the descriptor arrays replaced below
  cudaFree(c->d_probs), cudaFree(c->d_contribs), cudaFree(c->d_tiles), cudaFree(c->d_potrf), cudaFree(c->d_trsm), cudaFree(c->d_trsm_tiles);
  c->d_probs = nullptr, c->d_contribs = nullptr, c->d_tiles = nullptr, c->d_potrf = nullptr, c->d_trsm = nullptr, c->d_trsm_tiles = nullptr;
  if (upload(c, &c->d_probs, c->D.probs)) return -100;
  if (upload(c, &c->d_contribs, c->D.contribs)) return -100;
  if (upload(c, &c->d_tiles, c->D.tiles)) return -100;
  if (upload(c, &c->d_potrf, c->D.potrf)) return -100;
  if (upload(c, &c->d_trsm, c->D.trsm)) return -100;
  if (upload(c, &c->d_trsm_tiles, c->D.trsm_tiles)) return -100;
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

static int ensure_device(chol_t *c) {
  if (!c->analyzed) return fail(c, "analyze first");
  if (c->device_ready) return 0;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(c, "no CUDA device: the numeric factorization runs on the GPU only (no CPU fallback)");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {  // the chain stream gets the highest priority: its few CTAs must slip in between the CTAs of a
     // trailing update that fills the GPU, not queue behind them
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->stream1, cudaStreamNonBlocking, hi));
  }
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  c->cur = c->stream;
  CK(cudaMalloc((void **)&c->d_fac, (size_t)c->S.total_doubles * sizeof(double)));
  c->device_ready = true;
  if (upload(c, &c->d_vals, c->P.ev)) return -100;
  if (upload(c, &c->d_aoff, c->D.a_off)) return -100;
  if (upload_schedule(c)) return -100;
  CK(cudaMalloc((void **)&c->d_info, sizeof(int)));
  std::vector<int64_t> doff(c->P.n);
  for (int h = 1; h <= c->P.N; h++)
    for (int i = 0; i < c->P.sz[h]; i++) doff[c->P.start[h] + i] = c->S.poff[h] + i + (int64_t)i * c->S.ld[h];
  if (upload(c, &c->d_diag_off, doff)) return -100;
  CK(cudaMalloc((void **)&c->d_diag, std::max(1, c->P.n) * sizeof(double)));
  CK(cudaFuncSetAttribute(trsm_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmemBytes));
  CK(cudaFuncSetAttribute(trsm_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTrsmSmemBytes));
  CK(cudaMalloc((void **)&c->d_flags, kMaxPeers * sizeof(unsigned long long)));
  CK(cudaMemset(c->d_flags, 0, kMaxPeers * sizeof(unsigned long long)));
  c->peers.n = 1, c->peers.rank = 0;
  c->peers.fac[0] = c->d_fac, c->peers.flags[0] = c->d_flags;
  c->peers_ready = (c->world == 1);
  c->epoch = 0;
  return 0;
}

static int do_assemble(chol_t *c) {
  CK(cudaMemsetAsync(c->d_fac, 0, (size_t)c->S.total_doubles * sizeof(double), c->stream));
  int info0 = 0x7fffffff;
  CK(cudaMemcpyAsync(c->d_info, &info0, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (c->P.nz > 0) {
    int64_t nzv = c->P.nz;
    assemble_kernel<<<(unsigned)((nzv + 255) / 256), 256, 0, c->stream>>>(c->d_vals, c->d_aoff, nzv, c->d_fac);
  }
  CK(cudaGetLastError());
  c->assembled = true;
  c->h_fac_valid = false;
  return 0;
}

}  // extern "C"
template <int BM, int BN, int BK, int WM, int WN, int ST>
static void launch_gemm(chol_t *c, const Launch &l) {
  using Cfg = GemmCfg<BM, BN, BK, WM, WN, ST>;
  static bool attr[kMaxDevices][2] = {};  // the opt-in shared-memory size is a per-device function attribute
  bool &done = attr[c->device % kMaxDevices][(l.shared == 1) ? 1 : 0];
  if (!done) {
    if (l.shared == 1) cudaFuncSetAttribute(gemm_grouped<BM, BN, BK, WM, WN, ST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    else cudaFuncSetAttribute(gemm_grouped<BM, BN, BK, WM, WN, ST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    done = true;
  }
  if (l.shared == 1)
    gemm_grouped<BM, BN, BK, WM, WN, ST, true><<<(unsigned)l.count, Cfg::kThreads, Cfg::kSmemBytes, c->cur>>>(
        c->d_probs, c->d_contribs, c->d_tiles + l.begin, c->d_fac, c->peers);
  else
    gemm_grouped<BM, BN, BK, WM, WN, ST, false><<<(unsigned)l.count, Cfg::kThreads, Cfg::kSmemBytes, c->cur>>>(
        c->d_probs, c->d_contribs, c->d_tiles + l.begin, c->d_fac, c->peers);
}
template <int BM, int BN, int BK, int WM, int WN, int ST, int MINB>
static void launch_gemm_ws(chol_t *c, const Launch &l) {
  using Cfg = GemmWsCfg<BM, BN, BK, WM, WN, ST>;
  static bool attr[kMaxDevices][2] = {};
  bool &done = attr[c->device % kMaxDevices][(l.shared == 1) ? 1 : 0];
  if (!done) {
    if (l.shared == 1) cudaFuncSetAttribute(gemm_grouped_ws<BM, BN, BK, WM, WN, ST, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    else cudaFuncSetAttribute(gemm_grouped_ws<BM, BN, BK, WM, WN, ST, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    done = true;
  }
  if (l.shared == 1)
    gemm_grouped_ws<BM, BN, BK, WM, WN, ST, MINB, true><<<(unsigned)l.count, Cfg::kThreads, Cfg::kSmemBytes, c->cur>>>(
        c->d_probs, c->d_contribs, c->d_tiles + l.begin, c->d_fac, c->peers);
  else
    gemm_grouped_ws<BM, BN, BK, WM, WN, ST, MINB, false><<<(unsigned)l.count, Cfg::kThreads, Cfg::kSmemBytes, c->cur>>>(
        c->d_probs, c->d_contribs, c->d_tiles + l.begin, c->d_fac, c->peers);
}
extern "C" {
static void launch_barrier(chol_t *c) {
  c->epoch++;
  peer_barrier<<<1, 32, 0, c->cur>>>(c->peers, c->epoch);
}

static int run_launch(chol_t *c, const Launch &l) {
  switch (l.kind) {
    case K_POTRF:
      if (c->potrf_r == 3) potrf_tile_r2<3><<<(unsigned)l.count, 2 * kNB, 0, c->cur>>>(c->d_potrf + l.begin, c->d_fac, c->d_info);
      else if (c->potrf_r == 2) potrf_tile_r2<1><<<(unsigned)l.count, 2 * kNB, 0, c->cur>>>(c->d_potrf + l.begin, c->d_fac, c->d_info);
      else if (c->potrf_r) potrf_tile_r<<<(unsigned)l.count, kPotrfRThreads, 0, c->cur>>>(c->d_potrf + l.begin, c->d_fac, c->d_info);
      else if (c->potrf_w) potrf_tile_w<<<(unsigned)l.count, kPotrfThreads, 0, c->cur>>>(c->d_potrf + l.begin, c->d_fac, c->d_info);
      else potrf_tile<<<(unsigned)l.count, kPotrfThreads, 0, c->cur>>>(c->d_potrf + l.begin, c->d_fac, c->d_info);
      break;
    case K_TRSM:
      if (c->trsm_batch) trsm_tile<true><<<(unsigned)l.count, kSlab, kTrsmSmemBytes, c->cur>>>(c->d_trsm, c->d_trsm_tiles + l.begin, c->d_fac);
      else trsm_tile<false><<<(unsigned)l.count, kSlab, kTrsmSmemBytes, c->cur>>>(c->d_trsm, c->d_trsm_tiles + l.begin, c->d_fac);
      break;
    case K_GEMM:
      if (l.count <= 0) break;
      if (l.cfg == 3)
        gemm_small_warp<<<(unsigned)((l.count + kSmallWarps - 1) / kSmallWarps), kSmallWarps * 32, 0, c->cur>>>(
            c->d_probs, c->d_contribs, c->d_tiles + l.begin, l.count, c->d_fac);
      else if (l.cfg == 2)
        launch_gemm_ws<128, 64, 16, 32, 32, 4, 2>(c, l);
      else if (l.cfg == 1)
        launch_gemm_ws<128, 128, 16, 32, 32, 4, 1>(c, l);
      else if (c->gemm_ws && c->gemm_stages == 4)  // experimental: four-stage ring (power-of-two indexing), 3 CTAs/SM
        launch_gemm_ws<64, 64, 16, 32, 32, 4, 3>(c, l);
      else if (c->gemm_ws)
        launch_gemm_ws<64, 64, 16, 32, 32, 3, 4>(c, l);
      else
        launch_gemm<64, 64, 16, 32, 32, 3>(c, l);
      break;
    case K_BARRIER:
      launch_barrier(c);
      break;
    case K_ALLREDUCE:  // one top panel: shared = its heap index, cfg = mask of contributing ranks
      allreduce_top<<<148 * 4, 256, 0, c->cur>>>(c->peers, l.begin / 2, l.count / 2, c->S.ld[l.shared] / 2, c->P.sz[l.shared],
                                                 (unsigned)l.cfg);
      break;
    case K_NOP:
      break;
  }
  return 0;
}

static int run_levels(chol_t *c, int lvl_from, int lvl_to, int phase_mask, bool per_kernel_timing) {
  if (c->world > 1 && !c->peers_ready) return fail(c, "multi-GPU handle: exchange IPC handles first (chol_ipc_export / chol_ipc_import)");
  std::vector<cudaEvent_t> ev;
  std::vector<int> kinds;
  std::vector<double> fl;
  // cross-stream events of the launch list (look-ahead); the chain stream starts after whatever is
  // already queued on the main stream (assembly)
  while ((int)c->evs.size() < c->D.num_events) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->evs.push_back(e);
  }
  // graph replay / capture: whole level loop, single GPU, uninstrumented (see chol::use_graph)
  const bool whole = lvl_from >= c->P.levels - 1 && lvl_to <= 0 && phase_mask == 7;
  const bool graphable = c->use_graph && c->world == 1 && whole && !per_kernel_timing;
  if (graphable && c->graph_exec) {
    CK(cudaGraphLaunch(c->graph_exec, c->stream));
    return 0;
  }
  const bool capture = graphable && c->graph_runs >= 1;
  if (graphable) c->graph_runs++;
  if (capture) CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  if (c->D.lookahead) {
    CK(cudaEventRecord(c->ev_fork, c->stream));
    CK(cudaStreamWaitEvent(c->stream1, c->ev_fork, 0));
  }
  for (const Launch &l : c->D.launches) {
    if (l.level > lvl_from || l.level < lvl_to) continue;
    if (l.kind != K_NOP && !(l.phase & phase_mask)) continue;
    // the instrumented pass runs everything on one stream so that an event pair brackets one kernel alone
    c->cur = (l.stream && !per_kernel_timing) ? c->stream1 : c->stream;
    if (l.wait_ev >= 0 && !per_kernel_timing) cudaStreamWaitEvent(c->cur, c->evs[l.wait_ev], 0);
    if (per_kernel_timing) {
      cudaEvent_t a, b;
      cudaEventCreate(&a), cudaEventCreate(&b);
      cudaEventRecord(a, c->cur);
      run_launch(c, l);
      cudaEventRecord(b, c->cur);
      ev.push_back(a), ev.push_back(b), kinds.push_back(l.kind), fl.push_back(l.flops);
    } else
      run_launch(c, l);
    if (l.rec_ev >= 0 && !per_kernel_timing) cudaEventRecord(c->evs[l.rec_ev], c->cur);
  }
  c->cur = c->stream;
  if (c->D.lookahead) {  // partial runs (piecewise calls) may leave work on the chain stream: join it
    CK(cudaEventRecord(c->ev_join, c->stream1));
    CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  }
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
      cudaEventDestroy(ev[2 * i]), cudaEventDestroy(ev[2 * i + 1]);
    }
  }
  return 0;
}

int chol_assemble(chol_t *c) {
  if (ensure_device(c)) return -1;
  if (do_assemble(c)) return -1;
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// kernels one step launches (the all-reduce is a barrier, the reduction kernel and a barrier)
static int64_t count_kernels(chol_t *c) {
  int64_t k = 0;
  for (const Launch &l : c->D.launches) k += l.kind == K_NOP ? 0 : (l.kind == K_GEMM && l.count <= 0) ? 0 : 1;
  return k;
}

static int fetch_info(chol_t *c, int *info) {
  int v = 0;
  CK(cudaMemcpyAsync(&v, c->d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *info = (v == 0x7fffffff) ? 0 : v;
  return 0;
}

int chol_factor(chol_t *c, int iterations, int warmup, chol_stats_t *st) {
  if (ensure_device(c)) return -1;
  if (ensure_schedule(c, false)) return -1;
  if (iterations < 1) iterations = 1;
  std::vector<double> secs;
  double asm_s = 0;
  cudaEvent_t e0, e1, e2;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventCreate(&e2));
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
  cudaEventDestroy(e0), cudaEventDestroy(e1), cudaEventDestroy(e2);
  int info = 0;
  if (fetch_info(c, &info)) return -1;
  if (st) {
    std::vector<double> s = secs;
    std::sort(s.begin(), s.end());
    st->seconds_best = s.front();
    st->seconds_median = s[s.size() / 2];
    st->seconds_last = secs.back();
    st->assemble_seconds = asm_s;
    st->flops = c->S.flops();
    st->kernel_launches = count_kernels(c);
    st->info = info;
  }
  if (info != 0) return fail(c, "matrix is not positive definite: pivot at permuted column " + std::to_string(info));
  return 0;
}

static int piecewise(chol_t *c, int lvl, int phase) {
  if (ensure_device(c)) return -1;
  if (!c->assembled) return fail(c, "assemble first");
  if (lvl < 0 || lvl >= c->P.levels) return fail(c, "bad level");
  if (ensure_schedule(c, phase != PH_UPDATE || c->D.split_phases)) return -1;
  c->h_fac_valid = false;
  if (run_levels(c, lvl, lvl, phase, false)) return -1;
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
int chol_fused_dpotrf(chol_t *c, int lvl) { return piecewise(c, lvl, PH_POTRF); }
int chol_fused_dtrsm(chol_t *c, int lvl) { return piecewise(c, lvl, PH_TRSM); }
int chol_fused_update(chol_t *c, int lvl) { return piecewise(c, lvl, PH_UPDATE); }

int chol_factor_host(chol_t *c, const double *values, int64_t nz, double *diag_out, chol_stats_t *st) {
  if (ensure_device(c)) return -1;
  if (ensure_schedule(c, false)) return -1;
  if (values && nz != c->P.nz) return fail(c, "value count differs from the loaded pattern");
  size_t need = std::max((size_t)c->P.nz, (size_t)c->P.n) * sizeof(double);
  if (c->h_pinned_bytes < need) {
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    CK(cudaMallocHost((void **)&c->h_pinned, need));
    c->h_pinned_bytes = need;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
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
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  if (diag_out) memcpy(diag_out, c->h_pinned, (size_t)c->P.n * sizeof(double));
  int info = 0;
  if (fetch_info(c, &info)) return -1;
  if (st) {
    st->seconds_best = st->seconds_median = st->seconds_last = ms * 1e-3;
    st->assemble_seconds = 0;
    st->flops = c->S.flops();
    st->kernel_launches = count_kernels(c) + 2;
    st->info = info;
  }
  if (info != 0) return fail(c, "matrix is not positive definite: pivot at permuted column " + std::to_string(info));
  return 0;
}

/* ---- multi-GPU plumbing */
int chol_set_partition(chol_t *c, int rank, int world) {
  if (world < 1 || world > kMaxPeers || (world & (world - 1)) || rank < 0 || rank >= world) return fail(c, "world must be 1, 2, 4 or 8 and 0 <= rank < world");
  c->rank = rank, c->world = world;
  c->analyzed = false;
  return 0;
}
int chol_ipc_export(chol_t *c, void *handles128) {
  if (ensure_device(c)) return -1;
  cudaIpcMemHandle_t h[2];
  CK(cudaIpcGetMemHandle(&h[0], c->d_fac));
  CK(cudaIpcGetMemHandle(&h[1], c->d_flags));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(handles128, h, 128);
  return 0;
}
int chol_ipc_import(chol_t *c, const void *all_handles, int world) {
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
/* what this rank's schedule covers: [0] matrix entries it assembles, [1] GEMM flops it executes,
 * [2] shared (tile-split) launches, [3] doubles of the shared top region, [4] potrf tiles, [5] trsm slabs */
int chol_partition_stats(chol_t *c, double *out6) {
  if (!c->analyzed) return fail(c, "analyze first");
  double a = 0, f = 0, sh = 0, pt = 0, ts = 0;
  for (int64_t o : c->D.a_off) a += (o >= 0);
  for (const Launch &l : c->D.launches) {
    if (l.kind == K_GEMM) f += l.flops, sh += (l.shared == 1);
    if (l.kind == K_POTRF) pt += (double)l.count;
    if (l.kind == K_TRSM) ts += (double)l.count;
  }
  out6[0] = a, out6[1] = f, out6[2] = sh, out6[3] = (double)c->D.top_doubles, out6[4] = pt, out6[5] = ts;
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
      dest += 16.0 * ((double)g.M * g.N - (g.tri ? 0.5 * g.N * (g.N - 1.0) : 0.0));
    }
  }
  out3[0] = panel, out3[1] = oper, out3[2] = dest;
  return 0;
}
int chol_rank(chol_t *c) { return c->rank; }
int chol_world(chol_t *c) { return c->world; }

/* per-launch device time (ms) of the last chol_kernel_times pass, in launch-list order */
int64_t chol_launch_times(chol_t *c, float *ms, int64_t cap) {
  int64_t n = std::min<int64_t>(cap, (int64_t)c->launch_ms.size());
  for (int64_t i = 0; i < n; i++) ms[i] = c->launch_ms[i];
  return (int64_t)c->launch_ms.size();
}
int64_t chol_num_launches(chol_t *c) { return c->analyzed ? (int64_t)c->D.launches.size() : -1; }
int chol_get_launch(chol_t *c, int64_t i, int *kind, int *level, int *phase, int64_t *ctas, double *flops, int *cfg) {
  if (!c->analyzed || i < 0 || i >= (int64_t)c->D.launches.size()) return -1;
  const Launch &l = c->D.launches[i];
  *kind = l.kind, *level = l.level, *phase = l.phase, *ctas = l.count, *flops = l.flops, *cfg = l.cfg | (l.shared << 4);
  return 0;
}

int chol_synchronize(chol_t *c) {
  if (!c->device_ready) return 0;
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int chol_kernel_times(chol_t *c, double *potrf_ms, double *trsm_ms, double *gemm_ms, double *gemm_flops) {
  if (ensure_device(c)) return -1;
  if (ensure_schedule(c, false)) return -1;
  if (do_assemble(c)) return -1;
  if (run_levels(c, c->P.levels - 1, 0, 7, true)) return -1;
  if (potrf_ms) *potrf_ms = c->k_ms[K_POTRF];
  if (trsm_ms) *trsm_ms = c->k_ms[K_TRSM];
  if (gemm_ms) *gemm_ms = c->k_ms[K_GEMM];
  if (gemm_flops) *gemm_flops = c->k_gemm_flops;
  return 0;
}

// ------------------------------------------------------------------------------ results
static int fetch_factor(chol_t *c) {
  if (!c->device_ready || !c->assembled) return fail(c, "nothing factored yet");
  if (c->h_fac_valid) return 0;
  c->h_fac.resize((size_t)c->S.total_doubles);
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(c->h_fac.data(), c->d_fac, (size_t)c->S.total_doubles * sizeof(double), cudaMemcpyDeviceToHost));
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
      if (c->world > 1) {  // a rank reports its own subtree; rank 0 also the shared top panels
        int lvc = P.level_of(hc), own = lvc < c->D.depth ? -1 : (hc >> (lvc - c->D.depth)) - (1 << c->D.depth);
        if (!(own == c->rank || (own < 0 && c->rank == 0))) continue;
      }
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
  if (info != 0) return fail(c, "matrix is not positive definite: pivot at permuted column " + std::to_string(info));
  return 0;
}

int chol_residual(chol_t *c, int k, uint64_t seed, double *rel) {
  // host evaluation over the stored pattern: R = A W - L (L^T W)
  if (fetch_factor(c)) return -1;
  const Problem &P = c->P;
  const Symbolic &S = c->S;
  size_t n = (size_t)P.n;
  if (k < 1) k = 1;
  std::vector<double> W(n * k), Y(n * k, 0.0), Z(n * k, 0.0), AW(n * k, 0.0);
  uint64_t s = seed ? seed : 1;
  for (auto &w : W) {
    s = mix64(s);
    w = (s & 1) ? 1.0 : -1.0;
  }
  // W is indexed by permuted row
  std::vector<int> iperm(n);
  for (int p = 0; p < P.n; p++) iperm[P.perm[p]] = p;
  for (int64_t e = 0; e < P.nz; e++) {
    int pi = iperm[P.ei[e]], pj = iperm[P.ej[e]];
    double v = P.ev[e];
    for (int q = 0; q < k; q++) {
      AW[(size_t)pi * k + q] += v * W[(size_t)pj * k + q];
      if (pi != pj) AW[(size_t)pj * k + q] += v * W[(size_t)pi * k + q];
    }
  }
  // Y = L^T W (by columns of L), Z = L Y
  for (int pass = 0; pass < 2; pass++)
    for (int hc = 1; hc <= P.N; hc++) {
      const double *pan = c->h_fac.data() + S.poff[hc];
      int ld = S.ld[hc];
      for (int64_t sgi = S.seg_ptr[hc]; sgi < S.seg_ptr[hc + 1]; sgi++) {
        const Seg &sg = S.segs[sgi];
        bool diag = sg.anc == hc;
        for (int col = 0; col < P.sz[hc]; col++) {
          size_t gc = (size_t)P.start[hc] + col;
          for (int r = diag ? col : 0; r < sg.hi - sg.lo; r++) {
            double v = pan[sg.off + r + (size_t)col * ld];
            if (v == 0) continue;
            size_t gr = (size_t)P.start[sg.anc] + sg.lo + r;
            for (int q = 0; q < k; q++) {
              if (pass == 0) Y[gc * k + q] += v * W[gr * k + q];
              else Z[gr * k + q] += v * Y[gc * k + q];
            }
          }
        }
      }
    }
  double num = 0, den = 0;
  for (size_t i = 0; i < n * k; i++) num += (AW[i] - Z[i]) * (AW[i] - Z[i]), den += AW[i] * AW[i];
  *rel = std::sqrt(num / (den > 0 ? den : 1));
  return 0;
}

// ------------------------------------------------------------------------------ solve on the GPU
}  // extern "C"
struct SolveDev {
  bool ready = false;
  SolveSchedule V;
  SolveTile *tiles = nullptr;
  SolveGemv *gemv = nullptr;
  TileRef *gemv_tiles = nullptr, *pull_tiles = nullptr, *gather_tiles = nullptr;
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
      upload(c, &v->perm, c->P.perm))
    return -100;
  CK(cudaMalloc((void **)&v->x, std::max(1, c->P.n) * sizeof(double)));
  CK(cudaMalloc((void **)&v->io, std::max(1, c->P.n) * sizeof(double)));
  v->ready = true;
  return 0;
}
extern "C" {

// launches [from, to) of the solve schedule on the handle's stream (every step is a CUDA kernel on the
// factor as it sits in HBM, csrc/solve_kernels.cuh)
static void run_solve_launches(chol_t *c, size_t from, size_t to) {
  SolveDev *v = c->solve;
  cudaStream_t st = c->stream;
  for (size_t i = from; i < to; i++) {
    const SolveLaunch &l = v->V.launches[i];
    const unsigned g = (unsigned)l.count;
    switch (l.kind) {
      case SK_TILE_F:
        solve_tile<false><<<g, kSolveNB, 0, st>>>(v->tiles + l.begin, c->d_fac, v->x);
        break;
      case SK_TILE_B:
        solve_tile<true><<<g, kSolveNB, 0, st>>>(v->tiles + l.begin, c->d_fac, v->x);
        break;
      case SK_GEMV_F:
        solve_gemv_fwd<<<g, kSolveSlab, 0, st>>>(v->gemv, v->gemv_tiles + l.begin, c->d_fac, v->x);
        break;
      case SK_GEMV_B:
        solve_gemv_bwd<<<g, kSolveColG * 32, 0, st>>>(v->gemv, v->gemv_tiles + l.begin, c->d_fac, v->x);
        break;
      case SK_PULL:
        solve_pull<<<g, kSolveSlab, 0, st>>>(v->pull, v->pull_contrib, v->pull_tiles + l.begin, c->d_fac, v->x);
        break;
      case SK_GATHER:
        solve_gather<<<g, kSolveColG * 32, 0, st>>>(v->gather, v->gather_tiles + l.begin, v->rowmap, c->d_fac, v->x);
        break;
      case SK_EXCHANGE:
        break;
    }
  }
}
static size_t solve_exchange_index(const SolveSchedule &V) {
  for (size_t i = 0; i < V.launches.size(); i++)
    if (V.launches[i].kind == SK_EXCHANGE) return i;
  return V.launches.size();
}

/* mmat.rg:1364-1495: permute b, forward substitution leaves -> root, backward root -> leaves, un-permute. */
int chol_solve(chol_t *c, const double *b, double *x) {
  if (!c->device_ready || !c->assembled) return fail(c, "factor first");
  if (c->world > 1)
    return fail(c, "chol_solve runs on a single-GPU handle; on a partitioned handle use chol_solve_forward / chol_solve_backward "
                   "with a sum of the top part over the ranks in between");
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

/* The same sweeps on a partitioned handle (one process per GPU, section 6 of DESIGN.md).  Forward: the rank's
 * subtree; its pulls leave in the top rows only this rank's contributions (rank 0 starts them from b, the
 * others from zero), so the SUM of top_partial over the ranks is the right-hand side the top levels see. */
int64_t chol_solve_top_size(chol_t *c) {
  if (!c->analyzed) return -1;
  if (c->world == 1) return 0;
  return (int64_t)c->P.n - c->P.start[(1 << c->D.depth) - 1];
}
int chol_solve_forward(chol_t *c, const double *b, double *top_partial) {
  if (!c->device_ready || !c->assembled) return fail(c, "factor first");
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
/* top_sum: the sum of every rank's top_partial.  x_owned (original dof order): the entries of the rank's own
 * separators (rank 0: also of the shared top), zeros elsewhere, so that the sum over the ranks is the solution. */
int chol_solve_backward(chol_t *c, const double *top_sum, double *x_owned) {
  if (!c->device_ready || !c->assembled) return fail(c, "factor first");
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
      const int lv = P.level_of(h), own = lv < c->D.depth ? -1 : (h >> (lv - c->D.depth)) - (1 << c->D.depth);
      if (own == c->rank || (own < 0 && c->rank == 0)) continue;
      for (int i = 0; i < P.sz[h]; i++) x_owned[P.perm[P.start[h] + i]] = 0.0;
    }
  return 0;
}
/* what the rank's solve schedule covers, [0..5] on its subtree levels and [6..11] on the shared top levels:
 * forward tiles, forward gemv slabs, pull slabs, gather column groups, backward tiles, backward gemv groups */
int chol_solve_stats(chol_t *c, double *out12) {
  if (!c->analyzed) return fail(c, "analyze first");
  SolveSchedule V;
  if (build_solve(c->P, c->S, V, c->rank, c->world, c->err)) return -1;
  for (int i = 0; i < 12; i++) out12[i] = 0;
  for (const SolveLaunch &l : V.launches) {
    if (l.kind == SK_EXCHANGE) continue;
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
