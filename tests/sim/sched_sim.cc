// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product.
//
// Host interpreter of the engine's compiled launch lists, for validating the multi-GPU schedule without GPUs:
// every rank of a partition gets its own factor buffer and flag words; the launches of all ranks are executed
// one at a time in a RANDOM order that respects only what the hardware would respect -- stream order, the
// cross-stream events of the list, and the flag words in "peer memory" (K_SYNC waits, K_PUSH signals).  A
// missing dependency therefore shows up as a wrong factor for some seed, and a cyclic wait as a deadlock
// report, before any GPU time is spent.  Each launch kind is restated from its kernel in csrc/kernels.cuh
// (same tiles, same masks, same row-pair granularity of the pushes); arithmetic is plain loops.
//
// Built by tests/sim/Makefile against the product's own host code (schedule compiler, symbolic analysis,
// generators), so what is interpreted is exactly what the GPU would be handed.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../cholesky_b200/csrc/chol_internal.h"

using namespace chb;

namespace {

constexpr int kPeers = 8;
struct Rank {
  Schedule D;
  std::vector<double> fac;
  unsigned long long flags[kFlagSlots][kPeers];
  std::vector<std::vector<int>> q;   // launch indices per stream
  size_t head[4] = {0, 0, 0, 0};
  std::vector<char> ev;              // event recorded?
  std::vector<char> signalled;       // K_SYNC: signals already sent
};

void potrf(double *A, int ld, int nb, bool &bad) {
  for (int k = 0; k < nb; k++) {
    double d = A[k + (size_t)k * ld];
    if (!(d > 0)) bad = true, d = 1.0;
    d = std::sqrt(d);
    A[k + (size_t)k * ld] = d;
    for (int i = k + 1; i < nb; i++) A[i + (size_t)k * ld] /= d;
    for (int j = k + 1; j < nb; j++)
      for (int i = j; i < nb; i++) A[i + (size_t)j * ld] -= A[i + (size_t)k * ld] * A[j + (size_t)k * ld];
  }
}
void trsm(const double *Lm, double *B, int ld, int nb, int r0, int r1) {  // rows [r0, r1) of B <- B L^-T
  for (int r = r0; r < r1; r++)
    for (int c = 0; c < nb; c++) {
      double s = B[r + (size_t)c * ld];
      for (int k = 0; k < c; k++) s -= B[r + (size_t)k * ld] * Lm[c + (size_t)k * ld];
      B[r + (size_t)c * ld] = s / Lm[c + (size_t)c * ld];
    }
}

}  // namespace

extern "C" int sim_factor(int nx, int ny, int nz, int stencil, int levels, int world, uint64_t seed, double *dense_out, double *copy_diff,
                          double *stats4, char *errbuf, int errlen) {
  std::string err;
  auto fail = [&](const std::string &m) {
    snprintf(errbuf, errlen, "%s", m.c_str());
    return -1;
  };
  Problem P;
  Symbolic S;
  if (generate_problem(P, nx, ny, nz, stencil, levels, err)) return fail(err);
  if (analyze(P, S, false, err)) return fail(err);
  std::vector<Rank> R(world);
  int depth = 0;
  while ((1 << depth) < world) depth++;
  double pushes = 0, syncs = 0, reduces = 0, gemm_flops = 0;
  for (int r = 0; r < world; r++) {
    if (build_schedule(P, S, R[r].D, r, world, false, err)) return fail(err);
    Rank &k = R[r];
    k.fac.assign((size_t)S.total_doubles, 0.0);
    for (int64_t e = 0; e < P.nz; e++)
      if (k.D.a_off[e] >= 0) k.fac[k.D.a_off[e]] = P.ev[e];
    memset(k.flags, 0, sizeof k.flags);
    k.q.assign(4, {});
    for (size_t i = 0; i < k.D.launches.size(); i++) k.q[k.D.launches[i].stream].push_back((int)i);
    k.ev.assign(k.D.num_events, 0);
    k.signalled.assign(k.D.launches.size(), 0);
  }
  std::mt19937_64 rng(seed);
  const unsigned long long run = 1;
  bool bad_pivot = false;
  auto flagval = [&](const Launch &l) { return (run << 32) | (unsigned long long)l.seq; };
  auto signal = [&](int me, const Launch &l) {
    for (int p = 0; p < world; p++)
      if ((l.sig_mask >> p) & 1u) R[p].flags[l.slot][me] = std::max(R[p].flags[l.slot][me], flagval(l));
  };
  auto exec = [&](int me, const Launch &l) {
    Rank &k = R[me];
    double *fac = k.fac.data();
    switch (l.kind) {
      case K_PANEL:  // slabs in grid order: a slab only depends on diagonal slabs that precede it
        for (int64_t si = l.begin; si < l.begin + l.count; si++) {
          const PanelSlab &sl = k.D.pslabs[si];
          const PanelDesc &d = k.D.pdesc[sl.desc];
          double *G = fac + d.off + sl.row0 + (size_t)d.c0 * d.ld;            // entry (r, c) of the slab
          const double *Gd = fac + d.off + d.c0 + (size_t)d.c0 * d.ld;        // entry (n, kk) of the diagonal block
          const int ncol = sl.t >= 0 ? std::min(d.w, 64 * (sl.t + 1)) : d.w;
          for (int j = 0; j * 64 < ncol; j++) {
            const int d0 = 64 * j, dw = std::min(64, d.w - d0);
            for (int c = 0; c < dw; c++)  // left-looking update from the tile columns before j
              for (int r = 0; r < sl.rows; r++) {
                if (sl.t == j && r < c) continue;  // strictly upper part of the diagonal tile: never used
                double acc = 0;
                for (int kk = 0; kk < d0; kk++) acc += G[r + (size_t)kk * d.ld] * Gd[(d0 + c) + (size_t)kk * d.ld];
                G[r + (size_t)(d0 + c) * d.ld] -= acc;
              }
            if (sl.t == j) potrf(G + (size_t)d0 * d.ld, d.ld, dw, bad_pivot);
            else trsm(Gd + d0 + (size_t)d0 * d.ld, G + (size_t)d0 * d.ld, d.ld, dw, 0, sl.rows);
          }
        }
        break;
      case K_TRSM:
        for (int64_t i = l.begin; i < l.begin + l.count; i++) {
          const TileRef &t = k.D.trsm_tiles[i];
          const TrsmDesc &d = k.D.trsm[t.prob];
          const int slab = (int)t.tr | ((int)t.tc << 16);
          trsm(fac + d.l_off, fac + d.b_off, d.ld, d.nb, slab * 128, std::min(d.rows, (slab + 1) * 128));
        }
        break;
      case K_GEMM: {
        const int bm = l.cfg == 3 ? 32 : 64;
        gemm_flops += l.flops;
        for (int64_t i = l.begin; i < l.begin + l.count; i++) {
          const TileRef &t = k.D.tiles[i];
          const GemmProblem &g = k.D.probs[t.prob];
          const int r0 = t.tr * bm, c0 = t.tc * bm, r1 = std::min(g.M, r0 + bm), c1 = std::min(g.N, c0 + bm);
          for (int cc = c0; cc < c1; cc++)
            for (int r = r0; r < r1; r++) {
              if ((g.tri & 1) && r < cc) continue;
              if ((g.tri & 2) && r < 1) continue;
              double acc = 0;
              for (int c = 0; c < g.contrib_count; c++) {
                const GemmContrib &cb = k.D.contribs[g.contrib_begin + c];
                const double *A = fac + cb.a_off + r, *B = fac + cb.b_off + cc;
                for (int kk = 0; kk < cb.K; kk++) acc += A[(size_t)kk * cb.lda] * B[(size_t)kk * cb.ldb];
              }
              fac[g.c_off + r + (size_t)cc * g.ldc] -= acc;
            }
        }
        break;
      }
      case K_PUSH:
        for (int64_t i = l.begin; i < l.begin + l.count; i++) {
          const RectDesc &d = k.D.rects[i];
          pushes += 1;
          for (int c = 0; c < d.cols; c++)
            for (int r2 = 0; r2 < (d.rows + 1) / 2; r2++) {
              if (d.tri0 < (1 << 29) && c > d.tri0 + 2 * r2 + 1) continue;
              const int64_t o = d.off + 2 * r2 + (int64_t)c * d.ld;
              for (int p = 0; p < world; p++)
                if ((l.mask >> p) & 1u) R[p].fac[o] = fac[o], R[p].fac[o + 1] = fac[o + 1];
            }
        }
        signal(me, l);
        break;
      case K_REDUCE:
        reduces += 1;
        for (int64_t i = l.begin; i < l.begin + l.count; i++) {
          const RectDesc &d = k.D.rects[i];
          for (int c = 0; c < d.cols; c++)
            for (int r2 = 0; r2 < (d.rows + 1) / 2; r2++) {
              if (d.tri0 < (1 << 29) && c > d.tri0 + 2 * r2 + 1) continue;
              const int64_t o = d.off + 2 * r2 + (int64_t)c * d.ld;
              double s0 = 0, s1 = 0;
              for (int p = 0; p < world; p++)
                if ((l.mask & d.mask) >> p & 1u) s0 += R[p].fac[o], s1 += R[p].fac[o + 1];
              fac[o] = s0, fac[o + 1] = s1;
            }
        }
        break;
      default:
        break;
    }
  };
  // ---- the interleaving loop
  size_t remaining = 0;
  for (auto &k : R) remaining += k.D.launches.size();
  // Adversarial interleavings: every (rank, stream) queue gets a heavy-tailed weight per seed, so some queues
  // race far ahead of the others whenever nothing holds them back (uniform picks keep the ranks in near
  // lock step and hide missing waits).
  double weight[kPeers][4];
  {
    std::normal_distribution<double> nd(0.0, 1.0);
    for (int r = 0; r < kPeers; r++)
      for (int s = 0; s < 4; s++) weight[r][s] = std::exp(4.0 * nd(rng));
  }
  std::vector<std::pair<int, int>> ready;
  while (remaining) {
    ready.clear();
    for (int r = 0; r < world; r++)
      for (int s = 0; s < 4; s++) {
        Rank &k = R[r];
        if (k.head[s] >= k.q[s].size()) continue;
        const int li = k.q[s][k.head[s]];
        const Launch &l = k.D.launches[li];
        if (l.wait_ev >= 0 && !k.ev[l.wait_ev]) continue;
        if (l.kind == K_SYNC) {
          if (!k.signalled[li]) {  // a sync raises its flags as soon as it starts ...
            signal(r, l);
            k.signalled[li] = 1;
            syncs += 1;
          }
          bool ok = true;  // ... and finishes when its own words have been raised
          for (int p = 0; p < world; p++)
            if (((l.wait_mask >> p) & 1u) && k.flags[l.slot][p] < flagval(l)) ok = false;
          if (!ok) continue;
        }
        ready.push_back({r, s});
      }
    if (ready.empty()) {
      std::string m = "deadlock: heads";
      for (int r = 0; r < world; r++)
        for (int s = 0; s < 4; s++)
          if (R[r].head[s] < R[r].q[s].size()) {
            const Launch &l = R[r].D.launches[R[r].q[s][R[r].head[s]]];
            m += " [rank " + std::to_string(r) + " stream " + std::to_string(s) + " kind " + std::to_string(l.kind) + " level " + std::to_string(l.level) +
                 " slot " + std::to_string(l.slot) + " seq " + std::to_string((long long)l.seq) + " wait_ev " + std::to_string(l.wait_ev) + "]";
          }
      return fail(m);
    }
    double tot = 0;
    for (auto &p : ready) tot += weight[p.first][p.second];
    double x = std::uniform_real_distribution<double>(0.0, tot)(rng);
    auto pick = ready.back();
    for (auto &p : ready) {
      x -= weight[p.first][p.second];
      if (x <= 0) {
        pick = p;
        break;
      }
    }
    Rank &k = R[pick.first];
    const int li = k.q[pick.second][k.head[pick.second]++];
    const Launch &l = k.D.launches[li];
    if (l.kind != K_SYNC) exec(pick.first, l);
    if (l.rec_ev >= 0) k.ev[l.rec_ev] = 1;
    remaining--;
  }
  if (bad_pivot) return fail("non-positive pivot");
  // ---- results: dense L from the panels each rank reports; agreement of the copies of the top panels
  const size_t n = (size_t)P.n;
  memset(dense_out, 0, n * n * sizeof(double));
  double diff = 0;
  for (int h = 1; h <= P.N; h++) {
    const int lv = P.level_of(h);
    const int own = (world == 1) ? 0 : (lv < depth ? -1 : (h >> (lv - depth)) - (1 << depth));
    const Rank &src = R[own < 0 ? 0 : own];
    for (int64_t s = S.seg_ptr[h]; s < S.seg_ptr[h + 1]; s++) {
      const Seg &sg = S.segs[s];
      for (int r = 0; r < sg.hi - sg.lo; r++)
        for (int c = 0; c < P.sz[h]; c++) {
          if (sg.anc == h && c > sg.lo + r) continue;  // pivot block: lower triangle
          const size_t o = (size_t)S.poff[h] + sg.off + r + (size_t)c * S.ld[h];
          dense_out[(size_t)(P.start[sg.anc] + sg.lo + r) * n + P.start[h] + c] = src.fac[o];
          if (own < 0)
            for (int q = 1; q < world; q++) diff = std::max(diff, std::fabs(R[q].fac[o] - src.fac[o]));
        }
    }
  }
  *copy_diff = diff;
  stats4[0] = pushes, stats4[1] = syncs, stats4[2] = reduces, stats4[3] = gemm_flops;
  return 0;
}
