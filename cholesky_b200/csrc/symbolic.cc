// Host symbolic analysis: the reference's cluster-granular fill analysis (compute_filled_clusters,
// mmat.rg:896-1028; merge_filled_clusters 635-695; fill_block's flags 561-627; cluster rectangles
// 364-451), restated so that it scales to the BASELINE grids: blocks get a flag bitmap only when
// something touches them, and each separator's result is kept as a supernodal panel (the list of
// its filled row clusters at the moment it is eliminated) instead of dense per-block instances.
// The `Filled` records the reference would hold at every interval label are counted and hashed
// on the fly (and kept on request) so the pattern can be compared bit for bit with the oracle.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "chol_internal.h"

namespace chb {

uint64_t mix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
uint64_t filled_hash(const FilledRec &r) {
  const int64_t *w = &r.filled;
  uint64_t h = 0;
  for (int k = 0; k < 9; k++) h = mix64(h ^ (uint64_t)w[k]);
  return h;
}

namespace {
inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
}  // namespace

int host_threads() {
  static int n = [] {
    int v = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("CHOL_HOST_THREADS")) v = atoi(e);
    return std::max(1, std::min(v, 16));
  }();
  return n;
}
void parallel_chunks(int64_t n, const std::function<void(int64_t, int64_t, int)> &fn, int *workers_out) {
  const int w = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), n / 4096));
  if (workers_out) *workers_out = w;
  if (w == 1) {
    fn(0, n, 0);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < w; t++) th.emplace_back([&, t] { fn(n * t / w, n * (t + 1) / w, t); });
  for (auto &x : th) x.join();
}

int analyze(const Problem &P, Symbolic &S, bool keep, std::string &err) {
  const int L = P.levels, N = P.N;
  S = Symbolic();
  // ---- composed cluster boundaries: interval k lists index interval k-1 (mmat.rg:405-410, 417-422)
  S.cb.assign(N + 2, {});
  for (int h = 1; h <= N; h++) {
    int ni = (int)P.iv[h].size();
    S.cb[h].resize(ni);
    for (int k = 0; k < ni; k++) {
      const auto &raw = P.iv[h][k];
      auto &out = S.cb[h][k];
      out.resize(raw.size());
      for (size_t j = 0; j < raw.size(); j++) {
        int v = raw[j];
        for (int i = k - 1; i >= 0; i--) {
          if (v < 0 || v >= (int)P.iv[h][i].size())
            return err = "cluster interval " + std::to_string(k) + " of separator id " + std::to_string(P.label_of(h) - 1) + " indexes out of range", -1;
          v = P.iv[h][i][v];
        }
        out[j] = v;
      }
      if (out.empty() || out.front() != 0 || out.back() != P.sz[h] || !std::is_sorted(out.begin(), out.end()))
        return err = "cluster interval " + std::to_string(k) + " of separator id " + std::to_string(P.label_of(h) - 1) + " does not partition the separator", -1;
    }
  }
  auto nc = [&](int h, int k) { return (int)S.cb[h][k].size() - 1; };

  // ---- allocated blocks (find_index_space_2d, mmat.rg:740-767): (r, c) with r == c or r ancestor of c
  std::vector<int64_t> boff(N + 2, 0);
  std::vector<int> lev(N + 2, 0);
  int64_t nb = 0;
  for (int h = 1; h <= N; h++) {
    lev[h] = P.level_of(h);
    boff[h] = nb;
    nb += lev[h] + 1;
  }
  S.nblocks = nb;
  S.nclusters0 = 0;
  for (int hc = 1; hc <= N; hc++)
    for (int hr = hc; hr >= 1; hr >>= 1) S.nclusters0 += (int64_t)nc(hr, 0) * nc(hc, 0);
  std::vector<std::vector<uint8_t>> flag((size_t)nb);  // 1 = structurally filled (empty vector: nothing filled)
  auto blk = [&](int hr, int hc) { return boff[hc] + (lev[hc] - lev[hr]); };

  // ---- interval-0 flags from A's nonzeros (fill_block, mmat.rg:561-627)
  {
    std::vector<int> iperm(P.n), rowheap(P.n);
    for (int p = 0; p < P.n; p++) iperm[P.perm[p]] = p;
    for (int h = 1; h <= N; h++)
      for (int i = 0; i < P.sz[h]; i++) rowheap[P.start[h] + i] = h;
    auto cluster0 = [&](int h, int pos) {
      const auto &b = S.cb[h][0];
      return (int)(std::upper_bound(b.begin(), b.end(), pos) - b.begin()) - 1;
    };
    // pass 1 (parallel): block and interval-0 cluster of every entry; pass 2 (serial): set the flags, allocating
    // a block's bitmap on first touch
    std::vector<int64_t> eb((size_t)P.nz, -1);
    std::vector<int32_t> ez((size_t)P.nz, 0);
    std::vector<int> bad(64, 0);
    parallel_chunks(P.nz, [&](int64_t e0, int64_t e1, int w) {
      for (int64_t e = e0; e < e1; e++) {
        if (P.ev[e] == 0.0) continue;
        if (P.ei[e] < 0 || P.ei[e] >= P.n || P.ej[e] < 0 || P.ej[e] >= P.n) {
          bad[w] = 1;
          return;
        }
        int pi = iperm[P.ei[e]], pj = iperm[P.ej[e]];
        if (pi < pj) std::swap(pi, pj);
        int hr = rowheap[pi], hc = rowheap[pj];
        int d = lev[hc] - lev[hr];
        if (d < 0 || (hc >> d) != hr) continue;  // couples two unrelated separators: no block, dropped (mmat.rg:1191)
        eb[e] = blk(hr, hc);
        ez[e] = cluster0(hr, pi - P.start[hr]) * nc(hc, 0) + cluster0(hc, pj - P.start[hc]);
      }
    });
    for (int b : bad)
      if (b) return err = "matrix entry out of range", -1;
    std::vector<int> rows0(N + 2), cols0(N + 2);
    for (int64_t e = 0; e < P.nz; e++) {
      if (eb[e] < 0) continue;
      auto &f = flag[eb[e]];
      if (f.empty()) {
        // the block id encodes (hr, hc): hc owns the ids boff[hc] .. boff[hc] + lev[hc]
        int hc = (int)(std::upper_bound(boff.begin() + 1, boff.begin() + N + 1, eb[e]) - boff.begin()) - 1;
        int hr = hc >> (int)(eb[e] - boff[hc]);
        f.assign((size_t)nc(hr, 0) * nc(hc, 0), 0);
      }
      f[(size_t)ez[e]] = 1;
    }
  }

  S.seg_ptr.assign(N + 2, 0);
  std::vector<std::vector<Seg>> segs_of(N + 2);
  S.nfilled.assign(L, 0);
  S.checksum.assign(L, 0);
  if (keep) S.records.assign(L, {});
  S.f_potrf.assign(L, 0), S.f_trsm.assign(L, 0), S.f_syrk.assign(L, 0), S.f_gemm.assign(L, 0);

  int k = 0;  // cluster partition index ("interval", mmat.rg:1350-1354)
  std::vector<std::vector<int>> F(L + 1);
  for (int t = 0; t < L; t++) {
    const int lvl = L - 1 - t;
    for (int h = 1; h < (1 << (lvl + 1)); h++)
      if ((int)P.iv[h].size() <= k)
        return err = "separator id " + std::to_string(P.label_of(h) - 1) + " lacks cluster interval " + std::to_string(k) + " needed at tree level " + std::to_string(lvl), -1;
    // ---- fill propagation from every separator of this level (mmat.rg:926-998)
    for (int hs = 1 << lvl; hs < (1 << (lvl + 1)); hs++) {
      if (nc(hs, k) != 1)
        return err = "separator id " + std::to_string(P.label_of(hs) - 1) + " has " + std::to_string(nc(hs, k)) + " clusters when eliminated; the reference's fused tasks need exactly 1", -1;
      const int n = P.sz[hs];
      auto &sg = segs_of[hs];
      sg.push_back(Seg{hs, 0, n, 0, 0});
      const auto &fd = flag[blk(hs, hs)];
      if (!fd.empty() && fd[0]) {
        S.f_potrf[lvl] += (double)n * n * n / 3.0 + (double)n * n / 2.0 + (double)n / 6.0;
        S.calls[0]++;
      }
      for (int d = 1; d <= lvl; d++) {
        int a = hs >> d;
        F[d].clear();
        const auto &f = flag[blk(a, hs)];
        if (f.empty()) continue;
        for (int rc = 0; rc < (int)f.size(); rc++)
          if (f[rc]) {
            F[d].push_back(rc);
            int lo = S.cb[a][k][rc], hi = S.cb[a][k][rc + 1];
            sg.push_back(Seg{a, lo, hi, 0, rc});
            S.f_trsm[lvl] += (double)(hi - lo) * n * n;
            S.calls[1]++;
          }
      }
      for (int dp = 1; dp <= lvl; dp++) {
        if (F[dp].empty()) continue;
        int p = hs >> dp, ncp = nc(p, k);
        for (int dg = dp; dg <= lvl; dg++) {
          if (F[dg].empty()) continue;
          int g = hs >> dg;
          auto &fc = flag[blk(g, p)];
          if (fc.empty()) fc.assign((size_t)nc(g, k) * ncp, 0);
          for (int ia : F[dg]) {
            int m = S.cb[g][k][ia + 1] - S.cb[g][k][ia];
            for (int jb : F[dp]) {
              if (dg == dp && jb > ia) break;
              fc[(size_t)ia * ncp + jb] = 1;
              int nn = S.cb[p][k][jb + 1] - S.cb[p][k][jb];
              if (dg == dp && jb == ia) {
                S.f_syrk[lvl] += (double)n * m * (m + 1);
                S.calls[2]++;
              } else {
                S.f_gemm[lvl] += 2.0 * m * nn * n;
                S.calls[3]++;
              }
            }
          }
        }
      }
    }
    // ---- snapshot F[t] (mmat.rg:1000-1016): every flagged cluster of every block that still carries flags.
    // Count and checksum are order independent, so column separators are spread over the host threads.
    {
      std::vector<int64_t> cnt(64, 0);
      std::vector<uint64_t> sum(64, 0);
      std::vector<std::vector<FilledRec>> recs(64);
      parallel_chunks(N, [&](int64_t h0, int64_t h1, int w) {
        for (int hc = (int)h0 + 1; hc <= (int)h1; hc++)
          for (int hr = hc; hr >= 1; hr >>= 1) {
            const auto &f = flag[blk(hr, hc)];
            if (f.empty()) continue;
            int ncc = nc(hc, k);
            for (size_t z = 0; z < f.size(); z++)
              if (f[z]) {
                int rc = (int)(z / ncc), cc = (int)(z % ncc);
                FilledRec r;
                r.filled = 0, r.sep_x = P.label_of(hr), r.sep_y = P.label_of(hc), r.interval = t, r.cluster = (int64_t)z;
                r.lo_x = P.start[hr] + S.cb[hr][k][rc], r.hi_x = P.start[hr] + S.cb[hr][k][rc + 1] - 1;
                r.lo_y = P.start[hc] + S.cb[hc][k][cc], r.hi_y = P.start[hc] + S.cb[hc][k][cc + 1] - 1;
                cnt[w]++;
                sum[w] += filled_hash(r);
                if (keep) recs[w].push_back(r);
              }
          }
      });
      for (int w = 0; w < 64; w++) {
        S.nfilled[t] += cnt[w];
        S.checksum[t] += sum[w];
        if (keep) S.records[t].insert(S.records[t].end(), recs[w].begin(), recs[w].end());
      }
    }
    if (keep)
      std::sort(S.records[t].begin(), S.records[t].end(), [](const FilledRec &a, const FilledRec &b) {
        if (a.sep_x != b.sep_x) return a.sep_x < b.sep_x;
        if (a.sep_y != b.sep_y) return a.sep_y < b.sep_y;
        return a.cluster < b.cluster;
      });
    // ---- advance the interval and coarsen the flags (mmat.rg:1018-1026, 635-695)
    if (lvl <= L - 2) {
      k++;
      if (k < L) {
        std::vector<int> mr, mc;
        for (int hc = 1; hc <= N; hc++)
          for (int hr = hc; hr >= 1; hr >>= 1) {
            auto &f = flag[blk(hr, hc)];
            if (f.empty()) continue;
            if ((int)P.iv[hr].size() <= k || (int)P.iv[hc].size() <= k) {
              std::vector<uint8_t>().swap(f);  // no partition at this interval: all flags reset
              continue;
            }
            int pr = nc(hr, k - 1), pc = nc(hc, k - 1), nr = nc(hr, k), ncn = nc(hc, k);
            mr.assign(pr, 0), mc.assign(pc, 0);
            for (int row = 0; row < nr; row++)
              for (int i = P.iv[hr][k][row]; i < P.iv[hr][k][row + 1]; i++) mr[i] = row;
            for (int col = 0; col < ncn; col++)
              for (int j = P.iv[hc][k][col]; j < P.iv[hc][k][col + 1]; j++) mc[j] = col;
            std::vector<uint8_t> nf((size_t)nr * ncn, 0);
            for (int i = 0; i < pr; i++)
              for (int j = 0; j < pc; j++)
                if (f[(size_t)i * pc + j]) nf[(size_t)mr[i] * ncn + mc[j]] = 1;
            f.swap(nf);
          }
      }
    }
  }

  // ---- panel layout: diagonal block first, then the filled row clusters in ascending permuted row;
  // every segment starts on an even row so that 16-byte async copies of A/B tiles stay aligned.
  S.rows.assign(N + 2, 0), S.ld.assign(N + 2, 0), S.poff.assign(N + 2, 0);
  int64_t off = 0, nseg = 0;
  for (int h = 1; h <= N; h++) nseg += (int64_t)segs_of[h].size();
  S.segs.reserve((size_t)nseg);
  for (int h = 1; h <= N; h++) {
    S.seg_ptr[h] = (int64_t)S.segs.size();
    int cur = 0;
    for (auto &s : segs_of[h]) {
      s.off = cur;
      cur = round_up(cur + (s.hi - s.lo), 2);
      S.segs.push_back(s);
    }
    S.rows[h] = cur;
    S.ld[h] = std::max(cur, 2);
    S.poff[h] = off;
    off += (int64_t)S.ld[h] * P.sz[h];
    off = (off + 15) / 16 * 16;
    std::vector<Seg>().swap(segs_of[h]);
  }
  S.seg_ptr[N + 1] = (int64_t)S.segs.size();
  S.poff[N + 1] = off;  // so that [poff[h], poff[h + 1]) is the storage of panel h for every h
  S.total_doubles = off + 4096;  // slack: tile loads may run a few rows past the last panel
  return 0;
}

}  // namespace chb
