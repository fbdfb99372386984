// Synthetic inputs of BASELINE.json: grid Laplacians and a geometric nested-dissection ordering
// with the cluster-interval hierarchy the reference's clust files carry.  The
// reference ships no generator (its ord/clust files come from an external partitioner); this
// one produces the same three inputs in memory, and write_problem() dumps them in the
// reference's text formats so the same files can be fed to the CPU oracle.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "chol_internal.h"

namespace chb {

namespace {

struct Box {
  int lo[3], hi[3];  // hi exclusive
  int ext(int a) const { return hi[a] - lo[a]; }
  bool empty() const { return ext(0) <= 0 || ext(1) <= 0 || ext(2) <= 0; }
};

struct Node {
  Box box;
  int axis = -1, mid = 0;  // internal nodes: separator = plane coord[axis] == mid
};

struct Gen {
  int dim[3];
  int levels, N;
  std::vector<Node> node;  // by heap index
  int idx(const int *q) const { return q[0] + dim[0] * (q[1] + dim[1] * q[2]); }

  void build(int h, const Box &b, int depth) {
    node[h].box = b;
    if (depth == levels - 1) return;
    int a = 0;
    for (int i = 1; i < 3; i++)
      if (b.ext(i) > b.ext(a)) a = i;
    int mid = (b.lo[a] + b.hi[a]) / 2;
    node[h].axis = a, node[h].mid = mid;
    Box l = b, r = b;
    l.hi[a] = mid;
    r.lo[a] = std::min(mid + 1, b.hi[a]);
    if (b.empty()) l = r = b;
    build(2 * h, l, depth + 1);
    build(2 * h + 1, r, depth + 1);
  }

  // heap index of the leaf-depth box that holds q, descending from node c; a point that lands on
  // a deeper separator plane is snapped to that separator's left side.
  int leaf_code(int c, int *q) const {
    while (node[c].axis >= 0) {
      const Node &nd = node[c];
      int a = nd.axis;
      if (q[a] < nd.mid) c = 2 * c;
      else if (q[a] > nd.mid) c = 2 * c + 1;
      else if (nd.mid - 1 >= nd.box.lo[a]) q[a] = nd.mid - 1, c = 2 * c;
      else if (nd.mid + 1 < nd.box.hi[a]) q[a] = nd.mid + 1, c = 2 * c + 1;
      else c = 2 * c;
    }
    return c;
  }
};

}  // namespace

int generate_problem(Problem &P, int nx, int ny, int nz, int stencil, int levels, std::string &err) {
  if (nx < 1 || ny < 1 || nz < 1) return err = "bad grid", -1;
  if (stencil != 5 && stencil != 7 && stencil != 27) return err = "stencil must be 5, 7 or 27", -1;
  if (stencil == 5 && nz != 1) return err = "5-point stencil is 2-D (nz = 1)", -1;
  int64_t n64 = (int64_t)nx * ny * nz;
  if (n64 > 0x7fffffff) return err = "grid too large", -1;
  int n = (int)n64;
  if (levels <= 0) levels = std::max(1, (int)std::ceil(std::log2((double)n / 64.0)) + 1); /* utils.py:6-7 */
  P = Problem();
  P.n = P.ncols = n;
  P.levels = levels;
  P.N = (1 << levels) - 1;

  // ---- matrix: lower triangle, sorted by column then row (as the reference's fixtures)
  std::vector<int> offs;  // forward neighbour offsets (dx, dy, dz) with positive index offset
  struct D3 {
    int dx, dy, dz;
  };
  std::vector<D3> fw;
  if (stencil == 27) {
    for (int dz = -1; dz <= 1; dz++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++)
          if (dx + nx * (dy + ny * dz) > 0 && (dz > 0 || (dz == 0 && (dy > 0 || (dy == 0 && dx > 0))))) fw.push_back({dx, dy, dz});
  } else {
    fw.push_back({1, 0, 0});
    fw.push_back({0, 1, 0});
    if (stencil == 7) fw.push_back({0, 0, 1});
  }
  double diag = stencil == 5 ? 4.0 : stencil == 7 ? 6.0 : 26.0;
  P.ei.reserve((size_t)n * (fw.size() + 1));
  P.ej.reserve((size_t)n * (fw.size() + 1));
  P.ev.reserve((size_t)n * (fw.size() + 1));
  std::vector<int> nb;
  for (int z = 0; z < nz; z++)
    for (int y = 0; y < ny; y++)
      for (int x = 0; x < nx; x++) {
        int j = x + nx * (y + ny * z);
        P.ei.push_back(j), P.ej.push_back(j), P.ev.push_back(diag);
        nb.clear();
        for (auto &d : fw) {
          int X = x + d.dx, Y = y + d.dy, Z = z + d.dz;
          if (X < 0 || X >= nx || Y < 0 || Y >= ny || Z < 0 || Z >= nz) continue;
          nb.push_back(X + nx * (Y + ny * Z));
        }
        std::sort(nb.begin(), nb.end());
        for (int i : nb) P.ei.push_back(i), P.ej.push_back(j), P.ev.push_back(-1.0);
      }
  P.nz = (int64_t)P.ev.size();

  // ---- geometric nested dissection
  Gen G;
  G.dim[0] = nx, G.dim[1] = ny, G.dim[2] = nz;
  G.levels = levels, G.N = P.N;
  G.node.assign(P.N + 2, Node());
  Box root;
  root.lo[0] = root.lo[1] = root.lo[2] = 0;
  root.hi[0] = nx, root.hi[1] = ny, root.hi[2] = nz;
  G.build(1, root, 0);

  P.sz.assign(P.N + 2, 0);
  P.iv.assign(P.N + 2, {});
  std::vector<std::vector<int>> dofs(P.N + 2);
  struct Key {
    int code, idx;
  };
  std::vector<Key> keys;
  for (int h = 1; h <= P.N; h++) {
    const Node &nd = G.node[h];
    int lvl = P.level_of(h);
    Box b = nd.box;
    if (nd.axis >= 0) {  // the separator plane
      b.lo[nd.axis] = nd.mid;
      b.hi[nd.axis] = std::min(nd.mid + 1, nd.box.hi[nd.axis]);
    }
    keys.clear();
    bool clustered = nd.axis >= 0 && lvl <= levels - 3;
    if (!b.empty())
      for (int z = b.lo[2]; z < b.hi[2]; z++)
        for (int y = b.lo[1]; y < b.hi[1]; y++)
          for (int x = b.lo[0]; x < b.hi[0]; x++) {
            int q[3] = {x, y, z};
            int id = G.idx(q);
            int code = 0;
            if (clustered) {
              int a = nd.axis;
              int c;
              if (nd.mid - 1 >= nd.box.lo[a]) q[a] = nd.mid - 1, c = 2 * h;
              else q[a] = nd.mid + 1, c = 2 * h + 1;
              code = G.leaf_code(c, q);
            }
            keys.push_back({code, id});
          }
    if (keys.empty()) return err = "empty separator (too many levels for this grid): id " + std::to_string(P.label_of(h) - 1), -1;
    std::sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) { return a.code != b.code ? a.code < b.code : a.idx < b.idx; });
    int m = (int)keys.size();
    P.sz[h] = m;
    dofs[h].resize(m);
    for (int i = 0; i < m; i++) dofs[h][i] = keys[i].idx;
    int nint = std::max(1, levels - 1 - lvl);
    P.iv[h].resize(nint);
    if (!clustered) {
      P.iv[h][0] = {0, m};
      continue;
    }
    // interval k cuts where the depth-(levels-1-k) code changes; written as indices into k-1
    std::vector<int> prev;  // boundary positions of interval k-1
    for (int k = 0; k < nint; k++) {
      std::vector<int> cur;
      cur.push_back(0);
      if (k < nint - 1)
        for (int i = 1; i < m; i++)
          if ((keys[i].code >> k) != (keys[i - 1].code >> k)) cur.push_back(i);
      cur.push_back(m);  // the last interval is forced to (0, size): one cluster at elimination
      if (k == 0) P.iv[h][0] = cur;
      else {
        std::vector<int> ix;
        size_t p = 0;
        for (int v : cur) {
          while (p < prev.size() && prev[p] < v) p++;
          if (p >= prev.size() || prev[p] != v) return err = "internal: cluster intervals not nested", -1;
          ix.push_back((int)p);
        }
        P.iv[h][k] = ix;
      }
      prev = cur;
    }
  }
  P.perm.resize(n);
  int pos = 0;
  int mx = -1;
  for (int label = 1; label <= P.N; label++) {
    int h = P.heap_of(label);
    for (int d : dofs[h]) P.perm[pos++] = d;
    for (auto &l : P.iv[h]) mx = std::max(mx, (int)l.size() + 1);
  }
  P.max_int_size = mx;
  if (pos != n) return err = "internal: nested dissection does not cover the grid", -1;
  return finish_problem(P, err);
}

}  // namespace chb
