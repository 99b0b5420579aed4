// Host-side (CPU, multi-threaded) helpers of the task construction: the NumPy
// passes that dominate time-to-first-iteration on meshes of millions of elements
// (`get_elements_volume`, facet extraction).  They are plain
// C++ behind the same C ABI, compiled into the library next to the kernels, take
// HOST pointers and never touch the device.  Each one reproduces its NumPy
// counterpart bit for bit (same operations in the same order, no fused
// multiply-add: this file is compiled with -ffp-contract=off), which the CPU tests
// check (tests/test_host_helpers.py).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace sktb;

namespace {

template <typename F>
void parallel_for(int64_t n, F &&body) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::min<int64_t>(std::max(1u, std::min(hw, 32u)), std::max<int64_t>(1, n / 65536));
  if (nt <= 1) {
    body((int64_t)0, n);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nt);
  const int64_t chunk = (n + nt - 1) / nt;
  for (int k = 0; k < nt; ++k) {
    const int64_t lo = k * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([&body, lo, hi]() { body(lo, hi); });
  }
  for (auto &t : th) t.join();
}

// |det(v1, v2, v3)| / 6 with NumPy's operation order (fea/composer.py:_abs_tet_volume)
inline double abs_tet(const double *a, const double *b, const double *c, const double *d) {
  const double v1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  const double v2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  const double v3[3] = {d[0] - a[0], d[1] - a[1], d[2] - a[2]};
  const double c0 = v1[1] * v2[2] - v1[2] * v2[1];
  const double c1 = v1[2] * v2[0] - v1[0] * v2[2];
  const double c2 = v1[0] * v2[1] - v1[1] * v2[0];
  const double det = (c0 * v3[0] + c1 * v3[1]) + c2 * v3[2];
  return std::fabs(det) / 6.0;
}

}  // namespace

// Sum of six |tetrahedron volumes| on the fixed local quadruples of the reference
// (fea/composer.py:191-248); t = connectivity [8][n_elem], p = coordinates [3][n_nodes].
extern "C" int sktb_host_hex_volumes(int64_t n_elem, int64_t n_nodes, const int32_t *t,
                                     const double *p, double *vol) {
  SKTB_REQUIRE(t && p && vol && n_elem >= 0 && n_nodes > 0, "bad argument");
  static const int quads[6][4] = {{0, 1, 3, 4}, {1, 2, 3, 6}, {1, 5, 6, 4},
                                  {3, 6, 7, 4}, {1, 3, 6, 4}, {1, 6, 5, 4}};
  std::atomic<bool> bad(false);
  parallel_for(n_elem, [&](int64_t lo, int64_t hi) {
    for (int64_t e = lo; e < hi; ++e) {
      double X[8][3];
      for (int k = 0; k < 8; ++k) {
        const int64_t n = t[(int64_t)k * n_elem + e];
        if (n < 0 || n >= n_nodes) {
          bad = true;
          return;
        }
        X[k][0] = p[n];
        X[k][1] = p[n_nodes + n];
        X[k][2] = p[2 * n_nodes + n];
      }
      double v = 0.0;
      for (int q = 0; q < 6; ++q)
        v += abs_tet(X[quads[q][0]], X[quads[q][1]], X[quads[q][2]], X[quads[q][3]]);
      vol[e] = v;
    }
  });
  SKTB_REQUIRE(!bad.load(), "connectivity entry out of range");
  return 0;
}

// Facet table of a hexahedral mesh whose elements are cells of a lattice numbered
// like init_tensor (node = iy + npy*ix + P*iz, P = npy*npx; ny_ = npy, any geometry,
// any valid local vertex order): the O(n) construction of
// sktopt/_fem/mesh.py::MeshHex._build_facets_lattice.  lf = the six local faces
// [6][4].  Pass 1 (facets == NULL): validates the structure, fills t2f_key
// [6][n_elem] (work array, int64) and returns the number of facets in *n_facets
// (-1: not such a mesh).  Pass 2: fills facets [4][n_facets], t2f [6][n_elem],
// f2t [2][n_facets], f2lf [2][n_facets] from the keys of pass 1.
extern "C" int sktb_host_lattice_facets(int64_t n_elem, int64_t n_nodes, const int32_t *t,
                                        const int32_t *lf, int64_t ny_, int64_t P,
                                        int64_t *key, int64_t *n_facets, int32_t *facets,
                                        int32_t *t2f, int32_t *f2t, int8_t *f2lf) {
  SKTB_REQUIRE(t && lf && key && n_facets && n_elem > 0 && n_nodes > 0, "bad argument");
  SKTB_REQUIRE(ny_ >= 3 && P >= 3 * ny_, "lattice too thin for the closed form");
  const int64_t span_of[3] = {ny_ + 1, P + 1, P + ny_};
  const int64_t pattern[8] = {0, 1, ny_, ny_ + 1, P, P + 1, P + ny_, P + ny_ + 1};
  if (!facets) {
    std::atomic<bool> ok(true);
    parallel_for(n_elem, [&](int64_t lo, int64_t hi) {
      for (int64_t e = lo; e < hi && ok.load(std::memory_order_relaxed); ++e) {
        int64_t v[8];
        for (int k = 0; k < 8; ++k) v[k] = t[(int64_t)k * n_elem + e];
        int64_t s[8];
        std::memcpy(s, v, sizeof(s));
        std::sort(s, s + 8);
        if (s[0] < 0 || s[7] >= n_nodes) {
          ok = false;
          return;
        }
        for (int k = 0; k < 8; ++k)
          if (s[k] - s[0] != pattern[k]) {
            ok = false;
            return;
          }
        for (int i = 0; i < 6; ++i) {
          int64_t b = v[lf[4 * i]], mx = b, sum = 0;
          for (int j = 0; j < 4; ++j) {
            const int64_t n = v[lf[4 * i + j]];
            b = std::min(b, n);
            mx = std::max(mx, n);
            sum += n;
          }
          const int64_t span = mx - b;
          const int f = span == span_of[0] ? 0 : (span == span_of[1] ? 1 : 2);
          if (span != span_of[f] || sum != 4 * b + 2 * span) {
            ok = false;
            return;
          }
          key[(int64_t)i * n_elem + e] = 3 * b + f;
        }
      }
    });
    if (!ok.load()) {
      *n_facets = -1;
      return 0;
    }
    std::vector<uint8_t> present((size_t)3 * n_nodes, 0);
    const int64_t n_keys = 6 * n_elem;
    for (int64_t i = 0; i < n_keys; ++i) present[key[i]] = 1;
    int64_t c = 0;
    for (size_t i = 0; i < present.size(); ++i) c += present[i];
    *n_facets = c;
    return 0;
  }
  SKTB_REQUIRE(t2f && f2t && f2lf && *n_facets > 0, "bad argument");
  const int64_t nfac = *n_facets, n_keys = 6 * n_elem;
  std::vector<uint8_t> present((size_t)3 * n_nodes, 0);
  for (int64_t i = 0; i < n_keys; ++i) present[key[i]] = 1;
  std::vector<int32_t> rank((size_t)3 * n_nodes);
  int64_t c = 0;
  const int64_t fa[3] = {1, 1, ny_}, fb[3] = {ny_, P, P};
  for (int64_t id = 0; id < 3 * n_nodes; ++id) {
    rank[id] = (int32_t)c;
    if (present[id]) {
      SKTB_REQUIRE(c < nfac, "facet count changed between the passes");
      const int64_t b = id / 3, f = id % 3;
      facets[c] = (int32_t)b;
      facets[nfac + c] = (int32_t)(b + fa[f]);
      facets[2 * nfac + c] = (int32_t)(b + fb[f]);
      facets[3 * nfac + c] = (int32_t)(b + fa[f] + fb[f]);
      ++c;
    }
  }
  SKTB_REQUIRE(c == nfac, "facet count changed between the passes");
  for (int64_t i = 0; i < 2 * nfac; ++i) {
    f2t[i] = -1;
    f2lf[i] = -1;
  }
  // ascending stacked index (lface * n_elem + elem): slot 0 = first occurrence
  for (int64_t i = 0; i < n_keys; ++i) {
    const int32_t fid = rank[key[i]];
    t2f[i] = fid;
    const int32_t e = (int32_t)(i % n_elem);
    const int8_t l = (int8_t)(i / n_elem);
    if (f2t[fid] < 0) {
      f2t[fid] = e;
      f2lf[fid] = l;
    } else {
      SKTB_REQUIRE(f2t[nfac + fid] < 0, "a facet with more than two elements");
      f2t[nfac + fid] = e;
      f2lf[nfac + fid] = l;
    }
  }
  return 0;
}
