// ar_oracle.cpp -- CPU ORACLE (test infrastructure, NOT the product).
// See ar_oracle.hpp for scope, sources and parity status ("parity unpinned"
// against real Ceres; pinned by mpmath / scipy / finite differences).
//
// This file restates what happens inside ArSlamSolver::optimize
// (/root/reference/ar_slam/src/ar_slam_util.cpp:1001-1018) ==> ceres::Solve
// with Ceres 2.0.0 defaults + DENSE_SCHUR + max_num_iterations 50, and the
// per-capture problem built by localizeOne (ar_slam_util.cpp:903-979).
#include "ar_oracle.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <limits>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ar_oracle_capi.h"

namespace oracle {
namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// One Ceres parameter block as it appears in the reduced program.
struct ParamBlock {
  int type;    // 0 camera, 1 capture, 2 tag
  int index;   // capture / tag index
  int size;    // 3 or 6
  int offset;  // offset into the state vector x (after Schur ordering)
  int is_e;    // member of the eliminated independent set
  int f_off;   // offset inside the reduced system, if !is_e
};

struct Problem {
  int n_cap, n_tag, n_blk;
  const int32_t* cap_idx;
  const int32_t* tag_idx;
  const double* obs;  // 8 per block
  double tag_size;
  int model;
  int cam_const;
  const uint8_t* cap_const;  // may be null
  const uint8_t* tag_const;  // may be null
  // state
  double* cam;  // 3
  double* cap;  // 6*n_cap
  double* tag;  // 6*n_tag
  // program
  std::vector<ParamBlock> pbs;
  int pb_cam = -1;
  std::vector<int> pb_of_cap, pb_of_tag;
  int n_x = 0, n_e = 0, n_f = 0;
  std::vector<int> e_list;                 // pb ids of e-blocks
  std::vector<int> e_rb_start, e_rb_list;  // residual blocks per e-block (CSR)
  std::vector<int> rb_noe;                 // residual blocks without an e-block
};

bool is_const(const Problem& p, int type, int idx) {
  if (type == 0) return p.cam_const != 0;
  if (type == 1) return p.cap_const && p.cap_const[idx];
  return p.tag_const && p.tag_const[idx];
}

// Program order == order of first appearance in AddResidualBlock(camera,
// capture, tag) calls (ar_slam_util.cpp:723-727); constant blocks dropped.
void build_program(Problem& p, int elimination) {
  p.pb_cam = -1;
  p.pb_of_cap.assign(p.n_cap, -1);
  p.pb_of_tag.assign(p.n_tag, -1);
  p.pbs.clear();
  for (int b = 0; b < p.n_blk; ++b) {
    if (p.pb_cam < 0 && !is_const(p, 0, 0)) {
      p.pb_cam = (int)p.pbs.size();
      p.pbs.push_back({0, 0, 3, 0, 0, 0});
    }
    const int c = p.cap_idx[b], a = p.tag_idx[b];
    if (p.pb_of_cap[c] < 0 && !is_const(p, 1, c)) {
      p.pb_of_cap[c] = (int)p.pbs.size();
      p.pbs.push_back({1, c, 6, 0, 0, 0});
    }
    if (p.pb_of_tag[a] < 0 && !is_const(p, 2, a)) {
      p.pb_of_tag[a] = (int)p.pbs.size();
      p.pbs.push_back({2, a, 6, 0, 0, 0});
    }
  }
  const int n = (int)p.pbs.size();
  // co-occurrence graph
  std::vector<std::vector<int>> adj(n);
  for (int b = 0; b < p.n_blk; ++b) {
    int ids[3] = {p.pb_cam, p.pb_of_cap[p.cap_idx[b]], p.pb_of_tag[p.tag_idx[b]]};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        if (i != j && ids[i] >= 0 && ids[j] >= 0) adj[ids[i]].push_back(ids[j]);
  }
  for (auto& v : adj) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  }
  std::vector<int> color(n, 0);  // 0 white, 1 grey, 2 black
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  if (elimination == 0) {
    // reorder_program.cc / graph_algorithms.h StableIndependentSetOrdering:
    // stable sort by increasing degree, greedy maximal independent set.
    std::stable_sort(order.begin(), order.end(),
                     [&](int l, int r) { return adj[l].size() < adj[r].size(); });
    for (int v : order) {
      if (color[v] != 0) continue;
      color[v] = 2;
      for (int u : adj[v]) if (color[u] == 0) color[u] = 1;
    }
  } else if (elimination == 1 || elimination == 2) {
    const int want = (elimination == 1) ? 2 : 1;  // 1: tags, 2: captures
    for (int v = 0; v < n; ++v) if (p.pbs[v].type == want) color[v] = 2;
  }  // 3: nothing eliminated (plain dense normal equations)
  p.e_list.clear();
  int off = 0;
  for (int v : order)
    if (color[v] == 2) {
      p.pbs[v].is_e = 1;
      p.pbs[v].offset = off;
      off += p.pbs[v].size;
      p.e_list.push_back(v);
    }
  p.n_e = off;
  int foff = 0;
  for (int v = 0; v < n; ++v)
    if (color[v] != 2) {
      p.pbs[v].offset = off;
      p.pbs[v].f_off = foff;
      off += p.pbs[v].size;
      foff += p.pbs[v].size;
    }
  p.n_x = off;
  p.n_f = foff;
  // residual blocks per e-block
  std::vector<int> e_slot(n, -1);
  for (size_t i = 0; i < p.e_list.size(); ++i) e_slot[p.e_list[i]] = (int)i;
  std::vector<int> cnt(p.e_list.size() + 1, 0);
  std::vector<int> rb_e(p.n_blk, -1);
  p.rb_noe.clear();
  for (int b = 0; b < p.n_blk; ++b) {
    int ids[3] = {p.pb_cam, p.pb_of_cap[p.cap_idx[b]], p.pb_of_tag[p.tag_idx[b]]};
    for (int i = 0; i < 3; ++i)
      if (ids[i] >= 0 && p.pbs[ids[i]].is_e) rb_e[b] = e_slot[ids[i]];
    if (rb_e[b] >= 0) cnt[rb_e[b] + 1]++; else p.rb_noe.push_back(b);
  }
  for (size_t i = 0; i < p.e_list.size(); ++i) cnt[i + 1] += cnt[i];
  p.e_rb_start = cnt;
  p.e_rb_list.assign(p.n_blk - (int)p.rb_noe.size(), 0);
  std::vector<int> fill(cnt.begin(), cnt.end() - 1);
  for (int b = 0; b < p.n_blk; ++b) if (rb_e[b] >= 0) p.e_rb_list[fill[rb_e[b]]++] = b;
}

// Residuals (8) and the 8x15 Jacobian [cam 3 | cap 6 | tag 6] of one block by
// forward-mode Jets == AutoDiffCostFunction<ArucoReprojectionError,8,3,6,6>
// (ar_slam_util.cpp:720-722).
void eval_block_jets(const double* rect, const double* cam, const double* cap, const double* tag,
                     double tag_size, int model, double* r8, double* J /*8x15 row-major*/) {
  typedef Jet<15> J15;
  J15 jc[3], jp[6], ja[6];
  for (int i = 0; i < 3; ++i) jc[i] = J15(cam[i], i);
  for (int i = 0; i < 6; ++i) jp[i] = J15(cap[i], 3 + i);
  for (int i = 0; i < 6; ++i) ja[i] = J15(tag[i], 9 + i);
  J15 res[8];
  block_residuals<J15>(rect, jc, jp, ja, tag_size, model, res);
  for (int k = 0; k < 8; ++k) {
    r8[k] = res[k].a;
    if (J) for (int j = 0; j < 15; ++j) J[k * 15 + j] = res[k].v[j];
  }
}
void eval_block_values(const double* rect, const double* cam, const double* cap, const double* tag,
                       double tag_size, int model, double* r8) {
  block_residuals<double>(rect, cam, cap, tag, tag_size, model, r8);
}

struct State {
  std::vector<double> r;  // 8 n_blk
  std::vector<double> J;  // 8*15 n_blk (unscaled then scaled in place)
};

double evaluate(const Problem& p, const double* cam, const double* cap, const double* tag,
                std::vector<double>* r, std::vector<double>* J, int nthreads) {
  double cost = 0.0;
  std::vector<double> block_cost(p.n_blk);
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int b = 0; b < p.n_blk; ++b) {
    double r8[8];
    const double* rect = p.obs + 8 * b;
    const double* c = cap + 6 * p.cap_idx[b];
    const double* a = tag + 6 * p.tag_idx[b];
    if (J) eval_block_jets(rect, cam, c, a, p.tag_size, p.model, r8, J->data() + 120 * (size_t)b);
    else eval_block_values(rect, cam, c, a, p.tag_size, p.model, r8);
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += r8[k] * r8[k];
    block_cost[b] = 0.5 * s;
    if (r) std::memcpy(r->data() + 8 * (size_t)b, r8, sizeof(r8));
  }
  for (int b = 0; b < p.n_blk; ++b) cost += block_cost[b];  // fixed order
  return cost;
}

// column range of parameter block pb inside a block's 8x15 Jacobian
inline int jcol(const ParamBlock& pb) { return pb.type == 0 ? 0 : (pb.type == 1 ? 3 : 9); }

// x <-> parameter arrays
void gather_x(const Problem& p, const double* cam, const double* cap, const double* tag, double* x) {
  for (const auto& pb : p.pbs) {
    const double* src = pb.type == 0 ? cam : (pb.type == 1 ? cap + 6 * pb.index : tag + 6 * pb.index);
    for (int i = 0; i < pb.size; ++i) x[pb.offset + i] = src[i];
  }
}
void scatter_x(const Problem& p, const double* x, double* cam, double* cap, double* tag) {
  for (const auto& pb : p.pbs) {
    double* dst = pb.type == 0 ? cam : (pb.type == 1 ? cap + 6 * pb.index : tag + 6 * pb.index);
    for (int i = 0; i < pb.size; ++i) dst[i] = x[pb.offset + i];
  }
}

inline void block_pbs(const Problem& p, int b, int ids[3]) {
  ids[0] = p.pb_cam;
  ids[1] = p.pb_of_cap[p.cap_idx[b]];
  ids[2] = p.pb_of_tag[p.tag_idx[b]];
}

// g = J' r, and squared column norms of J, over the reduced program.
void jt_r_and_colnorm2(const Problem& p, const std::vector<double>& J, const std::vector<double>& r,
                       double* g, double* cn2) {
  if (g) std::fill(g, g + p.n_x, 0.0);
  if (cn2) std::fill(cn2, cn2 + p.n_x, 0.0);
  for (int b = 0; b < p.n_blk; ++b) {
    int ids[3];
    block_pbs(p, b, ids);
    const double* Jb = J.data() + 120 * (size_t)b;
    const double* rb = r.data() + 8 * (size_t)b;
    for (int t = 0; t < 3; ++t) {
      if (ids[t] < 0) continue;
      const ParamBlock& pb = p.pbs[ids[t]];
      const int c0 = jcol(pb);
      for (int k = 0; k < 8; ++k)
        for (int i = 0; i < pb.size; ++i) {
          const double v = Jb[k * 15 + c0 + i];
          if (g) g[pb.offset + i] += v * rb[k];
          if (cn2) cn2[pb.offset + i] += v * v;
        }
    }
  }
}

void scale_columns(const Problem& p, std::vector<double>& J, const double* scale) {
  for (int b = 0; b < p.n_blk; ++b) {
    int ids[3];
    block_pbs(p, b, ids);
    double* Jb = J.data() + 120 * (size_t)b;
    for (int t = 0; t < 3; ++t) {
      if (ids[t] < 0) continue;
      const ParamBlock& pb = p.pbs[ids[t]];
      const int c0 = jcol(pb);
      for (int k = 0; k < 8; ++k)
        for (int i = 0; i < pb.size; ++i) Jb[k * 15 + c0 + i] *= scale[pb.offset + i];
    }
  }
}

// y = J * x over the reduced program (residual space, 8 n_blk)
void right_multiply(const Problem& p, const std::vector<double>& J, const double* x, double* y) {
  for (int b = 0; b < p.n_blk; ++b) {
    int ids[3];
    block_pbs(p, b, ids);
    const double* Jb = J.data() + 120 * (size_t)b;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < 3; ++t) {
      if (ids[t] < 0) continue;
      const ParamBlock& pb = p.pbs[ids[t]];
      const int c0 = jcol(pb);
      for (int k = 0; k < 8; ++k)
        for (int i = 0; i < pb.size; ++i) acc[k] += Jb[k * 15 + c0 + i] * x[pb.offset + i];
    }
    for (int k = 0; k < 8; ++k) y[8 * (size_t)b + k] = acc[k];
  }
}

// In-place lower Cholesky of a dense row-major n x n matrix (only the lower
// triangle is referenced).  Fails like Eigen::LLT: first non-positive pivot.
bool dense_llt(double* A, int n, int nthreads) {
  for (int j = 0; j < n; ++j) {
    double* Aj = A + (size_t)j * n;
    double d = Aj[j];
    {
      double s = 0.0;
#pragma omp simd reduction(+ : s)
      for (int k = 0; k < j; ++k) s += Aj[k] * Aj[k];
      d -= s;
    }
    if (!(d > 0.0)) return false;
    const double ljj = std::sqrt(d);
    Aj[j] = ljj;
    const double inv = 1.0 / ljj;
#pragma omp parallel for num_threads(nthreads) schedule(static) if (n - j > 256)
    for (int i = j + 1; i < n; ++i) {
      double* Ai = A + (size_t)i * n;
      double s = 0.0;
#pragma omp simd reduction(+ : s)
      for (int k = 0; k < j; ++k) s += Ai[k] * Aj[k];
      Ai[j] = (Ai[j] - s) * inv;
    }
  }
  return true;
}
void dense_llt_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    const double* Li = L + (size_t)i * n;
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= Li[k] * b[k];
    b[i] = s / Li[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}
// small (<=6) SPD inverse through LLT.solve(Identity), like InvertPSDMatrix
bool small_inverse(const double* M, int n, double* inv) {
  double L[36];
  std::memcpy(L, M, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    double d = L[j * n + j];
    for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
    if (!(d > 0.0)) return false;
    const double ljj = std::sqrt(d);
    L[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double s = L[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / ljj;
    }
  }
  for (int c = 0; c < n; ++c) {
    double b[6] = {0, 0, 0, 0, 0, 0};
    b[c] = 1.0;
    for (int i = 0; i < n; ++i) {
      double s = b[i];
      for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
      b[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = b[i];
      for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
      b[i] = s / L[i * n + i];
    }
    for (int i = 0; i < n; ++i) inv[i * n + c] = b[i];
  }
  return true;
}

// Solve (J'J + D'D) y = J' r through the dense Schur complement:
// schur_eliminator_impl.h (Eliminate / BackSubstitute) +
// DenseSchurComplementSolver::SolveReducedLinearSystem (Eigen LLT).
// Returns false on LINEAR_SOLVER_FAILURE.
bool schur_solve(const Problem& p, const std::vector<double>& J, const std::vector<double>& r,
                 const double* D, double* y, int nthreads, std::vector<double>& S_buf) {
  const int nf = p.n_f;
  const int ne_blocks = (int)p.e_list.size();
  S_buf.assign((size_t)nf * nf, 0.0);
  double* S = S_buf.data();
  std::vector<double> rhs(nf, 0.0);
  // D_f^2 on the diagonal of the reduced system
  for (const auto& pb : p.pbs)
    if (!pb.is_e)
      for (int i = 0; i < pb.size; ++i) {
        const double d = D[pb.offset + i];
        S[(size_t)(pb.f_off + i) * nf + pb.f_off + i] = d * d;
      }
  const bool par = nthreads > 1;
  auto add = [&](double* dst, double v) {
    if (par) {
#pragma omp atomic
      *dst += v;
    } else {
      *dst += v;
    }
  };
  // F'F and F'b contributions of one residual block (lower triangle of S only)
  auto add_ff = [&](int b, int skip_pb) {
    int ids[3];
    block_pbs(p, b, ids);
    const double* Jb = J.data() + 120 * (size_t)b;
    const double* rb = r.data() + 8 * (size_t)b;
    for (int t = 0; t < 3; ++t) {
      if (ids[t] < 0 || ids[t] == skip_pb) continue;
      const ParamBlock& pf = p.pbs[ids[t]];
      const int c0 = jcol(pf);
      for (int i = 0; i < pf.size; ++i) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += Jb[k * 15 + c0 + i] * rb[k];
        add(&rhs[pf.f_off + i], s);
      }
      for (int u = 0; u < 3; ++u) {
        if (ids[u] < 0 || ids[u] == skip_pb) continue;
        const ParamBlock& pg = p.pbs[ids[u]];
        if (pg.f_off > pf.f_off) continue;  // lower triangle
        const int d0 = jcol(pg);
        for (int i = 0; i < pf.size; ++i)
          for (int j = 0; j < pg.size; ++j) {
            if (pg.f_off == pf.f_off && j > i) continue;
            double s = 0.0;
            for (int k = 0; k < 8; ++k) s += Jb[k * 15 + c0 + i] * Jb[k * 15 + d0 + j];
            add(&S[(size_t)(pf.f_off + i) * nf + pg.f_off + j], s);
          }
      }
    }
  };
  std::vector<double> inv_ete((size_t)ne_blocks * 36), g_e((size_t)ne_blocks * 6);
  bool ok = true;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
  for (int e = 0; e < ne_blocks; ++e) {
    const ParamBlock& pe = p.pbs[p.e_list[e]];
    const int es = pe.size, e0 = jcol(pe);
    double ete[36] = {0}, ge[6] = {0};
    for (int i = 0; i < es; ++i) ete[i * es + i] = D[pe.offset + i] * D[pe.offset + i];
    // buffer: E'F for every f block met in this chunk
    struct FB { int pb; double m[36]; };
    std::vector<FB> fbs;
    for (int q = p.e_rb_start[e]; q < p.e_rb_start[e + 1]; ++q) {
      const int b = p.e_rb_list[q];
      const double* Jb = J.data() + 120 * (size_t)b;
      const double* rb = r.data() + 8 * (size_t)b;
      for (int i = 0; i < es; ++i) {
        for (int j = 0; j < es; ++j) {
          double s = 0.0;
          for (int k = 0; k < 8; ++k) s += Jb[k * 15 + e0 + i] * Jb[k * 15 + e0 + j];
          ete[i * es + j] += s;
        }
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += Jb[k * 15 + e0 + i] * rb[k];
        ge[i] += s;
      }
      int ids[3];
      block_pbs(p, b, ids);
      for (int t = 0; t < 3; ++t) {
        if (ids[t] < 0 || ids[t] == p.e_list[e]) continue;
        const ParamBlock& pf = p.pbs[ids[t]];
        FB* fb = nullptr;
        for (auto& x : fbs) if (x.pb == ids[t]) fb = &x;
        if (!fb) { fbs.push_back(FB{ids[t], {0}}); fb = &fbs.back(); }
        const int c0 = jcol(pf);
        for (int i = 0; i < es; ++i)
          for (int j = 0; j < pf.size; ++j) {
            double s = 0.0;
            for (int k = 0; k < 8; ++k) s += Jb[k * 15 + e0 + i] * Jb[k * 15 + c0 + j];
            fb->m[i * 6 + j] += s;
          }
      }
      add_ff(b, p.e_list[e]);
    }
    double* inv = inv_ete.data() + (size_t)e * 36;
    if (!small_inverse(ete, es, inv)) {
      for (int i = 0; i < es * es; ++i) inv[i] = std::numeric_limits<double>::quiet_NaN();
    }
    for (int i = 0; i < es; ++i) g_e[(size_t)e * 6 + i] = ge[i];
    // S -= (E'F)' inv (E'F) ; rhs -= (E'F)' inv g_e
    double ig[6];
    for (int i = 0; i < es; ++i) {
      double s = 0.0;
      for (int j = 0; j < es; ++j) s += inv[i * es + j] * ge[j];
      ig[i] = s;
    }
    for (auto& f1 : fbs) {
      const ParamBlock& pf = p.pbs[f1.pb];
      for (int i = 0; i < pf.size; ++i) {
        double s = 0.0;
        for (int k = 0; k < es; ++k) s += f1.m[k * 6 + i] * ig[k];
        add(&rhs[pf.f_off + i], -s);
      }
      // tmp = inv * E'F2 then F1' tmp
      for (auto& f2 : fbs) {
        const ParamBlock& pg = p.pbs[f2.pb];
        if (pg.f_off > pf.f_off) continue;
        double tmp[36];
        for (int k = 0; k < es; ++k)
          for (int j = 0; j < pg.size; ++j) {
            double s = 0.0;
            for (int l = 0; l < es; ++l) s += inv[k * es + l] * f2.m[l * 6 + j];
            tmp[k * 6 + j] = s;
          }
        for (int i = 0; i < pf.size; ++i)
          for (int j = 0; j < pg.size; ++j) {
            if (pg.f_off == pf.f_off && j > i) continue;
            double s = 0.0;
            for (int k = 0; k < es; ++k) s += f1.m[k * 6 + i] * tmp[k * 6 + j];
            add(&S[(size_t)(pf.f_off + i) * nf + pg.f_off + j], -s);
          }
      }
    }
  }
  for (int b : p.rb_noe) add_ff(b, -1);
  (void)ok;
  // reduced system
  std::vector<double> yf(rhs);
  if (nf > 0) {
    if (!dense_llt(S, nf, nthreads)) return false;
    dense_llt_solve(S, nf, yf.data());
  }
  for (const auto& pb : p.pbs)
    if (!pb.is_e)
      for (int i = 0; i < pb.size; ++i) y[pb.offset + i] = yf[pb.f_off + i];
  // back substitution: y_e = inv (g_e - E'F y_f)
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int e = 0; e < ne_blocks; ++e) {
    const ParamBlock& pe = p.pbs[p.e_list[e]];
    const int es = pe.size, e0 = jcol(pe);
    double t[6];
    for (int i = 0; i < es; ++i) t[i] = g_e[(size_t)e * 6 + i];
    for (int q = p.e_rb_start[e]; q < p.e_rb_start[e + 1]; ++q) {
      const int b = p.e_rb_list[q];
      const double* Jb = J.data() + 120 * (size_t)b;
      int ids[3];
      block_pbs(p, b, ids);
      // F y_f in residual space, then E' of it
      double fy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int u = 0; u < 3; ++u) {
        if (ids[u] < 0 || ids[u] == p.e_list[e]) continue;
        const ParamBlock& pf = p.pbs[ids[u]];
        const int c0 = jcol(pf);
        for (int k = 0; k < 8; ++k)
          for (int j = 0; j < pf.size; ++j) fy[k] += Jb[k * 15 + c0 + j] * yf[pf.f_off + j];
      }
      for (int i = 0; i < es; ++i) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += Jb[k * 15 + e0 + i] * fy[k];
        t[i] -= s;
      }
    }
    const double* inv = inv_ete.data() + (size_t)e * 36;
    for (int i = 0; i < es; ++i) {
      double s = 0.0;
      for (int j = 0; j < es; ++j) s += inv[i * es + j] * t[j];
      y[pe.offset + i] = s;
    }
  }
  return true;
}

inline double norm2(const double* v, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return std::sqrt(s);
}

}  // namespace

// Trust-region Levenberg-Marquardt, Ceres 2.0.0 control flow
// (TrustRegionMinimizer::Minimize and helpers; LevenbergMarquardtStrategy).
int solve(Problem& p, const oracle_options& o, oracle_summary* sum, double* iter_log, int log_cap) {
  const double t_start = now_s();
  const int nthreads = std::max(1, o.num_threads);
  build_program(p, o.elimination);
  const int n = p.n_x;
  std::memset(sum, 0, sizeof(*sum));
  sum->n_e_blocks = (int)p.e_list.size();
  sum->reduced_dim = p.n_f;
  sum->num_parameters = n;
  std::vector<double> x(n), cand_x(n), g(n), scale(n, 1.0), diag(n), D(n), step(n), delta(n);
  std::vector<double> r(8 * (size_t)p.n_blk), J(120 * (size_t)p.n_blk), model_r(8 * (size_t)p.n_blk);
  std::vector<double> cam(p.cam, p.cam + 3), cap(p.cap, p.cap + 6 * (size_t)p.n_cap),
      tag(p.tag, p.tag + 6 * (size_t)p.n_tag);
  std::vector<double> S_buf;
  gather_x(p, cam.data(), cap.data(), tag.data(), x.data());
  double x_norm = norm2(x.data(), n);
  double radius = o.initial_trust_region_radius;
  double decrease_factor = 2.0;
  bool reuse_diagonal = false;
  bool have_scale = false;
  int consecutive_invalid = 0;
  double x_cost = 0.0, grad_max = 0.0, grad_norm = 0.0;
  double t_jac = 0.0, t_lin = 0.0;

  auto eval_grad_jac = [&]() {
    const double t0 = now_s();
    x_cost = evaluate(p, cam.data(), cap.data(), tag.data(), &r, &J, nthreads);
    jt_r_and_colnorm2(p, J, r, g.data(), nullptr);  // gradient from the UNSCALED Jacobian
    if (o.jacobi_scaling) {
      if (!have_scale) {
        jt_r_and_colnorm2(p, J, r, nullptr, scale.data());
        for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(scale[i]));
        have_scale = true;
      }
      scale_columns(p, J, scale.data());
    }
    // ||x - Plus(x, -g)||_inf  (no local parameterisation: equals ||g||_inf)
    grad_max = 0.0;
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
      const double d = x[i] - (x[i] - g[i]);
      grad_max = std::max(grad_max, std::fabs(d));
      s += d * d;
    }
    grad_norm = std::sqrt(s);
    t_jac += now_s() - t0;
  };

  auto log_iter = [&](int it, double cost, double cost_change, double step_norm, double rho,
                      int valid, int successful) {
    if (iter_log && it < log_cap) {
      double* L = iter_log + 8 * (size_t)it;
      L[0] = cost; L[1] = cost_change; L[2] = grad_max; L[3] = step_norm;
      L[4] = rho; L[5] = radius; L[6] = valid; L[7] = successful;
    }
  };

  // iteration 0
  eval_grad_jac();
  sum->initial_cost = x_cost;
  int iteration = 0;
  bool last_successful = true;
  log_iter(0, x_cost, 0.0, 0.0, 0.0, 1, 1);
  sum->num_successful_steps = 1;
  int termination = ORACLE_NO_CONVERGENCE, reason = ORACLE_REASON_MAX_ITERATIONS;

  while (true) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue (tests in Ceres' order)
    if (iteration >= o.max_num_iterations) {
      termination = ORACLE_NO_CONVERGENCE; reason = ORACLE_REASON_MAX_ITERATIONS; break; }
    if (last_successful && grad_max <= o.gradient_tolerance) {
      termination = ORACLE_CONVERGENCE; reason = ORACLE_REASON_GRADIENT; break; }
    if (radius <= o.min_trust_region_radius) {
      termination = ORACLE_CONVERGENCE; reason = ORACLE_REASON_MIN_RADIUS; break; }
    ++iteration;
    last_successful = false;

    // LevenbergMarquardtStrategy::ComputeStep
    const double t0 = now_s();
    if (!reuse_diagonal) {
      jt_r_and_colnorm2(p, J, r, nullptr, diag.data());
      for (int i = 0; i < n; ++i)
        diag[i] = std::min(std::max(diag[i], o.min_lm_diagonal), o.max_lm_diagonal);
    }
    for (int i = 0; i < n; ++i) D[i] = std::sqrt(diag[i] / radius);
    for (int i = 0; i < n; ++i) step[i] = std::numeric_limits<double>::quiet_NaN();
    bool lin_ok = schur_solve(p, J, r, D.data(), step.data(), nthreads, S_buf);
    if (lin_ok)
      for (int i = 0; i < n; ++i) if (!std::isfinite(step[i])) { lin_ok = false; break; }
    if (lin_ok) for (int i = 0; i < n; ++i) step[i] = -step[i];
    reuse_diagonal = true;
    t_lin += now_s() - t0;

    bool step_valid = false;
    double model_cost_change = 0.0;
    if (lin_ok) {
      right_multiply(p, J, step.data(), model_r.data());
      double s = 0.0;
      for (size_t k = 0; k < model_r.size(); ++k) s += model_r[k] * (r[k] + model_r[k] / 2.0);
      model_cost_change = -s;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {
      // HandleInvalidStep
      if (++consecutive_invalid >= o.max_num_consecutive_invalid_steps) {
        termination = ORACLE_FAILURE; reason = ORACLE_REASON_INVALID_STEPS;
        sum->num_unsuccessful_steps++;
        break;
      }
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      sum->num_unsuccessful_steps++;
      log_iter(iteration, x_cost, 0.0, 0.0, 0.0, 0, 0);
      continue;
    }
    consecutive_invalid = 0;
    for (int i = 0; i < n; ++i) delta[i] = step[i] * scale[i];

    // ComputeCandidatePointAndEvaluateCost
    for (int i = 0; i < n; ++i) cand_x[i] = x[i] + delta[i];
    std::vector<double> ccam(cam), ccap(cap), ctag(tag);
    scatter_x(p, cand_x.data(), ccam.data(), ccap.data(), ctag.data());
    double cand_cost = evaluate(p, ccam.data(), ccap.data(), ctag.data(), nullptr, nullptr, nthreads);
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();

    // ParameterToleranceReached
    double sn = 0.0;
    for (int i = 0; i < n; ++i) { const double d = x[i] - cand_x[i]; sn += d * d; }
    const double step_norm = std::sqrt(sn);
    const double cost_change = x_cost - cand_cost;
    if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) {
      termination = ORACLE_CONVERGENCE; reason = ORACLE_REASON_PARAMETER;
      log_iter(iteration, x_cost, cost_change, step_norm, 0.0, 1, 0);
      break;
    }
    // FunctionToleranceReached
    if (std::fabs(cost_change) <= o.function_tolerance * x_cost) {
      termination = ORACLE_CONVERGENCE; reason = ORACLE_REASON_FUNCTION;
      log_iter(iteration, x_cost, cost_change, step_norm, 0.0, 1, 0);
      break;
    }
    const double rho = cost_change / model_cost_change;
    if (rho > o.min_relative_decrease) {
      // HandleSuccessfulStep
      x = cand_x;
      x_norm = norm2(x.data(), n);
      cam = ccam; cap = ccap; tag = ctag;
      eval_grad_jac();
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      last_successful = true;
      sum->num_successful_steps++;
      log_iter(iteration, x_cost, cost_change, step_norm, rho, 1, 1);
    } else {
      // HandleUnsuccessfulStep
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      sum->num_unsuccessful_steps++;
      log_iter(iteration, cand_cost, cost_change, step_norm, rho, 1, 0);
    }
  }
  std::memcpy(p.cam, cam.data(), sizeof(double) * 3);
  std::memcpy(p.cap, cap.data(), sizeof(double) * 6 * (size_t)p.n_cap);
  std::memcpy(p.tag, tag.data(), sizeof(double) * 6 * (size_t)p.n_tag);
  sum->iterations = iteration;
  sum->final_cost = x_cost;
  sum->termination = termination;
  sum->reason = reason;
  sum->final_radius = radius;
  sum->gradient_max_norm = grad_max;
  sum->jacobian_seconds = t_jac;
  sum->linear_solver_seconds = t_lin;
  sum->total_seconds = now_s() - t_start;
  return 0;
}

}  // namespace oracle

// ---------------------------------------------------------------- C API ----
using namespace oracle;

extern "C" {

void oracle_default_options(oracle_options* o) {
  // ar_slam_util.cpp:1003-1012 + Ceres 2.0.0 Solver::Options defaults
  o->max_num_iterations = 50;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->max_num_consecutive_invalid_steps = 5;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->jacobi_scaling = 1;
  o->num_threads = 1;
  o->elimination = 0;
}

void oracle_project_block(const double* cam, const double* cap, const double* tag, double tag_size,
                          int model, double* uv8) {
  for (unsigned i = 0; i < 4; ++i) project_corner<double>(cam, cap, tag, i, tag_size, model, uv8 + 2 * i);
}

int oracle_evaluate(int n_blk, const int32_t* cap_idx, const int32_t* tag_idx, const double* obs,
                    const double* cam, const double* cap, const double* tag, double tag_size, int model,
                    int num_threads, double* cost, double* residuals, double* jac_cam, double* jac_cap,
                    double* jac_tag) {
  const int nt = std::max(1, num_threads);
  std::vector<double> bc(n_blk);
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int b = 0; b < n_blk; ++b) {
    double r8[8], J[120];
    const bool want_j = jac_cam || jac_cap || jac_tag;
    if (want_j)
      eval_block_jets(obs + 8 * (size_t)b, cam, cap + 6 * (size_t)cap_idx[b], tag + 6 * (size_t)tag_idx[b],
                      tag_size, model, r8, J);
    else
      eval_block_values(obs + 8 * (size_t)b, cam, cap + 6 * (size_t)cap_idx[b], tag + 6 * (size_t)tag_idx[b],
                        tag_size, model, r8);
    double s = 0.0;
    for (int k = 0; k < 8; ++k) {
      s += r8[k] * r8[k];
      if (residuals) residuals[8 * (size_t)b + k] = r8[k];
      if (jac_cam) for (int j = 0; j < 3; ++j) jac_cam[(size_t)b * 24 + k * 3 + j] = J[k * 15 + j];
      if (jac_cap) for (int j = 0; j < 6; ++j) jac_cap[(size_t)b * 48 + k * 6 + j] = J[k * 15 + 3 + j];
      if (jac_tag) for (int j = 0; j < 6; ++j) jac_tag[(size_t)b * 48 + k * 6 + j] = J[k * 15 + 9 + j];
    }
    bc[b] = 0.5 * s;
  }
  double c = 0.0;
  for (int b = 0; b < n_blk; ++b) c += bc[b];
  if (cost) *cost = c;
  return 0;
}

void oracle_init_capture_pose(const double* rect8, const double* cam, const double* tag_pose,
                              double tag_size, double* cap_pose_out) {
  init_capture_pose(rect8, cam, tag_pose, tag_size, cap_pose_out);
}
void oracle_init_tag_pose(const double* rect8, const double* cam, const double* cap_pose,
                          double tag_size, double* tag_pose_out) {
  init_ar_pose(rect8, cam, cap_pose, tag_size, tag_pose_out);
}
void oracle_compose_axis_angle(const double* r1, const double* r2, double* out) {
  compose_axis_angle(r1, r2, out);
}
void oracle_rotate_point(const double* aa, const double* pt, double* out) {
  angle_axis_rotate_point<double>(aa, pt, out);
}

int oracle_solve(int n_cap, int n_tag, int n_blk, const int32_t* cap_idx, const int32_t* tag_idx,
                 const double* obs, double tag_size, int model, int cam_const, const uint8_t* cap_const,
                 const uint8_t* tag_const, const oracle_options* opt, double* cam, double* cap, double* tag,
                 oracle_summary* summary, double* iter_log, int log_cap) {
  Problem p;
  p.n_cap = n_cap; p.n_tag = n_tag; p.n_blk = n_blk;
  p.cap_idx = cap_idx; p.tag_idx = tag_idx; p.obs = obs;
  p.tag_size = tag_size; p.model = model;
  p.cam_const = cam_const; p.cap_const = cap_const; p.tag_const = tag_const;
  p.cam = cam; p.cap = cap; p.tag = tag;
  return solve(p, *opt, summary, iter_log, log_cap);
}

// localizeOne (ar_slam_util.cpp:903-979) for a batch of independent captures:
// seed from blk_offsets[i] + seed_block[i] (initCapturePose), then a fresh
// problem with every tag and the camera constant.  seed_block[i] < 0 leaves
// the capture untouched (no tag shared with the map, :929-933).
int oracle_localize_batch(int n_loc, const int32_t* blk_offsets, const int32_t* tag_idx, const double* obs,
                          const int32_t* seed_block, int n_tag, const double* cam, const double* tag,
                          double tag_size, int model, const oracle_options* opt, int num_threads,
                          double* cap_pose, int32_t* iterations, double* final_cost, int32_t* termination) {
  const int nt = std::max(1, num_threads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 64)
  for (int i = 0; i < n_loc; ++i) {
    const int b0 = blk_offsets[i], nb = blk_offsets[i + 1] - b0;
    if (seed_block[i] < 0 || nb <= 0) {
      if (iterations) iterations[i] = -1;
      if (final_cost) final_cost[i] = 0.0;
      if (termination) termination[i] = -1;
      continue;
    }
    double* pose = cap_pose + 6 * (size_t)i;
    const int sb = b0 + seed_block[i];
    init_capture_pose(obs + 8 * (size_t)sb, cam, tag + 6 * (size_t)tag_idx[sb], tag_size, pose);
    // compact problem: one capture, its nb tags (all constant), constant camera
    std::vector<int32_t> ci(nb, 0), ti(nb);
    std::vector<double> cam_c(cam, cam + 3), tag_c(6 * (size_t)nb);
    std::vector<uint8_t> tconst(nb, 1);
    for (int j = 0; j < nb; ++j) {
      ti[j] = j;
      std::memcpy(tag_c.data() + 6 * (size_t)j, tag + 6 * (size_t)tag_idx[b0 + j], 6 * sizeof(double));
    }
    Problem p;
    p.n_cap = 1; p.n_tag = nb; p.n_blk = nb;
    p.cap_idx = ci.data(); p.tag_idx = ti.data(); p.obs = obs + 8 * (size_t)b0;
    p.tag_size = tag_size; p.model = model;
    p.cam_const = 1; p.cap_const = nullptr; p.tag_const = tconst.data();
    p.cam = cam_c.data(); p.cap = pose; p.tag = tag_c.data();
    oracle_options o = *opt;
    o.num_threads = 1;
    oracle_summary s;
    solve(p, o, &s, nullptr, 0);
    if (iterations) iterations[i] = s.iterations;
    if (final_cost) final_cost[i] = s.final_cost;
    if (termination) termination[i] = s.termination;
  }
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
