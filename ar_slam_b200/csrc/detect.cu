// detect.cu -- marker detection on the GPU: the step before the optimisation path.
//
// Replaces cv::aruco::detectMarkers as the reference calls it (ar_slam/src/aruco_detector.cpp:106,
// ar_slam/src/ar_slam_util.cpp:268).  Stages, in stream order (all frames of a batch at once):
//   (d1) gray_threshold_kernel   BGR -> grey (cvtColor's fixed point) + every adaptive-threshold window from ONE
//                                shared-memory integral tile; one byte per pixel out, bit k = window k
//   (d2) border_starts_kernel    every place a border can start (0->1 and 1->0 steps of each window), compacted
//        border_follow_kernel    one thread per start follows its border (Suzuki-Abe, as cv::findContours); a start
//                                that meets a raster-earlier start of the same border stops: exactly the borders the
//                                sequential scan would trace survive, with the same first point and direction
//   (d3) approx_quad_kernel      a warp per border: cv::approxPolyDP (farthest-point searches as warp arg-max
//                                reductions), convexity, corner-distance and image-border tests, clockwise order
//   (d4) identify_kernel         a warp per quad: homography, nearest-neighbour perspective removal, Otsu, cell
//                                votes, border check, dictionary match over all four rotations
// The grouping of near-duplicate quads (tens per frame, O(n^2) on float corners) is host code below, like the
// reference's own schedules.  tests/test_gpu_detect.py compares every stage with the CPU restatement of the same
// algorithm (test infrastructure, pinned against cv2) and the result with committed cv2 golden corners.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ar_slam_b200.h"

namespace ard {

constexpr int MAX_WIN = 8;        // adaptive-threshold windows per call
constexpr int MAX_RADIUS = 15;    // window <= 31
constexpr int TX = 64, TY = 64;   // output tile of the threshold kernel
constexpr int MAX_MARKER = 6;     // marker_size <= 6: (6 + 2) * 4 = 32 pixel wide canonical image at 4 pixels per cell
constexpr int LPAD = 16;          // zero bytes left of every row of the threshold-bit frames (keeps 8-byte loads aligned)
constexpr int CANON_SIDE = 48;     // canonical image side limit ((marker_size + 2 border) * pixels per cell)

struct Windows {
  int n;
  int radius[MAX_WIN];
  int rmax;
  int idelta;
};

// ---------------------------------------------------------------------------------------------- (d1)
template <int CH>
__device__ __forceinline__ int grey_at(const uint8_t* img, size_t idx) {
  if (CH == 1) return img[idx];
  const uint8_t* p = img + idx * 3;
  return (p[0] * 1868 + p[1] * 9617 + p[2] * 4899 + (1 << 13)) >> 14;   // cv::cvtColor BGR2GRAY, 8 bit
}

// One CTA per TX x TY tile of one frame.  Shared memory: the integral image of the tile with its halo (replicated
// image borders, like BORDER_REPLICATE) and the grey tile itself.  A warp loads a row, converts it to grey and scans it
// in the same pass; columns are then summed by a thread each; every thread finishes four neighbouring pixels.
// R = halo = the largest window radius the kernel serves (11 for the default windows 3 / 13 / 23, 15 in general).
template <int CH, int R>
__global__ void __launch_bounds__(256) gray_threshold_kernel(const uint8_t* __restrict__ images, int W, int H,
                                                             uint8_t* __restrict__ gray, uint8_t* __restrict__ mask,
                                                             int P, Windows wins) {
  constexpr int TW = TX + 2 * R, TH = TY + 2 * R, LD = TW + 1, GP = (TW + 3) / 4 * 4;
  extern __shared__ uint32_t integ[];          // (TH + 1) x LD, then the grey tile TH x GP bytes
  uint8_t* gt = reinterpret_cast<uint8_t*>(integ + (TH + 1) * LD);
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const uint8_t* img = images + (size_t)b * W * H * CH;
  uint8_t* g_out = gray + (size_t)b * W * H;
  uint8_t* m_out = mask + (size_t)b * (H + 2) * P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < LD; i += 256) integ[i] = 0;
  for (int j = tid; j < TH + 1; j += 256) integ[j * LD] = 0;
  for (int j = warp; j < TH; j += 8) {
    const int gy = min(max(y0 + j - R, 0), H - 1);
    const size_t row = (size_t)gy * W;
    uint32_t carry = 0;
#pragma unroll
    for (int c = 0; c < TW; c += 32) {
      const int i = c + lane;
      uint32_t v = 0;
      if (i < TW) {
        const int gx = min(max(x0 + i - R, 0), W - 1);
        v = (uint32_t)grey_at<CH>(img, row + gx);
        gt[j * GP + i] = (uint8_t)v;
      }
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
      }
      v += carry;
      if (i < TW) integ[(j + 1) * LD + i + 1] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  for (int i = tid; i < TW; i += 256) {       // columns
    uint32_t acc = 0;
#pragma unroll 8
    for (int j = 0; j < TH; ++j) {
      acc += integ[(j + 1) * LD + i + 1];
      integ[(j + 1) * LD + i + 1] = acc;
    }
  }
  __syncthreads();
  int radius[MAX_WIN], w2[MAX_WIN];
#pragma unroll
  for (int k = 0; k < MAX_WIN; ++k) { radius[k] = k < wins.n ? wins.radius[k] : 0; w2[k] = (2 * radius[k] + 1) * (2 * radius[k] + 1); }
  const bool vec_grey = (W & 3) == 0;
  for (int q = tid; q < TX * TY / 4; q += 256) {
    const int ty = q / (TX / 4), tx = (q - ty * (TX / 4)) * 4;
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= W || gy >= H) continue;
    uint32_t gpack = 0, mpack = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int cy = ty + R, cx = tx + e + R;     // tile coordinates; integral index = +1
      const int g = gt[cy * GP + cx];
      unsigned bits = 0;
#pragma unroll
      for (int k = 0; k < MAX_WIN; ++k) {
        if (k < wins.n) {
          const int r = radius[k];
          const int sum = (int)(integ[(cy + r + 1) * LD + cx + r + 1] - integ[(cy - r) * LD + cx + r + 1] -
                                integ[(cy + r + 1) * LD + cx - r] + integ[(cy - r) * LD + cx - r]);
          // THRESH_BINARY_INV: g - mean <= -delta with mean = floor((2 sum + w2) / (2 w2)), the rounded box mean
          // (w2 is odd: no ties); floor(a / b) >= c  <=>  a >= c b, so no division
          if (2 * sum + w2[k] >= (g + wins.idelta) * 2 * w2[k]) bits |= 1u << k;
        }
      }
      gpack |= (uint32_t)g << (8 * e);
      mpack |= bits << (8 * e);
    }
    uint8_t* gp_ = g_out + (size_t)gy * W + gx;
    uint8_t* mp_ = m_out + (size_t)(gy + 1) * P + LPAD + gx;
    if (gx + 3 < W) {
      *reinterpret_cast<uint32_t*>(mp_) = mpack;                    // P and LPAD are multiples of 16, gx of 4
      if (vec_grey) *reinterpret_cast<uint32_t*>(gp_) = gpack;
      else { gp_[0] = gpack; gp_[1] = gpack >> 8; gp_[2] = gpack >> 16; gp_[3] = gpack >> 24; }
    } else {
      for (int e = 0; gx + e < W; ++e) { gp_[e] = gpack >> (8 * e); mp_[e] = mpack >> (8 * e); }
    }
  }
}

// ---------------------------------------------------------------------------------------------- (d2)
// A border start: pixel index inside the padded frame, window k, kind 0 (its west neighbour is a zero of the
// border: outer borders start there) or 1 (east neighbour: hole borders).  Key order = raster order, kind 0 first.
__global__ void __launch_bounds__(256) border_starts_kernel(const uint8_t* __restrict__ mask, int W, int H, int P, int n_img,
                                                            int drop_isolated, unsigned long long* __restrict__ starts,
                                                            unsigned long long cap, unsigned long long* __restrict__ n_starts) {
  // a thread per 8 consecutive pixels: the tests are bitwise inside every byte, so they run on 64-bit words
  const int gpr = (W + 7) / 8;                      // groups per row; the last one may reach into the zero padding
  const size_t total = (size_t)n_img * H * gpr;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long outer = 0, hole = 0;
  size_t gp = 0;
  unsigned long long rec = 0;                       // image << 36 | pixel inside the padded frame << 4
  if (i < total) {
    const int b = (int)(i / ((size_t)H * gpr));
    const size_t r = i - (size_t)b * H * gpr;
    const int y = (int)(r / gpr), x = 8 * (int)(r - (size_t)y * gpr);
    gp = (size_t)b * (H + 2) * P + (size_t)(y + 1) * P + LPAD + x;
    rec = (unsigned long long)b << 36 | (unsigned long long)((y + 1) * P + LPAD + x) << 4;
    const unsigned long long cur = *reinterpret_cast<const unsigned long long*>(mask + gp);
    if (cur) {
      // The start the sequential scan uses for an outer border is the raster-first pixel of its 8-connected
      // component: nothing set at W, NW, N, NE.  For a hole it is the pixel left of the raster-first pixel of a
      // 4-connected background region: E is clear and the pixel above E is set.  Necessary, not sufficient: the
      // border following below still stops at any raster-earlier start of the same border.
      const unsigned long long up = *reinterpret_cast<const unsigned long long*>(mask + gp - P);
      const unsigned long long l = mask[gp - 1], rr = mask[gp + 8], ul = mask[gp - P - 1], ur = mask[gp - P + 8];
      const unsigned long long west = cur << 8 | l, east = cur >> 8 | rr << 56;
      const unsigned long long upw = up << 8 | ul, upe = up >> 8 | ur << 56;
      outer = cur & ~(west | upw | up | upe);
      hole = cur & ~east & upe;
      if (drop_isolated && outer) {
        // a pixel without any set neighbour is a border of one point: most of what noise produces, and never a
        // marker unless one-point borders are admissible (frames under 34 pixels)
        const unsigned long long dn = *reinterpret_cast<const unsigned long long*>(mask + gp + P);
        const unsigned long long dl = mask[gp + P - 1], dr = mask[gp + P + 8];
        outer &= east | dn | (dn << 8 | dl) | (dn >> 8 | dr << 56);
      }
    }
  }
  int cnt = __popcll(outer) + __popcll(hole);
  int incl = cnt;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  int tot = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if (lane == 31 && tot) base = atomicAdd(n_starts, (unsigned long long)tot);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (!cnt) return;
  unsigned long long o = base + incl - cnt;
  while (outer) {
    int bit = __ffsll((long long)outer) - 1;
    outer &= outer - 1;
    if (o < cap) starts[o] = (rec + ((unsigned long long)(bit >> 3) << 4)) | ((bit & 7) << 1);
    ++o;
  }
  while (hole) {
    int bit = __ffsll((long long)hole) - 1;
    hole &= hole - 1;
    if (o < cap) starts[o] = (rec + ((unsigned long long)(bit >> 3) << 4)) | ((bit & 7) << 1) | 1;
    ++o;
  }
}

struct Border {       // one traced border (contour) that can be a marker candidate
  int img, win;
  int disc;           // where the sequential scan would have discovered it: raster index in the padded frame
  int len, pts_off;
  // results of (d3) / (d4)
  int valid, near_border;
  float quad[8];
  int id, rot;
};

struct FollowArgs {
  const uint8_t* mask;
  int W, H, P, n_img;
  const unsigned long long* starts;
  unsigned long long n_starts;
  unsigned long long* next;          // work counter: warps take chunks of starts from it
  int min_len, max_len;
  Border* borders;
  int border_cap;
  int* n_borders;
  int* pts;           // x | y << 16
  unsigned long long pts_cap;
  unsigned long long* n_pts;
  int* overflow;
};

// Offsets of the eight Freeman directions (0 = east, counter-clockwise on the screen) from two packed tables.
__device__ __forceinline__ int dir_off(int d, int P) {
  const int dx = (int)((0x901au >> (2 * d)) & 3u) - 1;     // {1, 1, 0, -1, -1, -1, 0, 1} + 1, two bits each
  const int dy = (int)((0xa901u >> (2 * d)) & 3u) - 1;     // {0, -1, -1, -1, 0, 1, 1, 1} + 1
  return dx + dy * P;
}
__device__ __forceinline__ unsigned neighbours(const uint8_t* __restrict__ m, int p, int P, int k) {
  // bit d = the neighbour in direction d is set; eight independent byte loads in flight, then the window's bit of
  // each is gathered with one multiply per four bytes
  const unsigned lo = (unsigned)m[p + 1] | (unsigned)m[p + 1 - P] << 8 | (unsigned)m[p - P] << 16 | (unsigned)m[p - 1 - P] << 24;
  const unsigned hi = (unsigned)m[p - 1] | (unsigned)m[p - 1 + P] << 8 | (unsigned)m[p + P] << 16 | (unsigned)m[p + 1 + P] << 24;
  const unsigned a = (((lo >> k) & 0x01010101u) * 0x01020408u) >> 24;
  const unsigned b = (((hi >> k) & 0x01010101u) * 0x01020408u) >> 24;
  return (a & 15u) | (b & 15u) << 4;
}

// Follows borders as cv::findContours does (Suzuki-Abe): from a start pixel, first set neighbour clockwise from west
// (outer) or east (hole), then counter-clockwise from the direction after the one we came from.  A start survives only
// if no raster-earlier start lies on its border, so every border is reported once, from the pixel the sequential scan
// would have started at.  Persistent warps: a lane whose border ends (or whose start loses) takes the next start at
// once.  While following, a lane leaves a checkpoint (pixel, direction) every SEG = max_len / 32 steps (a power of two); when a border of admissible
// length closes, the whole warp writes its points: lane i replays segment i from checkpoint i, so the second pass
// over a border of L points is a chain of at most SEG steps instead of L.
__global__ void __launch_bounds__(128) border_follow_kernel(FollowArgs a) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int PROBE_AFTER = 8, PROBE_STEPS = 16;
  constexpr int CHUNK = 32;           // starts a warp takes from the global pool at a time: small, so that every warp keeps
                                      // refilling its idle lanes until the pool is empty and the tails of all warps coincide
  __shared__ unsigned ckpt[4][32][32];            // [warp][lane][checkpoint] = pixel | direction << 29
  const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, P = a.P;
  const size_t frame = (size_t)(a.H + 2) * P;
  const int max_len = a.max_len;
  int seg_shift = 0;
  while ((32 << seg_shift) < max_len) ++seg_shift;
  const int SEG = 1 << seg_shift;
  unsigned long long w_next = 0, w_end = 0;      // the warp's current chunk of starts (uniform)
  bool exhausted = false;
  // lane state
  bool active = false, finished = false;
  const uint8_t* m = a.mask;
  int p0 = 0, p1 = 0, p3 = 0, kind = 0, k = 0, s = 0, s0 = 0, len = 0;
  int probe = 0, q = 0, c = 0;                   // the backward probe: steps left, pixel, direction we came from
  unsigned nb = 0;
  unsigned long long done_po = 0;
  for (;;) {
    // idle lanes take the next starts of the warp's chunk (all control flow here is warp-uniform)
    for (;;) {
      const unsigned need = __ballot_sync(FULL, !active);
      if (!need || exhausted) break;
      if (w_next >= w_end) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.next, (unsigned long long)CHUNK);
        base = __shfl_sync(FULL, base, 0);
        if (base >= a.n_starts) { exhausted = true; break; }
        w_next = base;
        w_end = base + CHUNK < a.n_starts ? base + CHUNK : a.n_starts;
      }
      const unsigned long long avail = w_end - w_next;
      const int rank = __popc(need & ((1u << lane) - 1u));
      if (!active && (unsigned long long)rank < avail) {
        const unsigned long long st = a.starts[w_next + rank];
        kind = (int)(st & 1);
        k = (int)(st >> 1 & 7);
        const int img = (int)(st >> 36);
        p0 = (int)(st >> 4);                       // 32 bits: pixel inside the padded frame
        m = a.mask + (size_t)img * frame;
        nb = neighbours(m, p0, P, k);
        len = 0;
        p3 = p0;
        if (nb) {
          const int s_end = kind ? 0 : 4;
          unsigned r = 0;                          // bit j = direction (s_end - 1 - j) & 7: the clockwise search
#pragma unroll
          for (int j = 0; j < 8; ++j) r |= ((nb >> ((s_end - 1 - j) & 7)) & 1u) << j;
          s0 = s = (s_end - 1 - (__ffs(r) - 1)) & 7;
          p1 = p0 + dir_off(s, P);
          probe = -1;
          active = true;
        } else if (kind == 0 && a.min_len <= 1 && max_len >= 1) {
          // an isolated pixel is a border of one point (it only matters for images under 34 pixels)
          int slot = atomicAdd(a.n_borders, 1);
          unsigned long long po = atomicAdd(a.n_pts, 1ull);
          if (slot >= a.border_cap || po + 1 > a.pts_cap) atomicExch(a.overflow, 1);
          else {
            a.pts[po] = ((p0 % P) - LPAD) | (((p0 / P) - 1) << 16);
            Border& br = a.borders[slot];
            br.img = img; br.win = k; br.disc = p0; br.len = 1; br.pts_off = (int)po;
            br.valid = 0; br.near_border = 0; br.id = -1; br.rot = 0;
          }
        }
      }
      const unsigned long long asked = (unsigned long long)__popc(need);
      w_next += asked < avail ? asked : avail;
    }
    if (!__any_sync(FULL, active)) break;          // only possible once the starts are exhausted
    if (active) {
      if ((len & (SEG - 1)) == 0) ckpt[wl][lane][len >> seg_shift] = (unsigned)p3 | (unsigned)s << 29;
      // one step: counter-clockwise from the direction after the one we came from: dir, dir + 1, ...
      const int dir = (s + 1) & 7;
      const unsigned rot = ((nb >> dir) | (nb << (8 - dir))) & 0xffu;
      const int t = __ffs(rot) - 1;            // clear directions passed; the pixel we came from ends the search at the latest
      s = (dir + t) & 7;
      const bool west = ((4 - dir) & 7) < t, east = ((8 - dir) & 7) < t;
      // a start with a smaller key (2 * pixel + kind) on this border: this one is not the scan's
      if ((west && p3 < p0 + kind) || (east && p3 < p0) || len >= max_len) {
        active = false;
      } else {
        ++len;
        const int p4 = p3 + dir_off(s, P);
        if (p4 == p0 && p3 == p1) {            // closed
          active = false;
          if (len >= a.min_len) {
            int slot = atomicAdd(a.n_borders, 1);
            unsigned long long po = atomicAdd(a.n_pts, (unsigned long long)len);
            if (slot >= a.border_cap || po + len > a.pts_cap) atomicExch(a.overflow, 1);
            else {
              Border& br = a.borders[slot];
              br.img = (int)((size_t)(m - a.mask) / frame); br.win = k;
              br.disc = p0 + kind;             // a hole is discovered at the clear pixel right of its first pixel
              br.len = len; br.pts_off = (int)po;
              br.valid = 0; br.near_border = 0; br.id = -1; br.rot = 0;
              done_po = po;
              finished = true;
            }
          }
        } else {
          p3 = p4;
          nb = neighbours(m, p3, P, k);
          s = (s + 4) & 7;
        }
      }
      // Most starts that lose do so only after most of the border (a local top on a jagged edge walks away from
      // the rows above it).  The same border followed BACKWARDS from the start climbs towards them at once, so a
      // start that is still alive after PROBE_AFTER steps also walks PROBE_STEPS steps backwards (the mirrored rule:
      // clockwise from the direction before the one we came from visits the same (pixel, clear neighbour) states in
      // reverse order) and gives up as soon as either walk meets a smaller start key.
      if (active) {
        if (probe < 0 && len >= PROBE_AFTER) {
          const unsigned n0 = neighbours(m, p0, P, k);
          const int d0 = (s0 + 1) & 7;
          const unsigned r0 = ((n0 >> d0) | (n0 << (8 - d0))) & 0xffu;
          c = (d0 + __ffs(r0) - 1) & 7;          // the first forward direction: where the backward walk "comes from"
          q = p0;
          probe = PROBE_STEPS;
        }
        if (probe > 0) {
          const unsigned nq = neighbours(m, q, P, k);
          const unsigned rq = ((nq >> c) | (nq << (8 - c))) & 0xffu;     // bit i = direction c + i; bit 0 is set
          const int hb = 31 - __clz(rq);         // clockwise from c - 1 = c + 7 down to the first set direction
          const bool westq = ((4 - c) & 7) > hb, eastq = ((8 - c) & 7) > hb;
          if ((westq && q < p0 + kind) || (eastq && q < p0)) {
            active = false;
          } else {
            const int d = (c + hb) & 7;
            q += dir_off(d, P);
            c = (d + 4) & 7;
            probe = q == p0 ? 0 : probe - 1;      // all the way round: a small border, nothing more to learn
          }
        }
      }
    }
    // borders that closed in this round: the warp writes their points, a segment per lane
    unsigned fin = __ballot_sync(FULL, finished);
    while (fin) {
      const int src = __ffs(fin) - 1;
      fin &= fin - 1;
      const int L = __shfl_sync(FULL, len, src);
      const int kk = __shfl_sync(FULL, k, src);
      const unsigned long long po = __shfl_sync(FULL, done_po, src);
      const uint8_t* mm = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(m), src));
      __syncwarp();
      const int first = lane * SEG;
      if (first < L) {
        const unsigned c = ckpt[wl][src][lane];
        int q = (int)(c & 0x1fffffffu), sd = (int)(c >> 29);
        int* out = a.pts + po + first;
        const int cnt = min(SEG, L - first);
        unsigned nq = neighbours(mm, q, P, kk);
        for (int j = 0; j < cnt; ++j) {
          const int dir = (sd + 1) & 7;
          const unsigned rot = ((nq >> dir) | (nq << (8 - dir))) & 0xffu;
          sd = (dir + __ffs(rot) - 1) & 7;
          out[j] = ((q % P) - LPAD) | (((q / P) - 1) << 16);
          q += dir_off(sd, P);
          nq = neighbours(mm, q, P, kk);
          sd = (sd + 4) & 7;
        }
      }
      if (lane == src) finished = false;
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------- (d3)
struct ApproxArgs {
  Border* borders;
  int n_borders;
  const int* pts;
  int W, H;
  double accuracy_rate, min_corner_rate;
  int min_border;
};

__device__ __forceinline__ void warp_argmax(double& best, int& idx) {     // larger value, then smaller index
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    double ob = __shfl_xor_sync(0xffffffffu, best, d);
    int oi = __shfl_xor_sync(0xffffffffu, idx, d);
    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
  }
}

__global__ void __launch_bounds__(128) approx_quad_kernel(ApproxArgs a) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= a.n_borders) return;
  Border& br = a.borders[w];
  const int n = br.len;
  const int* src = a.pts + br.pts_off;
  auto px = [&](int i) { return src[i] & 0xffff; };
  auto py = [&](int i) { return src[i] >> 16; };
  const double eps = (double)n * a.accuracy_rate, eps2 = eps * eps;

  // 1. approximately the two farthest points (three sweeps)
  int pos = 0, rs = 0;
  bool le_eps = false;
  for (int it = 0; it < 3; ++it) {
    pos = (pos + rs) % n;
    const int sx = px(pos), sy = py(pos);
    double best = 0.0;
    int bj = 0x7fffffff;
    for (int j = 1 + lane; j < n; j += 32) {
      int q = pos + j;
      if (q >= n) q -= n;
      double dx = (double)(px(q) - sx), dy = (double)(py(q) - sy);
      double d = dx * dx + dy * dy;
      if (d > best) { best = d; bj = j; }
    }
    warp_argmax(best, bj);
    if (best > 0.0) rs = bj;
    le_eps = best <= eps2;
  }
  if (le_eps) return;                       // a single point: not a quad
  int st_s[12], st_e[12], top = 0;
  int dst[9], nd = 0;
  {
    int slice_start = pos, right_end = pos;
    int slice_end = (rs + slice_start) % n;
    st_s[top] = slice_end; st_e[top] = right_end; ++top;
    st_s[top] = slice_start; st_e[top] = slice_end; ++top;
  }
  // 2. recursive splitting; more than 8 raw points cannot be cleaned up to 4
  while (top > 0) {
    --top;
    const int s_start = st_s[top], s_end = st_e[top];
    const int ex = px(s_end), ey = py(s_end), sx = px(s_start), sy = py(s_start);
    int m = s_end - s_start - 1;
    if (m < 0) m += n;
    bool le = true;
    int r = 0;
    if (m > 0) {
      const double dx = (double)(ex - sx), dy = (double)(ey - sy), len2 = dx * dx + dy * dy;
      double best = 0.0;
      int bt = 0x7fffffff;
      for (int t = lane; t < m; t += 32) {
        int q = s_start + 1 + t;
        if (q >= n) q -= n;
        const double qx = (double)(px(q) - sx), qy = (double)(py(q) - sy);
        const double proj = qx * dx + qy * dy;
        double d;
        if (proj < 0.0 || len2 == 0.0) d = qx * qx + qy * qy;
        else if (proj > len2) { double fx = (double)(px(q) - ex), fy = (double)(py(q) - ey); d = fx * fx + fy * fy; }
        else { double cr = qy * dx - qx * dy; d = cr * cr / len2; }
        if (d > best) { best = d; bt = t; }
      }
      warp_argmax(best, bt);
      le = best <= eps2;
      if (!le) { r = s_start + 1 + bt; if (r >= n) r -= n; }
    }
    if (le) {
      if (nd >= 9) return;
      dst[nd++] = s_start;
    } else {
      if (nd + top + 2 > 9) return;          // every pending slice yields at least one point
      st_s[top] = r; st_e[top] = s_end; ++top;
      st_s[top] = s_start; st_e[top] = r; ++top;
    }
  }
  if (nd < 4 || nd > 8) return;
  // 3. clean-up of [almost] straight runs, as cv::approxPolyDP's last stage (every lane, same result)
  int qx[9], qy[9];
  for (int i = 0; i < nd; ++i) { qx[i] = px(dst[i]); qy[i] = py(dst[i]); }
  {
    int count = nd, new_count = nd;
    int p = count - 1;
    int sxp = qx[p], syp = qy[p];
    p = (p + 1) % count;
    int wpos = p;
    int ptx = qx[p], pty = qy[p];
    p = (p + 1) % count;
    int i = 0;
    while (i < count && new_count > 2) {
      int exx = qx[p], eyy = qy[p];
      p = (p + 1) % count;
      double dx = (double)(exx - sxp), dy = (double)(eyy - syp);
      double dist = fabs((double)(ptx - sxp) * dy - (double)(pty - syp) * dx);
      long long inner = (long long)(ptx - sxp) * (exx - ptx) + (long long)(pty - syp) * (eyy - pty);
      if (dist * dist <= 0.5 * eps2 * (dx * dx + dy * dy) && dx != 0 && dy != 0 && inner >= 0) {
        --new_count;
        qx[wpos] = sxp = exx; qy[wpos] = syp = eyy;
        wpos = (wpos + 1) % count;
        ptx = qx[p]; pty = qy[p];
        p = (p + 1) % count;
        i += 2;
        continue;
      }
      qx[wpos] = sxp = ptx; qy[wpos] = syp = pty;
      wpos = (wpos + 1) % count;
      ptx = exx; pty = eyy;
      ++i;
    }
    nd = new_count;
  }
  if (nd != 4) return;
  // 4. cv::isContourConvex
  {
    int prx = qx[2], pry = qy[2], cx = qx[3], cy = qy[3];
    int dx0 = cx - prx, dy0 = cy - pry, orientation = 0;
    for (int i = 0; i < 4; ++i) {
      prx = cx; pry = cy; cx = qx[i]; cy = qy[i];
      int dx = cx - prx, dy = cy - pry;
      long long dxdy0 = (long long)dx * dy0, dydx0 = (long long)dy * dx0;
      orientation |= dydx0 > dxdy0 ? 1 : (dydx0 < dxdy0 ? 2 : 3);
      if (orientation == 3) return;
      dx0 = dx; dy0 = dy;
    }
  }
  // 5. smallest side against the contour length, distance to the image border
  {
    const int big = max(a.W, a.H);
    double min_d2 = (double)big * (double)big;
    for (int j = 0; j < 4; ++j) {
      long long dx = qx[j] - qx[(j + 1) & 3], dy = qy[j] - qy[(j + 1) & 3];
      min_d2 = fmin(min_d2, (double)(dx * dx + dy * dy));
    }
    const double mc = (double)n * a.min_corner_rate;
    if (min_d2 < mc * mc) return;
  }
  bool near = false;
  for (int j = 0; j < 4; ++j)
    near |= qx[j] < a.min_border || qy[j] < a.min_border || qx[j] > a.W - 1 - a.min_border || qy[j] > a.H - 1 - a.min_border;
  // 6. clockwise (_reorderCandidatesCorners)
  {
    double dx1 = qx[1] - qx[0], dy1 = qy[1] - qy[0], dx2 = qx[2] - qx[0], dy2 = qy[2] - qy[0];
    if (dx1 * dy2 - dy1 * dx2 < 0.0) { int t = qx[1]; qx[1] = qx[3]; qx[3] = t; t = qy[1]; qy[1] = qy[3]; qy[3] = t; }
  }
  if (lane == 0) {
    for (int j = 0; j < 4; ++j) { br.quad[2 * j] = (float)qx[j]; br.quad[2 * j + 1] = (float)qy[j]; }
    br.near_border = near;
    br.valid = 1;
  }
}

// ---------------------------------------------------------------------------------------------- (d4)
struct IdentifyArgs {
  Border* borders;
  int n_borders;
  const uint8_t* gray;
  int W, H;
  int marker_size, border_bits, cell, margin;
  int max_border_err, max_corr;
  double min_otsu_std;
  const unsigned long long* codes;     // [n_markers][4]: rotation r = the marker turned counter-clockwise r times
  int n_markers;
};

__global__ void __launch_bounds__(128) identify_kernel(IdentifyArgs a) {
  __shared__ uint8_t canon[4][CANON_SIDE * CANON_SIDE];
  __shared__ int hist[4][256];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * 4 + wl;
  if (w >= a.n_borders) return;
  Border& br = a.borders[w];
  if (!br.valid) return;
  const int n = a.marker_size + 2 * a.border_bits, size = n * a.cell;
  const uint8_t* gray = a.gray + (size_t)br.img * a.W * a.H;
  uint8_t* can = canon[wl];
  int* hs = hist[wl];

  // cv::getPerspectiveTransform(quad -> canonical square), then its inverse (what cv::warpPerspective maps with)
  double inv[9];
  {
    // lanes 0..7 hold one row of the 8 x 9 system each, in registers; pivots and rows travel by shuffles
    double row[9];
    {
      const double c = (double)(size - 1);
      const int i = lane & 3;
      const double ddx = (i == 1 || i == 2) ? c : 0.0, ddy = i >= 2 ? c : 0.0;
      const double sx = br.quad[2 * i], sy = br.quad[2 * i + 1];
      const bool second = (lane & 7) >= 4;           // rows 4..7: the y equations
      row[0] = second ? 0.0 : sx; row[1] = second ? 0.0 : sy; row[2] = second ? 0.0 : 1.0;
      row[3] = second ? sx : 0.0; row[4] = second ? sy : 0.0; row[5] = second ? 1.0 : 0.0;
      const double dd = second ? ddy : ddx;
      row[6] = -sx * dd; row[7] = -sy * dd; row[8] = dd;
    }
    bool singular = false;
#pragma unroll
    for (int col = 0; col < 8; ++col) {             // Gaussian elimination with partial pivoting
      // pivot: the row >= col with the largest |a[col]|, first one on ties
      double best = (lane >= col && lane < 8) ? fabs(row[col]) : -1.0;
      int piv = lane;
#pragma unroll
      for (int d = 4; d; d >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, d);
        int op = __shfl_xor_sync(0xffffffffu, piv, d);
        if (ob > best || (ob == best && op < piv)) { best = ob; piv = op; }
      }
      piv = __shfl_sync(0xffffffffu, piv, 0);
      // swap rows col and piv
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const double from_piv = __shfl_sync(0xffffffffu, row[j], piv), from_col = __shfl_sync(0xffffffffu, row[j], col);
        if (piv != col) { if (lane == col) row[j] = from_piv; else if (lane == piv) row[j] = from_col; }
      }
      const double dgl = __shfl_sync(0xffffffffu, row[col], col);
      if (dgl == 0.0) singular = true;
      const double f = row[col] / dgl;
#pragma unroll
      for (int j = col; j < 9; ++j) {
        const double pj = __shfl_sync(0xffffffffu, row[j], col);
        if (lane > col && lane < 8) row[j] -= f * pj;
      }
    }
    if (singular) return;
    double m[9];
    m[8] = 1.0;
#pragma unroll
    for (int r = 7; r >= 0; --r) {
      double sacc = row[8];
#pragma unroll
      for (int j = r + 1; j < 8; ++j) sacc -= row[j] * m[j];
      m[r] = __shfl_sync(0xffffffffu, sacc / row[r], r);
    }
    double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    if (det == 0.0) return;
    double id = 1.0 / det;
    inv[0] = c00 * id; inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    inv[3] = c01 * id; inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    inv[6] = c02 * id; inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  }
  for (int i = lane; i < 256; i += 32) hs[i] = 0;
  __syncwarp();
  // nearest-neighbour warp; statistics of the inner region (half a cell trimmed)
  const int lo = a.cell / 2, hi = size - a.cell / 2;
  long long sum = 0, sumsq = 0;
  for (int i = lane; i < size * size; i += 32) {
    int y = i / size, x = i - y * size;
    double X = inv[0] * x + inv[1] * y + inv[2], Y = inv[3] * x + inv[4] * y + inv[5], Wd = inv[6] * x + inv[7] * y + inv[8];
    double iw = Wd != 0.0 ? 1.0 / Wd : 0.0;
    double fx = fmax(-2147483648.0, fmin(2147483647.0, X * iw)), fy = fmax(-2147483648.0, fmin(2147483647.0, Y * iw));
    long long sx = (long long)rint(fx), sy = (long long)rint(fy);
    int v = (sx >= 0 && sx < a.W && sy >= 0 && sy < a.H) ? gray[(size_t)sy * a.W + sx] : 0;
    can[i] = (uint8_t)v;
    atomicAdd(&hs[v], 1);
    if (x >= lo && x < hi && y >= lo && y < hi) { sum += v; sumsq += v * v; }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
    sumsq += __shfl_xor_sync(0xffffffffu, sumsq, d);
  }
  __syncwarp();
  const double cnt = (double)((hi - lo) * (hi - lo));
  const double mean = (double)sum / cnt;
  const double var = (double)sumsq / cnt - mean * mean;
  const double sd = sqrt(var > 0.0 ? var : 0.0);
  unsigned long long bits = 0;            // n x n cells, row-major, bit (n*n - 1 - cell index)
  if (sd < a.min_otsu_std) {
    bits = mean > 127.0 ? ~0ull : 0ull;
    if (n * n < 64) bits &= (1ull << (n * n)) - 1;
  } else {
    // Otsu's threshold of the whole canonical image (cv::threshold THRESH_OTSU), lane 0
    int thr = 0;
    if (lane == 0) {
      const double scale = 1.0 / (double)(size * size);
      double mu = 0.0;
      for (int i = 0; i < 256; ++i) mu += (double)i * (double)hs[i];
      mu *= scale;
      double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
      for (int i = 0; i < 256; ++i) {
        double p_i = hs[i] * scale;
        mu1 *= q1;
        q1 += p_i;
        double q2 = 1.0 - q1;
        if (fmin(q1, q2) < 1.1920928955078125e-07 || fmax(q1, q2) > 1.0 - 1.1920928955078125e-07) continue;
        mu1 = (mu1 + i * p_i) / q1;
        double mu2 = (mu - q1 * mu1) / q2;
        double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) { max_sigma = sigma; thr = i; }
      }
    }
    thr = __shfl_sync(0xffffffffu, thr, 0);
    const int side = a.cell - 2 * a.margin;
    for (int c0 = 0; c0 < n * n; c0 += 32) {
      int c = c0 + lane;
      bool one = false;
      if (c < n * n) {
        int cy = c / n, cx = c - cy * n, nz = 0;
        for (int yy = 0; yy < side; ++yy)
          for (int xx = 0; xx < side; ++xx)
            nz += can[(cy * a.cell + a.margin + yy) * size + cx * a.cell + a.margin + xx] > thr;
        one = nz > (side * side) / 2;
      }
      unsigned bal = __ballot_sync(0xffffffffu, one);
      for (int l = 0; l < 32 && c0 + l < n * n; ++l)
        if (bal >> l & 1) bits |= 1ull << (n * n - 1 - (c0 + l));
    }
  }
  // border cells must be black
  auto bit_at = [&](int y, int x) { return (int)(bits >> (n * n - 1 - (y * n + x)) & 1); };
  int err = 0;
  for (int k = 0; k < n; ++k)
    for (int b = 0; b < a.border_bits; ++b) err += bit_at(k, b) + bit_at(k, n - 1 - b);
  for (int k = a.border_bits; k < n - a.border_bits; ++k)
    for (int b = 0; b < a.border_bits; ++b) err += bit_at(b, k) + bit_at(n - 1 - b, k);
  if (err > a.max_border_err) return;
  const int ms = a.marker_size;
  unsigned long long code = 0;
  for (int y = 0; y < ms; ++y)
    for (int x = 0; x < ms; ++x) code = code << 1 | (unsigned long long)bit_at(y + a.border_bits, x + a.border_bits);
  // cv::aruco::Dictionary::identify: first marker within the corrected distance, its first best rotation
  int found = 0x7fffffff, frot = 0;
  for (int m = lane; m < a.n_markers && found == 0x7fffffff; m += 32) {
    int best = ms * ms + 1, rot = -1;
    for (int r = 0; r < 4; ++r) {
      int d = __popcll(a.codes[m * 4 + r] ^ code);
      if (d < best) { best = d; rot = r; }
    }
    if (best <= a.max_corr) { found = m; frot = rot; }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    int of = __shfl_xor_sync(0xffffffffu, found, d), orr = __shfl_xor_sync(0xffffffffu, frot, d);
    if (of < found) { found = of; frot = orr; }
  }
  if (lane == 0 && found != 0x7fffffff) { br.id = found; br.rot = frot; }
}

}  // namespace ard

// ================================================================================================ host side
using ard::Border;

struct arslam_detector {
  int device = 0, sm_count = 148;
  int max_images = 0, max_w = 0, max_h = 0;
  cudaStream_t stream = nullptr;
  uint8_t* d_images = nullptr;
  uint8_t* d_gray = nullptr;
  uint8_t* d_mask = nullptr;
  size_t mask_bytes = 0;
  int last_w = -1, last_h = -1;
  int call_n = 0, call_w = 0, call_h = 0, call_borders = 0;   // shape of the last arslam_detect_markers
  unsigned long long* d_starts = nullptr;
  unsigned long long starts_cap = 0;
  Border* d_borders = nullptr;
  int border_cap = 0;
  int* d_pts = nullptr;
  unsigned long long pts_cap = 0;
  unsigned long long* d_counters = nullptr;   // [0] starts, [1] points, then int n_borders, int overflow
  unsigned long long* h_counters = nullptr;   // pinned
  Border* h_borders = nullptr;                // pinned
  unsigned long long* d_codes = nullptr;
  int n_markers = 0, marker_size = 0, max_correction_bits = 0;
  cudaEvent_t ev[7] = {};
  double ms[6] = {};
  long long launches = 0;
  // candidates of the last call, in cv::aruco's order
  std::vector<Border> cand;
  std::string err;
};

static thread_local std::string g_detect_create_error;

static int dfail(arslam_detector* d, int code, const std::string& msg) {
  if (d) d->err = msg; else g_detect_create_error = msg;
  return code;
}
#define DCUDA(d, call)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return dfail(d, ARSLAM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

static const uint16_t kDict4x4_50[50] = {
#include "dict_4x4_50.inc"
};
#include "aruco_dictionaries.inc"

static unsigned long long rotate_code_ccw(unsigned long long code, int ms) {   // np.rot90(bits, 1)
  unsigned long long out = 0;
  for (int y = 0; y < ms; ++y)
    for (int x = 0; x < ms; ++x) {
      // rot90 counter-clockwise: out[y][x] = in[x][ms - 1 - y]
      int sy = x, sx = ms - 1 - y;
      unsigned long long b = code >> (ms * ms - 1 - (sy * ms + sx)) & 1;
      out |= b << (ms * ms - 1 - (y * ms + x));
    }
  return out;
}

static int upload_dictionary(arslam_detector* d, int n_markers, int ms, int max_corr, const std::vector<unsigned long long>& codes0) {
  std::vector<unsigned long long> all((size_t)n_markers * 4);
  for (int m = 0; m < n_markers; ++m) {
    unsigned long long c = codes0[m];
    for (int r = 0; r < 4; ++r) { all[(size_t)m * 4 + r] = c; c = rotate_code_ccw(c, ms); }
  }
  if (d->d_codes) cudaFree(d->d_codes);
  d->d_codes = nullptr;
  DCUDA(d, cudaMalloc(&d->d_codes, all.size() * sizeof(unsigned long long)));
  DCUDA(d, cudaMemcpy(d->d_codes, all.data(), all.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  d->n_markers = n_markers;
  d->marker_size = ms;
  d->max_correction_bits = max_corr;
  return ARSLAM_OK;
}

extern "C" {

void arslam_detect_default_params(arslam_detect_params* p) {
  if (!p) return;
  p->adaptive_thresh_win_size_min = 3;
  p->adaptive_thresh_win_size_max = 23;
  p->adaptive_thresh_win_size_step = 10;
  p->min_distance_to_border = 3;
  p->marker_border_bits = 1;
  p->perspective_remove_pixel_per_cell = 4;
  p->adaptive_thresh_constant = 7.0;
  p->min_marker_perimeter_rate = 0.03;
  p->max_marker_perimeter_rate = 4.0;
  p->polygonal_approx_accuracy_rate = 0.03;
  p->min_corner_distance_rate = 0.05;
  p->min_marker_distance_rate = 0.125;
  p->min_group_distance = (double)0.21f;
  p->perspective_remove_ignored_margin_per_cell = 0.13;
  p->max_erroneous_bits_in_border_rate = 0.35;
  p->min_otsu_std_dev = 5.0;
  p->error_correction_rate = 0.6;
}

const char* arslam_detector_last_error(const arslam_detector* d) { return d ? d->err.c_str() : g_detect_create_error.c_str(); }

void arslam_detector_destroy(arslam_detector* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  if (d->stream) cudaStreamSynchronize(d->stream);
  cudaFree(d->d_images); cudaFree(d->d_gray); cudaFree(d->d_mask); cudaFree(d->d_starts); cudaFree(d->d_borders);
  cudaFree(d->d_pts); cudaFree(d->d_counters); cudaFree(d->d_codes);
  if (d->h_counters) cudaFreeHost(d->h_counters);
  if (d->h_borders) cudaFreeHost(d->h_borders);
  for (auto& e : d->ev) if (e) cudaEventDestroy(e);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

int arslam_detector_create(int device, int32_t max_images, int32_t max_width, int32_t max_height, arslam_detector** out) {
  if (!out) return dfail(nullptr, ARSLAM_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (max_images < 1 || max_width < 8 || max_height < 8 || max_width > 16384 || max_height > 16384 ||
      (double)max_images * (max_width + 18.0) * (max_height + 2.0) > 1.0e9)
    return dfail(nullptr, ARSLAM_ERR_INVALID, "arslam_detector_create: bad sizes (frames up to 16384 x 16384, about 1e9 pixels per batch)");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev)
    return dfail(nullptr, ARSLAM_ERR_NO_DEVICE, "no CUDA device (the detector has no CPU path)");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return dfail(nullptr, ARSLAM_ERR_NO_DEVICE, "device is not sm_100 (the library is built for sm_100a only)");
  arslam_detector* d = new arslam_detector;
  d->device = device;
  d->sm_count = prop.multiProcessorCount;
  d->max_images = max_images; d->max_w = max_width; d->max_h = max_height;
  auto bail = [&](int rc) { std::string m = d->err; arslam_detector_destroy(d); g_detect_create_error = m; return rc; };
#define DC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { d->err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(ARSLAM_ERR_CUDA); } } while (0)
  DC(cudaSetDevice(device));
  DC(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  const size_t px = (size_t)max_images * max_width * max_height;
  const int P = (max_width + 2 * ard::LPAD + 15) / 16 * 16;
  d->mask_bytes = (size_t)max_images * (max_height + 2) * P;
  d->starts_cap = std::max<unsigned long long>(px / 2, 1 << 16);          // grown on demand, like the two below
  d->border_cap = (int)std::min<size_t>((size_t)max_images * 4096, (size_t)1 << 24);
  d->pts_cap = std::max<unsigned long long>(px / 2, 1ull << 18);
  DC(cudaMalloc(&d->d_images, px * 3));
  DC(cudaMalloc(&d->d_gray, px));
  DC(cudaMalloc(&d->d_mask, d->mask_bytes));
  DC(cudaMalloc(&d->d_starts, d->starts_cap * sizeof(unsigned long long)));
  DC(cudaMalloc(&d->d_borders, (size_t)d->border_cap * sizeof(Border)));
  DC(cudaMalloc(&d->d_pts, d->pts_cap * sizeof(int)));
  DC(cudaMalloc(&d->d_counters, 4 * sizeof(unsigned long long)));
  DC(cudaMallocHost(&d->h_counters, 4 * sizeof(unsigned long long)));
  DC(cudaMallocHost(&d->h_borders, (size_t)d->border_cap * sizeof(Border)));
  for (auto& e : d->ev) DC(cudaEventCreate(&e));
#undef DC
  std::vector<unsigned long long> codes(50);
  for (int i = 0; i < 50; ++i) codes[i] = kDict4x4_50[i];
  int rc = upload_dictionary(d, 50, 4, 1, codes);
  if (rc != ARSLAM_OK) return bail(rc);
  *out = d;
  return ARSLAM_OK;
}

int arslam_detector_set_dictionary(arslam_detector* d, int32_t n_markers, int32_t marker_size, int32_t max_correction_bits,
                                   const uint8_t* bits) {
  if (!d) return ARSLAM_ERR_INVALID;
  if (!bits || n_markers < 1 || marker_size < 1 || marker_size > ard::MAX_MARKER || max_correction_bits < 0)
    return dfail(d, ARSLAM_ERR_INVALID, "arslam_detector_set_dictionary: marker_size must be 1..6, n_markers >= 1");
  DCUDA(d, cudaSetDevice(d->device));
  DCUDA(d, cudaStreamSynchronize(d->stream));
  std::vector<unsigned long long> codes(n_markers);
  const int nb = marker_size * marker_size;
  for (int m = 0; m < n_markers; ++m) {
    unsigned long long c = 0;
    for (int i = 0; i < nb; ++i) c = c << 1 | (bits[(size_t)m * nb + i] ? 1ull : 0ull);
    codes[m] = c;
  }
  return upload_dictionary(d, n_markers, marker_size, max_correction_bits, codes);
}

/* The three names the reference's detector node accepts (aruco_detector.cpp:148-152). */
int arslam_detector_set_predefined_dictionary(arslam_detector* d, const char* name) {
  if (!d) return ARSLAM_ERR_INVALID;
  const std::string n = name ? name : "";
  std::vector<unsigned long long> codes;
  int ms = 0, mc = 0;
  if (n == "4X4_50") { codes.assign(kDict4x4_50, kDict4x4_50 + 50); ms = 4; mc = 1; }
  else if (n == "5X5_100") { codes.assign(kDICT_5X5_100, kDICT_5X5_100 + 100); ms = 5; mc = kDICT_5X5_100_max_correction; }
  else if (n == "6X6_250") { codes.assign(kDICT_6X6_250, kDICT_6X6_250 + 250); ms = 6; mc = kDICT_6X6_250_max_correction; }
  else return dfail(d, ARSLAM_ERR_INVALID, "invalid aruco_dict " + n + " (4X4_50, 5X5_100, 6X6_250)");
  DCUDA(d, cudaSetDevice(d->device));
  DCUDA(d, cudaStreamSynchronize(d->stream));
  return upload_dictionary(d, (int)codes.size(), ms, mc, codes);
}

static float side_sum(const float* q) {          // float32 like OpenCV's MarkerCandidateTree perimeter
  float s = 0.f;
  for (int i = 0; i < 4; ++i) {
    float dx = q[2 * i] - q[2 * ((i + 1) & 3)], dy = q[2 * i + 1] - q[2 * ((i + 1) & 3) + 1];
    s += std::sqrt(dx * dx + dy * dy);
  }
  return s;
}
static float average_distance(const float* a, const float* b) {     // getAverageDistance
  float best = 3.402823466e+38f;
  for (int first = 0; first < 4; ++first) {
    float dist = 0.f;
    for (int c = 0; c < 4; ++c) {
      int m = (first + c) & 3;
      float dx = a[2 * c] - b[2 * m], dy = a[2 * c + 1] - b[2 * m + 1];
      dist += std::sqrt(dx * dx + dy * dy);
    }
    dist /= 4.f;
    best = std::min(best, dist);
  }
  return best;
}

int arslam_detect_markers(arslam_detector* d, const uint8_t* images, int32_t n_images, int32_t W, int32_t H, int32_t channels,
                          int32_t on_device, const arslam_detect_params* params, int32_t max_markers, int32_t* n_found,
                          int32_t* ids, float* corners) {
  if (!d) return ARSLAM_ERR_INVALID;
  arslam_detect_params p;
  if (params) p = *params; else arslam_detect_default_params(&p);
  if (!images || !n_found || !ids || !corners || max_markers < 1)
    return dfail(d, ARSLAM_ERR_INVALID, "arslam_detect_markers: NULL argument or max_markers < 1");
  if (n_images < 1 || n_images > d->max_images || W < 8 || H < 8 || W > d->max_w || H > d->max_h ||
      (size_t)n_images * (H + 2) * ((W + 2 * ard::LPAD + 15) / 16 * 16) > d->mask_bytes)
    return dfail(d, ARSLAM_ERR_INVALID, "arslam_detect_markers: batch or frame size exceeds what arslam_detector_create reserved");
  if (channels != 1 && channels != 3) return dfail(d, ARSLAM_ERR_INVALID, "arslam_detect_markers: channels must be 1 (grey) or 3 (BGR)");
  ard::Windows wins;
  std::memset(&wins, 0, sizeof wins);
  if (p.adaptive_thresh_win_size_min < 3 || p.adaptive_thresh_win_size_max < p.adaptive_thresh_win_size_min ||
      p.adaptive_thresh_win_size_step < 1)
    return dfail(d, ARSLAM_ERR_INVALID, "arslam_detect_markers: adaptive threshold window range");
  wins.n = (p.adaptive_thresh_win_size_max - p.adaptive_thresh_win_size_min) / p.adaptive_thresh_win_size_step + 1;
  if (wins.n > ard::MAX_WIN) return dfail(d, ARSLAM_ERR_INVALID, "arslam_detect_markers: more than 8 adaptive threshold windows");
  for (int k = 0; k < wins.n; ++k) {
    int w = p.adaptive_thresh_win_size_min + k * p.adaptive_thresh_win_size_step;
    if (w % 2 == 0 || w / 2 > ard::MAX_RADIUS)
      return dfail(d, ARSLAM_ERR_INVALID, "arslam_detect_markers: adaptive threshold windows must be odd and <= 31");
    wins.radius[k] = w / 2;
    wins.rmax = std::max(wins.rmax, w / 2);
  }
  wins.idelta = (int)std::floor(p.adaptive_thresh_constant);
  const int ms = d->marker_size, nb = ms + 2 * p.marker_border_bits, cell = p.perspective_remove_pixel_per_cell;
  if (p.marker_border_bits < 1 || cell < 1 || nb * cell > ard::CANON_SIDE || nb * nb > 64)
    return dfail(d, ARSLAM_ERR_INVALID, "arslam_detect_markers: canonical marker image larger than 48 pixels or 64 cells");
  const int margin = (int)(p.perspective_remove_ignored_margin_per_cell * cell);
  if (cell - 2 * margin < 1) return dfail(d, ARSLAM_ERR_INVALID, "arslam_detect_markers: ignored margin leaves no pixel per cell");

  DCUDA(d, cudaSetDevice(d->device));
  cudaStream_t st = d->stream;
  const int P = (W + 2 * ard::LPAD + 15) / 16 * 16;
  const size_t px = (size_t)n_images * W * H;
  d->launches = 0;
  DCUDA(d, cudaEventRecord(d->ev[0], st));
  const uint8_t* d_in = images;
  if (!on_device) {
    DCUDA(d, cudaMemcpyAsync(d->d_images, images, px * channels, cudaMemcpyHostToDevice, st));
    d_in = d->d_images;
  }
  if (W != d->last_w || H != d->last_h) {            // the zero frame around every image
    DCUDA(d, cudaMemsetAsync(d->d_mask, 0, d->mask_bytes, st));
    d->last_w = W; d->last_h = H;
  }
  DCUDA(d, cudaMemsetAsync(d->d_counters, 0, 4 * sizeof(unsigned long long), st));
  DCUDA(d, cudaEventRecord(d->ev[1], st));
  {
    dim3 grid((W + ard::TX - 1) / ard::TX, (H + ard::TY - 1) / ard::TY, n_images);
    const int R = wins.rmax <= 11 ? 11 : 15;
    const int tw = ard::TX + 2 * R, th = ard::TY + 2 * R;
    size_t smem = (size_t)(th + 1) * (tw + 1) * sizeof(uint32_t) + (size_t)th * ((tw + 3) / 4 * 4);
    if (channels == 3 && R == 11) ard::gray_threshold_kernel<3, 11><<<grid, 256, smem, st>>>(d_in, W, H, d->d_gray, d->d_mask, P, wins);
    else if (channels == 3) ard::gray_threshold_kernel<3, 15><<<grid, 256, smem, st>>>(d_in, W, H, d->d_gray, d->d_mask, P, wins);
    else if (R == 11) ard::gray_threshold_kernel<1, 11><<<grid, 256, smem, st>>>(d_in, W, H, d->d_gray, d->d_mask, P, wins);
    else ard::gray_threshold_kernel<1, 15><<<grid, 256, smem, st>>>(d_in, W, H, d->d_gray, d->d_mask, P, wins);
    ++d->launches;
  }
  DCUDA(d, cudaEventRecord(d->ev[2], st));
  const int big = std::max(W, H);
  const int min_len = (int)(unsigned)(p.min_marker_perimeter_rate * big);
  const int max_len = (int)(unsigned)(p.max_marker_perimeter_rate * big);
  unsigned long long n_starts = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {      // a frame of pure noise has more starts than reserved: grow once
    unsigned blocks = (unsigned)(((size_t)n_images * H * ((W + 7) / 8) + 255) / 256);
    ard::border_starts_kernel<<<blocks, 256, 0, st>>>(d->d_mask, W, H, P, n_images, min_len > 1 ? 1 : 0, d->d_starts, d->starts_cap, d->d_counters);
    ++d->launches;
    if (attempt == 0) DCUDA(d, cudaEventRecord(d->ev[3], st));
    DCUDA(d, cudaMemcpyAsync(d->h_counters, d->d_counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    DCUDA(d, cudaStreamSynchronize(st));
    n_starts = d->h_counters[0];
    if (n_starts <= d->starts_cap) break;
    if (attempt == 1) return dfail(d, ARSLAM_ERR_CUDA, "arslam_detect_markers: border starts overflow after growing the workspace");
    DCUDA(d, cudaFree(d->d_starts));
    d->d_starts = nullptr;
    d->starts_cap = n_starts + n_starts / 4;
    DCUDA(d, cudaMalloc(&d->d_starts, d->starts_cap * sizeof(unsigned long long)));
    DCUDA(d, cudaMemsetAsync(d->d_counters, 0, 4 * sizeof(unsigned long long), st));
  }
  int* d_nb = reinterpret_cast<int*>(d->d_counters + 2);
  const int* h_nb = reinterpret_cast<const int*>(d->h_counters + 2);
  int n_borders = 0;
  d->call_n = n_images; d->call_w = W; d->call_h = H; d->call_borders = 0;
  d->h_counters[1] = 0;
  for (int attempt = 0; attempt < 2 && n_starts; ++attempt) {
    ard::FollowArgs fa;
    fa.mask = d->d_mask; fa.W = W; fa.H = H; fa.P = P; fa.n_img = n_images;
    fa.starts = d->d_starts; fa.n_starts = n_starts;
    fa.min_len = std::max(min_len, 1); fa.max_len = max_len;
    fa.borders = d->d_borders; fa.border_cap = d->border_cap; fa.n_borders = d_nb;
    fa.pts = d->d_pts; fa.pts_cap = d->pts_cap; fa.n_pts = d->d_counters + 1; fa.overflow = d_nb + 1;
    fa.next = d->d_counters + 3;
    const unsigned want = (unsigned)((n_starts + 127) / 128);      // a lane per start at most
    ard::border_follow_kernel<<<std::min(want, (unsigned)d->sm_count * 12u), 128, 0, st>>>(fa);
    ++d->launches;
    if (attempt == 0) DCUDA(d, cudaEventRecord(d->ev[4], st));
    DCUDA(d, cudaMemcpyAsync(d->h_counters, d->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    DCUDA(d, cudaStreamSynchronize(st));
    n_borders = h_nb[0];
    if (!h_nb[1]) break;
    if (attempt == 1) return dfail(d, ARSLAM_ERR_CUDA, "arslam_detect_markers: borders overflow after growing the workspace");
    // more (or longer) borders than reserved: grow to what this batch asked for and follow again
    if (n_borders > d->border_cap) {
      DCUDA(d, cudaFree(d->d_borders)); d->d_borders = nullptr;
      DCUDA(d, cudaFreeHost(d->h_borders)); d->h_borders = nullptr;
      d->border_cap = n_borders + n_borders / 4;
      DCUDA(d, cudaMalloc(&d->d_borders, (size_t)d->border_cap * sizeof(Border)));
      DCUDA(d, cudaMallocHost(&d->h_borders, (size_t)d->border_cap * sizeof(Border)));
    }
    if (d->h_counters[1] > d->pts_cap) {
      DCUDA(d, cudaFree(d->d_pts)); d->d_pts = nullptr;
      d->pts_cap = d->h_counters[1] + d->h_counters[1] / 4;
      DCUDA(d, cudaMalloc(&d->d_pts, d->pts_cap * sizeof(int)));
    }
    DCUDA(d, cudaMemsetAsync(d->d_counters + 1, 0, 3 * sizeof(unsigned long long), st));
  }
  if (!n_starts) DCUDA(d, cudaEventRecord(d->ev[4], st));
  d->call_borders = n_borders;
  if (n_borders) {
    ard::ApproxArgs aa;
    aa.borders = d->d_borders; aa.n_borders = n_borders; aa.pts = d->d_pts; aa.W = W; aa.H = H;
    aa.accuracy_rate = p.polygonal_approx_accuracy_rate; aa.min_corner_rate = p.min_corner_distance_rate;
    aa.min_border = p.min_distance_to_border;
    ard::approx_quad_kernel<<<(n_borders + 3) / 4, 128, 0, st>>>(aa);
    ++d->launches;
  }
  DCUDA(d, cudaEventRecord(d->ev[5], st));
  if (n_borders) {
    ard::IdentifyArgs ia;
    ia.borders = d->d_borders; ia.n_borders = n_borders; ia.gray = d->d_gray; ia.W = W; ia.H = H;
    ia.marker_size = ms; ia.border_bits = p.marker_border_bits; ia.cell = cell; ia.margin = margin;
    ia.max_border_err = (int)(ms * ms * p.max_erroneous_bits_in_border_rate);
    ia.max_corr = (int)(d->max_correction_bits * p.error_correction_rate);
    ia.min_otsu_std = p.min_otsu_std_dev;
    ia.codes = d->d_codes; ia.n_markers = d->n_markers;
    ard::identify_kernel<<<(n_borders + 3) / 4, 128, 0, st>>>(ia);
    ++d->launches;
    DCUDA(d, cudaMemcpyAsync(d->h_borders, d->d_borders, (size_t)n_borders * sizeof(Border), cudaMemcpyDeviceToHost, st));
  }
  DCUDA(d, cudaEventRecord(d->ev[6], st));
  DCUDA(d, cudaStreamSynchronize(st));
  DCUDA(d, cudaGetLastError());
  for (int i = 0; i < 5; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, d->ev[i + 1], d->ev[i + 2]);
    d->ms[i] = t;
  }
  { float t = 0.f; cudaEventElapsedTime(&t, d->ev[0], d->ev[6]); d->ms[5] = t; }

  // ---- host: candidates in cv::aruco's order, grouping (_filterTooCloseCandidates), output
  d->cand.clear();
  for (int i = 0; i < n_borders; ++i) if (d->h_borders[i].valid) d->cand.push_back(d->h_borders[i]);
  std::sort(d->cand.begin(), d->cand.end(), [](const Border& a, const Border& b) {
    if (a.img != b.img) return a.img < b.img;
    if (a.win != b.win) return a.win < b.win;
    return a.disc > b.disc;                 // cv::findContours lists the border found last first
  });
  const float rate = (float)p.min_marker_distance_rate, group_dist = (float)p.min_group_distance;
  size_t lo = 0;
  for (int b = 0; b < n_images; ++b) {
    size_t hi = lo;
    while (hi < d->cand.size() && d->cand[hi].img == b) ++hi;
    std::vector<const Border*> c;
    for (size_t i = lo; i < hi; ++i) c.push_back(&d->cand[i]);
    lo = hi;
    std::vector<float> per(c.size());
    std::vector<size_t> order(c.size());
    for (size_t i = 0; i < c.size(); ++i) order[i] = i;
    std::vector<float> per0(c.size());
    for (size_t i = 0; i < c.size(); ++i) per0[i] = side_sum(c[i]->quad);
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return per0[x] > per0[y]; });
    std::vector<const Border*> s(c.size());
    for (size_t i = 0; i < c.size(); ++i) { s[i] = c[order[i]]; per[i] = per0[order[i]]; }
    const size_t n = s.size();
    std::vector<int> group(n, -1);
    std::vector<std::vector<size_t>> groups;
    std::vector<char> selected(n, 1);
    for (size_t i = 0; i < n; ++i)
      for (size_t j = i + 1; j < n; ++j)
        if (average_distance(s[i]->quad, s[j]->quad) < per[j] * rate) {
          selected[i] = selected[j] = 0;
          if (group[i] < 0 && group[j] < 0) { group[i] = group[j] = (int)groups.size(); groups.push_back({i, j}); }
          else if (group[i] > -1 && group[j] == -1) { group[j] = group[i]; groups[group[i]].push_back(j); }
          else if (group[j] > -1 && group[i] == -1) { group[i] = group[j]; groups[group[j]].push_back(i); }
        }
    std::vector<std::vector<size_t>> close(n);
    for (auto& g : groups) {
      std::sort(g.begin(), g.end());
      size_t cur = g[0];
      selected[cur] = 1;
      for (size_t t = 1; t < g.size(); ++t) {
        size_t k = g[t];
        float dist = average_distance(s[k]->quad, s[cur]->quad);
        float module = side_sum(s[k]->quad) / (float)(4 * nb);
        if (dist > group_dist * module) { cur = k; if (!s[k]->near_border) close[g[0]].push_back(k); }
      }
    }
    int found = 0;
    for (size_t i = 0; i < n; ++i) {
      if (!selected[i] || s[i]->near_border) continue;
      const Border* hit = s[i]->id >= 0 ? s[i] : nullptr;
      if (!hit)
        for (size_t k : close[i]) if (s[k]->id >= 0) { hit = s[k]; break; }
      if (!hit) continue;
      if (found >= max_markers) return dfail(d, ARSLAM_ERR_INVALID, "arslam_detect_markers: more than max_markers markers in a frame");
      ids[(size_t)b * max_markers + found] = hit->id;
      float* o = corners + ((size_t)b * max_markers + found) * 8;
      for (int j = 0; j < 4; ++j) {          // std::rotate(begin, begin + 4 - rot, end)
        int src = (j + 4 - hit->rot) & 3;
        o[2 * j] = hit->quad[2 * src];
        o[2 * j + 1] = hit->quad[2 * src + 1];
      }
      ++found;
    }
    n_found[b] = found;
  }
  return ARSLAM_OK;
}

int arslam_detector_candidates(arslam_detector* d, int32_t cap, int32_t* image, int32_t* window, float* corners,
                               int32_t* near_border, int32_t* id, int32_t* rotation) {
  if (!d) return ARSLAM_ERR_INVALID;
  int n = (int)d->cand.size();
  for (int i = 0; i < n && i < cap; ++i) {
    const Border& b = d->cand[i];
    if (image) image[i] = b.img;
    if (window) window[i] = b.win;
    if (corners) std::memcpy(corners + (size_t)i * 8, b.quad, 8 * sizeof(float));
    if (near_border) near_border[i] = b.near_border;
    if (id) id[i] = b.id;
    if (rotation) rotation[i] = b.rot;
  }
  return n;
}

int64_t arslam_detector_read_stage(arslam_detector* d, int32_t what, void* out, int64_t cap_bytes) {
  if (!d) return ARSLAM_ERR_INVALID;
  if (cudaSetDevice(d->device) != cudaSuccess) return dfail(d, ARSLAM_ERR_CUDA, "cudaSetDevice");
  const int n = d->call_n, W = d->call_w, H = d->call_h, P = (W + 2 * ard::LPAD + 15) / 16 * 16;
  if (n == 0) return dfail(d, ARSLAM_ERR_INVALID, "arslam_detector_read_stage: no arslam_detect_markers call yet");
  int64_t need = 0;
  if (what == 0 || what == 1) need = (int64_t)n * W * H;
  else if (what == 2) need = (int64_t)d->call_borders * 5 * sizeof(int32_t);
  else if (what == 3) need = (int64_t)d->h_counters[1] * sizeof(int32_t);
  else return dfail(d, ARSLAM_ERR_INVALID, "arslam_detector_read_stage: what must be 0..3");
  if (!out) return need;
  if (cap_bytes < need) return dfail(d, ARSLAM_ERR_INVALID, "arslam_detector_read_stage: buffer too small");
  if (what == 0) {
    DCUDA(d, cudaMemcpy(out, d->d_gray, need, cudaMemcpyDeviceToHost));
  } else if (what == 1) {
    for (int b = 0; b < n; ++b)
      DCUDA(d, cudaMemcpy2D((uint8_t*)out + (size_t)b * W * H, W, d->d_mask + (size_t)b * (H + 2) * P + P + ard::LPAD, P, W, H,
                            cudaMemcpyDeviceToHost));
  } else if (what == 2) {
    std::vector<Border> tmp(d->call_borders);
    if (d->call_borders) DCUDA(d, cudaMemcpy(tmp.data(), d->d_borders, tmp.size() * sizeof(Border), cudaMemcpyDeviceToHost));
    int32_t* o = (int32_t*)out;
    for (size_t i = 0; i < tmp.size(); ++i) {
      o[5 * i] = tmp[i].img; o[5 * i + 1] = tmp[i].win; o[5 * i + 2] = tmp[i].disc; o[5 * i + 3] = tmp[i].len; o[5 * i + 4] = tmp[i].pts_off;
    }
  } else {
    if (need) DCUDA(d, cudaMemcpy(out, d->d_pts, need, cudaMemcpyDeviceToHost));
  }
  return need;
}

int arslam_detector_times(arslam_detector* d, double* ms6, int64_t* launches) {
  if (!d) return ARSLAM_ERR_INVALID;
  if (ms6) for (int i = 0; i < 6; ++i) ms6[i] = d->ms[i];
  if (launches) *launches = d->launches;
  return ARSLAM_OK;
}

}  // extern "C"
