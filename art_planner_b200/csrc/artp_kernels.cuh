// art_planner_b200/csrc/artp_kernels.cuh
// Pose-validity kernels (StateValidityChecker::isValid, validity_checker.cpp:39-45, on top of the ODE
// box-vs-heightfield decision, heightfield.cpp:973-1964) for sm_100a, as a three-stage pipeline:
//
//   A  classify_items_kernel     one THREAD per work item (pose / interpolated edge state): quaternion -> R,
//      dxOrthogonalizeR, the five box centres / map-inside tests / AABBs / zones, the zone min/max/all-finite from
//      exact range tables (no zone scan), and the collider's four early-outs. Items decided here (the majority)
//      never touch the heightfield; every box that needs the vertex / plane tests becomes a BoxRec in a queue.
//   B  box_tiles_warp_kernel (artp_tiles.cuh)   one WARP per queued box, the box's zone staged in shared memory by TMA:
//      vertex-in-box and plane tests by warp ballots. The reference's O(T^2) plane grouping is replaced by an
//      exact shortcut: only triangles under one of the 8 box corners can own a plane-contact point, so only those
//      "candidate" planes are built; a bloom filter on the (approximate) normal finds any earlier triangle that
//      could epsilon-merge with a live candidate. None => every candidate is its own group base (exact);
//      otherwise the box is deferred to C.
//   C  box_items_block_kernel    one CTA per deferred box: the full decision including the reference's greedy,
//      order-dependent epsilon grouping, with all planes staged in shared memory.
// The pose result is a pure AND over its boxes (torso free, every reach box touching), so B and C only ever
// clear the provisional 1 that A wrote. All three are bit-exact against oracle/ (tests/test_pose_gpu.py).
#pragma once

#include "artp_device.cuh"

namespace artp {

constexpr int kWarpsPerCta = 8;
constexpr int kMaxCand = 32;          // live candidates kept per box (more -> the box goes to the grouping stage)
constexpr int kBloomWords = 128;      // 4096-bit filter per warp
constexpr unsigned kFull = 0xffffffffu;

enum { R_FREE = 0, R_HIT = 1, R_DEFER = 2 };

struct WarpScratch {
  float cpl[kMaxCand][4];   // candidate planes (exact)
  int cidx[kMaxCand];       // emission index of a LIVE candidate, -1 otherwise
  uint32_t bloom[kBloomWords];    // level 1: approximate (n0, n2)
  uint32_t bloom2[kBloomWords];   // level 2: exact (n0, n2, d)
  uint32_t cand_bits[kBloomWords];   // bit i: triangle i of the zone is a live candidate
};

// Work description shared by K1/K2. EDGE mode: item w -> edge e = w / (steps+1), j = w % (steps+1);
// j == 0 checks s2, j >= 1 checks interp(s1, s2, j/(steps+1)) (OMPL SE3 interpolation).
struct Work {
  const double* s1;     // EDGE: start states; POSE: unused
  const double* s2;     // EDGE: end states;   POSE: the states
  const float* s2f;     // POSE only: states already cast to float (exactly what Pose3FromSE3 does first); else null
  uint8_t* valid;       // per pose / per edge
  uint32_t item_base;   // first work item of this launch (chunked calls)
  uint32_t n_items;     // one past the last work item of this launch
  int steps;            // EDGE: interior steps; POSE: 0
  int edge_mode;
  // INTERIOR mode (edge_mode == 0, item_off != null): item i is interior state j = i - item_off[e] + 1 of edge e at
  // t = j * (1.0 / (n_e + 1)), n_e = item_off[e+1] - item_off[e]  (prm_motion_cost.cpp:345-353); valid[] is per item.
  const uint32_t* item_off = nullptr;   // n_edges + 1 exclusive prefix sums of the per-edge interior-state counts
  uint32_t n_edges = 0;
  // SEGMENT mode (item_off != null, quotient = 1): OMPL DiscreteMotionValidator over nd_e = item_off[e+1] - item_off[e]
  // segments: item k of edge e is interpolate(s1, s2, (k + 1) / nd_e) for k < nd_e - 1 and s2 itself for k = nd_e - 1.
  int quotient = 0;
};

__device__ __forceinline__ uint32_t bloom_hash(int kx, int kz) {
  return (((uint32_t)kx * 0x9E3779B1u) ^ ((uint32_t)kz * 0x85EBCA77u)) >> 20;   // 12 bits
}
// ceil(2^32 / n) for the flattened-index division t / n = umulhi(t, magic) (exact while t * n < 2^32).
struct MagicTab {
  uint32_t v[129];
  constexpr MagicTab() : v() {
    for (int i = 2; i < 129; ++i) v[i] = 0xFFFFFFFFu / (uint32_t)i + 1u;
  }
};
__constant__ MagicTab kMagicTab = MagicTab();
__device__ __forceinline__ uint32_t magic_for(int n) {
  return n <= 1 ? 0u : (n <= 128 ? kMagicTab.v[n] : 0xFFFFFFFFu / (uint32_t)n + 1u);
}

constexpr float kKeyScale = 16384.0f;   // bucket width 2^-14 on n0, n2 in [-1, 1]
constexpr float kKeyMargin = 4e-6f;     // > eps + rsqrt.approx error + quantisation error

// OMPL 1.4.2 SE3StateSpace::interpolate (RealVector lerp + SO3 slerp), double.
__device__ __forceinline__ void se3_interpolate(const double* a, const double* b, double t, double* out) {
  for (int i = 0; i < 3; ++i) out[i] = a[i] + (b[i] - a[i]) * t;
  const double dq = a[3] * b[3] + a[4] * b[4] + a[5] * b[5] + a[6] * b[6];
  const double dqa = fabs(dq);
  const double theta = (dqa > 1.0 - 1e-9) ? 0.0 : acos(dqa);
  if (theta > 2.220446049250313e-16) {
    const double d = 1.0 / sin(theta);
    const double s0 = sin((1.0 - t) * theta);
    double s1 = sin(t * theta);
    if (dq < 0) s1 = -s1;
    out[3] = (a[3] * s0 + b[3] * s1) * d;
    out[4] = (a[4] * s0 + b[4] * s1) * d;
    out[5] = (a[5] * s0 + b[5] * s1) * d;
    out[6] = (a[6] * s0 + b[6] * s1) * d;
  } else {
    out[3] = a[3]; out[4] = a[4]; out[5] = a[5]; out[6] = a[6];
  }
}

__device__ __forceinline__ void load_item_state(const Work& w, uint32_t item, double s[7]) {
  if (w.item_off) {
    uint32_t lo = 0, hi = w.n_edges;          // largest e with item_off[e] <= item
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(w.item_off + mid) <= item) lo = mid; else hi = mid;
    }
    const uint32_t o0 = __ldg(w.item_off + lo), o1 = __ldg(w.item_off + lo + 1);
    const int n_e = (int)(o1 - o0), step = (int)(item - o0) + 1;
    double a[7], b[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { a[k] = w.s1[(size_t)lo * 7 + k]; b[k] = w.s2[(size_t)lo * 7 + k]; }
    if (w.quotient) {
      if (step == n_e) {
#pragma unroll
        for (int k = 0; k < 7; ++k) s[k] = b[k];
      } else {
        se3_interpolate(a, b, (double)step / (double)n_e, s);
      }
      return;
    }
    const double n_interp_div = 1.0 / (double)(n_e + 1);
    se3_interpolate(a, b, (double)step * n_interp_div, s);
    return;
  }
  if (!w.edge_mode) {
    if (w.s2f) {
#pragma unroll
      for (int k = 0; k < 7; ++k) s[k] = (double)w.s2f[(size_t)item * 7 + k];   // exact; cast back to float downstream
    } else {
#pragma unroll
      for (int k = 0; k < 7; ++k) s[k] = w.s2[(size_t)item * 7 + k];
    }
    return;
  }
  const uint32_t per = (uint32_t)w.steps + 1u;
  const uint32_t e = item / per, j = item - e * per;
  double b[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) b[k] = w.s2[(size_t)e * 7 + k];
  if (j == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) s[k] = b[k];
  } else {
    double a[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) a[k] = w.s1[(size_t)e * 7 + k];
    se3_interpolate(a, b, (double)j / (double)per, s);
  }
}

__device__ __forceinline__ uint32_t item_slot(const Work& w, uint32_t item) {
  return w.edge_mode ? item / ((uint32_t)w.steps + 1u) : item;
}

// Heights of the four corners of cell (cx, cz): A(x,z) B(x+1,z) C(x,z+1) D(x+1,z+1).
__device__ __forceinline__ void load_cell(const Field& f, int cx, int cz, float& hA, float& hB, float& hC, float& hD) {
  const float* p = f.H + (size_t)cz * f.pitch + cx;
  hA = __ldg(p); hB = __ldg(p + 1); hC = __ldg(p + f.pitch); hD = __ldg(p + f.pitch + 1);
}

// Exact plane of the Up / Down triangle of cell (cx, cz).
__device__ __forceinline__ void cell_plane(const Field& f, bool isUp, int cx, int cz, float hA, float hB, float hC,
                                           float hD, float pl[4]) {
  const float xA = cx * f.sW, xB = (cx + 1) * f.sW, zA = cz * f.sD, zC = (cz + 1) * f.sD;
  // (A, B, C) or (D, B, C): one instruction stream for both (lanes holding an Up and a Down triangle do not diverge)
  tri_plane(isUp, isUp ? xA : xB, isUp ? hA : hD, isUp ? zA : zC, xB, hB, zA, xA, hC, zC, pl);
}

// -------------------------------------------------------------------------------------------------
// K1: warp-level box-vs-heightfield decision. Returns R_FREE / R_HIT / R_DEFER (warp-uniform).
// -------------------------------------------------------------------------------------------------

// Early outs of dCollideHeightfieldZone (heightfield.cpp:1027-1064, 1139-1160) given the zone reductions.
// Returns R_FREE / R_HIT, or -1 if the vertex / plane stages are needed.
__device__ __forceinline__ int zone_early_out(const BoxCtx& b, float maxY, float minY, bool allFinite) {
  if (b.minB - maxY > -ARTP_EPS) return R_FREE;                                            // above
  if (minY - b.maxB > -ARTP_EPS) return R_FREE;                                            // under (art_planner mod)
  if (allFinite && minY - b.minB > -ARTP_EPS && b.maxB - maxY > -ARTP_EPS) return R_HIT;   // spans
  if (allFinite && maxY - minY < ARTP_EPS) {                                               // single plane
    const float pl[4] = {0.0f, 1.0f, 0.0f, minY};
    float cx[4], cz[4];
    return box_plane(b, pl, 1, cx, cz) > 0 ? R_HIT : R_FREE;
  }
  if (b.x1 - b.x0 < 1 || b.z1 - b.z0 < 1) return R_FREE;                                   // no cell, no triangle
  return -1;
}

// Zone min / max / all-finite (heightfield.cpp:1002-1026). Fast path: exact range tables -- the zone is covered
// by <= 32 overlapping 2^k x 2^k windows (one table entry per lane; max/min/or are idempotent so overlap is
// harmless). Fallback (tiny or very elongated zones, e.g. clipped at the map border): stride the zone itself.
__device__ __forceinline__ void zone_reduce(const Field& f, const BoxCtx& b, int lane, float& maxY, float& minY,
                                            bool& allFinite) {
  const int nX = b.x1 - b.x0 + 1, nZ = b.z1 - b.z0 + 1;
  float mx = -CUDART_INF_F, mn = CUDART_INF_F;
  bool fin = true;
  const int k = 31 - __clz(min(nX, nZ));
  const int cx = (nX + (1 << k) - 1) >> k, cz = (nZ + (1 << k) - 1) >> k;
  if (k >= 1 && k <= f.kmax && cx * cz <= 32) {
    if (lane < cx * cz) {
      const int iz = lane / cx, ix = lane - iz * cx, s = 1 << k;
      const int xs = min(b.x0 + ix * s, b.x1 - s + 1), zs = min(b.z0 + iz * s, b.z1 - s + 1);
      const size_t idx = (size_t)zs * f.pitch + xs;
      const float2 v = __ldg(f.T[k] + idx);
      mx = v.x; mn = v.y;
      fin = (window_flags(f.NF[k], idx - f.x_lo) & 1) == 0;
    }
  } else {
    const int nV = nX * nZ;
    const uint32_t magicX = (nX > 1) ? (0xFFFFFFFFu / (uint32_t)nX + 1u) : 0u;
    const float* base = f.H + (size_t)b.z0 * f.pitch + b.x0;
    for (int t = lane; t < nV; t += 32) {
      const int zi = (nX > 1) ? (int)__umulhi((uint32_t)t, magicX) : t;
      const int xi = t - zi * nX;
      const float h = __ldg(base + (size_t)zi * f.pitch + xi);
      mx = fmaxf(mx, h);
      if (finitef(h)) mn = fminf(mn, h); else fin = false;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    mn = fminf(mn, __shfl_xor_sync(kFull, mn, o));
  }
  allFinite = __all_sync(kFull, fin);
  maxY = mx; minY = mn;
}

// The zone as the warp stage reads it: element (xi, zi) -- vertex (x0 + xi, z0 + zi) -- at p[zi * stride + xi]; either the
// heightfield itself (p = H + z0 * pitch + x0, stride = pitch) or a shared-memory tile the zone was copied into by TMA.
struct ZoneView { const float* p; int stride; };
template <bool SMEM>
__device__ __forceinline__ float zload(const ZoneView& v, int xi, int zi) {
  const float* q = v.p + zi * v.stride + xi;
  return SMEM ? *q : __ldg(q);
}
template <bool SMEM>
__device__ __forceinline__ void zcell(const ZoneView& v, int lx, int lz, float& hA, float& hB, float& hC, float& hD) {
  const float* q = v.p + lz * v.stride + lx;
  if (SMEM) { hA = q[0]; hB = q[1]; hC = q[v.stride]; hD = q[v.stride + 1]; }
  else { hA = __ldg(q); hB = __ldg(q + 1); hC = __ldg(q + v.stride); hD = __ldg(q + v.stride + 1); }
}

// Bloom keys of the merge screen. Level 1 (cheap, every kept triangle): buckets of the APPROXIMATE normal (n0, n2);
// level 2 (flagged triangles only): buckets of the EXACT plane (n0, n2, d).
__device__ __forceinline__ uint32_t bloom_hash3(int kx, int kz, int kd) {
  return ((((uint32_t)kx * 0x9E3779B1u) ^ ((uint32_t)kz * 0x85EBCA77u)) + (uint32_t)kd * 0xC2B2AE3Du) >> 20;   // 12 bits
}
constexpr float kDScale = 524288.0f;    // level-2 bucket width 2^-19 on d (> 2 eps, so |d_m - d_c| < eps spans <= 2 buckets)
__device__ __forceinline__ int dkey(float d) { return (int)floorf(fminf(fmaxf(d, -2000.0f), 2000.0f) * kDScale); }

// Warp-level decision for one box; the zone is read through `zv`. Returns R_FREE / R_HIT / R_DEFER (warp-uniform).
//   (3) vertex stage: lane = vertex. Big all-finite zones (the torso's ~40 x 40 vertices) are walked window by window:
//       the level-3 range table holds the maximum of every 8 x 8 vertex window, and a window whose maximum does not
//       exceed the box bottom holds neither a colliding vertex nor a kept triangle, so it is skipped in the vertex
//       stage AND in the merge screen -- on rough terrain most torso boxes hover over all but a few peaks.
//   (4) plane stage: a plane contact point is always a box corner (up to a few ulps), so only the triangles in the
//       cells under the 8 corners (+- cell_margin) can report one: <= 64 candidates, lane = candidate.
//   (5) merge screen: does an EARLIER kept triangle epsilon-match a live candidate (greedy grouping,
//       heightfield.cpp:1511-1556)? Two bloom levels: the approximate normal of every kept triangle against the
//       candidates' (n0, n2) buckets; a flagged lane builds its exact plane and tests the (n0, n2, d) buckets -- matching
//       planes need |d_m - d_c| < eps, which on non-degenerate terrain never happens, so the exact compare against the
//       candidate list (shared memory) is reached only on flat / terraced ground. Live candidates themselves are
//       skipped by the screen (a bitmap over the triangle indices); candidate-vs-candidate matches are caught when the
//       level-2 keys are inserted. Any match => R_DEFER (exact stage C).
// Code size matters here (round 2: a 5 400-instruction version stalled on instruction fetch, `no_instruction` 9 of 16
// cycles per issue): one region loop serves both the windowed and the flat case, nothing is unrolled.
template <bool SMEM>
__device__ int box_collide_warp(const Field& f, const BoxCtx& b, const ZoneView& zv, WarpScratch& ws, int lane,
                                float cell_margin, bool needs_reduce, bool all_finite_known, bool merge_free) {
  const int nX = b.x1 - b.x0 + 1, nZ = b.z1 - b.z0 + 1;
  const int nV = nX * nZ;

  // (1)+(2) zone reductions and early outs -- normally already done by stage A
  bool allFinite = all_finite_known;
  if (needs_reduce) {
    const uint32_t magicX = magic_for(nX);
    float mx = -CUDART_INF_F, mn = CUDART_INF_F;
    bool fin = true;
#pragma unroll 1
    for (int t = lane; t < nV; t += 32) {
      const int zi = (nX > 1) ? (int)__umulhi((uint32_t)t, magicX) : t;
      const float h = zload<SMEM>(zv, t - zi * nX, zi);
      mx = fmaxf(mx, h);
      if (finitef(h)) mn = fminf(mn, h); else fin = false;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
      mn = fminf(mn, __shfl_xor_sync(kFull, mn, o));
    }
    allFinite = __all_sync(kFull, fin);
    const int e = zone_early_out(b, mx, mn, allFinite);
    if (e >= 0) return e;
  }

  const int nCZ = nZ - 1;
  // A vertex strictly inside the box lies within the box's vertical extent: h >= maxB + slack cannot be inside
  // (slack = 1e-4 + 4e-6 |maxB|: > 10x the rounding of the rotated coordinates and of maxB itself).
  const float top = b.maxB + (1e-4f + 4e-6f * fabsf(b.maxB));

  // Regions: the active 8 x 8-vertex windows of a big all-finite zone (bit w of `act`), else the whole zone.
  const int nWx = (nX + 5) / 7, nWz = (nZ + 5) / 7;       // windows advance by 7 cells
  const bool windowed = allFinite && nV > 256 && f.kmax >= 3 && nX >= 8 && nZ >= 8 && nWx * nWz <= 64;
  // ... and, for the vertex stage only, `actv`: the windows that can hold a vertex INSIDE the box. A window is the
  // axis-aligned block [xs, xs+7] x [min h, max h] x [zs, zs+7] (level-3 table: max and min of its 64 heights); if
  // that block lies entirely beyond one of the box's six faces -- its extent along the face normal, evaluated at the
  // block's corners, stays outside +-side/2 by more than 2 mm (>> the rounding of vertex_inside's fp32 arithmetic at
  // map coordinates of a few hundred metres) -- no vertex of it passes dGeomBoxPointDepth. For a tilted torso over rough
  // ground this is far tighter than "window max > box bottom": the AABB bottom lies below most of the zone, the
  // box's own bottom face does not. Kept triangles are a different question (h > minB): the plane stage keeps `act`.
  unsigned long long act = 1ull, actv = 1ull;
  if (windowed) {
    const float2* __restrict__ T3 = f.T[3];
    act = 0ull; actv = 0ull;
    const uint32_t magicWx = magic_for(nWx);
    // per face axis j: centre offset and the block's half extents along it (the x / z half extents are the same for
    // every window: 3.5 cells), in centre / radius form
    const float hx = 3.5f * f.sW, hz = 3.5f * f.sD;
    const float rxz0 = fabsf(b.R1[0]) * hx + fabsf(b.R1[6]) * hz, rxz1 = fabsf(b.R1[1]) * hx + fabsf(b.R1[7]) * hz,
                rxz2 = fabsf(b.R1[2]) * hx + fabsf(b.R1[8]) * hz;
    const float lim0 = 0.5f * b.side[0] + 2e-3f, lim1 = 0.5f * b.side[1] + 2e-3f, lim2 = 0.5f * b.side[2] + 2e-3f;
#pragma unroll 1
    for (int w0 = 0; w0 < nWx * nWz; w0 += 32) {
      const int wi = w0 + lane;
      bool a = false, av = false;
      if (wi < nWx * nWz) {
        const int wz = (nWx > 1) ? (int)__umulhi((uint32_t)wi, magicWx) : wi, wx = wi - wz * nWx;
        const int xs = min(b.x0 + 7 * wx, b.x1 - 7), zs = min(b.z0 + 7 * wz, b.z1 - 7);
        const float2 mm = __ldg(T3 + (size_t)zs * f.pitch + xs);
        a = mm.x > b.minB;
        if (a) {
          const float dx = ((float)xs + 3.5f) * f.sW - b.P[0], dz = ((float)zs + 3.5f) * f.sD - b.P[2];
          const float dy = 0.5f * (mm.x + mm.y) - b.P[1], hy = 0.5f * (mm.x - mm.y);
          const float c0 = b.R1[0] * dx + b.R1[3] * dy + b.R1[6] * dz, c1 = b.R1[1] * dx + b.R1[4] * dy + b.R1[7] * dz,
                      c2 = b.R1[2] * dx + b.R1[5] * dy + b.R1[8] * dz;
          const bool sep = fabsf(c0) - (rxz0 + fabsf(b.R1[3]) * hy) > lim0 || fabsf(c1) - (rxz1 + fabsf(b.R1[4]) * hy) > lim1 ||
                           fabsf(c2) - (rxz2 + fabsf(b.R1[5]) * hy) > lim2;
          av = !sep;
        }
      }
      act |= (unsigned long long)__ballot_sync(kFull, a) << w0;
      actv |= (unsigned long long)__ballot_sync(kFull, av) << w0;
    }
    if (act == 0ull) return R_FREE;   // no vertex above the box bottom: no colliding vertex, no kept triangle
  }
  // neighbouring windows share a row / column of vertices: when most of them are active one pass over the zone is cheaper
  const bool by_window = windowed && 3 * __popcll(act) <= 2 * nWx * nWz;
  const bool by_window_v = windowed && 3 * __popcll(actv) <= 2 * nWx * nWz;
  if (!by_window) act = 1ull;
  if (!by_window_v) actv = 1ull;          // (windowed && actv == 0 stays 0: the vertex stage has nothing to scan)
  // region r of the walk: origin (lx0, lz0) and extent (rw, rh) in vertices
  auto region = [&](bool by_window, int wi, int& lx0, int& lz0, int& rw, int& rh) {
    if (by_window) {
      const int wz = wi / nWx, wx = wi - wz * nWx;
      lx0 = min(7 * wx, nX - 8); lz0 = min(7 * wz, nZ - 8); rw = 8; rh = 8;
    } else { lx0 = 0; lz0 = 0; rw = nX; rh = nZ; }
  };

  // (3) vertex-in-box test of every colliding vertex of a kept triangle (heightfield.cpp:1306-1441)
  if (allFinite) {
    // every colliding vertex belongs to some kept triangle (all finite, >= 1 cell)
    if (nV > 128 && actv != 0ull) {
      // Probe pass: one vertex per lane on an 8 x 4 lattice over the box footprint (a hit anywhere in the zone is the
      // reference's answer) -- finds most intersecting torso boxes in one step.
      const float u0 = (((float)(lane & 7) + 0.5f) * 0.25f - 1.0f) * (0.5f * b.side[0]);
      const float u1 = (((float)(lane >> 3) + 0.5f) * 0.5f - 1.0f) * (0.5f * b.side[1]);
      const float qx = b.P[0] + u0 * b.R1[0] + u1 * b.R1[1], qz = b.P[2] + u0 * b.R1[6] + u1 * b.R1[7];
      const int vx = min(max(__float2int_rn(qx * f.iW), b.x0), b.x1), vz = min(max(__float2int_rn(qz * f.iD), b.z0), b.z1);
      const float h = zload<SMEM>(zv, vx - b.x0, vz - b.z0);
      const bool hit = h > b.minB && h < top && vertex_inside(b, vx * f.sW, h, vz * f.sD);
      if (__any_sync(kFull, hit)) return R_HIT;
    }
    // the whole-zone walk covers only the box's own xz extent: a point inside the box has |x - P.x| <= sum_j |R1[0][j]|
    // side_j / 2 (likewise z); the zone is that extent padded to whole cells, its outer ring cannot hold an inside vertex
    int ix0 = 0, iz0 = 0, iw = nX, ih = nZ;
    if (!by_window_v) {
      const float xr = 0.5f * (fabsf(b.R1[0] * b.side[0]) + fabsf(b.R1[1] * b.side[1]) + fabsf(b.R1[2] * b.side[2])) + 1e-4f;
      const float zr = 0.5f * (fabsf(b.R1[6] * b.side[0]) + fabsf(b.R1[7] * b.side[1]) + fabsf(b.R1[8] * b.side[2])) + 1e-4f;
      const int vx0 = max(b.x0, (int)ceilf((b.P[0] - xr) * f.iW)), vx1 = min(b.x1, (int)floorf((b.P[0] + xr) * f.iW));
      const int vz0 = max(b.z0, (int)ceilf((b.P[2] - zr) * f.iD)), vz1 = min(b.z1, (int)floorf((b.P[2] + zr) * f.iD));
      ix0 = vx0 - b.x0; iz0 = vz0 - b.z0; iw = max(vx1 - vx0 + 1, 0); ih = max(vz1 - vz0 + 1, 0);
    }
#pragma unroll 1
    for (unsigned long long m = actv; m; m &= m - 1) {
      int lx0, lz0, rw, rh;
      region(by_window_v, __ffsll((long long)m) - 1, lx0, lz0, rw, rh);
      if (!by_window_v) { lx0 = ix0; lz0 = iz0; rw = iw; rh = ih; }
      const uint32_t magicW = magic_for(rw);
      const int n = rw * rh;
#pragma unroll 1
      for (int t0 = 0; t0 < n; t0 += 32) {
        const int t = t0 + lane;
        bool hit = false;
        if (t < n) {
          const int q = (rw > 1) ? (int)__umulhi((uint32_t)t, magicW) : t;
          const int xi = lx0 + t - q * rw, zi = lz0 + q;
          const float h = zload<SMEM>(zv, xi, zi);
          hit = h > b.minB && h < top && vertex_inside(b, (b.x0 + xi) * f.sW, h, (b.z0 + zi) * f.sD);
        }
        if (__any_sync(kFull, hit)) return R_HIT;
      }
    }
  } else {
    const int nCX = nX - 1, nC = nCX * nCZ;
    const uint32_t magicC = magic_for(nCX);
#pragma unroll 1
    for (int t0 = 0; t0 < nC; t0 += 32) {
      const int t = t0 + lane;
      bool hit = false;
      if (t < nC) {
        const int czi = (nCX > 1) ? (int)__umulhi((uint32_t)t, magicC) : t;   // flattened, x fastest
        const int cxi = t - czi * nCX;
        const int cx = b.x0 + cxi, cz = b.z0 + czi;
        float hA, hB, hC, hD;
        zcell<SMEM>(zv, cxi, czi, hA, hB, hC, hD);
        const bool fA = finitef(hA), fB = finitef(hB), fC = finitef(hC), fD = finitef(hD);
        const bool cA = fA && hA > b.minB, cB = fB && hB > b.minB, cC = fC && hC > b.minB, cD = fD && hD > b.minB;
        const bool keepUp = (cA || cB || cC) && (fA && fB && fC);
        const bool keepDn = (cB || cC || cD) && (fB && fC && fD);
        // vertex v of the cell (0 A, 1 B, 2 C, 3 D) is tested if it collides and belongs to a kept triangle
#pragma unroll 1
        for (int v = 0; v < 4 && !hit; ++v) {
          const bool cv = v == 0 ? (keepUp && cA) : v == 1 ? ((keepUp || keepDn) && cB) : v == 2 ? ((keepUp || keepDn) && cC) : (keepDn && cD);
          if (cv) hit = vertex_inside(b, (cx + (v & 1)) * f.sW, v == 0 ? hA : v == 1 ? hB : v == 2 ? hC : hD, (cz + (v >> 1)) * f.sD);
        }
      }
      if (__any_sync(kFull, hit)) return R_HIT;
    }
  }

  // (4) plane stage (heightfield.cpp:1474-1617)
  const int T = 2 * (nX - 1) * nCZ;                      // triangles of the zone
  const bool use_bits = T <= 32 * kBloomWords;           // candidate bitmap over the triangle indices
  // merge_free: the map's plane tables (artp_set_map) say that no two triangles of this zone lie in one plane (within
  // eps): every kept triangle is its own group, no screen is needed -- the normal case on natural terrain.
  __syncwarp();   // the previous box's reads of this warp's scratch are done (no WAR across boxes)
  if (!merge_free) {
#pragma unroll 1
    for (int i = lane; i < kBloomWords; i += 32) { ws.bloom[i] = 0u; ws.bloom2[i] = 0u; ws.cand_bits[i] = 0u; }
  }
  __syncwarp();
  // Task t = lane + 32 * pass: corner (t >> 1) & 7, triangle t & 1 (Up / Down of a cell on neighbouring lanes), sub-cell
  // t >> 4. A corner has a second / third / fourth candidate cell only when it lies within cell_margin of a cell
  // boundary, so pass 0 is normally 16 busy lanes and pass 1 is skipped by the whole warp.
  const int corner = (lane >> 1) & 7, u = lane & 1;
  float px = b.P[0], py = b.P[1], pz = b.P[2];
  {
    const float h0 = 0.5f * b.side[0], h1 = 0.5f * b.side[1], h2 = 0.5f * b.side[2];
    if (corner & 1) { px += h0 * b.R1[0]; py += h0 * b.R1[3]; pz += h0 * b.R1[6]; } else { px -= h0 * b.R1[0]; py -= h0 * b.R1[3]; pz -= h0 * b.R1[6]; }
    if (corner & 2) { px += h1 * b.R1[1]; py += h1 * b.R1[4]; pz += h1 * b.R1[7]; } else { px -= h1 * b.R1[1]; py -= h1 * b.R1[4]; pz -= h1 * b.R1[7]; }
    if (corner & 4) { px += h2 * b.R1[2]; py += h2 * b.R1[5]; pz += h2 * b.R1[8]; } else { px -= h2 * b.R1[2]; py -= h2 * b.R1[5]; pz -= h2 * b.R1[8]; }
  }
  // Height of this lane's corner (the lower one of the vertical pair that may be folded into one lane below). Every
  // contact dCollideBoxPlane returns is a box corner lying BELOW the plane; if it also lies on the triangle, the plane
  // there is no higher than the triangle's highest vertex. So in a merge-free zone (each triangle is tested against its
  // own plane only) a candidate triangle whose vertices all lie more than 1 mm below the corner that selected it
  // cannot be hit through that corner -- and the lane of any other corner over the same cell tests it for itself.
  // A torso hovering over rough ground keeps nearly all candidates by the h > minB rule (its AABB bottom is low) and
  // drops nearly all of them by this one.
  const float py_pair = fminf(py, __shfl_xor_sync(kFull, py, 8));
  const float gx = px * f.iW, gz = pz * f.iD;
  const int cxl = (int)floorf(gx - cell_margin), cxh = (int)floorf(gx + cell_margin);
  const int czl = (int)floorf(gz - cell_margin), czh = (int)floorf(gz + cell_margin);
  bool hit_own = false, pair_possible = false;
  int nLive = 0, max_live = -1;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int sub = 2 * pass + (lane >> 4);
    const int ccx = (sub & 1) ? cxh : cxl, ccz = (sub & 2) ? czh : czl;
    bool actc = !((sub & 1) && cxh == cxl) && !((sub & 2) && czh == czl);
    actc = actc && ccx >= b.x0 && ccx < b.x1 && ccz >= b.z0 && ccz < b.z1;
    if (!__any_sync(kFull, actc)) continue;
    // an upright box projects its top corners into the cells of the bottom corners: the same triangle twice
    {
      const int pcx = __shfl_xor_sync(kFull, ccx, 8), pcz = __shfl_xor_sync(kFull, ccz, 8);
      const bool pact = __shfl_xor_sync(kFull, actc, 8);
      if ((corner & 4) && pact && pcx == ccx && pcz == ccz) actc = false;
    }
    bool live = false;
    float pl[4] = {0.f, 0.f, 0.f, 0.f};
    int idx = -1;
    if (actc) {
      float hA, hB, hC, hD;
      zcell<SMEM>(zv, ccx - b.x0, ccz - b.z0, hA, hB, hC, hD);
      const bool fA = finitef(hA), fB = finitef(hB), fC = finitef(hC), fD = finitef(hD);
      const bool cA = fA && hA > b.minB, cB = fB && hB > b.minB, cC = fC && hC > b.minB, cD = fD && hD > b.minB;
      const bool isUp = (u == 0);
      bool keep = isUp ? ((cA || cB || cC) && (fA && fB && fC)) : ((cB || cC || cD) && (fB && fC && fD));
      if (merge_free && keep) {
        const float hT = isUp ? fmaxf(hA, fmaxf(hB, hC)) : fmaxf(hD, fmaxf(hB, hC));
        keep = hT > py_pair - 1e-3f;
      }
      if (keep) {
        cell_plane(f, isUp, ccx, ccz, hA, hB, hC, hD, pl);
        // Liveness: any plane within eps of this one changes the box-plane depth by far less than tau.
        const float Q1 = pl[0] * b.R1[0] + pl[1] * b.R1[3] + pl[2] * b.R1[6];
        const float Q2 = pl[0] * b.R1[1] + pl[1] * b.R1[4] + pl[2] * b.R1[7];
        const float Q3 = pl[0] * b.R1[2] + pl[1] * b.R1[5] + pl[2] * b.R1[8];
        const float B1 = fabsf(b.side[0] * Q1), B2 = fabsf(b.side[1] * Q2), B3 = fabsf(b.side[2] * Q3);
        const float depth = pl[3] + 0.5f * (B1 + B2 + B3) - (pl[0] * b.P[0] + pl[1] * b.P[1] + pl[2] * b.P[2]);
        const float tau = 16.0f * ARTP_EPS * (1.0f + b.side[0] + b.side[1] + b.side[2] + fabsf(b.P[0]) +
                                              fabsf(b.P[1]) + fabsf(b.P[2]) + fabsf(pl[3]));
        if (depth >= -tau) {   // else dead: no plane of its would-be group can touch the box
          live = true;
          idx = ((ccx - b.x0) * nCZ + (ccz - b.z0)) * 2 + u;   // emission order: x outer, z inner, Up, Down
          if (!merge_free) {
          if (use_bits) atomicOr(&ws.cand_bits[idx >> 5], 1u << (idx & 31));
          // level-1 keys: every bucket the approximate normal of an eps-matching triangle may fall into;
          // level-2 keys: every bucket its exact (n0, n2, d) may fall into. The candidate's own level-2 bucket goes in
          // last (after every candidate of the pass has inserted its neighbour buckets): finding it occupied means
          // another live candidate may match this one.
          const int kxc = (int)floorf((pl[0] + 1.0f) * kKeyScale), kzc = (int)floorf((pl[2] + 1.0f) * kKeyScale), kdc = dkey(pl[3]);
          const int kx0 = (int)floorf((pl[0] - kKeyMargin + 1.0f) * kKeyScale), kx1 = (int)floorf((pl[0] + kKeyMargin + 1.0f) * kKeyScale);
          const int kz0 = (int)floorf((pl[2] - kKeyMargin + 1.0f) * kKeyScale), kz1 = (int)floorf((pl[2] + kKeyMargin + 1.0f) * kKeyScale);
          const int kd0 = dkey(pl[3] - 2.0f * ARTP_EPS), kd1 = dkey(pl[3] + 2.0f * ARTP_EPS);
#pragma unroll 1
          for (int kx = kx0; kx <= kx1; ++kx)
#pragma unroll 1
            for (int kz = kz0; kz <= kz1; ++kz) {
              const uint32_t hsh = bloom_hash(kx, kz);
              atomicOr(&ws.bloom[hsh >> 5], 1u << (hsh & 31));
#pragma unroll 1
              for (int kd = kd0; kd <= kd1; ++kd) {
                if (kx == kxc && kz == kzc && kd == kdc) continue;
                const uint32_t h3 = bloom_hash3(kx, kz, kd);
                atomicOr(&ws.bloom2[h3 >> 5], 1u << (h3 & 31));
              }
            }
          }
          // contact points with the triangle's OWN plane (valid if it turns out to be its group base)
          float cx[4], cz[4];
          const int nc = box_plane(b, pl, 4, cx, cz);
          const int tcx = isUp ? ccx : ccx + 1, tcz = isUp ? ccz : ccz + 1;
#pragma unroll 1
          for (int i = 0; i < nc; ++i) hit_own = hit_own || on_tri(f, isUp, tcx, tcz, cx[i], cz[i]);
        }
      }
    }
    __syncwarp();
    if (live && !merge_free) {
      const uint32_t h3 = bloom_hash3((int)floorf((pl[0] + 1.0f) * kKeyScale), (int)floorf((pl[2] + 1.0f) * kKeyScale), dkey(pl[3]));
      if ((atomicOr(&ws.bloom2[h3 >> 5], 1u << (h3 & 31)) >> (h3 & 31)) & 1u) pair_possible = true;
    }
    // append this pass's live candidates to the list
    const unsigned lm = __ballot_sync(kFull, live);
    if (nLive + __popc(lm) > kMaxCand) return R_DEFER;   // exact fallback (not seen in practice)
    if (live) {
      const int p = nLive + __popc(lm & ((1u << lane) - 1u));
      ws.cpl[p][0] = pl[0]; ws.cpl[p][1] = pl[1]; ws.cpl[p][2] = pl[2]; ws.cpl[p][3] = pl[3];
      ws.cidx[p] = idx;
    }
    nLive += __popc(lm);
    max_live = max(max_live, idx);
  }
  if (nLive == 0) return R_FREE;
  if (merge_free) return __any_sync(kFull, hit_own) ? R_HIT : R_FREE;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) max_live = max(max_live, __shfl_xor_sync(kFull, max_live, o));
  __syncwarp();
  if (!use_bits) pair_possible = true;
  if (__any_sync(kFull, pair_possible)) {
    // candidate against candidate, exactly: any two matching live candidates => one may absorb the other
    bool merge = false;
    if (lane < nLive) {
#pragma unroll 1
      for (int j = 0; j < nLive; ++j)
        if (ws.cidx[j] != ws.cidx[lane] && plane_match(ws.cpl[lane], ws.cpl[j])) merge = true;
    }
    if (__any_sync(kFull, merge)) return R_DEFER;
  }

  // (5) merge screen over the regions: one cell per lane
#pragma unroll 1
  for (unsigned long long m = act; m; m &= m - 1) {
    int lx0, lz0, rw, rh;
    region(by_window, __ffsll((long long)m) - 1, lx0, lz0, rw, rh);
    if ((lx0 * nCZ + lz0) * 2 >= max_live) continue;       // the whole region is emitted after the last live candidate
    const int cw = rw - 1, n = cw * (rh - 1);
    const uint32_t magicW = magic_for(cw);
#pragma unroll 1
    for (int t0 = 0; t0 < n; t0 += 32) {
      const int t = t0 + lane;
      bool merge = false;
      if (t < n) {
        const int q = (cw > 1) ? (int)__umulhi((uint32_t)t, magicW) : t;
        const int cxi = lx0 + t - q * cw, czi = lz0 + q;
        const int cell_idx = (cxi * nCZ + czi) * 2;
        if (cell_idx < max_live) {   // only triangles emitted before the last live candidate can absorb one
          const int cx = b.x0 + cxi, cz = b.z0 + czi;
          float hA, hB, hC, hD;
          zcell<SMEM>(zv, cxi, czi, hA, hB, hC, hD);
          const bool fA = finitef(hA), fB = finitef(hB), fC = finitef(hC), fD = finitef(hD);
          const bool cA = fA && hA > b.minB, cB = fB && hB > b.minB, cC = fC && hC > b.minB, cD = fD && hD > b.minB;
          const bool keepUp = (cA || cB || cC) && (fA && fB && fC), keepDn = (cB || cC || cD) && (fB && fC && fD);
          if (keepUp || keepDn) {
            const float xA = cx * f.sW, xB = (cx + 1) * f.sW, zA = cz * f.sD, zC = (cz + 1) * f.sD;
#pragma unroll 1
            for (int uu = 0; uu < 2; ++uu) {
              const int idx = cell_idx + uu;
              if (!(uu == 0 ? keepUp : keepDn) || idx >= max_live) continue;
              if (use_bits && ((ws.cand_bits[idx >> 5] >> (idx & 31)) & 1u)) continue;   // live candidates: handled above
              // approximate normal (cross product value-identical to tri_plane's, zero terms dropped):
              // Up (A,B,C): E1 = C-A, E2 = B-A, c = E1 x E2;  Down (D,B,C): E1 = C-D, E2 = B-D, c = E2 x E1
              const float ey = uu == 0 ? hC - hA : hB - hD, ez = uu == 0 ? zC - zA : zA - zC;
              const float gx2 = uu == 0 ? xB - xA : xA - xB, gy = uu == 0 ? hB - hA : hC - hD;
              const float c0 = -(ez * gy), c1 = ez * gx2, c2 = -(ey * gx2);
              const float r = rsqrtf(c0 * c0 + c1 * c1 + c2 * c2);
              const uint32_t hsh = bloom_hash((int)floorf((c0 * r + 1.0f) * kKeyScale), (int)floorf((c2 * r + 1.0f) * kKeyScale));
              if (ws.bloom[hsh >> 5] & (1u << (hsh & 31))) {
                float plm[4];   // flagged: exact plane, level-2 buckets
                cell_plane(f, uu == 0, cx, cz, hA, hB, hC, hD, plm);
                const uint32_t h3 = bloom_hash3((int)floorf((plm[0] + 1.0f) * kKeyScale), (int)floorf((plm[2] + 1.0f) * kKeyScale), dkey(plm[3]));
                if (ws.bloom2[h3 >> 5] & (1u << (h3 & 31))) {
#pragma unroll 1
                  for (int sidx = 0; sidx < nLive; ++sidx)
                    if (ws.cidx[sidx] > idx && plane_match(plm, ws.cpl[sidx])) merge = true;
                }
              }
            }
          }
        }
      }
      if (__any_sync(kFull, merge)) return R_DEFER;
    }
  }
  const int res = __any_sync(kFull, hit_own) ? R_HIT : R_FREE;
  __syncwarp();   // scratch is reused by the next box
  return res;
}

// One queued box (80 bytes): everything stage B/C need, so nothing is recomputed.
struct alignas(16) BoxRec {
  float R1[9];
  float P[3];
  float minB, maxB;
  int x0, x1, z0, z1;
  uint32_t item;     // verdict slot of the work item (pose index / edge index / interior-state index)
  uint32_t flags;    // bits 0-2: box (0 torso, 1..4 feet); bit 3: zone all finite; bit 4: zone not reduced yet;
                     // bit 5: no mergeable triangle pair in the zone (plane tables, artp_set_map)
};
static_assert(sizeof(BoxRec) == 80, "BoxRec is read as five 16-byte words");
enum { REC_ALLFINITE = 8, REC_NEEDS_REDUCE = 16, REC_MERGEFREE = 32 };
constexpr uint32_t kPendingGroup = 0x80000000u;   // classify-internal tag: the reach box goes to the 8-lane-group queue

__device__ __forceinline__ void rec_to_ctx(const Checker& c, const BoxRec& r, BoxCtx& b) {
#pragma unroll
  for (int i = 0; i < 9; ++i) b.R1[i] = r.R1[i];
  b.P[0] = r.P[0]; b.P[1] = r.P[1]; b.P[2] = r.P[2];
  b.minB = r.minB; b.maxB = r.maxB;
  b.x0 = r.x0; b.x1 = r.x1; b.z0 = r.z0; b.z1 = r.z1;
  const int w = (r.flags & 7) ? 1 : 0;
  b.side[0] = c.side[w][0]; b.side[1] = c.side[w][1]; b.side[2] = c.side[w][2];
}

// One box of one item in the classify stage: box pose, zone, range-table reductions, the collider's early outs and the
// vertex probes. Returns R_FREE / R_HIT when decided, -1 when the box needs the vertex / plane stages (b and fl are
// then complete), kBoxOutside when the box centre is outside the map (the caller applies the outside-map rule).
constexpr int kBoxOutside = -2;
constexpr int kBoxOutsideWindow = -3;   // map shards (artp_set_map_window): the item is reported invalid + sticky error
__device__ __forceinline__ int classify_box(const Checker& c, const float R[9], const float R1[9], const float t[3], int k,
                                            int force_all, int probe /* bit 0: reach boxes, bit 1: torso */, BoxCtx& b,
                                            uint32_t& fl) {
  const Field& g = c.f[0];      // both layers share the map geometry
  const bool foot = k > 0;
  const int fk = k - 1;
  const float ox = foot ? ((fk & 2) ? -c.feet_ox : c.feet_ox) : c.torso_off[0];
  const float oy = foot ? ((fk & 1) ? -c.feet_oy : c.feet_oy) : c.torso_off[1];
  const float oz = foot ? 0.0f : c.torso_off[2];
  const float sd0 = foot ? c.side[1][0] : c.side[0][0], sd1 = foot ? c.side[1][1] : c.side[0][1],
              sd2 = foot ? c.side[1][2] : c.side[0][2];
  float tt[3];
  compose_translation(R, t, ox, oy, oz, tt);
  if (!is_inside(c, tt[0], tt[1])) return kBoxOutside;   // validity_checker_body.cpp:29-32, validity_checker_feet.cpp:34-37
  // dCollideHeightfield prologue + dxBox::computeAABB (see box_setup())
  const float d0 = tt[0] - g.px, d1 = tt[1] - g.py, d2 = tt[2] - 0.0f;
  b.P[0] = -d0 + g.hW; b.P[1] = d2; b.P[2] = d1 + g.hD;
  const float xr = 0.5f * (fabsf(R1[0] * sd0) + fabsf(R1[1] * sd1) + fabsf(R1[2] * sd2));
  const float yr = 0.5f * (fabsf(R1[3] * sd0) + fabsf(R1[4] * sd1) + fabsf(R1[5] * sd2));
  const float zr = 0.5f * (fabsf(R1[6] * sd0) + fabsf(R1[7] * sd1) + fabsf(R1[8] * sd2));
  const float a0 = b.P[0] - xr, a1 = b.P[0] + xr, a4 = b.P[2] - zr, a5 = b.P[2] + zr;
  b.minB = b.P[1] - yr; b.maxB = b.P[1] + yr;
  int r;                        // R_FREE / R_HIT / -1 undecided
  fl = (uint32_t)k;
  if ((a0 > g.W || a4 > g.D) || (a1 < 0.0f || a5 < 0.0f)) {
    r = R_FREE;                 // dCollide returns 0 (heightfield.cpp:1870-1876)
  } else {
    b.x0 = max((int)floorf(next_down(a0 * g.iW)), 0);
    b.x1 = min((int)ceilf(next_up(a1 * g.iW)), g.nx - 1);
    b.z0 = max((int)floorf(next_down(a4 * g.iD)), 0);
    b.z1 = min((int)ceilf(next_up(a5 * g.iD)), g.nz - 1);
    if (b.x0 < g.x_lo || b.x1 > g.x_hi) return kBoxOutsideWindow;   // the zone leaves this handle's map window
#pragma unroll
    for (int i = 0; i < 9; ++i) b.R1[i] = R1[i];
    b.side[0] = sd0; b.side[1] = sd1; b.side[2] = sd2;
    // zone reductions from the range tables (exact: max/min/or are idempotent, windows may overlap)
    const Field& f = foot ? c.f[1] : c.f[0];
    const int nX = b.x1 - b.x0 + 1, nZ = b.z1 - b.z0 + 1;
    const int kk = 31 - __clz(min(nX, nZ));
    const int cx = (nX + (1 << kk) - 1) >> kk, cz = (nZ + (1 << kk) - 1) >> kk;
    if (force_all || kk < 1 || kk > f.kmax || cx * cz > 32) {
      r = -1; fl |= REC_NEEDS_REDUCE;
    } else {
      const float2* __restrict__ T = f.T[kk];
      const uint32_t* __restrict__ NF = f.NF[kk];
      const int sW = 1 << kk;
      float mx = -CUDART_INF_F, mn = CUDART_INF_F;
      int nf = 0;
      if (cx <= 2 && cz <= 2) {
        // the common case (window edge > half the zone edge): all four windows are requested before any is used
        const int xs0 = b.x0, xs1 = b.x1 - sW + 1, zs0 = b.z0, zs1 = b.z1 - sW + 1;
        const size_t i00 = (size_t)zs0 * f.pitch + xs0, i01 = (size_t)zs0 * f.pitch + xs1,
                     i10 = (size_t)zs1 * f.pitch + xs0, i11 = (size_t)zs1 * f.pitch + xs1;
        const float2 v0 = __ldg(T + i00), v1 = __ldg(T + i01), v2 = __ldg(T + i10), v3 = __ldg(T + i11);
        const int n0 = window_flags(NF, i00 - f.x_lo), n1 = window_flags(NF, i01 - f.x_lo), n2 = window_flags(NF, i10 - f.x_lo), n3 = window_flags(NF, i11 - f.x_lo);
        mx = fmaxf(fmaxf(v0.x, v1.x), fmaxf(v2.x, v3.x));
        mn = fminf(fminf(v0.y, v1.y), fminf(v2.y, v3.y));
        nf = n0 | n1 | n2 | n3;
      } else {
        for (int iz = 0; iz < cz; ++iz) {
          const int zs = min(b.z0 + iz * sW, b.z1 - sW + 1);
          for (int ix = 0; ix < cx; ++ix) {
            const int xs = min(b.x0 + ix * sW, b.x1 - sW + 1);
            const size_t idx = (size_t)zs * f.pitch + xs;
            const float2 v = __ldg(T + idx);
            mx = fmaxf(mx, v.x); mn = fminf(mn, v.y);
            nf |= window_flags(NF, idx - f.x_lo);
          }
        }
      }
      const bool allFinite = (nf & 1) == 0;
      r = zone_early_out(b, mx, mn, allFinite);
      if (allFinite) fl |= REC_ALLFINITE;
      if ((nf & 2) == 0) fl |= REC_MERGEFREE;   // no two triangles of the zone lie in one plane: greedy grouping is the identity
      if (r == -1 && allFinite && nX >= 2 && nZ >= 2 && (probe & (foot ? 1 : 2))) {
        // Vertex probes. In an all-finite zone every vertex with h > minB belongs to a kept triangle, and the
        // collider returns 1 as soon as ANY such vertex lies inside the box (heightfield.cpp:1344-1441), so a
        // hit found here is exactly the reference's answer; a miss decides nothing and the box is queued.
        // Reach box: the cell under the box centre; torso: a 3x3 pattern across the footprint.
        // (measured: 3x3 / 5x5 reach-box patterns remove another 25 % of the queue but cost the classify stage
        // twice what the warp stage saves -- one thread walks them serially)
        const int np = foot ? 1 : 9;
        const float h0 = 0.5f * sd0, h1 = 0.5f * sd1;
        for (int pi = 0; pi < np && r == -1; ++pi) {
          const float u0 = foot ? 0.0f : 0.7f * (float)(pi % 3 - 1), u1 = foot ? 0.0f : 0.7f * (float)(pi / 3 - 1);
          const float qx = b.P[0] + (u0 * h0) * R1[0] + (u1 * h1) * R1[1];
          const float qz = b.P[2] + (u0 * h0) * R1[6] + (u1 * h1) * R1[7];
          const int pcx = min(max((int)floorf(qx * g.iW), b.x0), b.x1 - 1);
          const int pcz = min(max((int)floorf(qz * g.iD), b.z0), b.z1 - 1);
          float hA, hB, hC, hD;
          load_cell(f, pcx, pcz, hA, hB, hC, hD);
          const float xA = pcx * f.sW, xB = (pcx + 1) * f.sW, zA = pcz * f.sD, zC = (pcz + 1) * f.sD;
          if ((hA > b.minB && vertex_inside(b, xA, hA, zA)) || (hB > b.minB && vertex_inside(b, xB, hB, zA)) ||
              (hC > b.minB && vertex_inside(b, xA, hC, zC)) || (hD > b.minB && vertex_inside(b, xB, hD, zC)))
            r = R_HIT;
        }
      }
    }
  }
  return r;
}

// ---------------------------------------------------------------------------------------------
// Stage A: one thread per work item.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 8)
classify_items_kernel(const Checker c, const Work w, BoxRec* __restrict__ recs_w, BoxRec* __restrict__ recs_f,
                      BoxRec* __restrict__ recs_g, uint32_t* __restrict__ count_w, uint32_t* __restrict__ count_f,
                      uint32_t* __restrict__ count_g, int flags) {
  const int force_all = flags & 1;      // every in-map box goes to the grouping stage (artp_set_mode 1)
  // Vertex probes here only for reach boxes (one cell, most lanes busy); an undecided torso is rare (a few lanes of a
  // warp) and its probes run lane-parallel at the head of the warp stage instead. ARTP_K0_FLAGS=2 turns probes off.
  const int probe = (flags & 2) ? 0 : 1;
  const uint32_t item = w.item_base + blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = item < w.n_items;
  const int lane = threadIdx.x & 31;
  int result = 1;                 // 1 valid so far, 0 invalid
  int n_w = 0, n_f = 0, n_g = 0;  // undecided boxes of this item: big-tile queue / reach-box queue (of which n_g for the group kernel)
  struct Pending { float P[3], minB, maxB; int x0, x1, z0, z1; uint32_t fl; };   // R1 is shared by the item's boxes
  Pending ub[5];                  // big-tile boxes from the front, reach boxes from the back
  float R1[9];
  uint32_t slot = 0;
  if (in_range) {
    slot = item_slot(w, item);
    double s[7];
    load_item_state(w, item, s);
    float t[3], R[9], Rb[9];
    pose3_from_se3(s, t, R);
#pragma unroll
    for (int i = 0; i < 9; ++i) Rb[i] = R[i];
    orthogonalize_r(Rb);          // dBodySetRotation of the same matrix for all five boxes
#pragma unroll
    for (int j = 0; j < 3; ++j) { R1[j] = -Rb[j]; R1[3 + j] = Rb[6 + j]; R1[6 + j] = Rb[3 + j]; }
    // Reach boxes first: this stage can only INVALIDATE an item through a reach box that touches nothing (a torso box is
    // decided free here or queued, never found colliding), so on an invalid pose the torso box is usually never looked at.
#pragma unroll 1
    for (int kk = 0; kk < 5 && result; ++kk) {
      const int k = kk == 4 ? 0 : kk + 1;
      BoxCtx b;
      uint32_t fl;
      const bool foot = k > 0;
      const int r = classify_box(c, R, R1, t, k, force_all, probe, b, fl);
      if (r == kBoxOutside) {
        if (foot && c.unknown_untraversable) result = 0;
        continue;
      }
      if (r == kBoxOutsideWindow) { *(volatile uint32_t*)c.err_word = 2u; result = 0; continue; }   // fail closed
      if (r == -1) {
        // reach boxes go to their own queue (small TMA tiles); the torso -- and a reach box whose zone would not fit the
        // small tile -- to the big-tile queue
        const bool thread_path = foot && (b.x1 - b.x0) + 4 <= c.reach_tw && (b.z1 - b.z0) + 1 <= c.reach_th;
        // ... and of those, the common kind (all-finite, merge-free zone, reduced by the tables) to the 8-lane-group kernel
        const bool group_path = thread_path && recs_g != nullptr &&
                                (fl & (REC_ALLFINITE | REC_MERGEFREE | REC_NEEDS_REDUCE)) == (REC_ALLFINITE | REC_MERGEFREE);
        const int q = thread_path ? 4 - n_f++ : n_w++;
        if (group_path) ++n_g;
        Pending& pd = ub[q];
        pd.P[0] = b.P[0]; pd.P[1] = b.P[1]; pd.P[2] = b.P[2]; pd.minB = b.minB; pd.maxB = b.maxB;
        pd.x0 = b.x0; pd.x1 = b.x1; pd.z0 = b.z0; pd.z1 = b.z1; pd.fl = fl | (group_path ? kPendingGroup : 0u);
      }
      else if (!foot) { if (r == R_HIT) result = 0; }      // torso must be free
      else { if (r == R_FREE) result = 0; }                // every reach box must touch
    }
    if (!result) { n_w = 0; n_f = 0; n_g = 0; }
    // provisional result; the later stages clear it if an undecided box fails
    if (w.edge_mode) { if (!result) w.valid[slot] = 0; }
    else w.valid[slot] = (uint8_t)result;
  }
  // queue the undecided boxes: one atomic per warp and queue (inclusive scans of the per-lane counts)
  const int n_fp = n_f - n_g;     // reach boxes for the one-warp-per-box kernel
  int incl_w = n_w, incl_f = n_fp, incl_g = n_g;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int yw = __shfl_up_sync(kFull, incl_w, o), yf = __shfl_up_sync(kFull, incl_f, o), yg = __shfl_up_sync(kFull, incl_g, o);
    if (lane >= o) { incl_w += yw; incl_f += yf; incl_g += yg; }
  }
  const int total_w = __shfl_sync(kFull, incl_w, 31), total_f = __shfl_sync(kFull, incl_f, 31), total_g = __shfl_sync(kFull, incl_g, 31);
  if (total_w == 0 && total_f == 0 && total_g == 0) return;
  uint32_t base_w = 0, base_f = 0, base_g = 0;
  if (lane == 31) {
    if (total_w) base_w = atomicAdd(count_w, (uint32_t)total_w);
    if (total_f) base_f = atomicAdd(count_f, (uint32_t)total_f);
    if (total_g) base_g = atomicAdd(count_g, (uint32_t)total_g);
  }
  base_w = __shfl_sync(kFull, base_w, 31) + (uint32_t)(incl_w - n_w);
  base_f = __shfl_sync(kFull, base_f, 31) + (uint32_t)(incl_f - n_fp);
  base_g = __shfl_sync(kFull, base_g, 31) + (uint32_t)(incl_g - n_g);
#pragma unroll 1
  for (int q = 0; q < n_w + n_f; ++q) {
    const bool fq = q >= n_w;
    const int src = fq ? 4 - (q - n_w) : q;
    const bool gq = fq && (ub[src].fl & kPendingGroup);
    BoxRec& o = gq ? recs_g[base_g++] : (fq ? recs_f[base_f++] : recs_w[base_w + q]);
    const Pending& b = ub[src];
#pragma unroll
    for (int i = 0; i < 9; ++i) o.R1[i] = R1[i];
    o.P[0] = b.P[0]; o.P[1] = b.P[1]; o.P[2] = b.P[2];
    o.minB = b.minB; o.maxB = b.maxB;
    o.x0 = b.x0; o.x1 = b.x1; o.z0 = b.z0; o.z1 = b.z1;
    o.item = slot; o.flags = b.fl & ~kPendingGroup;   // later stages only need the verdict slot
  }
}

// -------------------------------------------------------------------------------------------------
// K2: block-level exact decision including the greedy epsilon grouping.
// Shared memory: planes[T][4] floats, group[T] ints, state[T] bytes, T = 2 * max cells of a zone.
// -------------------------------------------------------------------------------------------------
struct BlockShared {
  float* planes;     // [T][4]
  int* group;        // [T] base index (== own index for bases); -1 for not-kept
  uint8_t* state;    // [T] 1 = assigned
  int T_cap;
};

__device__ __forceinline__ float block_reduce_max(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float a = __shfl_xor_sync(kFull, v, o);
    v = is_max ? ((v > a) ? v : a) : ((v > a) ? a : v);
  }
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
    const float a = red[i];
    r = is_max ? ((r > a) ? r : a) : ((r > a) ? a : r);
  }
  return r;
}

// Returns R_FREE / R_HIT, or R_DEFER if the zone does not fit the shared-memory plane store (host
// sizes it so that this cannot happen for the configured boxes).
__device__ int box_collide_block(const Field& f, const BoxCtx& b, const BlockShared& sh, float* red, int* s_next) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int nX = b.x1 - b.x0 + 1, nZ = b.z1 - b.z0 + 1, nV = nX * nZ;
  const float* base = f.H + (size_t)b.z0 * f.pitch + b.x0;
  float mx = -CUDART_INF_F, mn = CUDART_INF_F;
  int fin = 1;
  for (int t = tid; t < nV; t += nthr) {
    const int zi = t / nX, xi = t - zi * nX;
    const float h = __ldg(base + (size_t)zi * f.pitch + xi);
    mx = (mx > h) ? mx : h;
    if (finitef(h)) mn = (mn > h) ? h : mn; else fin = 0;
  }
  const float maxY = block_reduce_max(mx, red, true);
  const float minY = block_reduce_max(mn, red, false);
  const bool allFinite = __syncthreads_and(fin) != 0;
  if (b.minB - maxY > -ARTP_EPS) return R_FREE;
  if (minY - b.maxB > -ARTP_EPS) return R_FREE;
  if (allFinite && minY - b.minB > -ARTP_EPS && b.maxB - maxY > -ARTP_EPS) return R_HIT;
  if (allFinite && maxY - minY < ARTP_EPS) {
    const float pl[4] = {0.0f, 1.0f, 0.0f, minY};
    float cx[4], cz[4];
    return box_plane(b, pl, 1, cx, cz) > 0 ? R_HIT : R_FREE;
  }
  if (nX < 2 || nZ < 2) return R_FREE;
  const int nCX = nX - 1, nCZ = nZ - 1, nC = nCX * nCZ, T = 2 * nC;
  if (T > sh.T_cap) return R_DEFER;
  // triangles in emission order (x outer, z inner, Up then Down): vertex tests + planes
  int hit = 0;
  for (int t = tid; t < nC; t += nthr) {
    const int cxi = t / nCZ, czi = t - cxi * nCZ;          // emission order index t = cxi*nCZ + czi
    const int cx = b.x0 + cxi, cz = b.z0 + czi;
    float hA, hB, hC, hD;
    load_cell(f, cx, cz, hA, hB, hC, hD);
    const bool fA = finitef(hA), fB = finitef(hB), fC = finitef(hC), fD = finitef(hD);
    const bool cA = fA && hA > b.minB, cB = fB && hB > b.minB, cC = fC && hC > b.minB, cD = fD && hD > b.minB;
    const bool keepUp = (cA || cB || cC) && (fA && fB && fC);
    const bool keepDn = (cB || cC || cD) && (fB && fC && fD);
    const float xA = cx * f.sW, xB = (cx + 1) * f.sW, zA = cz * f.sD, zC = (cz + 1) * f.sD;
    if (keepUp && cA) hit |= vertex_inside(b, xA, hA, zA);
    if ((keepUp || keepDn) && cB) hit |= vertex_inside(b, xB, hB, zA);
    if ((keepUp || keepDn) && cC) hit |= vertex_inside(b, xA, hC, zC);
    if (keepDn && cD) hit |= vertex_inside(b, xB, hD, zC);
    const int i0 = 2 * t;
    sh.state[i0] = 0; sh.state[i0 + 1] = 0;
    sh.group[i0] = -1; sh.group[i0 + 1] = -1;
    if (keepUp) { cell_plane(f, true, cx, cz, hA, hB, hC, hD, sh.planes + 4 * i0); sh.group[i0] = i0; }
    if (keepDn) { cell_plane(f, false, cx, cz, hA, hB, hC, hD, sh.planes + 4 * (i0 + 1)); sh.group[i0 + 1] = i0 + 1; }
  }
  if (__syncthreads_or(hit)) return R_HIT;
  // Singleton screen. The greedy grouping below is sequential in the number of groups (~T when nothing merges:
  // 2000 block-wide steps, 60 us for a torso zone). A triangle can absorb or be absorbed only if some OTHER kept
  // triangle epsilon-matches it, and matching planes have normals within eps, i.e. the other's (n0, n2) bucket lies in
  // this one's +-kKeyMargin neighbourhood. Two bit tables over the hashed buckets (occupied, occupied twice) find the
  // triangles that cannot have a partner: they are their own group, marked assigned up front; the sequential loop only
  // walks the rest (normally none).
  {
    __shared__ uint32_t occ1[kBloomWords], occ2[kBloomWords];
    for (int i = tid; i < kBloomWords; i += nthr) { occ1[i] = 0u; occ2[i] = 0u; }
    __syncthreads();
    for (int m = tid; m < T; m += nthr) {
      if (sh.group[m] < 0) continue;
      const float* pm = sh.planes + 4 * m;
      const uint32_t hsh = bloom_hash((int)floorf((pm[0] + 1.0f) * kKeyScale), (int)floorf((pm[2] + 1.0f) * kKeyScale));
      const uint32_t bit = 1u << (hsh & 31);
      if (atomicOr(&occ1[hsh >> 5], bit) & bit) atomicOr(&occ2[hsh >> 5], bit);
    }
    __syncthreads();
    int any_pot = 0;
    for (int m = tid; m < T; m += nthr) {
      if (sh.group[m] < 0) continue;
      const float* pm = sh.planes + 4 * m;
      const int kxc = (int)floorf((pm[0] + 1.0f) * kKeyScale), kzc = (int)floorf((pm[2] + 1.0f) * kKeyScale);
      const int kx0 = (int)floorf((pm[0] - kKeyMargin + 1.0f) * kKeyScale), kx1 = (int)floorf((pm[0] + kKeyMargin + 1.0f) * kKeyScale);
      const int kz0 = (int)floorf((pm[2] - kKeyMargin + 1.0f) * kKeyScale), kz1 = (int)floorf((pm[2] + kKeyMargin + 1.0f) * kKeyScale);
      bool pot = false;
      for (int kx = kx0; kx <= kx1; ++kx)
        for (int kz = kz0; kz <= kz1; ++kz) {
          const uint32_t hsh = bloom_hash(kx, kz);
          const uint32_t* tab = (kx == kxc && kz == kzc) ? occ2 : occ1;   // own bucket: someone else must be there too
          pot = pot || ((tab[hsh >> 5] >> (hsh & 31)) & 1u);
        }
      if (pot) any_pot = 1; else sh.state[m] = 1;
    }
    any_pot = __syncthreads_or(any_pot);
    if (!any_pot) goto groups_done;
  }
  // greedy grouping (heightfield.cpp:1511-1556)
  {
  int k = -1;
  for (;;) {
    if (tid == 0) {
      int q = k + 1;
      while (q < T && (sh.group[q] < 0 || sh.state[q])) ++q;
      *s_next = q;
    }
    __syncthreads();
    k = *s_next;
    if (k >= T) break;
    const float* pk = sh.planes + 4 * k;
    const float p0 = pk[0], p1 = pk[1], p2 = pk[2], p3 = pk[3];
    for (int m = k + 1 + tid; m < T; m += nthr) {
      if (sh.group[m] < 0 || sh.state[m]) continue;
      const float* pm = sh.planes + 4 * m;
      if (fabsf(p1 - pm[1]) < ARTP_EPS && fabsf(p3 - pm[3]) < ARTP_EPS && fabsf(p0 - pm[0]) < ARTP_EPS &&
          fabsf(p2 - pm[2]) < ARTP_EPS) {
        sh.state[m] = 1;
        sh.group[m] = k;
      }
    }
    if (tid == 0) sh.state[k] = 1;
    __syncthreads();
  }
  }
groups_done:
  // per-group plane contacts vs member triangles (heightfield.cpp:1573-1617), evaluated per member
  hit = 0;
  for (int m = tid; m < T; m += nthr) {
    const int g = sh.group[m];
    if (g < 0) continue;
    float cx[4], cz[4];
    const int nc = box_plane(b, sh.planes + 4 * g, 4, cx, cz);
    if (nc == 0) continue;
    const int t = m >> 1;
    const int cxi = t / nCZ, czi = t - cxi * nCZ;
    const bool isUp = (m & 1) == 0;
    const int tcx = b.x0 + cxi + (isUp ? 0 : 1), tcz = b.z0 + czi + (isUp ? 0 : 1);
    for (int i = 0; i < nc; ++i) hit |= on_tri(f, isUp, tcx, tcz, cx[i], cz[i]);
  }
  return __syncthreads_or(hit) ? R_HIT : R_FREE;
}

constexpr int kBlockStageThreads = 256;   // threads per deferred box in the grouping stage
__global__ void __launch_bounds__(kBlockStageThreads)
box_items_block_kernel(const Checker c, const Work w, const BoxRec* __restrict__ recs, const BoxRec* __restrict__ recs_f,
                       const uint32_t* __restrict__ defer_count, const uint32_t* __restrict__ defer_list, int T_cap,
                       uint32_t* __restrict__ overflow) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[kBlockStageThreads / 32];
  __shared__ int s_next;
  BlockShared sh;
  sh.T_cap = T_cap;
  sh.planes = reinterpret_cast<float*>(smem_raw);
  sh.group = reinterpret_cast<int*>(smem_raw + (size_t)T_cap * 16);
  sh.state = reinterpret_cast<uint8_t*>(smem_raw + (size_t)T_cap * 20);
  const uint32_t count = *defer_count;
  for (uint32_t q = blockIdx.x; q < count; q += gridDim.x) {
    const uint32_t e = defer_list[q];     // bit 31: record of the reach-box queue
    const BoxRec r = (e & 0x80000000u) ? recs_f[e & 0x7fffffffu] : recs[e];
    const uint32_t slot = r.item;
    const bool foot = (r.flags & 7) != 0;
    BoxCtx b;
    rec_to_ctx(c, r, b);
    const int res = box_collide_block(foot ? c.f[1] : c.f[0], b, sh, red, &s_next);
    if (threadIdx.x == 0) {
      // a zone that does not fit the plane store cannot be decided: fail closed (item invalid) and raise the sticky
      // error word (mapped host memory; the host returns ARTP_E_LIMIT from the call / artp_poll_error)
      if (res == R_DEFER) { *(volatile uint32_t*)overflow = 1u; w.valid[slot] = 0; }
      else if ((!foot && res == R_HIT) || (foot && res == R_FREE)) w.valid[slot] = 0;
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
// Latency path: the planner's one-state-at-a-time isValid calls (ompl::base::StateValidityChecker::isValid). One launch
// does everything for up to kSmallBatch poses -- the states travel in the kernel parameter block and the verdicts are
// written straight to mapped host memory, so a call is one launch + one stream synchronise instead of memset, H2D,
// three launches and D2H. One CTA per pose: threads 0..4 classify the five boxes, warps 0..4 run the warp-stage
// decision of the undecided ones, the whole CTA handles a deferred box with the grouping-stage code.
// -------------------------------------------------------------------------------------------------
constexpr int kSmallBatch = 64;
struct SmallBatch {
  double s[kSmallBatch][7];
};

__global__ void __launch_bounds__(256)
pose_small_kernel(const Checker c, const SmallBatch sb, uint8_t* __restrict__ out, int T_cap, uint32_t* __restrict__ overflow,
                  int force_all, int steps) {   // steps < 0: CTA b checks sb.s[b]; else edges: see below
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[8];
  __shared__ int s_next;
  __shared__ BoxCtx s_box[5];
  __shared__ uint32_t s_fl[5];
  __shared__ int s_res[5];          // per box after classify: -1 undecided, R_FREE / R_HIT decided, kBoxOutside
  __shared__ int s_res2[5];         // warp-stage result of an undecided box (separate: other warps may still read s_res)
  __shared__ WarpScratch s_ws[5];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < 5) {
    double st[7];
    if (steps < 0) {
#pragma unroll
      for (int i = 0; i < 7; ++i) st[i] = sb.s[blockIdx.x][i];
    } else {
      // edge e = (s1 = sb.s[2e], s2 = sb.s[2e+1]); CTA e*(steps+1)+j checks s2 (j == 0) or interpolate(s1, s2, j/(steps+1))
      const uint32_t per = (uint32_t)steps + 1u, e = blockIdx.x / per, j = blockIdx.x - e * per;
      double a[7], bb[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) { a[i] = sb.s[2 * e][i]; bb[i] = sb.s[2 * e + 1][i]; }
      if (j == 0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) st[i] = bb[i];
      } else {
        se3_interpolate(a, bb, (double)j / (double)per, st);
      }
    }
    float t[3], R[9], Rb[9], R1[9];
    pose3_from_se3(st, t, R);
#pragma unroll
    for (int i = 0; i < 9; ++i) Rb[i] = R[i];
    orthogonalize_r(Rb);
#pragma unroll
    for (int j = 0; j < 3; ++j) { R1[j] = -Rb[j]; R1[3 + j] = Rb[6 + j]; R1[6 + j] = Rb[3 + j]; }
    BoxCtx b;
    uint32_t fl = 0;
    const int r = classify_box(c, R, R1, t, tid, force_all, 3, b, fl);
    s_res[tid] = r;
    if (r == -1) { s_box[tid] = b; s_fl[tid] = fl; }
  }
  __syncthreads();
  // item verdict from the decided boxes (validity_checker.cpp:39-45 and the outside-map rules)
  bool valid = true;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int r = s_res[k];
    if (r == kBoxOutsideWindow) { if (tid == 0) *(volatile uint32_t*)c.err_word = 2u; valid = false; }
    else if (r == kBoxOutside) { if (k > 0 && c.unknown_untraversable) valid = false; }
    else if (k == 0) { if (r == R_HIT) valid = false; }
    else if (r == R_FREE) valid = false;
  }
  if (valid && !force_all && wid < 5 && s_res[wid] == -1) {
    const bool foot = wid > 0;
    const Field& fw = foot ? c.f[1] : c.f[0];
    const ZoneView zv{fw.H + (size_t)s_box[wid].z0 * fw.pitch + s_box[wid].x0, fw.pitch};   // straight from the heightfield
    const int res = box_collide_warp<false>(fw, s_box[wid], zv, s_ws[wid], lane, c.cell_margin,
                                            (s_fl[wid] & REC_NEEDS_REDUCE) != 0, (s_fl[wid] & REC_ALLFINITE) != 0,
                                            (s_fl[wid] & REC_MERGEFREE) != 0);
    if (lane == 0) s_res2[wid] = res;
  }
  __syncthreads();
  if (valid) {
    BlockShared sh;
    sh.T_cap = T_cap;
    sh.planes = reinterpret_cast<float*>(smem_raw);
    sh.group = reinterpret_cast<int*>(smem_raw + (size_t)T_cap * 16);
    sh.state = reinterpret_cast<uint8_t*>(smem_raw + (size_t)T_cap * 20);
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
      int r = s_res[k];                       // block-uniform
      if (r == -1 && !force_all) r = s_res2[k];
      if (r == -1 || r == R_DEFER) {          // -1 only in force_all mode
        const bool foot = k > 0;
        r = box_collide_block(foot ? c.f[1] : c.f[0], s_box[k], sh, red, &s_next);
        if (r == R_DEFER) { if (tid == 0) *(volatile uint32_t*)overflow = 1u; valid = false; }   // fail closed
        __syncthreads();
      }
      if (r == kBoxOutside || r == R_DEFER) continue;
      if ((k == 0 && r == R_HIT) || (k > 0 && r == R_FREE)) valid = false;
    }
  }
  if (tid == 0) out[blockIdx.x] = valid ? 1 : 0;
}

}  // namespace artp
