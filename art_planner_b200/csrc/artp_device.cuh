// art_planner_b200/csrc/artp_device.cuh
// Exact fp32 building blocks of the box-vs-heightfield decision, shared by the warp kernel and the
// block-level grouping kernel. Arithmetic contract: IEEE fp32, round-to-nearest,
// NO FMA contraction, left-to-right association exactly as the reference's ODE source writes it.
// This translation unit is compiled with -fmad=false and default -prec-div/-prec-sqrt/-ftz=false;
// 1/sqrt is spelled as two correctly rounded operations (ode/include/ode/common.h:285).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#define ARTP_EPS 1.1920928955078125e-07f  // dEpsilon = FLT_EPSILON (ode/ode/src/common.h:42)

namespace artp {

// One heightfield layer as ODE sees it (dxHeightfieldData::SetData, ode/ode/src/heightfield.cpp:130-169).
constexpr int kMaxLevel = 6;   // range tables for windows up to 64 x 64 vertices

struct Field {
  const float* H;  // H[x + z*pitch] = layer(x, nz-1-z): column-reversed copy (height_map_box_checker.cpp:44)
  int nx, nz;      // m_nWidthSamples (rows), m_nDepthSamples (cols)
  int pitch;       // row stride in floats (multiple of 4 -> 16 B aligned rows)
  // Range tables (exact, idempotent reductions): level k holds, for every (x,z), the reduction over the
  // 2^k x 2^k vertex window starting there: T[k][x + z*pitch] = (max h, min over finite h or +inf),
  // NF[k], entry x + z*pitch (bit-packed, see window_flags): bit 0 = the window holds a non-finite height, bit 1 = a cell starting in the window has a
  // triangle whose plane matches (within eps) the plane of another triangle of the map. Built at artp_set_map, k = 1..kmax.
  const float2* T[kMaxLevel + 1];
  const uint32_t* NF[kMaxLevel + 1];   // 2 flag bits per entry, 16 entries per word: 32x smaller than the (max, min) tables, cache resident
  int kmax;
  float W, D, hW, hD, sW, sD, asp, iW, iD;
  float px, py;    // heightfield body position (float casts of the map centre)
  // Map window (artp_set_map_window): only vertices x in [x_lo, x_hi] are stored (H, T, NF are shifted by -x_lo so that
  // global indices keep working); nx and all geometry are those of the full map. Whole map: 0, nx - 1.
  int x_lo, x_hi;
};

struct Checker {
  Field f[2];             // 0: `elevation` (torso), 1: `elevation_masked` (feet)
  float side[2][3];       // torso box, reach box
  float torso_off[3];     // (off.x, off.y, float(off.z - feet.off.z)), validity_checker.cpp:41-43
  float feet_ox, feet_oy;
  int unknown_untraversable;
  double Lx, Ly, cx, cy;  // grid_map length / position (doubles) for isInside
  float cell_margin;      // candidate-cell search margin in cells (plane stage)
  uint32_t* err_word;     // sticky error word (mapped host memory): bit 0 plane-store overflow, bit 1 box outside the map window
  // extent of the reach-box queue's TMA tile in floats (artp_tiles.cuh; 0: no such queue). The tile starts at column
  // x0 & ~3, so a zone may be at most reach_tw - 3 wide.
  int reach_tw, reach_th;
};

// Box pose in heightfield space + AABB + zone.
struct BoxCtx {
  float R1[9];   // rows: -R.row0, R.row2, R.row1 of the orthogonalised box rotation (3x3, row-major)
  float P[3];    // box centre in heightfield space
  float side[3];
  float minB, maxB;
  int x0, x1, z0, z1;
};

// flags of range-table entry idx (bit 0 non-finite, bit 1 mergeable); idx is LOCAL to the handle's map window
// (global entry - x_lo: the packed words cannot be shifted by a pointer offset the way the float2 tables are)
__device__ __forceinline__ int window_flags(const uint32_t* __restrict__ nf, size_t idx) {
  return (int)((__ldg(nf + (idx >> 4)) >> ((idx & 15) * 2)) & 3u);
}

// nextafterf(x, -inf) / nextafterf(x, +inf) for finite x (dNextAfter, ode/include/ode/common.h:296), as integer ops.
__device__ __forceinline__ float next_down(float x) {
  if (x == 0.0f) return __int_as_float(0x80000001);
  return __int_as_float(__float_as_int(x) + ((x > 0.0f) ? -1 : 1));
}
__device__ __forceinline__ float next_up(float x) {
  if (x == 0.0f) return __int_as_float(0x00000001);
  return __int_as_float(__float_as_int(x) + ((x > 0.0f) ? 1 : -1));
}

// 1.0f / sqrtf(x) as two correctly rounded operations; __frcp_rn(y) is the correctly rounded 1/y, i.e. bit-identical to
// __fdiv_rn(1.0f, y), in fewer instructions.
__device__ __forceinline__ float rsqrt_exact(float x) { return __frcp_rn(__fsqrt_rn(x)); }

// dxSafeNormalize3, ode/ode/src/odemath.cpp:95-161: scale by the largest component m, then u = lower-index other
// component / m, v = higher-index other component / m, l = 1 / sqrt(1 + u*u + v*v). Written without a branch per
// largest-component case (which component is largest depends on the yaw, so a warp would run all three copies): the
// operands are selected, the arithmetic is the same sequence of correctly rounded operations.
__device__ __forceinline__ void safe_normalize3(float& a0, float& a1, float& a2) {
  const float b0 = fabsf(a0), b1 = fabsf(a1), b2 = fabsf(a2);
  int idx;
  if (b1 > b0) idx = (b2 > b1) ? 2 : 1;
  else if (b2 > b0) idx = 2;
  else { if (!(b0 > 0.0f)) return; idx = 0; }
  const float bm = idx == 0 ? b0 : (idx == 1 ? b1 : b2);
  const float am = idx == 0 ? a0 : (idx == 1 ? a1 : a2);
  const float au = idx == 0 ? a1 : a0, av = idx == 2 ? a1 : a2;
  const float r = __fdiv_rn(1.0f, bm);
  const float u = au * r, v = av * r;
  const float l = rsqrt_exact(1.0f + u * u + v * v);
  const float nu = u * l, nv = v * l, nm = copysignf(l, am);
  a0 = idx == 0 ? nm : nu;
  a1 = idx == 1 ? nm : (idx == 0 ? nu : nv);
  a2 = idx == 2 ? nm : nv;
}

// dBodySetRotation -> dxOrthogonalizeR (ode/ode/src/ode.cpp:358-374, odemath.cpp:260-313) on the 3x3
// row-major m; quirk kept: with proj != 0 the Gram-Schmidt row goes to a temporary, stored row 1 untouched.
__device__ __forceinline__ void orthogonalize_r(float m[9]) {
  if (!(m[0] != 0.0f || m[1] != 0.0f || m[2] != 0.0f)) return;
  const float n0 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
  float r0 = m[3], r1 = m[4], r2 = m[5];
  const float proj = m[0] * m[3] + m[1] * m[4] + m[2] * m[5];
  const bool tmp = (proj != 0.0f);
  if (tmp) {
    const float pd = __fdiv_rn(proj, n0);
    r0 = m[3] - pd * m[0];
    r1 = m[4] - pd * m[1];
    r2 = m[5] - pd * m[2];
  }
  if (!(r0 != 0.0f || r1 != 0.0f || r2 != 0.0f)) return;
  if (n0 != 1.0f) safe_normalize3(m[0], m[1], m[2]);
  const float n1 = r0 * r0 + r1 * r1 + r2 * r2;
  if (n1 != 1.0f) safe_normalize3(r0, r1, r2);
  if (!tmp) { m[3] = r0; m[4] = r1; m[5] = r2; }   // alias case: row 1 normalised in place
  m[6] = m[1] * r2 - m[2] * r1;                      // dCalcVectorCross3(row2, row0, row1')
  m[7] = m[2] * r0 - m[0] * r2;
  m[8] = m[0] * r1 - m[1] * r0;
}

// Eigen::Quaternion<float>(w,x,y,z).toRotationMatrix() on double->float casts (utils.h:25-38).
__device__ __forceinline__ void pose3_from_se3(const double* __restrict__ s, float t[3], float R[9]) {
  t[0] = (float)s[0]; t[1] = (float)s[1]; t[2] = (float)s[2];
  const float x = (float)s[3], y = (float)s[4], z = (float)s[5], w = (float)s[6];
  const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0f - (tyy + tzz); R[1] = txy - twz;          R[2] = txz + twy;
  R[3] = txy + twz;          R[4] = 1.0f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;          R[7] = tyz + twx;          R[8] = 1.0f - (txx + tyy);
}

// (pose * Pose3FromXYZ(o)).translation(): R*o + t, 3-term dot reduced as a0 + (a1 + a2) (Eigen redux).
__device__ __forceinline__ void compose_translation(const float R[9], const float t[3], float o0, float o1,
                                                    float o2, float out[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a0 = R[3 * i] * o0, a1 = R[3 * i + 1] * o1, a2 = R[3 * i + 2] * o2;
    out[i] = (a0 + (a1 + a2)) + t[i];
  }
}

// grid_map checkIfPositionWithinMap (double).
__device__ __forceinline__ bool is_inside(const Checker& c, float px, float py) {
  const double tx = -(((double)px - c.cx) - 0.5 * c.Lx);
  const double ty = -(((double)py - c.cy) - 0.5 * c.Ly);
  return tx >= 0.0 && ty >= 0.0 && tx < c.Lx && ty < c.Ly;
}

// dCollideHeightfield prologue (ode/ode/src/heightfield.cpp:1841-1892) + dxBox::computeAABB (box.cpp:60-77).
// Rb = orthogonalised box rotation (3x3). Returns false if rejected by the AABB-vs-extent test.
__device__ __forceinline__ bool box_setup(const Field& f, const float side[3], const float origin[3],
                                          const float Rb[9], BoxCtx& b) {
  // pos1 = Rf^T (origin - field pos), Rf = rows [-1,0,0],[0,0,1],[0,1,-0]; R1 = Rf^T * Rb.
  // Literal products with the 0/+-1 constants only change signs of zeros, which no comparison below sees.
  const float d0 = origin[0] - f.px, d1 = origin[1] - f.py, d2 = origin[2] - 0.0f;
  b.P[0] = -d0 + f.hW;
  b.P[1] = d2;
  b.P[2] = d1 + f.hD;
#pragma unroll
  for (int j = 0; j < 3; ++j) { b.R1[j] = -Rb[j]; b.R1[3 + j] = Rb[6 + j]; b.R1[6 + j] = Rb[3 + j]; }
  b.side[0] = side[0]; b.side[1] = side[1]; b.side[2] = side[2];
  const float xr = 0.5f * (fabsf(b.R1[0] * side[0]) + fabsf(b.R1[1] * side[1]) + fabsf(b.R1[2] * side[2]));
  const float yr = 0.5f * (fabsf(b.R1[3] * side[0]) + fabsf(b.R1[4] * side[1]) + fabsf(b.R1[5] * side[2]));
  const float zr = 0.5f * (fabsf(b.R1[6] * side[0]) + fabsf(b.R1[7] * side[1]) + fabsf(b.R1[8] * side[2]));
  const float a0 = b.P[0] - xr, a1 = b.P[0] + xr, a4 = b.P[2] - zr, a5 = b.P[2] + zr;
  b.minB = b.P[1] - yr;
  b.maxB = b.P[1] + yr;
  if (a0 > f.W || a4 > f.D) return false;
  if (a1 < 0.0f || a5 < 0.0f) return false;
  int nMinX = (int)floorf(next_down(a0 * f.iW));
  int nMaxX = (int)ceilf(next_up(a1 * f.iW));
  int nMinZ = (int)floorf(next_down(a4 * f.iD));
  int nMaxZ = (int)ceilf(next_up(a5 * f.iD));
  b.x0 = max(nMinX, 0);
  b.x1 = min(nMaxX, f.nx - 1);
  b.z0 = max(nMinZ, 0);
  b.z1 = min(nMaxZ, f.nz - 1);
  return true;
}

// dGeomBoxPointDepth(v) > dEpsilon (ode/ode/src/box.cpp:109-173): depth > eps  <=>  all six face
// distances > eps (outside => depth <= 0; inside => depth = min of the six).
__device__ __forceinline__ bool vertex_inside(const BoxCtx& b, float vx, float vy, float vz) {
  const float p0 = vx - b.P[0], p1 = vy - b.P[1], p2 = vz - b.P[2];
  const float q0 = b.R1[0] * p0 + b.R1[3] * p1 + b.R1[6] * p2;   // dMultiply1_331
  const float q1 = b.R1[1] * p0 + b.R1[4] * p1 + b.R1[7] * p2;
  const float q2 = b.R1[2] * p0 + b.R1[5] * p1 + b.R1[8] * p2;
  const float s0 = b.side[0] * 0.5f, s1 = b.side[1] * 0.5f, s2 = b.side[2] * 0.5f;
  return (s0 - q0 > ARTP_EPS) && (s0 + q0 > ARTP_EPS) && (s1 - q1 > ARTP_EPS) && (s1 + q1 > ARTP_EPS) &&
         (s2 - q2 > ARTP_EPS) && (s2 + q2 > ARTP_EPS);
}

// Plane of a heightfield triangle (ode/ode/src/heightfield.cpp:1474-1501).
// v0 = vertices[0], v1 = vertices[1], v2 = vertices[2]; Up: (A,B,C), Down: (D,B,C).
__device__ __forceinline__ void tri_plane(bool isUp, float v0x, float v0y, float v0z, float v1x, float v1y,
                                          float v1z, float v2x, float v2y, float v2z, float pl[4]) {
  const float e1x = v2x - v0x, e1y = v2y - v0y, e1z = v2z - v0z;   // Edge1 = v2 - v0
  const float e2x = v1x - v0x, e2y = v1y - v0y, e2z = v1z - v0z;   // Edge2 = v1 - v0
  float ax, ay, az, bx, by, bz;
  if (isUp) { ax = e1x; ay = e1y; az = e1z; bx = e2x; by = e2y; bz = e2z; }
  else      { ax = e2x; ay = e2y; az = e2z; bx = e1x; by = e1y; bz = e1z; }
  float c0 = ay * bz - az * by;
  float c1 = az * bx - ax * bz;
  float c2 = ax * by - ay * bx;
  const float inv = rsqrt_exact(c0 * c0 + c1 * c1 + c2 * c2);
  c0 *= inv; c1 *= inv; c2 *= inv;
  pl[0] = c0; pl[1] = c1; pl[2] = c2;
  pl[3] = c0 * v0x + c1 * v0y + c2 * v0z;
}

__device__ __forceinline__ bool plane_match(const float a[4], const float b[4]) {   // heightfield.cpp:1541-1546
  return fabsf(a[1] - b[1]) < ARTP_EPS && fabsf(a[3] - b[3]) < ARTP_EPS && fabsf(a[0] - b[0]) < ARTP_EPS &&
         fabsf(a[2] - b[2]) < ARTP_EPS;
}

// dCollideBoxPlane (ode/ode/src/box.cpp:745-878) with maxc clamped to `maxc` (1 or 4): contact positions.
// cx/cz receive the X and Z of each contact (Y is never used by IsOnHeightfield2).
__device__ __forceinline__ int box_plane(const BoxCtx& b, const float n[4], int maxc, float cx[4], float cz[4]) {
  const float* R = b.R1;
  const float Q1 = n[0] * R[0] + n[1] * R[3] + n[2] * R[6];
  const float Q2 = n[0] * R[1] + n[1] * R[4] + n[2] * R[7];
  const float Q3 = n[0] * R[2] + n[1] * R[5] + n[2] * R[8];
  const float A1 = b.side[0] * Q1, A2 = b.side[1] * Q2, A3 = b.side[2] * Q3;
  const float B1 = fabsf(A1), B2 = fabsf(A2), B3 = fabsf(A3);
  const float depth = n[3] + 0.5f * (B1 + B2 + B3) - (n[0] * b.P[0] + n[1] * b.P[1] + n[2] * b.P[2]);
  if (depth < 0.0f) return 0;
  float px = b.P[0], pz = b.P[2];
  {
    const float h0 = 0.5f * b.side[0], h1 = 0.5f * b.side[1], h2 = 0.5f * b.side[2];
    if (A1 > 0.0f) { px -= h0 * R[0]; pz -= h0 * R[6]; } else { px += h0 * R[0]; pz += h0 * R[6]; }
    if (A2 > 0.0f) { px -= h1 * R[1]; pz -= h1 * R[7]; } else { px += h1 * R[1]; pz += h1 * R[7]; }
    if (A3 > 0.0f) { px -= h2 * R[2]; pz -= h2 * R[8]; } else { px += h2 * R[2]; pz += h2 * R[8]; }
  }
  cx[0] = px; cz[0] = pz;
  int ret = 1;
  if (maxc == 1) return ret;
  int first, second;
  if (B1 < B2) {
    if (B3 < B1) { first = 2; second = 0; }
    else         { first = 0; second = (B2 < B3) ? 1 : 2; }
  } else {
    if (B3 < B2) { first = 2; second = 1; }
    else         { first = 1; second = (B1 < B3) ? 0 : 2; }
  }
  float d1 = 0.0f, d2 = 0.0f;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int j = (c == 0) ? first : second;
    const float Bj = (j == 0) ? B1 : (j == 1 ? B2 : B3);
    const float Aj = (j == 0) ? A1 : (j == 1 ? A2 : A3);
    const float sj = (j == 0) ? b.side[0] : (j == 1 ? b.side[1] : b.side[2]);
    const float rx = (j == 0) ? R[0] : (j == 1 ? R[1] : R[2]);
    const float rz = (j == 0) ? R[6] : (j == 1 ? R[7] : R[8]);
    if (depth - Bj < 0.0f) break;
    if (Aj > 0.0f) { cx[ret] = px + sj * rx; cz[ret] = pz + sj * rz; }
    else           { cx[ret] = px - sj * rx; cz[ret] = pz - sj * rz; }
    if (c == 0) d1 = depth - Bj; else d2 = depth - Bj;
    ret++;
  }
  if (ret == 3) {
    const float d4 = d1 + d2 - depth;
    if (d4 > 0.0f) {
      cx[3] = cx[1] + cx[2] - px;
      cz[3] = cz[1] + cz[2] - pz;
      ret++;
    }
  }
  return ret;
}

// dxHeightfieldData::IsOnHeightfield2 (ode/ode/src/heightfield.cpp:264-321). (cx,cz): integer coords of the
// triangle's first vertex (A for Up, D for Down).
__device__ __forceinline__ bool on_tri(const Field& f, bool isUp, int cx, int cz, float X, float Z) {
  // Up:   MinX = cx*sW,     MaxX = (cx+1)*sW, ... inside && (MaxZ - Z) >  (X - MinX) * asp
  // Down: MinX = (cx-1)*sW, MaxX = cx*sW,     ... inside && (MaxZ - Z) <= (X - MinX) * asp
  // written without an isUp branch (same products, same comparisons) so mixed warps do not diverge.
  const int lx = isUp ? cx : cx - 1, lz = isUp ? cz : cz - 1;
  const float MinX = lx * f.sW, MaxX = (lx + 1) * f.sW, MinZ = lz * f.sD, MaxZ = (lz + 1) * f.sD;
  const bool inside = (X >= MinX) && (X < MaxX) && (Z >= MinZ) && (Z < MaxZ);
  const bool above = (MaxZ - Z) > (X - MinX) * f.asp;
  return inside && (isUp ? above : !above);
}

__device__ __forceinline__ bool finitef(float h) { return fabsf(h) < CUDART_INF_F; }   // no NaN by contract

}  // namespace artp
