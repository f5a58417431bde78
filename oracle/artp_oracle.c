/*
 * oracle/artp_oracle.c -- TEST INFRASTRUCTURE ONLY: plain-C restatement ("port") of the reference's
 * box-vs-heightfield decision procedure, i.e. dCollide(box, heightfield, flags=1) != 0 as art_planner's
 * modified, vendored ODE 0.16.1 computes it in single precision (dReal = float, no FMA).
 *
 * Pinning: this file is checked bit-for-bit against the reference's own compiled ODE
 * (oracle/_ref/liborc_ref.so, built by oracle/Makefile from /root/reference/ode) by
 * tests/test_oracle_vs_reference.py (runs where /root/reference exists) and against the committed
 * golden masks in tests/golden/ that oracle/make_golden.py generated from that library.
 *
 * Build: gcc -O2 -ffp-contract=off (no -march, no fast-math): x86-64 SSE scalar fp32, like the reference.
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "artp_oracle.h"
#include "artp_wrappers.h"

#define ORC_EPS FLT_EPSILON /* dEpsilon, ode/ode/src/common.h:42 */

typedef struct orc_field {
  int nx, nz;            /* m_nWidthSamples = rows, m_nDepthSamples = cols */
  float W, D;            /* m_fWidth, m_fDepth */
  float hW, hD;          /* m_fHalfWidth, m_fHalfDepth */
  float sW, sD;          /* m_fSampleWidth, m_fSampleDepth */
  float asp;             /* m_fSampleZXAspect */
  float iW, iD;          /* m_fInvSampleWidth, m_fInvSampleDepth */
  float px, py;          /* heightfield body position (float), z = 0 */
  float* H;              /* H[x + z*nx] = layer(x, nz-1-z)  (rowwise().reverse(), col-major) */
} orc_field;

struct orc_handle {
  orc_params p;
  orc_geom g;
  orc_field f[2];        /* 0: elevation (torso), 1: elevation_masked (feet) */
  float side[2][3];      /* box sides as float (HeightMapBoxChecker ctor takes floats) */
};

/* ------------------------------------------------------------------------------------------------
 * dxSafeNormalize3, ode/ode/src/odemath.cpp:95-161
 * ---------------------------------------------------------------------------------------------- */
static int safe_normalize3(float a[3]) {
  const float abs_a0 = fabsf(a[0]), abs_a1 = fabsf(a[1]), abs_a2 = fabsf(a[2]);
  int idx;
  if (abs_a1 > abs_a0) {
    idx = (abs_a2 > abs_a1) ? 2 : 1;
  } else if (abs_a2 > abs_a0) {
    idx = 2;
  } else {
    if (!(abs_a0 > 0.0f)) return 0;
    idx = 0;
  }
  if (idx == 0) {
    const float r = 1.0f / abs_a0;
    const float a1 = a[1] * r, a2 = a[2] * r;
    const float l = 1.0f / sqrtf(1.0f + a1 * a1 + a2 * a2);
    a[1] = a1 * l; a[2] = a2 * l; a[0] = copysignf(l, a[0]);
  } else if (idx == 1) {
    const float r = 1.0f / abs_a1;
    const float a0 = a[0] * r, a2 = a[2] * r;
    const float l = 1.0f / sqrtf(1.0f + a0 * a0 + a2 * a2);
    a[0] = a0 * l; a[2] = a2 * l; a[1] = copysignf(l, a[1]);
  } else {
    const float r = 1.0f / abs_a2;
    const float a0 = a[0] * r, a1 = a[1] * r;
    const float l = 1.0f / sqrtf(1.0f + a0 * a0 + a1 * a1);
    a[0] = a0 * l; a[1] = a1 * l; a[2] = copysignf(l, a[2]);
  }
  return 1;
}

/* dxCouldBeNormalized3: any non-zero component (odemath.cpp:60-77). */
static int could_be_normalized3(const float a[3]) {
  return a[0] != 0.0f || a[1] != 0.0f || a[2] != 0.0f;
}

/* ------------------------------------------------------------------------------------------------
 * dBodySetRotation -> dxOrthogonalizeR, ode/ode/src/ode.cpp:358-374, odemath.cpp:260-313.
 * m is the 3x4 row-major dMatrix3. Note the quirk: with proj != 0 the Gram-Schmidt result goes to a
 * temporary, the stored row 1 is left untouched.
 * ---------------------------------------------------------------------------------------------- */
static int orthogonalize_r(float m[12]) {
  if (!could_be_normalized3(m)) return 0;
  const float n0 = m[0] * m[0] + m[1] * m[1] + m[2] * m[2];
  float row2_store[3];
  float* row2 = m + 4;
  const float proj = m[0] * m[4] + m[1] * m[5] + m[2] * m[6];
  if (proj != 0) {
    const float proj_div_n0 = proj / n0;
    row2_store[0] = m[4] - proj_div_n0 * m[0];
    row2_store[1] = m[5] - proj_div_n0 * m[1];
    row2_store[2] = m[6] - proj_div_n0 * m[2];
    row2 = row2_store;
  }
  if (!could_be_normalized3(row2)) return 0;
  if (n0 != 1.0f) safe_normalize3(m);
  const float n1 = row2[0] * row2[0] + row2[1] * row2[1] + row2[2] * row2[2];
  if (n1 != 1.0f) safe_normalize3(row2);
  /* dCalcVectorCross3(row[2], row[0], row2), ode/include/ode/odemath.h:234-244 */
  const float r0 = m[1] * row2[2] - m[2] * row2[1];
  const float r1 = m[2] * row2[0] - m[0] * row2[2];
  const float r2 = m[0] * row2[1] - m[1] * row2[0];
  m[8] = r0; m[9] = r1; m[10] = r2;
  m[3] = m[7] = m[11] = 0.0f;
  return 1;
}

/* dGeomBoxPointDepth, ode/ode/src/box.cpp:109-173 (R = box rotation 3x4, pos = box centre). */
static float box_point_depth(const float R[12], const float pos[3], const float side[3],
                             float x, float y, float z) {
  float p[3], q[3], dist[6];
  p[0] = x - pos[0]; p[1] = y - pos[1]; p[2] = z - pos[2];
  /* dMultiply1_331: q_i = R[i]*p0 + R[4+i]*p1 + R[8+i]*p2 */
  q[0] = R[0] * p[0] + R[4] * p[1] + R[8] * p[2];
  q[1] = R[1] * p[0] + R[5] * p[1] + R[9] * p[2];
  q[2] = R[2] * p[0] + R[6] * p[1] + R[10] * p[2];
  int inside = 1;
  for (int i = 0; i < 3; ++i) {
    const float s = side[i] * 0.5f;
    dist[i] = s - q[i];
    dist[i + 3] = s + q[i];
    if (dist[i] < 0 || dist[i + 3] < 0) inside = 0;
  }
  if (inside) {
    float smallest = (float)(unsigned)-1;
    for (int i = 0; i < 6; ++i) if (dist[i] < smallest) smallest = dist[i];
    return smallest;
  }
  float largest = 0;
  for (int i = 0; i < 6; ++i) if (dist[i] > largest) largest = dist[i];
  return -largest;
}

/* dCollideBoxPlane, ode/ode/src/box.cpp:745-878. n[4] = plane (normal, d). Returns contact count and
 * the contact positions (up to maxc <= 4). */
static int box_plane(const float R[12], const float pos[3], const float side[3], const float n[4],
                     int maxc, float cpos[4][3]) {
  const float Q1 = n[0] * R[0] + n[1] * R[4] + n[2] * R[8];   /* dCalcVectorDot3_14(n, R+0) */
  const float Q2 = n[0] * R[1] + n[1] * R[5] + n[2] * R[9];
  const float Q3 = n[0] * R[2] + n[1] * R[6] + n[2] * R[10];
  const float A1 = side[0] * Q1, A2 = side[1] * Q2, A3 = side[2] * Q3;
  const float B1 = fabsf(A1), B2 = fabsf(A2), B3 = fabsf(A3);
  const float A[3] = {A1, A2, A3};
  const float B[3] = {B1, B2, B3};
  const float depth = n[3] + 0.5f * (B1 + B2 + B3) - (n[0] * pos[0] + n[1] * pos[1] + n[2] * pos[2]);
  if (depth < 0) return 0;
  if (maxc > 4) maxc = 4;
  float p[3] = {pos[0], pos[1], pos[2]};
  for (int i = 0; i < 3; ++i) {
    if (A[i] > 0) {
      p[0] -= 0.5f * side[i] * R[0 + i]; p[1] -= 0.5f * side[i] * R[4 + i]; p[2] -= 0.5f * side[i] * R[8 + i];
    } else {
      p[0] += 0.5f * side[i] * R[0 + i]; p[1] += 0.5f * side[i] * R[4 + i]; p[2] += 0.5f * side[i] * R[8 + i];
    }
  }
  cpos[0][0] = p[0]; cpos[0][1] = p[1]; cpos[0][2] = p[2];
  int ret = 1;
  float cdepth[4];
  cdepth[0] = depth;
  if (maxc == 1) return ret;
  /* second and third contact: along the two sides with the smallest projected length. */
  int first, second;
  if (B1 < B2) {
    if (B3 < B1) { first = 2; second = (B1 < B2) ? 0 : 1; }
    else         { first = 0; second = (B2 < B3) ? 1 : 2; }
  } else {
    if (B3 < B2) { first = 2; second = (B1 < B2) ? 0 : 1; }
    else         { first = 1; second = (B1 < B3) ? 0 : 2; }
  }
  const int order[2] = {first, second};
  for (int c = 0; c < 2; ++c) {
    const int j = order[c];
    if (depth - B[j] < 0) goto done;
    if (A[j] > 0) {
      cpos[ret][0] = p[0] + side[j] * R[0 + j]; cpos[ret][1] = p[1] + side[j] * R[4 + j]; cpos[ret][2] = p[2] + side[j] * R[8 + j];
    } else {
      cpos[ret][0] = p[0] - side[j] * R[0 + j]; cpos[ret][1] = p[1] - side[j] * R[4 + j]; cpos[ret][2] = p[2] - side[j] * R[8 + j];
    }
    cdepth[ret] = depth - B[j];
    ret++;
    if (maxc == 2) goto done;
  }
done:
  if (maxc == 4 && ret == 3) {
    const float d4 = cdepth[1] + cdepth[2] - depth;
    if (d4 > 0) {
      cpos[3][0] = cpos[1][0] + cpos[2][0] - p[0];
      cpos[3][1] = cpos[1][1] + cpos[2][1] - p[1];
      cpos[3][2] = cpos[1][2] + cpos[2][2] - p[2];
      ret++;
    }
  }
  return ret;
}

/* dxHeightfieldData::IsOnHeightfield2, ode/ode/src/heightfield.cpp:264-321.
 * (cx,cz) = integer coords of the triangle's first vertex; vx, vz = that vertex' position. */
static int is_on_heightfield2(const orc_field* f, int cx, int cz, float vx, float vz,
                              const float pos[3], int isABC) {
  float MaxX, MinX, MaxZ, MinZ;
  if (isABC) {
    MinX = vx;
    if (pos[0] < MinX) return 0;
    MaxX = (cx + 1) * f->sW;
    if (pos[0] >= MaxX) return 0;
    MinZ = vz;
    if (pos[2] < MinZ) return 0;
    MaxZ = (cz + 1) * f->sD;
    if (pos[2] >= MaxZ) return 0;
    return (MaxZ - pos[2]) > (pos[0] - MinX) * f->asp;
  } else {
    MaxX = vx;
    if (pos[0] >= MaxX) return 0;
    MinX = (cx - 1) * f->sW;
    if (pos[0] < MinX) return 0;
    MaxZ = vz;
    if (pos[2] >= MaxZ) return 0;
    MinZ = (cz - 1) * f->sD;
    if (pos[2] < MinZ) return 0;
    return (MaxZ - pos[2]) <= (pos[0] - MinX) * f->asp;
  }
}

typedef struct orc_tri {
  int vx[3], vz[3];      /* integer coords of vertices[0..2] */
  int isUp;
  int state;
  float plane[4];
} orc_tri;

typedef struct orc_scratch {
  orc_tri* tri; size_t tri_cap;
  int* group; size_t group_cap;
  /* diagnostics of the last call (port only): exit stage, kept triangles, plane groups */
  int stage; uint32_t num_tri, num_groups;
} orc_scratch;

enum { ORC_ST_AABB = 0, ORC_ST_ABOVE = 1, ORC_ST_UNDER = 2, ORC_ST_SPAN = 3, ORC_ST_SINGLE = 4,
       ORC_ST_VERTEX = 5, ORC_ST_PLANE = 6, ORC_ST_NONE = 7 };

/* dxHeightfield::dCollideHeightfieldZone for a box with numMaxContactsPossible = 1,
 * ode/ode/src/heightfield.cpp:973-1789. Returns 0/1. R1/P = box pose in heightfield space. */
static int collide_zone(const orc_field* f, int minX, int maxX, int minZ, int maxZ,
                        const float R1[12], const float P[3], const float side[3],
                        const float aabb[6], orc_scratch* sc) {
  const float minO2Height = aabb[2], maxO2Height = aabb[3];
  float maxY = -INFINITY, minY = INFINITY;
  int allFinite = 1;
  /* :1002-1026 zone scan */
  for (int x = minX; x <= maxX; ++x) {
    for (int z = minZ; z <= maxZ; ++z) {
      const float h = f->H[x + (size_t)z * f->nx];   /* GetHeight(x,z): (h*1)+0, :325-384 */
      maxY = (maxY > h) ? maxY : h;                   /* dMAX(maxY, h) */
      if (isfinite(h)) minY = (minY > h) ? h : minY;  /* dMIN(minY, h) */
      else allFinite = 0;
    }
  }
  sc->num_tri = sc->num_groups = 0;
  sc->stage = ORC_ST_ABOVE;
  if (minO2Height - maxY > -ORC_EPS) return 0;                                   /* :1027 above */
  sc->stage = ORC_ST_UNDER;
  if (minY - maxO2Height > -ORC_EPS) return 0;                                   /* :1032-1058 under */
  sc->stage = ORC_ST_SPAN;
  if (allFinite && minY - minO2Height > -ORC_EPS && maxO2Height - maxY > -ORC_EPS) return 1; /* :1059 */
  sc->stage = ORC_ST_SINGLE;
  if (allFinite && maxY - minY < ORC_EPS) {                                      /* :1139-1160 */
    const float pl[4] = {0, 1, 0, minY};
    float cp[4][3];
    return box_plane(R1, P, side, pl, 1, cp);
  }
  sc->stage = ORC_ST_VERTEX;
  /* :1306-1460 triangle emission with the art_planner vertex-depth early return */
  const size_t numTriMax = (size_t)(maxX - minX) * (size_t)(maxZ - minZ) * 2;
  if (sc->tri_cap < numTriMax) {
    free(sc->tri); sc->tri = (orc_tri*)malloc(sizeof(orc_tri) * (numTriMax ? numTriMax : 1)); sc->tri_cap = numTriMax;
  }
  size_t numTri = 0;
  for (int x = minX; x < maxX; ++x) {
    for (int z = minZ; z < maxZ; ++z) {
      const float hA = f->H[x + (size_t)z * f->nx];
      const float hB = f->H[(x + 1) + (size_t)z * f->nx];
      const float hC = f->H[x + (size_t)(z + 1) * f->nx];
      const float hD = f->H[(x + 1) + (size_t)(z + 1) * f->nx];
      const int fA = isfinite(hA), fB = isfinite(hB), fC = isfinite(hC), fD = isfinite(hD);
      const int cA = (hA > minO2Height) && fA, cB = (hB > minO2Height) && fB;
      const int cC = (hC > minO2Height) && fC, cD = (hD > minO2Height) && fD;
      const float xA = x * f->sW, xB = (x + 1) * f->sW;         /* :1004 */
      const float zA = z * f->sD, zC = (z + 1) * f->sD;         /* :1010 */
      if ((cA || cB || cC) && (fA && fB && fC)) {
        if (cA && box_point_depth(R1, P, side, xA, hA, zA) > ORC_EPS) return 1;
        if (cB && box_point_depth(R1, P, side, xB, hB, zA) > ORC_EPS) return 1;
        if (cC && box_point_depth(R1, P, side, xA, hC, zC) > ORC_EPS) return 1;
        orc_tri* t = &sc->tri[numTri++];
        t->vx[0] = x; t->vz[0] = z; t->vx[1] = x + 1; t->vz[1] = z; t->vx[2] = x; t->vz[2] = z + 1;
        t->isUp = 1; t->state = 0;
      }
      if ((cB || cC || cD) && (fB && fC && fD)) {
        if (cB && box_point_depth(R1, P, side, xB, hB, zA) > ORC_EPS) return 1;
        if (cC && box_point_depth(R1, P, side, xA, hC, zC) > ORC_EPS) return 1;
        if (cD && box_point_depth(R1, P, side, xB, hD, zC) > ORC_EPS) return 1;
        orc_tri* t = &sc->tri[numTri++];
        t->vx[0] = x + 1; t->vz[0] = z + 1; t->vx[1] = x + 1; t->vz[1] = z; t->vx[2] = x; t->vz[2] = z + 1;
        t->isUp = 0; t->state = 0;
      }
    }
  }
  sc->stage = ORC_ST_PLANE;
  sc->num_tri = (uint32_t)numTri;
  /* :1474-1501 plane of every kept triangle */
  for (size_t k = 0; k < numTri; ++k) {
    orc_tri* t = &sc->tri[k];
    float v[3][3];
    for (int i = 0; i < 3; ++i) {
      v[i][0] = t->vx[i] * f->sW;
      v[i][1] = f->H[t->vx[i] + (size_t)t->vz[i] * f->nx];
      v[i][2] = t->vz[i] * f->sD;
    }
    float E1[3], E2[3], c[3];
    for (int i = 0; i < 3; ++i) { E1[i] = v[2][i] - v[0][i]; E2[i] = v[1][i] - v[0][i]; }
    const float* a = t->isUp ? E1 : E2;
    const float* b = t->isUp ? E2 : E1;
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
    const float dinv = 1.0f / sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    c[0] *= dinv; c[1] *= dinv; c[2] *= dinv;
    t->plane[0] = c[0]; t->plane[1] = c[1]; t->plane[2] = c[2];
    t->plane[3] = c[0] * v[0][0] + c[1] * v[0][1] + c[2] * v[0][2];
  }
  /* :1511-1556 greedy epsilon grouping in emission order, then :1573-1617 per-group plane test.
   * (sortPlanes :1559-1560 only reorders groups and cannot change the boolean.) */
  if (sc->group_cap < numTri) {
    free(sc->group); sc->group = (int*)malloc(sizeof(int) * (numTri ? numTri : 1)); sc->group_cap = numTri;
  }
  for (size_t k = 0; k < numTri; ++k) {
    orc_tri* base = &sc->tri[k];
    if (base->state) continue;
    size_t ng = 0;
    sc->group[ng++] = (int)k;
    const float normx = base->plane[0], normy = base->plane[1], normz = base->plane[2], dist = base->plane[3];
    for (size_t m = k + 1; m < numTri; ++m) {
      orc_tri* tt = &sc->tri[m];
      if (tt->state) continue;
      if (fabsf(normy - tt->plane[1]) < ORC_EPS && fabsf(dist - tt->plane[3]) < ORC_EPS &&
          fabsf(normx - tt->plane[0]) < ORC_EPS && fabsf(normz - tt->plane[2]) < ORC_EPS) {
        sc->group[ng++] = (int)m;
        tt->state = 1;
      }
    }
    base->state = 1;
    sc->num_groups++;
    /* planeTestFlags -> HEIGHTFIELDMAXCONTACTPERCELL(10), dCollideBoxPlane clamps to 4 */
    float cp[4][3];
    const int nc = box_plane(R1, P, side, base->plane, 10, cp);
    for (int i = 0; i < nc; ++i) {
      for (size_t b = 0; b < ng; ++b) {
        const orc_tri* t = &sc->tri[sc->group[b]];
        const float vx = t->vx[0] * f->sW, vz = t->vz[0] * f->sD;
        if (is_on_heightfield2(f, t->vx[0], t->vz[0], vx, vz, cp[i], t->isUp)) return 1;
      }
    }
  }
  /* pass 2 (:1651-1719) cannot add a contact for a box: every vertex it could visit was already
   * depth-tested <= eps above; edge pass (:1721) is compiled out. */
  sc->stage = ORC_ST_NONE;
  return 0;
}

/* dCollide -> dCollideHeightfield, collision_kernel.cpp:292-338, heightfield.cpp:1791-1964,
 * with HeightMapBoxChecker::checkCollision's dBodySetPosition/dBodySetRotation in front
 * (height_map_box_checker.cpp:58-72). */
static int box_collide(const orc_field* f, const float side[3], const float origin[3],
                       const float rot12[12], orc_scratch* sc, uint32_t* zv) {
  float R[12];
  memcpy(R, rot12, sizeof(R));
  orthogonalize_r(R);
  if (zv) *zv = 0;
  /* heightfield body rotation = dRFrom2Axes(-1,0,0, 0,0,1) = rows [-1,0,0],[0,0,1],[0,1,0]
   * (height_map_box_checker.cpp:22, rotation.cpp:94-133); exact, unchanged by dxOrthogonalizeR. */
  static const float Rf[12] = {-1, 0, 0, 0, 0, 0, 1, 0, 0, 1, -0.0f, 0};   /* [10] is -0: -bx*ay + ax*by */
  float pos0[3], P[3], R1[12];
  pos0[0] = origin[0] - f->px; pos0[1] = origin[1] - f->py; pos0[2] = origin[2] - 0.0f;
  /* dMultiply1_331(pos1, Rf, pos0): pos1_i = Rf[i]*p0 + Rf[4+i]*p1 + Rf[8+i]*p2 */
  for (int i = 0; i < 3; ++i) P[i] = Rf[i] * pos0[0] + Rf[4 + i] * pos0[1] + Rf[8 + i] * pos0[2];
  /* dMultiply1_333(R1, Rf, R): R1[4i+j] = Rf[i]*R[j] + Rf[4+i]*R[4+j] + Rf[8+i]*R[8+j] */
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R1[4 * i + j] = Rf[i] * R[j] + Rf[4 + i] * R[4 + j] + Rf[8 + i] * R[8 + j];
    R1[4 * i + 3] = 0.0f;
  }
  P[0] += f->hW; P[2] += f->hD;                                                 /* :1851-1852 */
  /* dxBox::computeAABB, box.cpp:60-77 */
  const float xr = 0.5f * (fabsf(R1[0] * side[0]) + fabsf(R1[1] * side[1]) + fabsf(R1[2] * side[2]));
  const float yr = 0.5f * (fabsf(R1[4] * side[0]) + fabsf(R1[5] * side[1]) + fabsf(R1[6] * side[2]));
  const float zr = 0.5f * (fabsf(R1[8] * side[0]) + fabsf(R1[9] * side[1]) + fabsf(R1[10] * side[2]));
  float aabb[6];
  aabb[0] = P[0] - xr; aabb[1] = P[0] + xr; aabb[2] = P[1] - yr; aabb[3] = P[1] + yr;
  aabb[4] = P[2] - zr; aabb[5] = P[2] + zr;
  sc->stage = ORC_ST_AABB; sc->num_tri = sc->num_groups = 0;
  if (aabb[0] > f->W || aabb[4] > f->D) return 0;                               /* :1870-1872 */
  if (aabb[1] < 0 || aabb[5] < 0) return 0;                                     /* :1874-1876 */
  int nMinX = (int)floorf(nextafterf(aabb[0] * f->iW, -INFINITY));              /* :1880-1885 */
  int nMaxX = (int)ceilf(nextafterf(aabb[1] * f->iW, INFINITY));
  int nMinZ = (int)floorf(nextafterf(aabb[4] * f->iD, -INFINITY));
  int nMaxZ = (int)ceilf(nextafterf(aabb[5] * f->iD, INFINITY));
  nMinX = nMinX > 0 ? nMinX : 0;                                                /* :1889-1892 */
  nMaxX = nMaxX > f->nx - 1 ? f->nx - 1 : nMaxX;
  nMinZ = nMinZ > 0 ? nMinZ : 0;
  nMaxZ = nMaxZ > f->nz - 1 ? f->nz - 1 : nMaxZ;
  if (zv) *zv = (uint32_t)((nMaxX - nMinX + 1) * (nMaxZ - nMinZ + 1));
  return collide_zone(f, nMinX, nMaxX, nMinZ, nMaxZ, R1, P, side, aabb, sc);
}

/* ------------------------------------------------------------------------------------------------
 * public interface
 * ---------------------------------------------------------------------------------------------- */
const char* orc_kind(void) { return "port"; }

orc_handle* orc_create(const orc_params* p) {
  orc_handle* h = (orc_handle*)calloc(1, sizeof(orc_handle));
  h->p = *p;
  /* ValidityCheckerBody ctor (validity_checker_body.cpp:9-13), Feet ctor (validity_checker_feet.cpp:13-18):
   * HeightMapBoxChecker(float,float,float) */
  h->side[0][0] = (float)p->torso_length; h->side[0][1] = (float)p->torso_width; h->side[0][2] = (float)p->torso_height;
  h->side[1][0] = (float)p->reach_x; h->side[1][1] = (float)p->reach_y; h->side[1][2] = (float)p->reach_z;
  return h;
}

void orc_destroy(orc_handle* h) {
  if (!h) return;
  free(h->f[0].H); free(h->f[1].H); free(h);
}

/* HeightMapBoxChecker::setHeightField (height_map_box_checker.cpp:38-54) +
 * dxHeightfieldData::SetData (heightfield.cpp:130-169). */
static void set_field(orc_field* f, const float* layer, int rows, int cols, double Lx, double Ly,
                      double cx, double cy) {
  free(f->H);
  f->H = (float*)malloc(sizeof(float) * (size_t)rows * cols);
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i)
      f->H[i + (size_t)j * rows] = layer[i + (size_t)(cols - 1 - j) * rows];   /* rowwise().reverse() */
  f->nx = rows; f->nz = cols;
  f->W = (float)Lx; f->D = (float)Ly;
  f->hW = f->W / 2.0f; f->hD = f->D / 2.0f;
  f->sW = f->W / (f->nx - 1.0f);
  f->sD = f->D / (f->nz - 1.0f);
  f->asp = f->sD / f->sW;
  f->iW = 1.0f / f->sW;
  f->iD = 1.0f / f->sD;
  f->px = (float)cx; f->py = (float)cy;
}

int orc_set_map(orc_handle* h, const float* elevation, const float* elevation_masked,
                int rows, int cols, double res, double cx, double cy) {
  if (!h || rows < 2 || cols < 2) return 1;
  const double Lx = rows * res, Ly = cols * res;     /* grid_map: length_ = size * resolution */
  set_field(&h->f[0], elevation, rows, cols, Lx, Ly, cx, cy);
  set_field(&h->f[1], elevation_masked, rows, cols, Lx, Ly, cx, cy);
  h->g.Lx = Lx; h->g.Ly = Ly; h->g.cx = cx; h->g.cy = cy; h->g.has_map = 1;
  return 0;
}

typedef struct port_ctx { const orc_handle* h; orc_scratch sc; } port_ctx;

static int port_collide(void* vctx, int which, const float origin[3], const float rot12[12], uint32_t* zv) {
  port_ctx* c = (port_ctx*)vctx;
  return box_collide(&c->h->f[which], c->h->side[which], origin, rot12, &c->sc, zv);
}

int orc_box_collide(orc_handle* h, int which, const float* origins, const float* rots, size_t n,
                    uint8_t* hit, uint32_t* zone_verts) {
  if (!h || !h->g.has_map) return 1;
  port_ctx c = {h, {0, 0, 0, 0, 0, 0, 0}};
  for (size_t i = 0; i < n; ++i) {
    uint32_t zv = 0;
    hit[i] = (uint8_t)(port_collide(&c, which, origins + 3 * i, rots + 12 * i, &zv) ? 1 : 0);
    if (zone_verts) zone_verts[i] = zv;
  }
  free(c.sc.tri); free(c.sc.group);
  return 0;
}

/* Port-only diagnostics: exit stage (ORC_ST_*), kept triangles and plane groups of each box call.
 * NOTE num_groups counts groups formed before an early plane hit. */
int orc_port_box_stats(orc_handle* h, int which, const float* origins, const float* rots, size_t n,
                       uint8_t* hit, uint8_t* stage, uint32_t* num_tri, uint32_t* num_groups) {
  if (!h || !h->g.has_map) return 1;
  port_ctx c = {h, {0, 0, 0, 0, 0, 0, 0}};
  for (size_t i = 0; i < n; ++i) {
    hit[i] = (uint8_t)(port_collide(&c, which, origins + 3 * i, rots + 12 * i, NULL) ? 1 : 0);
    stage[i] = (uint8_t)c.sc.stage; num_tri[i] = c.sc.num_tri; num_groups[i] = c.sc.num_groups;
  }
  free(c.sc.tri); free(c.sc.group);
  return 0;
}

/* Port-only diagnostics: all five boxes of every pose WITHOUT the reference's short-circuits (what the GPU's
 * classify stage sees): exit stage, hit flag and zone vertex count per box; box 0 = torso, 1..4 = feet.
 * stage 255 = box centre outside the map (no collider call). */
int orc_port_pose_box_stats(orc_handle* h, const double* states, size_t n, uint8_t* stage, uint8_t* hit, uint32_t* zv_out) {
  if (!h || !h->g.has_map) return 1;
  port_ctx c = {h, {0, 0, 0, 0, 0, 0, 0}};
  for (size_t i = 0; i < n; ++i) {
    float t[3], R[9], rot[12], o[3], tt[3];
    orc_pose3_from_se3(states + 7 * i, t, R);
    orc_fill_rot12(R, rot);
    const float fx = (float)h->p.feet_off_x, fy = (float)h->p.feet_off_y;
    for (int k = 0; k < 5; ++k) {
      if (k == 0) { o[0] = (float)h->p.torso_off_x; o[1] = (float)h->p.torso_off_y; o[2] = (float)(h->p.torso_off_z - h->p.feet_off_z); }
      else { o[0] = ((k - 1) & 2) ? -fx : fx; o[1] = ((k - 1) & 1) ? -fy : fy; o[2] = 0.0f; }
      orc_compose_translation(R, t, o, tt);
      uint32_t zv = 0;
      if (!orc_is_inside(&h->g, tt[0], tt[1])) { stage[5 * i + k] = 255; hit[5 * i + k] = 0; zv_out[5 * i + k] = 0; continue; }
      hit[5 * i + k] = (uint8_t)(port_collide(&c, k ? 1 : 0, tt, rot, &zv) ? 1 : 0);
      stage[5 * i + k] = (uint8_t)c.sc.stage;
      zv_out[5 * i + k] = zv;
    }
  }
  free(c.sc.tri); free(c.sc.group);
  return 0;
}

int orc_check_poses(orc_handle* h, const double* states, size_t n, uint8_t* valid, uint32_t* zone_verts) {
  if (!h || !h->g.has_map) return 1;
  port_ctx c = {h, {0, 0, 0, 0, 0, 0, 0}};
  for (size_t i = 0; i < n; ++i) {
    uint32_t zv = 0;
    valid[i] = (uint8_t)orc_state_valid(&h->p, &h->g, port_collide, &c, states + 7 * i, &zv);
    if (zone_verts) zone_verts[i] = zv;
  }
  free(c.sc.tri); free(c.sc.group);
  return 0;
}

int orc_check_motions(orc_handle* h, const double* s1, const double* s2, size_t n, int n_steps,
                      uint8_t* valid, uint32_t* zone_verts) {
  if (!h || !h->g.has_map) return 1;
  port_ctx c = {h, {0, 0, 0, 0, 0, 0, 0}};
  for (size_t i = 0; i < n; ++i) {
    uint32_t zv = 0, zsum = 0;
    int ok = orc_state_valid(&h->p, &h->g, port_collide, &c, s2 + 7 * i, &zv);
    zsum += zv;
    for (int j = 1; j <= n_steps && ok; ++j) {
      double st[7];
      orc_se3_interpolate(s1 + 7 * i, s2 + 7 * i, (double)j / (double)(n_steps + 1), st);
      ok = orc_state_valid(&h->p, &h->g, port_collide, &c, st, &zv);
      zsum += zv;
    }
    valid[i] = (uint8_t)ok;
    if (zone_verts) zone_verts[i] = zsum;
  }
  free(c.sc.tri); free(c.sc.group);
  return 0;
}

int orc_check_edge_interiors(orc_handle* h, const double* s1, const double* s2, size_t n, const int32_t* n_interp,
                             double max_lateral, int32_t* valid_prefix) {
  if (!h || !h->g.has_map) return 1;
  port_ctx c = {h, {0, 0, 0, 0, 0, 0, 0}};
  for (size_t i = 0; i < n; ++i) {
    const int32_t ni = n_interp ? n_interp[i] : orc_n_interp(s1 + 7 * i, s2 + 7 * i, max_lateral);
    valid_prefix[i] = orc_edge_interior_prefix(&h->p, &h->g, port_collide, &c, s1 + 7 * i, s2 + 7 * i, ni);
  }
  free(c.sc.tri); free(c.sc.group);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * SE3FromSE2Sampler::sampleUniform, art_planner/src/sampler.cpp:40-131 (uniform variates are inputs)
 * ---------------------------------------------------------------------------------------------- */
/* grid_map::getPositionFromIndex (GridMapMath.cpp): position = mapPosition + (0.5*length - 0.5*res) + res * (-index) */
static void gm_position_of_index(const orc_sampler_map* m, int row, int col, double pos[2]) {
  const double offx = 0.5 * (m->rows * m->res) - 0.5 * m->res, offy = 0.5 * (m->cols * m->res) - 0.5 * m->res;
  pos[0] = (m->cx + offx) + m->res * (-(double)row);
  pos[1] = (m->cy + offy) + m->res * (-(double)col);
}

/* grid_map::getIndexFromPosition: indexVector = (position - 0.5*length - mapPosition) / res; index = (int)(-indexVector);
 * valid iff checkIfPositionWithinMap && index in range. */
static int gm_index_of_position(const orc_sampler_map* m, const double pos[2], int* row, int* col) {
  const double Lx = m->rows * m->res, Ly = m->cols * m->res;
  const double vx = ((pos[0] - 0.5 * Lx) - m->cx) / m->res, vy = ((pos[1] - 0.5 * Ly) - m->cy) / m->res;
  *row = (int)(-vx);
  *col = (int)(-vy);
  orc_geom g = {Lx, Ly, m->cx, m->cy, 1};
  const double tx = -((pos[0] - g.cx) - 0.5 * Lx), ty = -((pos[1] - g.cy) - 0.5 * Ly);
  const int inside = tx >= 0.0 && ty >= 0.0 && tx < Lx && ty < Ly;
  return inside && *row >= 0 && *col >= 0 && *row < m->rows && *col < m->cols;
}

static void cross3d(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

static void sample_one(const orc_sampler_map* m, const orc_sampler_params* p, const double u[6], double s[7], int32_t rc[2]) {
  double pos[2];
  if (p->sample_from_distribution) {
    /* samplePositionInMapFromDist, sampler.cpp:54-77: linear scans, float CDF against double variate */
    const double samp_col = u[0], samp_row = u[1];
    int row, col;
    for (row = 0; row < m->rows - 1; ++row)
      if ((double)m->cum_prob_rowwise[row] > samp_row) break;
    for (col = 0; col < m->cols - 1; ++col)
      if ((double)m->cum_prob[row + (size_t)col * m->rows] > samp_col) break;
    gm_position_of_index(m, row, col, pos);
  } else {
    /* samplePositionInMap, sampler.cpp:40-50: one attempt; uniformReal(a,b) = (b-a)*u + a */
    pos[0] = (p->high[0] - p->low[0]) * u[0] + p->low[0];
    pos[1] = (p->high[1] - p->low[1]) * u[1] + p->low[1];
  }
  int row, col;
  if (!gm_index_of_position(m, pos, &row, &col)) {      /* :91 getIndexOfPosition throws / :50 loop repeats */
    for (int k = 0; k < 7; ++k) s[k] = NAN;
    rc[0] = rc[1] = -1;
    return;
  }
  rc[0] = row; rc[1] = col;
  const size_t at = row + (size_t)col * m->rows;
  double x = pos[0], y = pos[1], z = (double)m->elevation[at];                 /* :93-95 */
  const double nw[3] = {(double)m->normal_x[at], (double)m->normal_y[at], (double)m->normal_z[at]};   /* :98 */
  const float sd = m->plane_fit_std_dev[at];                                    /* :100 */
  const double pert = ((2.0 * u[2] + -1.0) * (double)(sd < 0.5f ? sd : 0.5f)) * p->reach_z;   /* :103 */
  x += nw[0] * pert; y += nw[1] * pert; z += nw[2] * pert;                      /* :105-107 */
  /* RNG::eulerRPY (OMPL 1.4.2 RandomNumbers.cpp) */
  const double pi = 3.14159265358979323846;
  double v0 = pi * (-2.0 * u[3] + 1.0);
  double v1 = acos(1.0 - 2.0 * u[4]) - pi / 2.0;
  const double v2 = pi * (-2.0 * u[5] + 1.0);
  /* Quaterniond(AngleAxisd(yaw, UnitZ)).inverse() * normal_w   (:118-121; Eigen Quaternion.h) */
  const double ha = 0.5 * v2, sn = sin(ha), cs = cos(ha);
  const double q[4] = {sn * 0.0, sn * 0.0, sn * 1.0, cs};                       /* x y z w */
  const double n2 = (q[0] * q[0] + q[2] * q[2]) + (q[1] * q[1] + q[3] * q[3]);  /* squaredNorm, 2-wide packets */
  const double qi[3] = {-q[0] / n2, -q[1] / n2, -q[2] / n2}, qw = q[3] / n2;
  double uv[3], t2[3], nb[3];
  cross3d(qi, nw, uv);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  cross3d(qi, uv, t2);
  for (int k = 0; k < 3; ++k) nb[k] = (nw[k] + qw * uv[k]) + t2[k];
  v0 = -atan2(nb[1], nb[2]) + v0 * p->max_roll_pert / M_PI_2;                   /* :123-124 */
  v1 = atan2(nb[0], nb[2]) + v1 * p->max_pitch_pert / M_PI_4;                   /* :125-126 */
  /* setSO3FromRPY, utils.h:101-115 */
  const double r2 = v0 * 0.5, p2 = v1 * 0.5, y2 = v2 * 0.5;
  const double cr = cos(r2), cp = cos(p2), cy = cos(y2), sr = sin(r2), sp = sin(p2), sy = sin(y2);
  s[0] = x; s[1] = y; s[2] = z;
  s[6] = cy * cp * cr + sy * sp * sr;
  s[3] = cy * cp * sr - sy * sp * cr;
  s[4] = sy * cp * sr + cy * sp * cr;
  s[5] = sy * cp * cr - cy * sp * sr;
}

int orc_sample_states(const orc_sampler_map* m, const orc_sampler_params* p, const double* u, size_t n, double* states,
                      int32_t* rowcol) {
  if (!m || !p || (p->sample_from_distribution && (!m->cum_prob || !m->cum_prob_rowwise))) return 1;
  for (size_t i = 0; i < n; ++i) {
    int32_t rc[2];
    sample_one(m, p, u + 6 * i, states + 7 * i, rc);
    if (rowcol) { rowcol[2 * i] = rc[0]; rowcol[2 * i + 1] = rc[1]; }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * art_planner::estimateNormals, art_planner/src/utils.cpp:213-324
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float v[3]; } f3;

static f3 en_sub(f3 a, f3 b) { f3 r = {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; return r; }

/* vec_x.cross(vec_y).normalized() added into sum (utils.cpp:268 etc.) */
static void en_accumulate(f3 vx, f3 vy, f3* sum) {
  f3 c;
  c.v[0] = vx.v[1] * vy.v[2] - vx.v[2] * vy.v[1];
  c.v[1] = vx.v[2] * vy.v[0] - vx.v[0] * vy.v[2];
  c.v[2] = vx.v[0] * vy.v[1] - vx.v[1] * vy.v[0];
  const float z = c.v[0] * c.v[0] + (c.v[1] * c.v[1] + c.v[2] * c.v[2]);
  if (z > 0.0f) { const float n = sqrtf(z); c.v[0] /= n; c.v[1] /= n; c.v[2] /= n; }
  sum->v[0] += c.v[0]; sum->v[1] += c.v[1]; sum->v[2] += c.v[2];
}

int orc_estimate_normals(const float* elevation, int rows, int cols, double res, double cx, double cy,
                         double estimation_radius, float* normal_x, float* normal_y, float* normal_z,
                         float* plane_fit_std_dev) {
  const int r_cells = (int)(estimation_radius / res);                         /* :226 */
  const int r_diag = (int)(estimation_radius * 0.70710678118 / res);          /* :227 */
  /* map_3d (:236-249): cell positions cast to float (grid_map::getPosition, see gm_position_of_index) */
  float* px = (float*)malloc(sizeof(float) * rows);
  float* py = (float*)malloc(sizeof(float) * cols);
  if (!px || !py) { free(px); free(py); return 1; }
  const double offx = 0.5 * (rows * res) - 0.5 * res, offy = 0.5 * (cols * res) - 0.5 * res;
  for (int i = 0; i < rows; ++i) px[i] = (float)((cx + offx) + res * (-(double)i));
  for (int j = 0; j < cols; ++j) py[j] = (float)((cy + offy) + res * (-(double)j));
#define EN_P(i, j) ((f3){{px[i], py[j], elevation[(i) + (size_t)(j) * rows]}})
#define EN_DZ(q_) do { const float a_ = fabsf((q_).v[2]); if (a_ > max_z_diff) max_z_diff = a_; } while (0)
  for (int i = 0; i < rows; ++i) {
    for (int j = 0; j < cols; ++j) {
      f3 sum = {{0.0f, 0.0f, 0.0f}};
      unsigned int n_vec = 0;
      float max_z_diff = 0.0f;
      const f3 center = EN_P(i, j);
      for (int o = 1; o < r_cells; ++o) {                                      /* :260-271 */
        if (i + o >= rows || j + o >= cols) continue;
        const f3 vx = en_sub(EN_P(i + o, j), center), vy = en_sub(EN_P(i, j + o), center);
        EN_DZ(vx); EN_DZ(vy);
        en_accumulate(vx, vy, &sum); ++n_vec;
      }
      for (int o = 1; o < r_cells; ++o) {                                      /* :272-282 */
        if (i - o < 0 || j - o < 0) continue;
        const f3 vx = en_sub(EN_P(i - o, j), center), vy = en_sub(EN_P(i, j - o), center);
        EN_DZ(vx); EN_DZ(vy);
        en_accumulate(vx, vy, &sum); ++n_vec;
      }
      for (int o = 1; o < r_diag; ++o) {                                       /* :283-297 */
        if (i + o >= rows || j + o >= cols || i - o < 0) continue;            /* j_offset_2 = j + o: same bound */
        const f3 vx = en_sub(EN_P(i + o, j + o), center), vy = en_sub(EN_P(i - o, j + o), center);
        EN_DZ(vx); EN_DZ(vy);
        en_accumulate(vx, vy, &sum); ++n_vec;
      }
      for (int o = 1; o < r_diag; ++o) {                                       /* :298-312 */
        if (i - o < 0 || j - o < 0 || i + o >= rows) continue;
        const f3 vx = en_sub(EN_P(i - o, j - o), center), vy = en_sub(EN_P(i + o, j - o), center);
        EN_DZ(vx); EN_DZ(vy);
        en_accumulate(vx, vy, &sum); ++n_vec;
      }
      if (n_vec > 0) { const float d = (float)n_vec; sum.v[0] /= d; sum.v[1] /= d; sum.v[2] /= d; }   /* :315-317 */
      const size_t at = i + (size_t)j * rows;
      plane_fit_std_dev[at] = max_z_diff;                                      /* :318 */
      const float z = sum.v[0] * sum.v[0] + (sum.v[1] * sum.v[1] + sum.v[2] * sum.v[2]);   /* normalize(), :320 */
      if (z > 0.0f) { const float n = sqrtf(z); sum.v[0] /= n; sum.v[1] /= n; sum.v[2] /= n; }
      normal_x[at] = sum.v[0]; normal_y[at] = sum.v[1]; normal_z[at] = sum.v[2];
    }
  }
#undef EN_P
#undef EN_DZ
  free(px); free(py);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * computeCumulativeProbabilityDistribution, probability_distribution.cpp:20-46
 * ---------------------------------------------------------------------------------------------- */
int orc_compute_cdf(const float* prob, int rows, int cols, float* cum_prob, float* cum_row) {
  float total = 0.0f;
  for (int i = 0; i < rows; ++i) {                       /* :23 prob.rowwise().sum() */
    float s = prob[i];
    for (int j = 1; j < cols; ++j) s = s + prob[i + (size_t)j * rows];
    cum_row[i] = s;
  }
  total = cum_row[0];
  for (int i = 1; i < rows; ++i) total = total + cum_row[i];   /* :26 prob_rowwise.sum() */
  for (int i = 0; i < rows; ++i) {
    const float rs = cum_row[i];
    /* :28 cum_prob.array().colwise() /= rowwise sums ; :37-39 cumulate along the columns */
    float run = prob[i] / rs;
    cum_prob[i] = run;
    for (int j = 1; j < cols; ++j) {
      run = prob[i + (size_t)j * rows] / rs + run;
      cum_prob[i + (size_t)j * rows] = run;
    }
  }
  /* :26 normalise, :32-35 cumulate the row distribution */
  float run = cum_row[0] / total;
  cum_row[0] = run;
  for (int i = 1; i < rows; ++i) { run = cum_row[i] / total + run; cum_row[i] = run; }
  return 0;
}

int orc_path_length_cost(orc_handle* h, const double* s1, const double* s2, size_t n, double* cost) {
  if (!h) return 1;
  for (size_t i = 0; i < n; ++i) cost[i] = orc_path_length(&h->p, s1 + 7 * i, s2 + 7 * i);
  return 0;
}

typedef struct mt_job { orc_handle* h; const double* states; size_t lo, hi; uint8_t* valid; } mt_job;

static void* mt_worker(void* v) {
  mt_job* j = (mt_job*)v;
  port_ctx c = {j->h, {0, 0, 0, 0, 0, 0, 0}};
  for (size_t i = j->lo; i < j->hi; ++i)
    j->valid[i] = (uint8_t)orc_state_valid(&j->h->p, &j->h->g, port_collide, &c, j->states + 7 * i, NULL);
  free(c.sc.tri); free(c.sc.group);
  return NULL;
}

int orc_check_poses_mt(orc_handle* h, const double* states, size_t n, uint8_t* valid, int n_threads) {
  if (!h || !h->g.has_map) return 1;
  if (n_threads < 1) n_threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  mt_job* jobs = (mt_job*)malloc(sizeof(mt_job) * n_threads);
  for (int t = 0; t < n_threads; ++t) {
    jobs[t].h = h; jobs[t].states = states; jobs[t].valid = valid;
    jobs[t].lo = n * (size_t)t / n_threads; jobs[t].hi = n * (size_t)(t + 1) / n_threads;
    pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return 0;
}

typedef struct mt_edge_job { orc_handle* h; const double *s1, *s2; size_t lo, hi; int n_steps; uint8_t* valid; } mt_edge_job;

static void* mt_edge_worker(void* v) {
  mt_edge_job* j = (mt_edge_job*)v;
  port_ctx c = {j->h, {0, 0, 0, 0, 0, 0, 0}};
  for (size_t i = j->lo; i < j->hi; ++i) {
    int ok = orc_state_valid(&j->h->p, &j->h->g, port_collide, &c, j->s2 + 7 * i, NULL);
    for (int k = 1; k <= j->n_steps && ok; ++k) {
      double st[7];
      orc_se3_interpolate(j->s1 + 7 * i, j->s2 + 7 * i, (double)k / (double)(j->n_steps + 1), st);
      ok = orc_state_valid(&j->h->p, &j->h->g, port_collide, &c, st, NULL);
    }
    j->valid[i] = (uint8_t)ok;
  }
  free(c.sc.tri); free(c.sc.group);
  return NULL;
}

/* orc_check_motions on n_threads workers. */
int orc_check_motions_mt(orc_handle* h, const double* s1, const double* s2, size_t n, int n_steps, uint8_t* valid,
                         int n_threads) {
  if (!h || !h->g.has_map) return 1;
  if (n_threads < 1) n_threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  mt_edge_job* jobs = (mt_edge_job*)malloc(sizeof(mt_edge_job) * n_threads);
  for (int t = 0; t < n_threads; ++t) {
    jobs[t].h = h; jobs[t].s1 = s1; jobs[t].s2 = s2; jobs[t].valid = valid; jobs[t].n_steps = n_steps;
    jobs[t].lo = n * (size_t)t / n_threads; jobs[t].hi = n * (size_t)(t + 1) / n_threads;
    pthread_create(&th[t], NULL, mt_edge_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return 0;
}

int orc_valid_segment_count(const double low[3], const double high[3], double frac, const double* s1, const double* s2, size_t n,
                            int32_t* nd) {
  for (size_t i = 0; i < n; ++i) nd[i] = orc_valid_segment_count_1(low, high, frac, s1 + 7 * i, s2 + 7 * i);
  return 0;
}

int orc_check_motions_segments(orc_handle* h, const double* s1, const double* s2, size_t n, const int32_t* nd, uint8_t* valid,
                               double* last_t) {
  if (!h || !h->g.has_map) return 1;
  port_ctx c = {h, {0, 0, 0, 0, 0, 0, 0}};
  for (size_t i = 0; i < n; ++i) {
    double t = 1.0;
    valid[i] = (uint8_t)orc_motion_last_valid(&h->p, &h->g, port_collide, &c, s1 + 7 * i, s2 + 7 * i, nd[i], &t);
    if (last_t) last_t[i] = t;
  }
  free(c.sc.tri); free(c.sc.group);
  return 0;
}

int orc_edge_matrix(const double* s_start, const double* s_target, size_t n, float* edges) {
  for (size_t i = 0; i < n; ++i) orc_edge_matrix_row(s_start + 7 * i, s_target + 7 * i, edges + 6 * i);
  return 0;
}
