/*
 * oracle/artp_wrappers.h -- TEST INFRASTRUCTURE ONLY. Pose-level wrapper arithmetic of the hot path,
 * shared by the C port (artp_oracle.c) and the compiled-reference harness (ref_harness.cpp).
 *
 * Restates (the reference's wrappers need Eigen, grid_map_core and OMPL, none of which are in the
 * reference tree or this image -- "parity unpinned" at this level):
 *   - Pose3FromSE3 / Pose3FromXYZ             art_planner/include/art_planner/utils.h:25-48
 *   - StateValidityChecker::isValid            art_planner/src/validity_checker/validity_checker.cpp:39-45
 *   - ValidityCheckerBody::isValid             art_planner/src/validity_checker/validity_checker_body.cpp:27-42
 *   - ValidityCheckerFeet::isValid & friends   art_planner/src/validity_checker/validity_checker_feet.cpp:32-70
 *   - grid_map::GridMap::isInside              (grid_map_core, checkIfPositionWithinMap; call sites
 *                                               validity_checker_body.cpp:29, validity_checker_feet.cpp:34)
 *   - ompl::base::SE3StateSpace::interpolate   (OMPL 1.4.2; call sites prm_motion_cost.cpp:353,652)
 *   - PathLengthObjective::motionCost          art_planner/src/objectives/path_length_objective.cpp:26-70
 *
 * All fp32 arithmetic must be compiled without FMA contraction (-ffp-contract=off).
 */
#ifndef ARTP_WRAPPERS_H
#define ARTP_WRAPPERS_H

#include <math.h>
#include <stdint.h>
#include "artp_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* box-vs-heightfield callback: returns dCollide(...) != 0; *zv = zone vertices scanned. */
typedef int (*orc_collide_fn)(void* ctx, int which, const float origin[3], const float rot12[12],
                              uint32_t* zv);

typedef struct orc_geom {
  double Lx, Ly, cx, cy;      /* grid_map length and position (doubles) */
  int has_map;
} orc_geom;

/* Eigen::Quaternion<float>(w,x,y,z).toRotationMatrix() on double->float casts; utils.h:25-38.
 * No normalisation. R is row-major 3x3. */
static inline void orc_pose3_from_se3(const double s[7], float t[3], float R[9]) {
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

/* Translation of (pose * Pose3FromXYZ(o)): Eigen Transform product = R*o + t, with the fixed-size
 * 3-term dot product reduced as a0 + (a1 + a2) (Eigen redux_novec_unroller halves the range). */
static inline void orc_compose_translation(const float R[9], const float t[3], const float o[3],
                                           float out[3]) {
  for (int i = 0; i < 3; ++i) {
    const float a0 = R[3 * i + 0] * o[0];
    const float a1 = R[3 * i + 1] * o[1];
    const float a2 = R[3 * i + 2] * o[2];
    const float d = a0 + (a1 + a2);
    out[i] = d + t[i];
  }
}

/* grid_map checkIfPositionWithinMap: p' = -(p - c - L/2); inside iff 0 <= p' < L per axis (double). */
static inline int orc_is_inside(const orc_geom* g, float px, float py) {
  const double ox = 0.5 * g->Lx, oy = 0.5 * g->Ly;
  const double tx = -(((double)px - g->cx) - ox);
  const double ty = -(((double)py - g->cy) - oy);
  return tx >= 0.0 && ty >= 0.0 && tx < g->Lx && ty < g->Ly;
}

static inline void orc_fill_rot12(const float R[9], float rot[12]) {
  rot[0] = R[0]; rot[1] = R[1]; rot[2]  = R[2]; rot[3]  = 0.0f;
  rot[4] = R[3]; rot[5] = R[4]; rot[6]  = R[5]; rot[7]  = 0.0f;
  rot[8] = R[6]; rot[9] = R[7]; rot[10] = R[8]; rot[11] = 0.0f;
}

/* StateValidityChecker::isValid. Returns 0/1; *zv accumulates scanned zone vertices. */
static inline int orc_state_valid(const orc_params* p, const orc_geom* g, orc_collide_fn collide,
                                  void* ctx, const double s[7], uint32_t* zv_out) {
  float t[3], R[9], rot[12], o[3], tt[3];
  uint32_t zv = 0, zv_sum = 0;
  orc_pose3_from_se3(s, t, R);
  orc_fill_rot12(R, rot);
  /* torso: validity_checker.cpp:41-43, body check validity_checker_body.cpp:27-42 */
  o[0] = (float)p->torso_off_x; o[1] = (float)p->torso_off_y;
  o[2] = (float)(p->torso_off_z - p->feet_off_z);
  orc_compose_translation(R, t, o, tt);
  int body_ok = 1;
  if (orc_is_inside(g, tt[0], tt[1])) {
    body_ok = !collide(ctx, 0, tt, rot, &zv);
    zv_sum += zv;
  }
  int ok = body_ok;
  if (ok) {
    /* feet: validity_checker_feet.cpp:59-66, order (+,+),(+,-),(-,+),(-,-), early break :49-56 */
    const float fx = (float)p->feet_off_x, fy = (float)p->feet_off_y;
    const float sx[4] = {fx, fx, -fx, -fx};
    const float sy[4] = {fy, -fy, fy, -fy};
    for (int k = 0; k < 4 && ok; ++k) {
      o[0] = sx[k]; o[1] = sy[k]; o[2] = 0.0f;
      orc_compose_translation(R, t, o, tt);
      int v;
      if (!orc_is_inside(g, tt[0], tt[1])) {
        v = !p->unknown_space_untraversable;          /* validity_checker_feet.cpp:34-37 */
      } else {
        v = collide(ctx, 1, tt, rot, &zv) ? 1 : 0;
        zv_sum += zv;
      }
      ok = ok && v;
    }
  }
  if (zv_out) *zv_out = zv_sum;
  return ok;
}

/* OMPL 1.4.2 SE3StateSpace::interpolate = RealVectorStateSpace lerp + SO3StateSpace slerp. */
static inline void orc_se3_interpolate(const double a[7], const double b[7], double t, double out[7]) {
  for (int i = 0; i < 3; ++i) out[i] = a[i] + (b[i] - a[i]) * t;
  const double dq_raw = a[3] * b[3] + a[4] * b[4] + a[5] * b[5] + a[6] * b[6];
  const double dqa = fabs(dq_raw);
  double theta = (dqa > 1.0 - 1e-9) ? 0.0 : acos(dqa);       /* SO3StateSpace arcLength */
  if (theta > 2.220446049250313e-16) {
    const double d = 1.0 / sin(theta);
    const double s0 = sin((1.0 - t) * theta);
    double s1 = sin(t * theta);
    if (dq_raw < 0) s1 = -s1;
    out[3] = (a[3] * s0 + b[3] * s1) * d;
    out[4] = (a[4] * s0 + b[4] * s1) * d;
    out[5] = (a[5] * s0 + b[5] * s1) * d;
    out[6] = (a[6] * s0 + b[6] * s1) * d;
  } else {
    out[3] = a[3]; out[4] = a[4]; out[5] = a[5]; out[6] = a[6];
  }
}

/* lateralDistance (utils.h:52-61) and the interior-state count of addValidMilestone (prm_motion_cost.cpp:341-343). */
static inline int32_t orc_n_interp(const double a[7], const double b[7], double max_lateral) {
  const double dx = b[0] - a[0], dy = b[1] - a[1];
  return (int32_t)(unsigned int)(sqrt(dx * dx + dy * dy) / max_lateral);
}

/* prm_motion_cost.cpp:345-372: leading valid interior states of one edge. */
static inline int32_t orc_edge_interior_prefix(const orc_params* p, const orc_geom* g, orc_collide_fn collide, void* ctx,
                                               const double a[7], const double b[7], int32_t n_interp) {
  const double n_interp_div = 1.0 / (n_interp + 1);
  int32_t k = 0;
  for (int32_t step = 1; step < n_interp + 1; ++step) {
    double st[7];
    orc_se3_interpolate(a, b, step * n_interp_div, st);
    if (!orc_state_valid(p, g, collide, ctx, st, NULL)) break;
    ++k;
  }
  return k;
}

/* getYawFromSO3 (utils.h:80-88): atan2 in double, returned through `Scalar` = float. */
static inline double orc_yaw_from_quat(const double s[7]) {
  const double x = s[3], y = s[4], z = s[5], w = s[6];
  return (double)(float)atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z));
}

static inline double orc_angle_diff(double x, double y) {
  const double d = fabs(y - x);
  return (d > M_PI) ? 2.0 * M_PI - d : d;
}

/* PathLengthObjective::motionCost / motionCostHeuristic (path_length_objective.cpp:26-70). */
static inline double orc_path_length(const orc_params* p, const double s1[7], const double s2[7]) {
  const double x_dif = s2[0] - s1[0], y_dif = s2[1] - s1[1], z_dif = s2[2] - s1[2];
  if (!p->use_directional_cost) {
    return sqrt(x_dif * x_dif + y_dif * y_dif + z_dif * z_dif) / p->max_lon_vel;
  }
  const double yaw1 = orc_yaw_from_quat(s1), yaw2 = orc_yaw_from_quat(s2);
  const double yaw_dif = orc_angle_diff(yaw2, yaw1);
  const double lon_dif = cos(yaw1) * x_dif + sin(yaw1) * y_dif;
  const double lat_dif = -sin(yaw1) * x_dif + cos(yaw1) * y_dif;
  const double t_yaw = fabs(yaw_dif) / p->max_ang_vel;
  const double t_lon = fabs(lon_dif) / p->max_lon_vel;
  const double t_lat = fabs(lat_dif) / p->max_lat_vel;
  const double m = t_lon > t_lat ? t_lon : t_lat;      /* std::max(std::max(t_lon,t_lat),t_yaw) */
  return m > t_yaw ? m : t_yaw;
}

/* ompl::base::CompoundStateSpace::validSegmentCount for SE3StateSpace (OMPL 1.4.2, not in the reference tree; call
 * sites via si_->checkMotion, prm_motion_cost.cpp:652, lazy_prm_star_min_update.cpp:725): max over the sub-spaces of
 * longestValidSegmentCountFactor (1) * (unsigned)ceil(distance / longestValidSegment), longestValidSegment = maximum
 * extent * fraction; R^3: Euclidean distance, extent |high - low| (bounds of planner.cpp:148-156); SO(3): arc length
 * acos(|q1.q2|) (0 above 1 - 1e-9), extent pi/2. */
static inline int32_t orc_valid_segment_count_1(const double low[3], const double high[3], double frac, const double a[7],
                                                const double b[7]) {
  double e2 = 0;
  for (int i = 0; i < 3; ++i) e2 += (high[i] - low[i]) * (high[i] - low[i]);
  const double seg_r3 = sqrt(e2) * frac, seg_so3 = 0.5 * M_PI * frac;
  const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  const double d3 = sqrt(dx * dx + dy * dy + dz * dz);
  const double dq = fabs(a[3] * b[3] + a[4] * b[4] + a[5] * b[5] + a[6] * b[6]);
  const double ds = (dq > 1.0 - 1e-9) ? 0.0 : acos(dq);
  const unsigned n3 = (unsigned)ceil(d3 / seg_r3), ns = (unsigned)ceil(ds / seg_so3);
  return (int32_t)(n3 > ns ? n3 : ns);
}

/* ompl::base::DiscreteMotionValidator::checkMotion(s1, s2, lastValid) (OMPL 1.4.2 DiscreteMotionValidator.cpp): interior
 * states j = 1 .. nd-1 at t = j / nd IN ORDER, the first invalid one sets lastValid.second = (j-1)/nd; then s2, whose
 * failure sets (nd-1)/nd. Returns validity; *last_t untouched for valid motions. nd < 1 is treated as 1. */
static inline int orc_motion_last_valid(const orc_params* p, const orc_geom* g, orc_collide_fn collide, void* ctx,
                                        const double a[7], const double b[7], int32_t nd, double* last_t) {
  if (nd < 1) nd = 1;
  for (int32_t j = 1; j < nd; ++j) {
    double st[7];
    orc_se3_interpolate(a, b, (double)j / (double)nd, st);
    if (!orc_state_valid(p, g, collide, ctx, st, NULL)) { *last_t = (double)(j - 1) / (double)nd; return 0; }
  }
  if (!orc_state_valid(p, g, collide, ctx, b, NULL)) { *last_t = (double)(nd - 1) / (double)nd; return 0; }
  return 1;
}

/* One row [tx, ty, tyaw, sx, sy, syaw] of the MotionCostFunc edge matrix as PRMMotionCostMaintainer::updateEdges fills it
 * (prm_motion_cost.cpp:27-47): doubles assigned into a float matrix, yaw through getYawFromSO3's float. */
static inline void orc_edge_matrix_row(const double s_start[7], const double s_target[7], float row[6]) {
  row[0] = (float)s_target[0]; row[1] = (float)s_target[1]; row[2] = (float)orc_yaw_from_quat(s_target);
  row[3] = (float)s_start[0]; row[4] = (float)s_start[1]; row[5] = (float)orc_yaw_from_quat(s_start);
}

#ifdef __cplusplus
}
#endif
#endif
