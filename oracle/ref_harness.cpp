/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's own compiled ODE (the vendored, modified ODE 0.16.1 under
 * /root/reference/ode, compiled unmodified by oracle/Makefile) through exactly the public-API call
 * sequence of art_planner's HeightMapBoxChecker
 *   (art_planner/src/validity_checker/height_map_box_checker.cpp:11-72),
 * and exports the interface of artp_oracle.h. The pose-level wrappers (Eigen/grid_map/OMPL
 * restatements) are the shared artp_wrappers.h, so liborc_ref.so and liborc_port.so differ exactly in
 * the collider: compiled reference ODE here vs. the C restatement there.
 */
#include <ode/ode.h>

#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

#include "artp_oracle.h"
#include "artp_wrappers.h"

namespace {

std::once_flag g_ode_once;

// Mirrors art_planner::HeightMapBoxChecker (height_map_box_checker.h:17-65).
struct RefChecker {
  dWorldID world;
  dSpaceID space;
  dBodyID body_field, body_box;
  dGeomID geom_field, geom_box;
  dHeightfieldDataID data;
  std::vector<float> mat;
  dReal rot[12];
  dContactGeom contact;

  RefChecker(float lx, float ly, float lz) {
    std::call_once(g_ode_once, [] { dInitODE(); });
    world = dWorldCreate();
    space = dHashSpaceCreate(0);
    body_box = dBodyCreate(world);
    body_field = dBodyCreate(world);
    geom_box = dCreateBox(space, lx, ly, lz);
    data = dGeomHeightfieldDataCreate();
    geom_field = dCreateHeightfield(space, data, 1);
    dRFrom2Axes(rot, -1, 0, 0, 0, 0, 1);
    dGeomSetBody(geom_box, body_box);
    dGeomSetBody(geom_field, body_field);
    dBodySetRotation(body_field, rot);
  }
  ~RefChecker() {
    dSpaceDestroy(space);
    dWorldDestroy(world);
    dGeomHeightfieldDataDestroy(data);
  }
  // setHeightField (height_map_box_checker.cpp:38-54). layer: col-major rows x cols.
  void setHeightField(const float* layer, int rows, int cols, double Lx, double Ly, double cx, double cy) {
    mat.resize((size_t)rows * cols);
    float mn = std::numeric_limits<float>::infinity(), mx = -std::numeric_limits<float>::infinity();
    for (int j = 0; j < cols; ++j)
      for (int i = 0; i < rows; ++i) {
        const float v = layer[i + (size_t)(cols - 1 - j) * rows];   // rowwise().reverse()
        mat[i + (size_t)j * rows] = v;
        if (std::isfinite(v)) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
      }
    dGeomHeightfieldDataBuildSingle(data, mat.data(), 0, Lx, Ly, rows, cols, 1, 0, 0, 0);
    dGeomHeightfieldDataSetBounds(data, mn, mx);
    dGeomHeightfieldSetHeightfieldData(geom_field, data);
    dBodySetPosition(body_field, cx, cy, 0);
  }
  // checkCollision for one pose (height_map_box_checker.cpp:58-72).
  int collide(const float origin[3], const float rot12[12]) {
    dBodySetPosition(body_box, origin[0], origin[1], origin[2]);
    dBodySetRotation(body_box, rot12);
    return dCollide(geom_box, geom_field, 1, &contact, sizeof(dContactGeom));
  }
  void fieldRotation(float out[12]) const {
    const dReal* R = dBodyGetRotation(body_field);
    for (int i = 0; i < 12; ++i) out[i] = R[i];
  }
};

struct LayerCopy { std::vector<float> elev, masked; int rows = 0, cols = 0; double res = 0, cx = 0, cy = 0; };

struct CheckerPair {
  RefChecker torso, foot;
  CheckerPair(const orc_params& p)
      : torso((float)p.torso_length, (float)p.torso_width, (float)p.torso_height),
        foot((float)p.reach_x, (float)p.reach_y, (float)p.reach_z) {}
  void setMap(const LayerCopy& m) {
    const double Lx = m.rows * m.res, Ly = m.cols * m.res;
    torso.setHeightField(m.elev.data(), m.rows, m.cols, Lx, Ly, m.cx, m.cy);
    foot.setHeightField(m.masked.data(), m.rows, m.cols, Lx, Ly, m.cx, m.cy);
  }
};

int ref_collide(void* ctx, int which, const float origin[3], const float rot12[12], uint32_t* zv) {
  CheckerPair* c = static_cast<CheckerPair*>(ctx);
  if (zv) *zv = 0;   // the compiled reference does not expose the zone size
  return (which == 0 ? c->torso : c->foot).collide(origin, rot12) != 0;
}

}  // namespace

struct orc_handle {
  orc_params p;
  orc_geom g;
  LayerCopy map;
  CheckerPair* pair;
  std::vector<CheckerPair*> mt_pairs;   // one ODE world + heightfield pair per worker thread (built lazily)
};

static void drop_mt_pairs(orc_handle* h) {
  for (auto* p : h->mt_pairs) delete p;
  h->mt_pairs.clear();
}

extern "C" {

const char* orc_kind(void) { return "reference"; }

orc_handle* orc_create(const orc_params* p) {
  orc_handle* h = new orc_handle();
  h->p = *p;
  std::memset(&h->g, 0, sizeof(h->g));
  h->pair = new CheckerPair(*p);
  return h;
}

void orc_destroy(orc_handle* h) {
  if (!h) return;
  drop_mt_pairs(h);
  delete h->pair;
  delete h;
}

int orc_set_map(orc_handle* h, const float* elevation, const float* elevation_masked, int rows, int cols,
                double res, double cx, double cy) {
  if (!h || rows < 2 || cols < 2) return 1;
  h->map.elev.assign(elevation, elevation + (size_t)rows * cols);
  h->map.masked.assign(elevation_masked, elevation_masked + (size_t)rows * cols);
  h->map.rows = rows; h->map.cols = cols; h->map.res = res; h->map.cx = cx; h->map.cy = cy;
  h->pair->setMap(h->map);
  drop_mt_pairs(h);
  h->g.Lx = rows * res; h->g.Ly = cols * res; h->g.cx = cx; h->g.cy = cy; h->g.has_map = 1;
  return 0;
}

int orc_box_collide(orc_handle* h, int which, const float* origins, const float* rots, size_t n,
                    uint8_t* hit, uint32_t* zone_verts) {
  if (!h || !h->g.has_map) return 1;
  for (size_t i = 0; i < n; ++i) {
    hit[i] = (uint8_t)ref_collide(h->pair, which, origins + 3 * i, rots + 12 * i, nullptr);
    if (zone_verts) zone_verts[i] = 0;
  }
  return 0;
}

int orc_check_poses(orc_handle* h, const double* states, size_t n, uint8_t* valid, uint32_t* zone_verts) {
  if (!h || !h->g.has_map) return 1;
  for (size_t i = 0; i < n; ++i) {
    valid[i] = (uint8_t)orc_state_valid(&h->p, &h->g, ref_collide, h->pair, states + 7 * i, nullptr);
    if (zone_verts) zone_verts[i] = 0;
  }
  return 0;
}

int orc_check_motions(orc_handle* h, const double* s1, const double* s2, size_t n, int n_steps,
                      uint8_t* valid, uint32_t* zone_verts) {
  if (!h || !h->g.has_map) return 1;
  for (size_t i = 0; i < n; ++i) {
    int ok = orc_state_valid(&h->p, &h->g, ref_collide, h->pair, s2 + 7 * i, nullptr);
    for (int j = 1; j <= n_steps && ok; ++j) {
      double st[7];
      orc_se3_interpolate(s1 + 7 * i, s2 + 7 * i, (double)j / (double)(n_steps + 1), st);
      ok = orc_state_valid(&h->p, &h->g, ref_collide, h->pair, st, nullptr);
    }
    valid[i] = (uint8_t)ok;
    if (zone_verts) zone_verts[i] = 0;
  }
  return 0;
}

int orc_check_edge_interiors(orc_handle* h, const double* s1, const double* s2, size_t n, const int32_t* n_interp,
                             double max_lateral, int32_t* valid_prefix) {
  if (!h || !h->g.has_map) return 1;
  for (size_t i = 0; i < n; ++i) {
    const int32_t ni = n_interp ? n_interp[i] : orc_n_interp(s1 + 7 * i, s2 + 7 * i, max_lateral);
    valid_prefix[i] = orc_edge_interior_prefix(&h->p, &h->g, ref_collide, h->pair, s1 + 7 * i, s2 + 7 * i, ni);
  }
  return 0;
}

int orc_path_length_cost(orc_handle* h, const double* s1, const double* s2, size_t n, double* cost) {
  if (!h) return 1;
  for (size_t i = 0; i < n; ++i) cost[i] = orc_path_length(&h->p, s1 + 7 * i, s2 + 7 * i);
  return 0;
}

// "all cores": one independent ODE world + heightfield pair per thread.
int orc_check_poses_mt(orc_handle* h, const double* states, size_t n, uint8_t* valid, int n_threads) {
  if (!h || !h->g.has_map) return 1;
  if (n_threads < 1) n_threads = 1;
  // worker checkers persist across calls (the per-thread world/heightfield setup is a one-time cost per map)
  while ((int)h->mt_pairs.size() < n_threads) {
    CheckerPair* cp = new CheckerPair(h->p);
    cp->setMap(h->map);
    h->mt_pairs.push_back(cp);
  }
  std::vector<CheckerPair*>& pairs = h->mt_pairs;
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) {
    th.emplace_back([=, &pairs] {
      const size_t lo = n * (size_t)t / n_threads, hi = n * (size_t)(t + 1) / n_threads;
      for (size_t i = lo; i < hi; ++i)
        valid[i] = (uint8_t)orc_state_valid(&h->p, &h->g, ref_collide, pairs[t], states + 7 * i, nullptr);
    });
  }
  for (auto& x : th) x.join();
  return 0;
}

// orc_check_motions on n_threads workers (same per-thread worlds as orc_check_poses_mt).
int orc_check_motions_mt(orc_handle* h, const double* s1, const double* s2, size_t n, int n_steps, uint8_t* valid,
                         int n_threads) {
  if (!h || !h->g.has_map) return 1;
  if (n_threads < 1) n_threads = 1;
  while ((int)h->mt_pairs.size() < n_threads) {
    CheckerPair* cp = new CheckerPair(h->p);
    cp->setMap(h->map);
    h->mt_pairs.push_back(cp);
  }
  std::vector<CheckerPair*>& pairs = h->mt_pairs;
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) {
    th.emplace_back([=, &pairs] {
      const size_t lo = n * (size_t)t / n_threads, hi = n * (size_t)(t + 1) / n_threads;
      for (size_t i = lo; i < hi; ++i) {
        int ok = orc_state_valid(&h->p, &h->g, ref_collide, pairs[t], s2 + 7 * i, nullptr);
        for (int j = 1; j <= n_steps && ok; ++j) {
          double st[7];
          orc_se3_interpolate(s1 + 7 * i, s2 + 7 * i, (double)j / (double)(n_steps + 1), st);
          ok = orc_state_valid(&h->p, &h->g, ref_collide, pairs[t], st, nullptr);
        }
        valid[i] = (uint8_t)ok;
      }
    });
  }
  for (auto& x : th) x.join();
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
  for (size_t i = 0; i < n; ++i) {
    double t = 1.0;
    valid[i] = (uint8_t)orc_motion_last_valid(&h->p, &h->g, ref_collide, h->pair, s1 + 7 * i, s2 + 7 * i, nd[i], &t);
    if (last_t) last_t[i] = t;
  }
  return 0;
}

int orc_edge_matrix(const double* s_start, const double* s_target, size_t n, float* edges) {
  for (size_t i = 0; i < n; ++i) orc_edge_matrix_row(s_start + 7 * i, s_target + 7 * i, edges + 6 * i);
  return 0;
}

// Extra (reference library only): the heightfield body's rotation after dBodySetRotation, so the test
// can pin the constant matrix the port hard-codes.
void orc_ref_field_rotation(orc_handle* h, float out[12]) { h->pair->torso.fieldRotation(out); }

}  // extern "C"
