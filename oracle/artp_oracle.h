/*
 * oracle/artp_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU oracle for the art_planner hot path).
 *
 * Nothing in the product path (art_planner_b200/, include/) may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Two shared libraries export the SAME C interface declared here:
 *   oracle/liborc_port.so        -- plain-C restatement of the reference algorithm (artp_oracle.c)
 *   oracle/_ref/liborc_ref.so    -- the reference's own vendored ODE (compiled from the sources
 *                                   where they lie under /root/reference/ode) driven by
 *                                   ref_harness.cpp; exists only where /root/reference exists.
 * The pose-level wrapper arithmetic (Eigen / grid_map / OMPL restatements, none of which are in the
 * reference tree) lives in artp_wrappers.h and is shared by both, so the two libraries differ exactly
 * in the box-vs-heightfield collider: restated C vs. the reference's compiled ODE.
 */
#ifndef ARTP_ORACLE_H
#define ARTP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors the fields of art_planner::Params the hot path reads
 * (art_planner/include/art_planner/params.h:14-123). All doubles, as in the reference. */
typedef struct orc_params {
  double torso_length, torso_width, torso_height;     /* params.h:92-94  */
  double torso_off_x, torso_off_y, torso_off_z;       /* params.h:96-100 */
  double feet_off_x, feet_off_y, feet_off_z;          /* params.h:106-110 */
  double reach_x, reach_y, reach_z;                   /* params.h:112-116 */
  int    unknown_space_untraversable;                 /* params.h:26 */
  int    use_directional_cost;                        /* params.h:73 */
  double max_lon_vel, max_lat_vel, max_ang_vel;       /* params.h:74-76 */
} orc_params;

typedef struct orc_handle orc_handle;

orc_handle* orc_create(const orc_params* p);
void        orc_destroy(orc_handle* h);

/* Layers in grid_map layout: column-major rows x cols float, index (i,j) at data[i + j*rows].
 * length = rows*res, cols*res (grid_map: length_ = size * resolution); centre (cx, cy).
 * Mirrors HeightMapBoxChecker::setHeightField (height_map_box_checker.cpp:38-54). */
int orc_set_map(orc_handle* h, const float* elevation, const float* elevation_masked,
                int rows, int cols, double res, double cx, double cy);

/* Raw dCollide(box, heightfield, 1, ...) != 0 for n box poses (height_map_box_checker.cpp:58-72).
 * which: 0 = torso box vs `elevation`, 1 = reach box vs `elevation_masked`.
 * origins n x 3 floats, rots n x 12 floats (row-major 3x4, dPose::rotation). out: 0/1 per pose.
 * zone_verts (nullable): number of heightfield vertices the zone scan visits (0 if rejected before). */
int orc_box_collide(orc_handle* h, int which, const float* origins, const float* rots, size_t n,
                    uint8_t* hit, uint32_t* zone_verts);

/* StateValidityChecker::isValid for n SE(3) states, each 7 doubles x y z qx qy qz qw
 * (validity_checker.cpp:39-45). zone_verts (nullable): sum over the boxes the reference would
 * actually execute (short-circuits honoured) of the zone vertex count. */
int orc_check_poses(orc_handle* h, const double* states, size_t n, uint8_t* valid, uint32_t* zone_verts);

/* Discrete motion check: valid(s2) && for j in 1..n_steps: valid(interp(s1, s2, j/(n_steps+1)))
 * (OMPL DiscreteMotionValidator semantics with a fixed segment count). */
int orc_check_motions(orc_handle* h, const double* s1, const double* s2, size_t n, int n_steps,
                      uint8_t* valid, uint32_t* zone_verts);

/* PRMMotionCost::addValidMilestone's edge interpolation (prm_motion_cost.cpp:341-372): edge e has n_interp[e] interior
 * states at t = step * (1.0 / (n_interp[e] + 1)), step = 1..n_interp[e]; valid_prefix[e] = number of leading interior
 * states that are valid (the loop stops at the first invalid one). n_interp == NULL: computed as
 * (unsigned)(lateralDistance(s1, s2) / max_lateral) like :341-343. */
int orc_check_edge_interiors(orc_handle* h, const double* s1, const double* s2, size_t n, const int32_t* n_interp,
                             double max_lateral, int32_t* valid_prefix);

/* ---- SE3FromSE2Sampler::sampleUniform (art_planner/src/sampler.cpp:40-131), liborc_port.so only ------------------
 * A deterministic restatement: the reference draws its uniforms from OMPL's RNG (std::mt19937, not in the tree); here
 * the six uniform01 variates one sample consumes are INPUTS, in the order the reference draws them:
 *   sample_from_distribution:  u0 = samp_col, u1 = samp_row (sampler.cpp:56-57), u2 -> uniformReal(-1,1) (:103),
 *                              u3,u4,u5 -> RNG::eulerRPY roll, pitch, yaw (:114; OMPL 1.4.2 RandomNumbers.cpp)
 *   otherwise:                 u0 -> x, u1 -> y in [low, high] (RealVectorStateSampler::sampleUniform; its z draw is
 *                              overwritten at :97 and is not an input), one attempt of the :46-50 loop; a position
 *                              outside the map makes the sample "rejected" (rowcol = -1, state = NaN) -- the reference
 *                              simply draws again, which leaves the distribution of accepted samples unchanged.
 * Layers are grid_map matrices (column-major rows x cols). grid_map / Eigen / OMPL are not in the reference tree:
 * parity unpinned at this level (restated from their published sources). */
typedef struct orc_sampler_map {
  const float *elevation, *normal_x, *normal_y, *normal_z, *plane_fit_std_dev;
  const float *cum_prob;           /* "cum_prob" layer (probability_distribution.cpp:20-46); NULL in uniform mode */
  const float *cum_prob_rowwise;   /* column 0 of "cum_prob_rowwise_hack": rows floats */
  int rows, cols;
  double res, cx, cy;
} orc_sampler_map;

typedef struct orc_sampler_params {
  double max_roll_pert, max_pitch_pert;   /* params.h:79-80 (radians) */
  int sample_from_distribution;           /* params.h:81 */
  double low[2], high[2];                 /* SE3 position bounds x, y (planner.cpp:148-160) */
  double reach_z;                         /* robot.feet.reach.z (sampler.cpp:103) */
} orc_sampler_params;

/* u: n x 6 doubles in [0,1); states: n x 7 doubles; rowcol (nullable): n x 2 ints (sampled cell, -1 if rejected). */
int orc_sample_states(const orc_sampler_map* m, const orc_sampler_params* p, const double* u, size_t n, double* states,
                      int32_t* rowcol);

/* art_planner::estimateNormals (art_planner/src/utils.cpp:213-324), liborc_port.so only: per-cell surface normal
 * (average of normalised cross products of axis / diagonal neighbour pairs within estimation_radius) and the
 * "plane_fit_std_dev" layer (largest |dz| seen). All float arithmetic as Eigen evaluates it (3-vectors are not
 * vectorised: cross = a1*b2-a2*b1,..; squaredNorm = x*x + (y*y + z*z); normalized() divides by sqrt when > 0).
 * Layers column-major rows x cols. */
int orc_estimate_normals(const float* elevation, int rows, int cols, double res, double cx, double cy,
                         double estimation_radius, float* normal_x, float* normal_y, float* normal_z,
                         float* plane_fit_std_dev);

/* computeCumulativeProbabilityDistribution (art_planner/src/map/processors/probability_distribution.cpp:20-46),
 * liborc_port.so only: "sample_probability" (rows x cols, column-major) -> "cum_prob" (each row divided by its sum, then
 * cumulated along the columns) and column 0 of "cum_prob_rowwise_hack" (row sums divided by their total, cumulated).
 * Sums are taken left to right (Eigen's rowwise partial reduction on a column-major matrix does the same per row; the
 * grand total uses Eigen's packet reduction in the reference, whose order depends on its build flags -- restated here
 * as the plain sequential sum: parity unpinned at this level). Rows without mass become NaN rows (0/0), as there. */
int orc_compute_cdf(const float* sample_probability, int rows, int cols, float* cum_prob, float* cum_prob_rowwise);

/* PathLengthObjective::motionCost (path_length_objective.cpp:26-70). */
int orc_path_length_cost(orc_handle* h, const double* s1, const double* s2, size_t n, double* cost);

/* Same as orc_check_poses but with T worker threads, each with its own collider state
 * ("all cores" CPU baseline). */
int orc_check_poses_mt(orc_handle* h, const double* states, size_t n, uint8_t* valid, int n_threads);

/* orc_check_motions with T worker threads (bench.py: mask check of the full configs[2] batch). */
int orc_check_motions_mt(orc_handle* h, const double* s1, const double* s2, size_t n, int n_steps, uint8_t* valid,
                         int n_threads);

/* OMPL 1.4.2 SE3StateSpace::validSegmentCount and DiscreteMotionValidator::checkMotion(s1, s2, lastValid) restated
 * (artp_wrappers.h; OMPL is not in the reference tree: parity unpinned at this level), and the MotionCostFunc edge-matrix
 * rows of PRMMotionCostMaintainer::updateEdges (prm_motion_cost.cpp:27-47). last_t[i] = lastValid.second (1.0 if valid). */
int orc_valid_segment_count(const double low[3], const double high[3], double frac, const double* s1, const double* s2, size_t n,
                            int32_t* nd);
int orc_check_motions_segments(orc_handle* h, const double* s1, const double* s2, size_t n, const int32_t* nd, uint8_t* valid,
                               double* last_t);
int orc_edge_matrix(const double* s_start, const double* s_target, size_t n, float* edges);

/* Identifies the implementation: "port" or "reference". */
const char* orc_kind(void);

#ifdef __cplusplus
}
#endif
#endif
