// tests/host_cpp/host_check.cpp -- drives the C++ host mirror (include/artp_host.hpp) the way art_planner's facade
// drives its plugins: construct the checker from Params, setMap + updateHeightField, isValid / checkMotion / motionCost.
//   host_check --expect-no-gpu            : construction must fail loudly (no CPU fallback)
//   host_check <in.bin> <out.bin>         : run the cases in in.bin, write the results (see tests/test_host_cpp.py)
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "artp_host.hpp"

using namespace artp_host;

template <class T> static void rd(std::ifstream& f, T* p, size_t n) { f.read(reinterpret_cast<char*>(p), sizeof(T) * n); }
template <class T> static void wr(std::ofstream& f, const T* p, size_t n) { f.write(reinterpret_cast<const char*>(p), sizeof(T) * n); }

int main(int argc, char** argv) {
  auto params = std::make_shared<Params>();
  if (argc == 2 && !std::strcmp(argv[1], "--expect-no-gpu")) {
    try {
      StateValidityChecker c(params);
    } catch (const std::runtime_error& e) {
      std::cout << "failed loudly: " << e.what() << "\n";
      return 0;
    }
    std::cout << "a handle was created: a CUDA device is present\n";
    return 3;
  }
  if (argc != 3) { std::cerr << "usage\n"; return 2; }
  std::ifstream in(argv[1], std::ios::binary);
  int32_t hdr[6];   // rows, cols, n_poses, n_edges, n_segments, preset (0 yaml, 1 header defaults)
  double geo[3];    // res, cx, cy
  rd(in, hdr, 6); rd(in, geo, 3);
  auto map = std::make_shared<Map>();
  map->rows = hdr[0]; map->cols = hdr[1]; map->resolution = geo[0]; map->position_x = geo[1]; map->position_y = geo[2];
  map->elevation.resize((size_t)hdr[0] * hdr[1]); map->elevation_masked.resize(map->elevation.size());
  rd(in, map->elevation.data(), map->elevation.size()); rd(in, map->elevation_masked.data(), map->elevation_masked.size());
  std::vector<State> poses(hdr[2]), s1(hdr[3]), s2(hdr[3]);
  rd(in, poses.data(), poses.size()); rd(in, s1.data(), s1.size()); rd(in, s2.data(), s2.size());
  const size_t ncell = map->elevation.size();
  for (auto* layer : {&map->normal_x, &map->normal_y, &map->normal_z, &map->plane_fit_std_dev, &map->cum_prob}) {
    layer->resize(ncell); rd(in, layer->data(), ncell);
  }
  map->cum_prob_rowwise.resize(hdr[0]); rd(in, map->cum_prob_rowwise.data(), map->cum_prob_rowwise.size());
  if (hdr[5] == 0) {   // art_planner_ros/config/params.yaml:55-71
    params->robot.torso.length = 1.31; params->robot.torso.width = 0.65; params->robot.torso.height = 0.30;
    params->robot.torso.offset.z = 0.04;
    params->robot.feet.offset.x = 0.51; params->robot.feet.offset.y = 0.20; params->robot.feet.offset.z = -0.475;
    params->robot.feet.reach.x = 0.2; params->robot.feet.reach.y = 0.2; params->robot.feet.reach.z = 0.2;
    params->objectives.custom_path_length.use_directional_cost = true;
    params->planner.prm_motion_cost.risk_threshold = 0.5f;
  }
  if (hdr[5] == 1) params->planner.prm_motion_cost.risk_threshold = 0.55f;   // seeded random weights give risks around 0.5
  auto checker = std::make_shared<StateValidityChecker>(params);
  if (checker->hasMap()) return 4;
  checker->setMap(map);
  checker->updateHeightField();
  if (!checker->hasMap()) return 5;
  std::vector<uint8_t> valid, single(std::min<size_t>(poses.size(), 16)), motion, motion1(std::min<size_t>(s1.size(), 8));
  checker->isValidBatch(poses, &valid);
  for (size_t i = 0; i < single.size(); ++i) single[i] = checker->isValid(&poses[i]);
  MotionValidator mv(checker, hdr[4]);
  mv.checkMotionBatch(s1, s2, &motion);
  for (size_t i = 0; i < motion1.size(); ++i) motion1[i] = mv.checkMotion(&s1[i], &s2[i]);
  PathLengthObjective plo(checker);
  std::vector<double> cost;
  plo.motionCostBatch(s1, s2, &cost);
  std::vector<int32_t> n_interp, prefix;
  mv.checkEdgeInteriors(s1, s2, 0.5, &n_interp, &prefix);                    // prm_motion_cost.cpp:341-372
  size_t cursor = 0;                                                          // "sampler": replays the pose file in order
  std::vector<State> sampled;
  const uint64_t drawn = checker->sampleValidBatch([&](State* st) { *st = poses[cursor++ % poses.size()]; },
                                                   /*n_wanted=*/100, /*batch=*/64, /*max_draws=*/poses.size(), &sampled);
  const uint64_t n_sampled = sampled.size();
  // SE3FromSE2Sampler: 63 + 1 candidates of the stream, then a fused sample -> check -> compact batch
  const double Lx = map->rows * map->resolution, Ly = map->cols * map->resolution;
  const double low[2] = {map->position_x - Lx, map->position_y - Ly}, high[2] = {map->position_x + Lx, map->position_y + Ly};
  SE3FromSE2Sampler smp(checker, map, /*seed=*/99, low, high);
  std::vector<State> drawn64, accepted;
  smp.sampleUniformBatch(63, &drawn64);
  drawn64.emplace_back(); smp.sampleUniform(&drawn64.back());
  smp.sampleValidBatch(2000, &accepted);
  const uint64_t n_accepted = accepted.size(), next_index = smp.nextIndex();
  std::ofstream out(argv[2], std::ios::binary);
  wr(out, valid.data(), valid.size()); wr(out, single.data(), single.size());
  wr(out, motion.data(), motion.size()); wr(out, motion1.data(), motion1.size()); wr(out, cost.data(), cost.size());
  wr(out, n_interp.data(), n_interp.size()); wr(out, prefix.data(), prefix.size());
  wr(out, &drawn, 1); wr(out, &n_sampled, 1);
  if (n_sampled) wr(out, &sampled[0].x, 7 * n_sampled);
  wr(out, &drawn64[0].x, 7 * drawn64.size()); wr(out, &n_accepted, 1); wr(out, &next_index, 1);
  if (n_accepted) wr(out, &accepted[0].x, 7 * n_accepted);
  // OMPL's per-edge segment rule + lastValid, then (if the input carries network weights) the learned edge cost:
  // MotionCostObjective::motionCost with its edge splitting and the updateEdges batch
  {
    artp_se3_space sp{};
    double bnd[6];
    rd(in, bnd, 6);
    for (int i = 0; i < 3; ++i) { sp.low[i] = bnd[i]; sp.high[i] = bnd[3 + i]; }
    sp.longest_valid_segment_fraction = 0.01;
    std::vector<uint8_t> seg_valid; std::vector<double> seg_t; std::vector<int32_t> nd;
    mv.checkMotionSegments(s1, s2, sp, &seg_valid, &seg_t, &nd);
    double t0 = -1.0; State lv{};
    const uint8_t one = mv.checkMotion(&s1[0], &s2[0], &t0, &lv) ? 1 : 0;      // nseg segments, single edge + lastValid state
    wr(out, nd.data(), nd.size()); wr(out, seg_valid.data(), seg_valid.size()); wr(out, seg_t.data(), seg_t.size());
    wr(out, &one, 1); wr(out, &t0, 1); wr(out, &lv.x, 7);
  }
  int32_t nw = 0;
  rd(in, &nw, 1);
  if (in && nw > 0) {
    std::vector<float> blob(nw);
    rd(in, blob.data(), blob.size());
    int32_t ne = 0;
    rd(in, &ne, 1);
    MotionCostObjective mco(checker);
    mco.setWeights(blob);
    mco.updateFeatures();
    std::vector<double> mc(ne);
    for (int i = 0; i < ne; ++i) mc[i] = mco.motionCost(&s1[i], &s2[i]);        // motion_cost_objective.cpp:36-95
    std::vector<double> ec; std::vector<uint8_t> ef;
    const uint8_t ok = mco.updateEdgesBatch(s1, s2, &ec, &ef) ? 1 : 0;          // prm_motion_cost.cpp:27-73
    wr(out, mc.data(), mc.size()); wr(out, &ok, 1); wr(out, ec.data(), ec.size()); wr(out, ef.data(), ef.size());
  }
  std::cout << "ok " << valid.size() << " poses, " << motion.size() << " edges\n";
  return 0;
}
