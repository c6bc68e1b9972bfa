// fpfh_manager.hpp -- source-compatible replacement of the reference's include/fpfh_manager.hpp:
// normals + FPFH-33 for both clouds, mutual-NN matching with the tuple test, packing of the matched
// pairs -- one C-ABI call (qb200_match_and_pack).  The PCD cache (save/loadFeaturePair, :179-232) is a
// "next" row and is not provided.
#pragma once

#include <iostream>
#include <stdexcept>
#include <utility>
#include <vector>

#include "quatro.hpp"

class FPFHManager {
 public:
  std::vector<std::pair<int, int>> corr;
  Eigen::Matrix3Xd src_matched, tgt_matched;
  pcl::PointCloud<PointType> src_matched_pcl, tgt_matched_pcl;

  FPFHManager(double normal_radius, double fpfh_radius, int interval = 1)
      : normal_radius_(normal_radius), fpfh_radius_(fpfh_radius), interval_(interval) {}
  FPFHManager() {}

  void flushAllFeatures() { is_initial_ = true; }
  void setParams(float normal_radius, float fpfh_radius, int interval) {
    normal_radius_ = normal_radius; fpfh_radius_ = fpfh_radius; interval_ = interval;
  }
  void clearInputs() { is_initial_ = true; corr.clear(); }
  // lattice cell of the neighbour search (only fixes the accumulation order); 0 = library default ((1 + 2^-9) fpfh_radius)
  void setGridCell(float cell) { grid_cell_ = cell; }
  void setSeed(uint64_t seed) { seed_ = seed; }  // tuple-test RNG (the reference seeds with time(NULL))

  void setFeaturePair(pcl::PointCloud<PointType>::Ptr src, pcl::PointCloud<PointType>::Ptr target) {
    if (normal_radius_ > fpfh_radius_) {  // fpfh_manager.hpp:99-102
      std::cout << normal_radius_ << " <-> " << fpfh_radius_ << std::endl;
      throw std::invalid_argument("[FPFHManager]: Normal should be lower than fpfh_radius!!!!");
    }
    qb200_handle* h = qb200::shared_handle();
    qb200_params p;
    qb200_default_params(&p);
    p.normal_radius = (float)normal_radius_;
    p.fpfh_radius = (float)fpfh_radius_;
    p.grid_cell = grid_cell_;
    p.seed = seed_;
    const int32_t cap = (int32_t)std::min(src->points.size(), target->points.size());
    std::vector<int32_t> c(2 * (size_t)std::max(cap, 1));
    std::vector<PointType> sm((size_t)std::max(cap, 1)), tm((size_t)std::max(cap, 1));
    int32_t n = 0;
    const int st = qb200_match_and_pack(h, qb200::as_float4(*src), (int32_t)src->points.size(), qb200::as_float4(*target),
                                        (int32_t)target->points.size(), &p, c.data(), reinterpret_cast<float*>(sm.data()),
                                        reinterpret_cast<float*>(tm.data()), cap, &n);
    if (st < 0) throw std::runtime_error(std::string("qb200_match_and_pack: ") + qb200_last_error(h));
    corr.resize((size_t)n);
    src_matched.resize(3, n);
    tgt_matched.resize(3, n);
    src_matched_pcl.clear();
    tgt_matched_pcl.clear();
    for (int i = 0; i < n; ++i) {
      corr[i] = {c[2 * i], c[2 * i + 1]};
      src_matched(0, i) = sm[i].x; src_matched(1, i) = sm[i].y; src_matched(2, i) = sm[i].z;
      tgt_matched(0, i) = tm[i].x; tgt_matched(1, i) = tm[i].y; tgt_matched(2, i) = tm[i].z;
      src_matched_pcl.push_back(PointType(sm[i].x, sm[i].y, sm[i].z));
      tgt_matched_pcl.push_back(PointType(tm[i].x, tm[i].y, tm[i].z));
    }
    is_initial_ = false;
  }
  Eigen::Matrix3Xd getSrcMatched() { return src_matched; }
  Eigen::Matrix3Xd getTgtMatched() { return tgt_matched; }
  pcl::PointCloud<PointType> getSrcKps() { return src_matched_pcl; }
  pcl::PointCloud<PointType> getTgtKps() { return tgt_matched_pcl; }
  std::vector<std::pair<int, int>> getCorrespondences() { return corr; }

 private:
  double normal_radius_ = 0.5, fpfh_radius_ = 0.75;
  int interval_ = 1;
  float grid_cell_ = 0.f;
  uint64_t seed_ = 0x5EED;
  bool is_initial_ = true;
};
