// patchwork.hpp -- source-compatible replacement of the reference's include/patchwork.hpp for the call the example makes:
//   patchwork.reset(new PatchWork<PointType>(...));
//   patchwork->estimate_ground(*srcRaw, srcGround, *ptrSrcNonground, tSrc);      (examples/run_global_registration.cpp:143-145)
// The reference reads its parameters from the ROS parameter server inside the constructor (patchwork.hpp:46-139); this class takes
// the same numbers as a plain struct (defaults = config/patchwork_params.yaml) -- a ROS caller fills it with its own
// nh.param(...) lines.  No computation happens on the host: estimate_ground is one qb200_patchwork call.
#pragma once

#include <chrono>
#include <stdexcept>
#include <string>
#include <vector>

#include "quatro.hpp"

template <typename PointT>
class PatchWork {
 public:
  using Params = qb200_patchwork_params;

  PatchWork() { qb200_default_patchwork_params(&params_); }
  explicit PatchWork(const Params& p) : params_(p) { check_input_parameters_are_correct(); }

  Params& params() { return params_; }
  const Params& params() const { return params_; }

  // reference: patchwork.hpp:329-455.  cloud_out = estimated ground, cloud_nonground = the rest, time_taken in seconds.
  void estimate_ground(const pcl::PointCloud<PointT>& cloud_in, pcl::PointCloud<PointT>& cloud_out,
                       pcl::PointCloud<PointT>& cloud_nonground, double& time_taken) {
    static_assert(sizeof(PointT) == 16, "point type must be a 16-byte xyz+pad record");
    check_input_parameters_are_correct();
    const auto t0 = std::chrono::steady_clock::now();
    const int32_t n = (int32_t)cloud_in.points.size();
    std::vector<PointT> g((size_t)n), ng((size_t)n);
    int32_t n_g = 0, n_ng = 0;
    const int st = qb200_patchwork(qb200::shared_handle(), qb200::as_float4(cloud_in), n, &params_, reinterpret_cast<float*>(g.data()), &n_g,
                                   reinterpret_cast<float*>(ng.data()), &n_ng);
    if (st < 0) throw std::runtime_error(std::string("qb200_patchwork: ") + qb200_last_error(qb200::shared_handle()));
    if (st == QB200_CAPACITY_EXCEEDED) throw std::runtime_error("qb200_patchwork: a patch holds more than 16384 points");
    cloud_out.points.assign(g.begin(), g.begin() + n_g);
    cloud_nonground.points.assign(ng.begin(), ng.begin() + n_ng);
    cloud_out.width = (uint32_t)n_g; cloud_out.height = 1;
    cloud_nonground.width = (uint32_t)n_ng; cloud_nonground.height = 1;
    time_taken = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }

 private:
  Params params_;

  void check_input_parameters_are_correct() const {  // patchwork.hpp:592-616, same messages
    if (params_.num_zones != 4) throw std::invalid_argument("Some parameters are wrong! the size of parameters should be same");
    if (params_.min_range != params_.min_ranges_each_zone[0])
      throw std::invalid_argument("Setting min. ranges are weired! The first term should be eqaul to min_range_");
    if (params_.num_thresholds < 0 || params_.num_thresholds > QB200_PW_MAX_THRESHOLDS)
      throw std::invalid_argument("Some parameters are wrong! Check the elevation/flatness_thresholds");
  }
};
