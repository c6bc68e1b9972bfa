// imageProjection.hpp -- source-compatible replacement of the reference's include/imageProjection.hpp for the calls the example
// makes in "Patchwork" mode (examples/run_global_registration.cpp:124-125, 153-162):
//   ImageProjection IPSrc(lidarType, neighborSelectionMode, groundSegMode);
//   IPSrc.segmentCloud(ptrSrcNonground);  IPSrc.getValidSegments(*srcValidSegments);  IPSrc.getOutliers(srcInvalidSegments);
// The per-sensor constants of the reference's constructor (imageProjection.hpp:86-131) are reproduced; the range image, the
// labelling and the extraction run on the device (one qb200_segment_cloud call).  "LeGO-LOAM" ground removal (:365-422) is the
// reference's alternative to Patchwork and is not provided: the constructor says so.
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "quatro.hpp"

class ImageProjection {
 public:
  int N_SCAN = 64, Horizon_SCAN = 1800, groundScanInd = 60;
  float ang_res_x = 0.2f, ang_res_y = 0.427f, ang_bottom = 25.0f;
  std::string groundSegMode = "Patchwork";

  ImageProjection(const std::string& lidarType, const std::string& neighborSelectionMode, const std::string& groundSegmentationMode,
                  int numSubclusteringCriteria = 30) {
    qb200_default_segment_params(&params_);
    if (lidarType == "Velodyne-64-HDE") set(64, 1800, 360.0 / float(1800), 26.9 / float(64 - 1), 25.0, 60);
    else if (lidarType == "VLP-16") set(16, 1800, 0.2, 2.0, 15.0 + 0.1, 7);
    else if (lidarType == "HDL-32E") set(32, 1800, 360.0 / float(1800), 41.33 / float(32 - 1), 30.67, 20);
    else if (lidarType == "Ouster-OS1-16") set(16, 1024, 360.0 / float(1024), 33.2 / float(16 - 1), 16.6 + 0.1, 7);
    else if (lidarType == "Ouster-OS1-64") set(64, 1024, 360.0 / float(1024), 33.2 / float(64 - 1), 16.6 + 0.1, 15);
    else throw std::invalid_argument("[ImageProjection]:Check your paramter. Lidar Type is wrong!");
    if (neighborSelectionMode == "4Neighbor") params_.neighbor_mode = QB200_NEIGHBORS_4;
    else if (neighborSelectionMode == "8Neighbor") params_.neighbor_mode = QB200_NEIGHBORS_8;
    else if (neighborSelectionMode == "4CrossNeighbor") params_.neighbor_mode = QB200_NEIGHBORS_4_CROSS;
    else throw std::invalid_argument("[ImageProjection]:Check your paramter. Neighbor selection mode is wrong!");
    groundSegMode = groundSegmentationMode;
    if (groundSegMode == "LeGO-LOAM")
      throw std::invalid_argument("[ImageProjection]: the device path provides the \"Patchwork\" ground mode only (run PatchWork::estimate_ground first)");
    if (groundSegMode != "Patchwork") throw std::invalid_argument("[ImageProjection]: Check your paramter. Ground Segmentation mode is wrong!");
    params_.min_pts_for_subclustering = numSubclusteringCriteria;
  }

  const qb200_segment_params& params() const { return params_; }

  // imageProjection.hpp:273-294
  template <class CloudPtr>
  void segmentCloud(const CloudPtr& pcPtr) {
    const auto& cloud = *pcPtr;
    const size_t npix = (size_t)N_SCAN * (size_t)Horizon_SCAN;
    valid_.assign(npix, PointType());
    outliers_.assign(npix, PointType());
    int32_t nv = 0, no = 0;
    const int st = qb200_segment_cloud(qb200::shared_handle(), qb200::as_float4(cloud), (int32_t)cloud.points.size(), &params_,
                                       reinterpret_cast<float*>(valid_.data()), &nv, reinterpret_cast<float*>(outliers_.data()), &no);
    if (st < 0) throw std::runtime_error(std::string("qb200_segment_cloud: ") + qb200_last_error(qb200::shared_handle()));
    valid_.resize((size_t)nv);
    outliers_.resize((size_t)no);
  }

  void getValidSegments(pcl::PointCloud<PointType>& output) const { fill(output, valid_); }   // :214-216
  void getOutliers(pcl::PointCloud<PointType>& output) const { fill(output, outliers_); }       // :230-232

 private:
  qb200_segment_params params_;
  std::vector<PointType> valid_, outliers_;

  void set(int n_scan, int horizon, double rx, double ry, double bottom, int ground_ind) {
    N_SCAN = n_scan; Horizon_SCAN = horizon; groundScanInd = ground_ind;
    ang_res_x = (float)rx; ang_res_y = (float)ry; ang_bottom = (float)bottom;
    params_.n_scan = n_scan; params_.horizon_scan = horizon;
    params_.ang_res_x = ang_res_x; params_.ang_res_y = ang_res_y; params_.ang_bottom = ang_bottom;
  }
  static void fill(pcl::PointCloud<PointType>& out, const std::vector<PointType>& pts) {
    out.points.assign(pts.begin(), pts.end());
    out.width = (uint32_t)pts.size();
    out.height = 1;
  }
};
