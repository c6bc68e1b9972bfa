// ROS-free mirror of the reference's examples/run_global_registration.cpp:94-108, 202-251: load two scans,
// voxelize, FPFHManager::setFeaturePair, Quatro::computeTransformation, print the same console table.
// The lines between the ===== markers are the caller contract of the hot path and read like the reference.
//
//   g++ -std=c++17 -Iinclude examples/run_global_registration.cpp -Lquatro_b200/lib -lquatro_b200
//       -Wl,-rpath,$PWD/quatro_b200/lib -o run_example
//   ./run_example src.bin tgt.bin        (KITTI .bin: float32 x,y,z,intensity records)
#include <chrono>
#include <cstdio>
#include <iomanip>
#include <iostream>

#include "quatro_b200/fpfh_manager.hpp"
#include "quatro_b200/quatro.hpp"

using namespace std;

void setParams(double noise_bound_of_each_measurement, double square_of_the_ratio_btw_noise_and_noise_bound, double estimating_scale,
               int num_max_iter, double control_parameter_for_gnc, double rot_cost_thr, const string& reg_type_name,
               Quatro<PointType, PointType>::Params& params) {
  params.noise_bound = noise_bound_of_each_measurement;
  params.cbar2 = square_of_the_ratio_btw_noise_and_noise_bound;
  params.estimate_scaling = estimating_scale;
  params.rotation_max_iterations = num_max_iter;
  params.rotation_gnc_factor = control_parameter_for_gnc;
  params.rotation_estimation_algorithm = Quatro<PointType, PointType>::ROTATION_ESTIMATION_ALGORITHM::GNC_TLS;
  params.rotation_cost_threshold = rot_cost_thr;
  params.reg_name = reg_type_name;
  params.inlier_selection_mode = Quatro<PointType, PointType>::INLIER_SELECTION_MODE::PMC_HEU;
}

pcl::PointCloud<PointType>::ConstPtr getCloud(std::string filename) {  // run_global_registration.cpp:377-402
  FILE* file = fopen(filename.c_str(), "rb");
  if (!file) {
    std::cerr << "error: failed to load " << filename << std::endl;
    return nullptr;
  }
  std::vector<float> buffer(1000000);
  size_t num_points = fread(reinterpret_cast<char*>(buffer.data()), sizeof(float), buffer.size(), file) / 4;
  fclose(file);
  pcl::PointCloud<PointType>::Ptr cloud(new pcl::PointCloud<PointType>());
  cloud->resize(num_points);
  for (size_t i = 0; i < num_points; i++) {
    auto& pt = cloud->at(i);
    pt.x = buffer[i * 4];
    pt.y = buffer[i * 4 + 1];
    pt.z = buffer[i * 4 + 2];
  }
  return cloud;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: " << argv[0] << " <src.bin> <tgt.bin>" << std::endl;
    return 2;
  }
  // config/params.yaml
  double voxel_size = 0.3, normal_radius = 0.5, fpfh_radius = 0.75;
  double noise_bound = 0.3, noise_bound_coeff = 1.0, gnc_factor = 1.4, rot_cost_diff_thr = 0.00011;
  bool estimating_scale = false;
  int num_max_iter = 50;

  pcl::PointCloud<PointType>::ConstPtr srcRaw = getCloud(argv[1]);
  pcl::PointCloud<PointType>::ConstPtr tgtRaw = getCloud(argv[2]);
  if (!srcRaw || !tgtRaw) return 1;

  // ===================================================================================================
  Quatro<PointType, PointType> quatro;
  Quatro<PointType, PointType>::Params params;
  setParams(noise_bound, noise_bound_coeff, estimating_scale, num_max_iter, gnc_factor, rot_cost_diff_thr, "Quatro", params);
  quatro.reset(params);

  std::chrono::system_clock::time_point start = std::chrono::system_clock::now();

  pcl::PointCloud<PointType>::Ptr srcFeat(new pcl::PointCloud<PointType>);
  pcl::PointCloud<PointType>::Ptr tgtFeat(new pcl::PointCloud<PointType>);
  voxelize(srcRaw, srcFeat, voxel_size);
  voxelize(tgtRaw, tgtFeat, voxel_size);

  FPFHManager fpfhmanager(normal_radius, fpfh_radius);
  fpfhmanager.flushAllFeatures();
  fpfhmanager.setFeaturePair(srcFeat, tgtFeat);

  pcl::PointCloud<PointType>::Ptr srcMatched(new pcl::PointCloud<PointType>);
  pcl::PointCloud<PointType>::Ptr tgtMatched(new pcl::PointCloud<PointType>);
  *srcMatched = fpfhmanager.getSrcKps();
  *tgtMatched = fpfhmanager.getTgtKps();

  cout << "# after voxelization | " << srcFeat->size() << " | " << tgtFeat->size() << endl;
  cout << "# after matching     | " << srcMatched->size() << " | " << tgtMatched->size() << endl;

  std::chrono::system_clock::time_point before_optim = std::chrono::system_clock::now();
  quatro.setInputSource(srcMatched);
  quatro.setInputTarget(tgtMatched);
  Eigen::Matrix4d output = Eigen::Matrix4d::Identity();
  quatro.computeTransformation(output);

  std::chrono::duration<double> sec = std::chrono::system_clock::now() - start;
  std::chrono::duration<double> optim_sec = std::chrono::system_clock::now() - before_optim;
  std::cout << setprecision(4) << "Total takes: " << sec.count() << " sec. (Setting matching pairs: " << sec.count() - optim_sec.count()
            << " sec. + Quatro: " << optim_sec.count() << " sec.)" << std::endl;
  // ===================================================================================================

  pcl::PointCloud<PointType> aligned;
  pcl::transformPointCloud(*srcRaw, aligned, output);
  pcl::PointCloud<PointType> srcMaxCliques, tgtMaxCliques;
  quatro.getMaxCliques(srcMaxCliques, tgtMaxCliques);
  std::cout << "valid: " << quatro.solution_.valid << "  max clique: " << quatro.getNumMaxCliqueInliers()
            << "  rotation inliers: " << quatro.getNumRotaionInliers() << "  final inliers: " << quatro.getFinalInliersIndices().size() << std::endl;
  std::cout << std::setprecision(9);
  for (int r = 0; r < 4; ++r) {
    std::cout << "T";
    for (int c = 0; c < 4; ++c) std::cout << " " << output(r, c);
    std::cout << std::endl;
  }
  return quatro.solution_.valid ? 0 : 3;
}
