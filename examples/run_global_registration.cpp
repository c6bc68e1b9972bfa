// ROS-free mirror of the reference's examples/run_global_registration.cpp:94-108, 202-251: load two scans,
// voxelize, FPFHManager::setFeaturePair, Quatro::computeTransformation, print the same console table.
// The lines between the ===== markers are the caller contract of the hot path and read like the reference.
//
//   g++ -std=c++17 -Iinclude examples/run_global_registration.cpp -Lquatro_b200/lib -lquatro_b200
//       -Wl,-rpath,$PWD/quatro_b200/lib -o run_example
//   ./run_example src.bin tgt.bin              (KITTI .bin: float32 x,y,z,intensity records; the scans are used as they are)
//   ./run_example src.bin tgt.bin --preprocess (ground removal + sub-cluster rejection first, the reference's STEP 2 / STEP 3)
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>

#include "quatro_b200/fpfh_manager.hpp"
#include "quatro_b200/imageProjection.hpp"
#include "quatro_b200/patchwork.hpp"
#include "quatro_b200/quatro.hpp"

using namespace std;

using QuatroReg = Quatro<PointType, PointType>;

// Solver parameters of the "Quatro" configuration of the reference's caller (run_global_registration.cpp:357-375 selects GNC-TLS
// rotation + PMC_HEU inlier selection; the numbers come from config/params.yaml).
static QuatroReg::Params quatro_params(double noise_bound, double noise_bound_coeff, bool estimating_scale, int num_max_iter,
                                       double gnc_factor, double rot_cost_diff_thr) {
  QuatroReg::Params prm;
  prm.reg_name = "Quatro";
  prm.noise_bound = noise_bound;
  prm.cbar2 = noise_bound_coeff;
  prm.estimate_scaling = estimating_scale;
  prm.rotation_estimation_algorithm = QuatroReg::ROTATION_ESTIMATION_ALGORITHM::GNC_TLS;
  prm.rotation_max_iterations = (size_t)num_max_iter;
  prm.rotation_gnc_factor = gnc_factor;
  prm.rotation_cost_threshold = rot_cost_diff_thr;
  prm.inlier_selection_mode = QuatroReg::INLIER_SELECTION_MODE::PMC_HEU;
  return prm;
}

// KITTI velodyne .bin: float32 records (x, y, z, intensity).  Like the reference's loader (:377-402) at most 250 000 records are read.
static pcl::PointCloud<PointType>::ConstPtr read_kitti_bin(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) {
    std::cerr << "error: failed to load " << path << std::endl;
    return nullptr;
  }
  constexpr size_t kMaxRecords = 250000;
  std::vector<float> rec(4 * kMaxRecords);
  in.read(reinterpret_cast<char*>(rec.data()), (std::streamsize)(rec.size() * sizeof(float)));
  const size_t n = (size_t)in.gcount() / (4 * sizeof(float));
  auto cloud = std::make_shared<pcl::PointCloud<PointType>>();
  cloud->reserve(n);
  for (size_t i = 0; i < n; ++i) cloud->push_back(PointType(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2]));
  return cloud;
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: " << argv[0] << " <src.bin> <tgt.bin> [--preprocess]" << std::endl;
    return 2;
  }
  // config/params.yaml
  double voxel_size = 0.3, normal_radius = 0.5, fpfh_radius = 0.75;
  double noise_bound = 0.3, noise_bound_coeff = 1.0, gnc_factor = 1.4, rot_cost_diff_thr = 0.00011;
  bool estimating_scale = false;
  int num_max_iter = 50;

  pcl::PointCloud<PointType>::ConstPtr srcRaw = read_kitti_bin(argv[1]);
  pcl::PointCloud<PointType>::ConstPtr tgtRaw = read_kitti_bin(argv[2]);
  if (!srcRaw || !tgtRaw) return 1;

  if (argc > 3 && std::string(argv[3]) == "--preprocess") {
    // ===== reference examples/run_global_registration.cpp:124-162 ("Patchwork" ground mode) =====
    std::string lidarType = "Velodyne-64-HDE", neighborSelectionMode = "4CrossNeighbor", groundSegMode = "Patchwork";
    pcl::PointCloud<PointType> srcGround, tgtGround, srcInvalidSegments, tgtInvalidSegments;
    pcl::PointCloud<PointType>::Ptr ptrSrcNonground(new pcl::PointCloud<PointType>), ptrTgtNonground(new pcl::PointCloud<PointType>);
    pcl::PointCloud<PointType>::Ptr srcValidSegments(new pcl::PointCloud<PointType>), tgtValidSegments(new pcl::PointCloud<PointType>);
    ImageProjection IPSrc(lidarType, neighborSelectionMode, groundSegMode);
    ImageProjection IPTgt(lidarType, neighborSelectionMode, groundSegMode);
    std::unique_ptr<PatchWork<PointType>> patchwork;
    double tSrc = 0, tTgt = 0;
    patchwork.reset(new PatchWork<PointType>());
    patchwork->estimate_ground(*(srcRaw), srcGround, *ptrSrcNonground, tSrc);
    patchwork->estimate_ground(*(tgtRaw), tgtGround, *ptrTgtNonground, tTgt);
    IPSrc.segmentCloud(ptrSrcNonground);
    IPTgt.segmentCloud(ptrTgtNonground);
    IPSrc.getValidSegments(*srcValidSegments);
    IPTgt.getValidSegments(*tgtValidSegments);
    IPSrc.getOutliers(srcInvalidSegments);
    IPTgt.getOutliers(tgtInvalidSegments);
    cout << "# of raw cloud       | " << srcRaw->size() << " | " << tgtRaw->size() << endl;
    cout << "# of ground          | " << srcGround.size() << " | " << tgtGround.size() << endl;
    cout << "# of valid segments  | " << srcValidSegments->size() << " | " << tgtValidSegments->size() << endl;
    cout << "# of outliers        | " << srcInvalidSegments.size() << " | " << tgtInvalidSegments.size() << endl;
    srcRaw = srcValidSegments;
    tgtRaw = tgtValidSegments;
  }

  // ===================================================================================================
  Quatro<PointType, PointType> quatro;
  Quatro<PointType, PointType>::Params params = quatro_params(noise_bound, noise_bound_coeff, estimating_scale, num_max_iter, gnc_factor, rot_cost_diff_thr);
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
