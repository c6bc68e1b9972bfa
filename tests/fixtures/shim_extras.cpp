// GPU check of the FPFHManager surface beyond setFeaturePair (reference include/fpfh_manager.hpp:74-77, 111-118, 161-232):
// descriptor / normal getters, the odometry reuse of the previous target as the next source, the matched-pair PCD cache.
//   shim_extras a.bin b.bin c.bin tmpdir      (KITTI-style float32 xyzw records)
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "quatro_b200/fpfh_manager.hpp"

static pcl::PointCloud<PointType>::Ptr load(const char* path) {
  std::ifstream in(path, std::ios::binary);
  std::vector<float> rec((std::istreambuf_iterator<char>(in)), {});
  in.close();
  std::ifstream in2(path, std::ios::binary | std::ios::ate);
  const size_t bytes = (size_t)in2.tellg();
  in2.seekg(0);
  std::vector<float> v(bytes / 4);
  in2.read(reinterpret_cast<char*>(v.data()), (std::streamsize)bytes);
  auto c = std::make_shared<pcl::PointCloud<PointType>>();
  for (size_t i = 0; i + 3 < v.size(); i += 4) c->push_back(PointType(v[i], v[i + 1], v[i + 2]));
  return c;
}

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { std::cerr << "CHECK failed: " #cond << " (line " << __LINE__ << ")" << std::endl; return 1; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  pcl::PointCloud<PointType>::Ptr raw[3], vox[3];
  for (int i = 0; i < 3; ++i) {
    raw[i] = load(argv[1 + i]);
    vox[i].reset(new pcl::PointCloud<PointType>);
    voxelize(raw[i], vox[i], 0.3);
    CHECK(vox[i]->size() > 100);
  }
  FPFHManager m(0.5, 0.75);
  m.flushAllFeatures();
  m.setFeaturePair(vox[0], vox[1]);
  const auto corr01 = m.getCorrespondences();
  CHECK(!corr01.empty());
  // getters: one descriptor per voxel point, one normal per correspondence
  const auto obj = m.getObjDescriptor(), scene = m.getSceneDescriptor();
  CHECK(obj.size() == vox[0]->size() && scene.size() == vox[1]->size());
  double sum = 0;
  for (int b = 0; b < 11; ++b) sum += obj.points[obj.size() / 2].histogram[b];
  CHECK(sum == 0.0 || std::fabs(sum - 100.0) < 1e-2);
  const auto tn = m.getTgtNormals();
  CHECK(tn.cols() == (long)corr01.size() && m.getSrcMatched().cols() == (long)corr01.size());
  int unit = 0;
  for (long i = 0; i < tn.cols(); ++i) {
    const double n2 = tn(0, i) * tn(0, i) + tn(1, i) * tn(1, i) + tn(2, i) * tn(2, i);
    if (std::fabs(n2 - 1.0) < 1e-3) ++unit;
  }
  CHECK(unit * 10 >= (int)tn.cols() * 8);  // NaN normals (< 3 neighbours) aside
  // PCD cache round trip
  m.setSaveDir(argv[4]);
  m.setLoadDir(argv[4]);
  const auto src_kps = m.getSrcKps(), tgt_kps = m.getTgtKps();
  m.saveFeaturePair(540, 1319);
  FPFHManager m2(0.5, 0.75);
  m2.setLoadDir(argv[4]);
  m2.loadFeaturePair(540, 1319);
  CHECK(m2.getSrcKps().size() == src_kps.size() && m2.getTgtKps().size() == tgt_kps.size());
  for (size_t i = 0; i < src_kps.size(); ++i)
    CHECK(m2.getSrcKps().points[i].x == src_kps.points[i].x && m2.getTgtKps().points[i].z == tgt_kps.points[i].z);
  // odometry chain: without flushAllFeatures() the previous target is the next source, whatever is passed as src
  m.setFeaturePair(vox[0], vox[2]);
  const auto chained = m.getCorrespondences();
  const auto chained_obj = m.getObjDescriptor();
  FPFHManager fresh(0.5, 0.75);
  fresh.flushAllFeatures();
  fresh.setFeaturePair(vox[1], vox[2]);
  CHECK(chained == fresh.getCorrespondences());
  CHECK(chained_obj.size() == vox[1]->size());
  for (int b = 0; b < 33; ++b) CHECK(chained_obj.points[7].histogram[b] == scene.points[7].histogram[b]);
  std::cout << "SHIM_EXTRAS_OK " << corr01.size() << " " << chained.size() << std::endl;
  return 0;
}
