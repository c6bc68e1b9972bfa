// synth.cpp -- deterministic synthetic scan-pair generator (host only; bench / test input, not
// part of the registration path).  Builds libqb200_synth.so.
//
// "64-ring KITTI-shape" pairs of BASELINE.json / SURVEY.md 8(d): HDL-64E ray geometry taken from the
// reference's own sensor model (include/imageProjection.hpp:85-92: 64 rings, 1800 azimuth steps,
// bottom angle -25 deg, vertical resolution 26.9/63 deg) and range gates / sensor height from
// config/patchwork_params.yaml:1,14-15 (1.723 m, 2.7 .. 80 m).  A seeded street scene (ground plane,
// building boxes, poles, car boxes) is ray-cast analytically from two sensor poses; the target pose
// is T_gt = yaw U(-180,180) deg, |t_xy| U(0,10) m, roll/pitch U(-1,1) deg, z U(-0.1,0.1) m, so that
// p_tgt = T_gt * p_src for a static world point.  Range noise sigma = 0.02 m along the ray.
// Ground returns carry w = -1 (every other point w = +1): the reference removes ground with
// Patchwork + image projection before voxelisation (examples/run_global_registration.cpp:143-162),
// which is out of scope here, so the ray-caster's flag stands in for it and the voxel stage drops
// flagged points in its load pass (qb200_params.skip_flagged).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct Rng {  // Philox4x32-10 stream
  uint64_t seed, ctr = 0;
  uint32_t buf[4];
  int have = 0;
  explicit Rng(uint64_t s) : seed(s) {}
  static void philox(uint64_t seed, uint64_t ctr, uint32_t out[4]) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x5eed5eedu, c3 = 0;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
      const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
      const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
  uint32_t u32() {
    if (!have) { philox(seed, ctr++, buf); have = 4; }
    return buf[--have];
  }
  double uni() { return (u32() + 0.5) * (1.0 / 4294967296.0); }
  double uni(double a, double b) { return a + (b - a) * uni(); }
};

struct Box { double lo[3], hi[3]; };
struct Cyl { double cx, cy, r, z0, z1; };

struct Scene {
  std::vector<Box> boxes;
  std::vector<Cyl> cyls;
};

struct Pose {  // world_from_sensor
  double R[9];
  double t[3];
};

Pose make_pose(double yaw, double pitch, double roll, double x, double y, double z) {
  const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch), cr = std::cos(roll), sr = std::sin(roll);
  Pose P;  // Rz(yaw) * Ry(pitch) * Rx(roll)
  P.R[0] = cy * cp; P.R[1] = cy * sp * sr - sy * cr; P.R[2] = cy * sp * cr + sy * sr;
  P.R[3] = sy * cp; P.R[4] = sy * sp * sr + cy * cr; P.R[5] = sy * sp * cr - cy * sr;
  P.R[6] = -sp;     P.R[7] = cp * sr;               P.R[8] = cp * cr;
  P.t[0] = x; P.t[1] = y; P.t[2] = z;
  return P;
}

double footprint_dist(const Box& b, double x, double y) {
  const double dx = std::max({b.lo[0] - x, 0.0, x - b.hi[0]});
  const double dy = std::max({b.lo[1] - y, 0.0, y - b.hi[1]});
  return std::sqrt(dx * dx + dy * dy);
}

Scene make_outdoor_scene(Rng& g, const double sensors[2][2], int n_build, int n_pole, int n_car, double extent, int n_clutter) {
  Scene s;
  auto clear_of_sensors = [&](const Box& b, double margin) {
    for (int k = 0; k < 2; ++k)
      if (footprint_dist(b, sensors[k][0], sensors[k][1]) < margin) return false;
    return true;
  };
  int guard = 0;
  while ((int)s.boxes.size() < n_build && guard++ < 100000) {
    const double cx = g.uni(-extent, extent), cy = g.uni(-extent, extent);
    const double hx = g.uni(2.5, 12.0), hy = g.uni(2.5, 12.0), h = g.uni(4.0, 15.0);
    Box b{{cx - hx, cy - hy, 0.0}, {cx + hx, cy + hy, h}};
    if (!clear_of_sensors(b, 4.0)) continue;
    s.boxes.push_back(b);
  }
  const int n_b = (int)s.boxes.size();
  guard = 0;
  while ((int)s.boxes.size() < n_b + n_car && guard++ < 100000) {
    const double cx = g.uni(-0.6 * extent, 0.6 * extent), cy = g.uni(-0.6 * extent, 0.6 * extent);
    const bool along_x = g.uni() < 0.5;
    const double hx = along_x ? 2.1 : 0.9, hy = along_x ? 0.9 : 2.1;
    Box b{{cx - hx, cy - hy, 0.0}, {cx + hx, cy + hy, g.uni(1.4, 1.9)}};
    if (!clear_of_sensors(b, 3.0)) continue;
    s.boxes.push_back(b);
  }
  guard = 0;
  while ((int)s.cyls.size() < n_pole && guard++ < 100000) {
    Cyl c{g.uni(-0.7 * extent, 0.7 * extent), g.uni(-0.7 * extent, 0.7 * extent), g.uni(0.1, 0.4), 0.0, g.uni(3.0, 8.0)};
    bool ok = true;
    for (int k = 0; k < 2; ++k)
      if (std::hypot(c.cx - sensors[k][0], c.cy - sensors[k][1]) < 2.0) ok = false;
    if (ok) s.cyls.push_back(c);
  }
  const int n_bc = (int)s.boxes.size();
  guard = 0;
  while ((int)s.boxes.size() < n_bc + n_clutter && guard++ < 1000000) {
    const double cx = g.uni(-extent, extent), cy = g.uni(-extent, extent);
    const double hx = g.uni(0.1, 0.6), hy = g.uni(0.1, 0.6);
    Box b{{cx - hx, cy - hy, 0.0}, {cx + hx, cy + hy, g.uni(0.2, 2.5)}};
    if (!clear_of_sensors(b, 2.0)) continue;
    s.boxes.push_back(b);
  }
  return s;
}

// nearest hit distance along o + t d (t > 0); kind 0 = ground, 1 = object, -1 = none
double cast(const Scene& s, const double o[3], const double d[3], int& kind) {
  double best = 1e30;
  kind = -1;
  if (d[2] < -1e-12) {
    const double t = -o[2] / d[2];
    if (t > 0 && t < best) { best = t; kind = 0; }
  }
  for (const Box& b : s.boxes) {
    double t0 = 0.0, t1 = best;
    bool hit = true;
    for (int a = 0; a < 3 && hit; ++a) {
      if (std::fabs(d[a]) < 1e-12) {
        if (o[a] < b.lo[a] || o[a] > b.hi[a]) hit = false;
      } else {
        double ta = (b.lo[a] - o[a]) / d[a], tb = (b.hi[a] - o[a]) / d[a];
        if (ta > tb) std::swap(ta, tb);
        t0 = std::max(t0, ta); t1 = std::min(t1, tb);
        if (t0 > t1) hit = false;
      }
    }
    if (hit && t0 > 1e-9 && t0 < best) { best = t0; kind = 1; }
  }
  for (const Cyl& c : s.cyls) {
    const double ox = o[0] - c.cx, oy = o[1] - c.cy;
    const double A = d[0] * d[0] + d[1] * d[1];
    if (A < 1e-14) continue;
    const double B = ox * d[0] + oy * d[1], C = ox * ox + oy * oy - c.r * c.r;
    const double disc = B * B - A * C;
    if (disc < 0) continue;
    const double t = (-B - std::sqrt(disc)) / A;
    if (t <= 1e-9 || t >= best) continue;
    const double z = o[2] + t * d[2];
    if (z < c.z0 || z > c.z1) continue;
    best = t; kind = 1;
  }
  return best;
}

int scan(const Scene& s, const Pose& P, uint64_t noise_seed, int rings, int azimuths, double v_bottom_deg, double v_span_deg,
         double sigma, double rmin, double rmax, float* out4, int cap) {
  int n = 0;
  const double o[3] = {P.t[0], P.t[1], P.t[2]};
  for (int r = 0; r < rings; ++r) {
    const double v = (-v_bottom_deg + (rings > 1 ? v_span_deg * r / (rings - 1) : 0.0)) * M_PI / 180.0;
    const double cv = std::cos(v), sv = std::sin(v);
    for (int a = 0; a < azimuths; ++a) {
      const double az = 2.0 * M_PI * a / azimuths;
      const double dl[3] = {cv * std::cos(az), cv * std::sin(az), sv};
      const double dw[3] = {P.R[0] * dl[0] + P.R[1] * dl[1] + P.R[2] * dl[2], P.R[3] * dl[0] + P.R[4] * dl[1] + P.R[5] * dl[2],
                            P.R[6] * dl[0] + P.R[7] * dl[1] + P.R[8] * dl[2]};
      int kind;
      double t = cast(s, o, dw, kind);
      if (kind < 0) continue;
      uint32_t rb[4];
      Rng::philox(noise_seed, (uint64_t)r * azimuths + a, rb);
      const double u1 = (rb[0] + 0.5) / 4294967296.0, u2 = (rb[1] + 0.5) / 4294967296.0;
      t += sigma * std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
      if (t < rmin || t > rmax) continue;
      if (n < cap) {
        out4[4 * n + 0] = (float)(t * dl[0]);
        out4[4 * n + 1] = (float)(t * dl[1]);
        out4[4 * n + 2] = (float)(t * dl[2]);
        out4[4 * n + 3] = kind == 0 ? -1.0f : 1.0f;
      }
      ++n;
    }
  }
  return n;
}

// nearest hit from INSIDE the hull box (the ray leaves it through a wall, the floor or the ceiling) or on an interior object
double cast_indoor(const Scene& s, const Box& hull, const double o[3], const double d[3]) {
  double best = 1e30;
  for (int a = 0; a < 3; ++a) {
    if (std::fabs(d[a]) < 1e-12) continue;
    const double t = ((d[a] > 0 ? hull.hi[a] : hull.lo[a]) - o[a]) / d[a];
    if (t > 1e-9 && t < best) best = t;
  }
  int kind;
  const double t_obj = cast(s, o, d, kind);   // boxes + cylinders (the ground-plane test of cast() never wins: it lies below the floor)
  if (kind == 1 && t_obj < best) best = t_obj;
  return best;
}

int scan_indoor(const Scene& s, const Box& hull, const Pose& P, uint64_t noise_seed, int n_rays, double sigma, double rmin, float* out4, int cap) {
  int n = 0;
  const double o[3] = {P.t[0], P.t[1], P.t[2]};
  const double golden = M_PI * (3.0 - std::sqrt(5.0));
  for (int r = 0; r < n_rays; ++r) {
    // Fibonacci sphere: uniform directions in the SENSOR frame
    const double z = 1.0 - 2.0 * (r + 0.5) / n_rays, rad = std::sqrt(std::max(0.0, 1.0 - z * z)), th = golden * r;
    const double dl[3] = {rad * std::cos(th), rad * std::sin(th), z};
    const double dw[3] = {P.R[0] * dl[0] + P.R[1] * dl[1] + P.R[2] * dl[2], P.R[3] * dl[0] + P.R[4] * dl[1] + P.R[5] * dl[2],
                          P.R[6] * dl[0] + P.R[7] * dl[1] + P.R[8] * dl[2]};
    double t = cast_indoor(s, hull, o, dw);
    if (t > 1e29) continue;
    uint32_t rb[4];
    Rng::philox(noise_seed, (uint64_t)r, rb);
    const double u1 = (rb[0] + 0.5) / 4294967296.0, u2 = (rb[1] + 0.5) / 4294967296.0;
    t += sigma * std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
    if (t < rmin) continue;
    if (n < cap) {
      out4[4 * n + 0] = (float)(t * dl[0]);
      out4[4 * n + 1] = (float)(t * dl[1]);
      out4[4 * n + 2] = (float)(t * dl[2]);
      out4[4 * n + 3] = 1.0f;   // indoors the floor stays in (no Patchwork): nothing is flagged
    }
    ++n;
  }
  return n;
}

}  // namespace

extern "C" {

// Dense indoor pair (BASELINE configs[4] / SURVEY.md 8d config 5): a hall of `extent` x `extent` x 3 m with interior wall slabs
// (doorways left open), furniture boxes and columns, scanned from two poses with n_rays uniformly distributed rays each
// (sigma = 5 mm).  T_gt: column-major 4x4 with p_tgt = T_gt * p_src.  Returns 0, or 1 if cap was too small.
int qb200_synth_indoor_pair(uint64_t seed, int n_rays, double extent, int n_furniture, float* src4, int* n_src, float* tgt4, int* n_tgt, int cap,
                            double* T_gt) {
  Rng g(seed * 0x9E3779B97F4A7C15ull + 0x7654321ull);
  const double h = 1.2;
  const double yaw = g.uni(-M_PI, M_PI), dist = g.uni(0.0, 1.5), dir = g.uni(-M_PI, M_PI);
  const double roll = g.uni(-1.0, 1.0) * M_PI / 180.0, pitch = g.uni(-1.0, 1.0) * M_PI / 180.0, dz = g.uni(-0.05, 0.05);
  const Pose Ps = make_pose(0, 0, 0, 0, 0, h);
  const Pose Pt = make_pose(yaw, pitch, roll, dist * std::cos(dir), dist * std::sin(dir), h + dz);
  const double sensors[2][2] = {{Ps.t[0], Ps.t[1]}, {Pt.t[0], Pt.t[1]}};
  const Box hull{{-extent, -extent, 0.0}, {extent, extent, 3.0}};
  Scene s;
  auto clear_of_sensors = [&](const Box& b, double margin) {
    for (int k = 0; k < 2; ++k)
      if (footprint_dist(b, sensors[k][0], sensors[k][1]) < margin) return false;
    return true;
  };
  // interior walls: slabs 0.15 m thick along x or y, floor to ceiling, each with a doorway gap
  int guard = 0;
  const int n_walls = 4 + (int)(g.uni() * 4.0);
  while ((int)s.boxes.size() < 2 * n_walls && guard++ < 10000) {
    const bool along_x = g.uni() < 0.5;
    const double c = g.uni(-0.8 * extent, 0.8 * extent), a0 = g.uni(-extent, 0.0), a1 = g.uni(0.0, extent), door = g.uni(a0 + 0.5, a1 - 1.5);
    Box b1, b2;
    if (along_x) { b1 = Box{{a0, c - 0.075, 0.0}, {door, c + 0.075, 3.0}}; b2 = Box{{door + 1.0, c - 0.075, 0.0}, {a1, c + 0.075, 3.0}}; }
    else { b1 = Box{{c - 0.075, a0, 0.0}, {c + 0.075, door, 3.0}}; b2 = Box{{c - 0.075, door + 1.0, 0.0}, {c + 0.075, a1, 3.0}}; }
    if (!clear_of_sensors(b1, 0.8) || !clear_of_sensors(b2, 0.8)) continue;
    s.boxes.push_back(b1); s.boxes.push_back(b2);
  }
  const int n_wall_boxes = (int)s.boxes.size();
  guard = 0;
  while ((int)s.boxes.size() < n_wall_boxes + n_furniture && guard++ < 100000) {
    const double cx = g.uni(-extent, extent), cy = g.uni(-extent, extent);
    const double hx = g.uni(0.15, 0.9), hy = g.uni(0.15, 0.9), z0 = g.uni() < 0.8 ? 0.0 : g.uni(0.5, 1.5), hh = g.uni(0.3, 1.8);
    Box b{{cx - hx, cy - hy, z0}, {cx + hx, cy + hy, std::min(3.0, z0 + hh)}};
    if (!clear_of_sensors(b, 0.6)) continue;
    s.boxes.push_back(b);
  }
  guard = 0;
  const int n_col = n_furniture / 4;
  while ((int)s.cyls.size() < n_col && guard++ < 100000) {
    Cyl c{g.uni(-extent, extent), g.uni(-extent, extent), g.uni(0.05, 0.3), 0.0, g.uni(0.8, 3.0)};
    bool ok = true;
    for (int k = 0; k < 2; ++k)
      if (std::hypot(c.cx - sensors[k][0], c.cy - sensors[k][1]) < 0.8) ok = false;
    if (ok) s.cyls.push_back(c);
  }
  const int ns = scan_indoor(s, hull, Ps, seed * 2 + 2000003ull, n_rays, 0.005, 0.3, src4, cap);
  const int nt = scan_indoor(s, hull, Pt, seed * 2 + 2000004ull, n_rays, 0.005, 0.3, tgt4, cap);
  *n_src = std::min(ns, cap);
  *n_tgt = std::min(nt, cap);
  double Rt[9], tt[3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Rt[3 * r + c] = Pt.R[3 * c + r];
  for (int r = 0; r < 3; ++r) tt[r] = -(Rt[3 * r] * Pt.t[0] + Rt[3 * r + 1] * Pt.t[1] + Rt[3 * r + 2] * Pt.t[2]);
  for (int i = 0; i < 16; ++i) T_gt[i] = 0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T_gt[4 * c + r] = Rt[3 * r] * Ps.R[c] + Rt[3 * r + 1] * Ps.R[3 + c] + Rt[3 * r + 2] * Ps.R[6 + c];
    T_gt[12 + r] = Rt[3 * r] * Ps.t[0] + Rt[3 * r + 1] * Ps.t[1] + Rt[3 * r + 2] * Ps.t[2] + tt[r];
  }
  T_gt[15] = 1;
  return (ns > cap || nt > cap) ? 1 : 0;
}

// Scene / motion knobs of the outdoor generator.  {30, 50, 15, 70, 10, 0.02} is the street scene of BASELINE configs[1..3];
// the "dense" preset (more clutter, a revisit within ~1 m, see synth.py) yields the ~3 k correspondences per pair that
// BASELINE.json quotes for the back end.
struct qb200_synth_scene {
  int n_build, n_pole, n_car;
  double extent;      // half size of the square the objects are placed in [m]
  double max_dist;    // |t_xy| ~ U(0, max_dist)
  double sigma;       // range noise [m]
  int n_clutter;      // small boxes (0.2 .. 1.2 m footprints, up to 2.5 m high): shrubs, bins, street furniture
};

// Outdoor 64-ring pair.  rings/azimuths let tests ask for a smaller sensor (e.g. 16 x 450).
// T_gt: column-major 4x4 with p_tgt = T_gt * p_src.  Returns 0, or 1 if cap was too small.
int qb200_synth_outdoor_pair_ex(uint64_t seed, int rings, int azimuths, const qb200_synth_scene* sc, float* src4, int* n_src, float* tgt4,
                                int* n_tgt, int cap, double* T_gt) {
  Rng g(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull);
  const double h = 1.723;
  const double yaw = g.uni(-M_PI, M_PI), dist = g.uni(0.0, sc->max_dist), dir = g.uni(-M_PI, M_PI);
  const double roll = g.uni(-1.0, 1.0) * M_PI / 180.0, pitch = g.uni(-1.0, 1.0) * M_PI / 180.0, dz = g.uni(-0.1, 0.1);
  const Pose Ps = make_pose(0, 0, 0, 0, 0, h);
  const Pose Pt = make_pose(yaw, pitch, roll, dist * std::cos(dir), dist * std::sin(dir), h + dz);
  const double sensors[2][2] = {{Ps.t[0], Ps.t[1]}, {Pt.t[0], Pt.t[1]}};
  const Scene s = make_outdoor_scene(g, sensors, sc->n_build, sc->n_pole, sc->n_car, sc->extent, sc->n_clutter);
  const int ns = scan(s, Ps, seed * 2 + 1000003ull, rings, azimuths, 25.0, 26.9, sc->sigma, 2.7, 80.0, src4, cap);
  const int nt = scan(s, Pt, seed * 2 + 1000004ull, rings, azimuths, 25.0, 26.9, sc->sigma, 2.7, 80.0, tgt4, cap);
  *n_src = std::min(ns, cap);
  *n_tgt = std::min(nt, cap);
  // T_gt = Pt^-1 * Ps
  double Rt[9], tt[3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Rt[3 * r + c] = Pt.R[3 * c + r];  // Pt.R^T
  for (int r = 0; r < 3; ++r) tt[r] = -(Rt[3 * r] * Pt.t[0] + Rt[3 * r + 1] * Pt.t[1] + Rt[3 * r + 2] * Pt.t[2]);
  double R[9], t[3];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) R[3 * r + c] = Rt[3 * r] * Ps.R[c] + Rt[3 * r + 1] * Ps.R[3 + c] + Rt[3 * r + 2] * Ps.R[6 + c];
    t[r] = Rt[3 * r] * Ps.t[0] + Rt[3 * r + 1] * Ps.t[1] + Rt[3 * r + 2] * Ps.t[2] + tt[r];
  }
  for (int i = 0; i < 16; ++i) T_gt[i] = 0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T_gt[4 * c + r] = R[3 * r + c];
    T_gt[12 + r] = t[r];
  }
  T_gt[15] = 1;
  return (ns > cap || nt > cap) ? 1 : 0;
}

int qb200_synth_outdoor_pair(uint64_t seed, int rings, int azimuths, float* src4, int* n_src, float* tgt4, int* n_tgt, int cap,
                             double* T_gt) {
  const qb200_synth_scene street = {30, 50, 15, 70.0, 10.0, 0.02, 0};
  return qb200_synth_outdoor_pair_ex(seed, rings, azimuths, &street, src4, n_src, tgt4, n_tgt, cap, T_gt);
}

}  // extern "C"
