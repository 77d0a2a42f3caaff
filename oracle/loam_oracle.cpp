// TEST INFRASTRUCTURE -- CPU restatement of the reference's per-sweep registration hot path.  NOT shipped code:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load liboracle.so.
//
// Every function cites the reference lines it restates (paths relative to laboshinl/loam_velodyne @ 25db5dd).
// Third-party arithmetic that the reference does not vendor is restated in oracle/shim (PCL VoxelGrid, Eigen
// ColPivHouseholderQR / SelfAdjointEigenSolver / inverse) -- parity for those boundaries is "unpinned" upstream.
// The k-NN here is an own exact k-d tree with nanoflann's observable semantics (ascending fp32 L2 accumulated
// x->y->z, strict '<' admission, nanoflann.hpp:115-139,372-379,1354-1412); the vendored nanoflann itself is only
// linked into oracle/_ref/libloam_ref.so, built from /root/reference where that exists.
//
// Pinning: tests/test_oracle.py checks this restatement against oracle/_ref (the unmodified reference sources) on
// the synthetic sweeps bit for bit wherever k-NN ties do not occur, and against tests/golden/*.npz recorded from
// oracle/_ref.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include <Eigen/Eigenvalues>
#include <Eigen/QR>
#include <pcl/filters/voxel_grid.h>
#include <pcl/point_cloud.h>

#include "driver_api.h"

namespace orc {

typedef pcl::PointXYZI Pt;
typedef pcl::PointCloud<Pt> Cloud;

// ------------------------------------------------------------------------------------------------ value types
// Angle.h:16-67 -- float radians with cached float cos/sin; negation flips the sine only.
struct Ang {
  float rad, c, s;
  Ang() : rad(0.f), c(1.f), s(0.f) {}
  Ang(float r) : rad(r), c(std::cos(r)), s(std::sin(r)) {}
  Ang neg() const { Ang a; a.rad = -rad; a.c = c; a.s = -s; return a; }
};
struct V3 { float x, y, z; };
struct Pose {  // Twist.h:15-27
  Ang rx, ry, rz;
  V3 t;
  Pose() { t.x = t.y = t.z = 0.f; }
};

// math_utils.h:129-201 -- elementary rotations, math_utils.h:212-275 -- their ZXY / YXZ chains
template <typename P> inline void spinX(P& p, const Ang& a) { float y = p.y; p.y = a.c * y - a.s * p.z; p.z = a.s * y + a.c * p.z; }
template <typename P> inline void spinY(P& p, const Ang& a) { float x = p.x; p.x = a.c * x + a.s * p.z; p.z = a.c * p.z - a.s * x; }
template <typename P> inline void spinZ(P& p, const Ang& a) { float x = p.x; p.x = a.c * x - a.s * p.y; p.y = a.s * x + a.c * p.y; }
template <typename P> inline void zxy(P& p, const Ang& z, const Ang& x, const Ang& y) { spinZ(p, z); spinX(p, x); spinY(p, y); }
template <typename P> inline void yxz(P& p, const Ang& y, const Ang& x, const Ang& z) { spinY(p, y); spinX(p, x); spinZ(p, z); }

// math_utils.h:68-76, 87-95, 103-107, 116-120
template <typename A, typename B> inline float d2(const A& a, const B& b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}
inline float d2w(const Pt& a, const Pt& b, float wb) {
  float dx = a.x - b.x * wb, dy = a.y - b.y * wb, dz = a.z - b.z * wb;
  return dx * dx + dy * dy + dz * dz;
}
inline float norm(const Pt& p) { return std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z); }
inline float norm2(const Pt& p) { return p.x * p.x + p.y * p.y + p.z * p.z; }
inline float deg(float r) { return (float)(r * 180.0 / M_PI); }  // math_utils.h:30-33

// ------------------------------------------------------------------------------------------------ exact k-NN
// Observable behaviour of nanoflann::KdTreeFLANN<PointXYZI>::nearestKSearch (nanoflann_pcl.h:141-152).
class KdTree {
 public:
  void build(const Cloud* c) {
    cloud_ = c;
    nodes_.clear();
    const int n = c ? (int)c->points.size() : 0;
    order_.resize(n);
    for (int i = 0; i < n; i++) order_[i] = i;
    if (n > 0) {
      nodes_.reserve(2 * (n / 8 + 1));
      split(0, n);
    }
  }
  int size() const { return (int)order_.size(); }
  // fills idx/dist (capacity k) ascending; returns count found
  int knn(const Pt& q, int k, int* idx, float* dist) const {
    for (int i = 0; i < k; i++) { idx[i] = -1; dist[i] = 0.f; }
    if (order_.empty()) return 0;
    dist[k - 1] = std::numeric_limits<float>::max();  // KNNResultSet::init (nanoflann.hpp:92-98)
    int count = 0;
    const float qv[3] = {q.x, q.y, q.z};
    walk(0, qv, k, idx, dist, count);
    return count;
  }

 private:
  struct Node { int lo, hi, axis, left, right; float cut_lo, cut_hi; };
  float coord(int i, int a) const { const Pt& p = cloud_->points[i]; return a == 0 ? p.x : a == 1 ? p.y : p.z; }
  int split(int lo, int hi) {
    const int id = (int)nodes_.size();
    nodes_.push_back(Node{lo, hi, -1, -1, -1, 0.f, 0.f});
    if (hi - lo <= 10) return id;  // leaf_max_size 10 (nanoflann.hpp:496)
    float mn[3], mx[3];
    for (int a = 0; a < 3; a++) { mn[a] = mx[a] = coord(order_[lo], a); }
    for (int i = lo + 1; i < hi; i++)
      for (int a = 0; a < 3; a++) { float v = coord(order_[i], a); mn[a] = std::min(mn[a], v); mx[a] = std::max(mx[a], v); }
    int axis = 0;
    for (int a = 1; a < 3; a++) if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
    const int mid = (lo + hi) / 2;
    std::nth_element(order_.begin() + lo, order_.begin() + mid, order_.begin() + hi,
                     [&](int a, int b) { return coord(a, axis) < coord(b, axis); });
    float left_max = coord(order_[lo], axis), right_min = coord(order_[mid], axis);
    for (int i = lo; i < mid; i++) left_max = std::max(left_max, coord(order_[i], axis));
    for (int i = mid; i < hi; i++) right_min = std::min(right_min, coord(order_[i], axis));
    const int l = split(lo, mid);
    const int r = split(mid, hi);
    nodes_[id].axis = axis; nodes_[id].left = l; nodes_[id].right = r;
    nodes_[id].cut_lo = left_max; nodes_[id].cut_hi = right_min;
    return id;
  }
  // admission rule of KNNResultSet::addPoint (nanoflann.hpp:115-139): strictly closer than the current worst,
  // inserted after every stored entry that is not farther
  static void admit(float d, int index, int k, int* idx, float* dist, int& count) {
    int i;
    for (i = count; i > 0; --i) {
      if (dist[i - 1] > d) {
        if (i < k) { dist[i] = dist[i - 1]; idx[i] = idx[i - 1]; }
      } else break;
    }
    if (i < k) { dist[i] = d; idx[i] = index; }
    if (count < k) count++;
  }
  void walk(int id, const float* q, int k, int* idx, float* dist, int& count) const {
    const Node& nd = nodes_[id];
    if (nd.axis < 0) {
      const float worst = dist[k - 1];
      for (int i = nd.lo; i < nd.hi; i++) {
        const Pt& p = cloud_->points[order_[i]];
        // L2_Simple_Adaptor::evalMetric (nanoflann.hpp:372-379): result += diff*diff for x, y, z in order
        float r = 0.f;
        float df = q[0] - p.x; r += df * df;
        df = q[1] - p.y; r += df * df;
        df = q[2] - p.z; r += df * df;
        if (r < worst) admit(r, order_[i], k, idx, dist, count);
      }
      return;
    }
    const float v = q[nd.axis];
    const float dl = v - nd.cut_lo, dh = v - nd.cut_hi;
    int first, second;
    float cut;
    if (dl + dh < 0) { first = nd.left; second = nd.right; cut = (v - nd.cut_hi) * (v - nd.cut_hi); }
    else { first = nd.right; second = nd.left; cut = (v - nd.cut_lo) * (v - nd.cut_lo); }
    walk(first, q, k, idx, dist, count);
    // conservative bound: distance to the separating slab along this axis only
    const bool far_possible = (first == nd.left) ? (v >= nd.cut_hi ? true : cut <= dist[k - 1])
                                                 : (v <= nd.cut_lo ? true : cut <= dist[k - 1]);
    if (far_possible) walk(second, q, k, idx, dist, count);
  }
  const Cloud* cloud_ = nullptr;
  std::vector<int> order_;
  std::vector<Node> nodes_;
};

// ------------------------------------------------------------------------------------------------ scan registration
struct RegCfg {  // RegistrationParams, BasicScanRegistration.h:37-71 / BasicScanRegistration.cpp:9-26
  float scanPeriod = 0.1f;
  int nFeatureRegions = 6, curvatureRegion = 5, maxCornerSharp = 2, maxCornerLessSharp = 20, maxSurfaceFlat = 4;
  float lessFlatFilterSize = 0.2f, surfaceCurvatureThreshold = 0.1f;
};

struct ScanReg {
  RegCfg cfg;
  Cloud full, sharp, lessSharp, flat, lessFlat;
  std::vector<std::pair<size_t, size_t> > ranges;
  std::vector<float> curv;
  std::vector<int> label;
  std::vector<size_t> sorted;
  std::vector<int> picked;

  // BasicScanRegistration.cpp:367-386
  void suppress(size_t cloudIdx, size_t scanIdx) {
    picked[scanIdx] = 1;
    for (int i = 1; i <= cfg.curvatureRegion; i++) {
      if (d2(full[cloudIdx + i], full[cloudIdx + i - 1]) > 0.05) break;
      picked[scanIdx + i] = 1;
    }
    for (int i = 1; i <= cfg.curvatureRegion; i++) {
      if (d2(full[cloudIdx - i], full[cloudIdx - i + 1]) > 0.05) break;
      picked[scanIdx - i] = 1;
    }
  }

  // BasicScanRegistration.cpp:321-363
  void maskUnreliable(size_t s, size_t e) {
    picked.assign(e - s + 1, 0);
    const int cr = cfg.curvatureRegion;
    for (size_t i = s + cr; i < e - cr; i++) {
      const Pt& prev = full[i - 1];
      const Pt& cur = full[i];
      const Pt& nxt = full[i + 1];
      float diffNext = d2(nxt, cur);
      if (diffNext > 0.1) {
        float depth1 = norm(cur), depth2 = norm(nxt);
        if (depth1 > depth2) {
          float wd = std::sqrt(d2w(nxt, cur, depth2 / depth1)) / depth2;
          if (wd < 0.1) {
            std::fill_n(&picked[i - s - cr], cr + 1, 1);
            continue;
          }
        } else {
          float wd = std::sqrt(d2w(cur, nxt, depth1 / depth2)) / depth1;
          if (wd < 0.1) std::fill_n(&picked[i - s + 1], cr + 1, 1);
        }
      }
      float diffPrev = d2(cur, prev);
      float dis = norm2(cur);
      if (diffNext > 0.0002 * dis && diffPrev > 0.0002 * dis) picked[i - s] = 1;
    }
  }

  // BasicScanRegistration.cpp:284-318: curvature then the stable ascending order the no-early-exit insertion sort yields
  void prepareRegion(size_t sp, size_t ep) {
    const size_t n = ep - sp + 1;
    curv.resize(n);
    sorted.resize(n);
    label.assign(n, 0);
    const float w = -2 * cfg.curvatureRegion;
    for (size_t i = sp, r = 0; i <= ep; i++, r++) {
      float dx = w * full[i].x, dy = w * full[i].y, dz = w * full[i].z;
      for (int j = 1; j <= cfg.curvatureRegion; j++) {
        dx += full[i + j].x + full[i - j].x;
        dy += full[i + j].y + full[i - j].y;
        dz += full[i + j].z + full[i - j].z;
      }
      curv[r] = dx * dx + dy * dy + dz * dz;
      sorted[r] = i;
    }
    // equivalent to the reference's bubble-style insertion with strict '<': stable sort by curvature
    std::stable_sort(sorted.begin(), sorted.end(), [&](size_t a, size_t b) { return curv[a - sp] < curv[b - sp]; });
  }

  // BasicScanRegistration.cpp:28-46 + 155-254
  void process(const std::vector<Cloud>& rings) {
    full.clear(); sharp.clear(); lessSharp.clear(); flat.clear(); lessFlat.clear();
    ranges.clear();
    size_t cloudSize = 0;
    for (size_t i = 0; i < rings.size(); i++) {
      full += rings[i];
      std::pair<size_t, size_t> range(cloudSize, 0);
      cloudSize += rings[i].size();
      range.second = cloudSize > 0 ? cloudSize - 1 : 0;
      ranges.push_back(range);
    }
    for (size_t ring = 0; ring < ranges.size(); ring++) {
      Cloud::Ptr ringLessFlat(new Cloud);
      const size_t s = ranges[ring].first, e = ranges[ring].second;
      if (e <= s + 2 * cfg.curvatureRegion) continue;
      maskUnreliable(s, e);
      for (int j = 0; j < cfg.nFeatureRegions; j++) {
        size_t sp = ((s + cfg.curvatureRegion) * (cfg.nFeatureRegions - j) + (e - cfg.curvatureRegion) * j) / cfg.nFeatureRegions;
        size_t ep = ((s + cfg.curvatureRegion) * (cfg.nFeatureRegions - 1 - j) + (e - cfg.curvatureRegion) * (j + 1)) / cfg.nFeatureRegions - 1;
        if (ep <= sp) continue;
        const size_t n = ep - sp + 1;
        prepareRegion(sp, ep);
        int nCorner = 0;
        for (size_t k = n; k > 0 && nCorner < cfg.maxCornerLessSharp;) {
          size_t idx = sorted[--k];
          size_t scanIdx = idx - s, regionIdx = idx - sp;
          if (picked[scanIdx] == 0 && curv[regionIdx] > cfg.surfaceCurvatureThreshold) {
            nCorner++;
            if (nCorner <= cfg.maxCornerSharp) { label[regionIdx] = 2; sharp.push_back(full[idx]); }
            else label[regionIdx] = 1;
            lessSharp.push_back(full[idx]);
            suppress(idx, scanIdx);
          }
        }
        int nFlat = 0;
        for (size_t k = 0; k < n && nFlat < cfg.maxSurfaceFlat; k++) {
          size_t idx = sorted[k];
          size_t scanIdx = idx - s, regionIdx = idx - sp;
          if (picked[scanIdx] == 0 && curv[regionIdx] < cfg.surfaceCurvatureThreshold) {
            nFlat++;
            label[regionIdx] = -1;
            flat.push_back(full[idx]);
            suppress(idx, scanIdx);
          }
        }
        for (size_t k = 0; k < n; k++)
          if (label[k] <= 0) ringLessFlat->push_back(full[sp + k]);
      }
      Cloud ds;
      pcl::VoxelGrid<Pt> f;
      f.setInputCloud(ringLessFlat);
      f.setLeafSize(cfg.lessFlatFilterSize, cfg.lessFlatFilterSize, cfg.lessFlatFilterSize);
      f.filter(ds);
      lessFlat += ds;
    }
  }
};

// ------------------------------------------------------------------------------------------------ shared pieces
// closed-form point-to-line residual, BasicLaserOdometry.cpp:319-337 == BasicLaserMapping.cpp:712-730
inline void lineResidual(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2,
                         float& la, float& lb, float& lc, float& ld2) {
  float a012 = std::sqrt(((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
                         ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
                         ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)));
  float l12 = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
  la = ((y1 - y2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) + (z1 - z2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))) / a012 / l12;
  lb = -((x1 - x2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) - (z1 - z2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
  lc = -((x1 - x2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) + (y1 - y2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
  ld2 = a012 / l12;
}

// solve + degeneracy projection, BasicLaserOdometry.cpp:555-597 == BasicLaserMapping.cpp:864-905
struct GaussNewton {
  bool degenerate = false;
  Eigen::Matrix<float, 6, 6> P;
  Eigen::Matrix<float, 6, 6> lastAtA;
  Eigen::Matrix<float, 6, 1> lastAtB;
  Eigen::Matrix<float, 6, 1> step(const std::vector<float>& rows, const std::vector<float>& rhs, bool first, float thr) {
    const int n = (int)rhs.size();
    Eigen::Matrix<float, Eigen::Dynamic, 6> A(n, 6);
    Eigen::Matrix<float, 6, Eigen::Dynamic> At(6, n);
    Eigen::VectorXf B(n);
    for (int i = 0; i < n; i++) {
      for (int c = 0; c < 6; c++) A(i, c) = rows[6 * i + c];
      B(i, 0) = rhs[i];
    }
    At = A.transpose();
    Eigen::Matrix<float, 6, 6> AtA = At * A;
    Eigen::Matrix<float, 6, 1> AtB = At * B;
    lastAtA = AtA;
    lastAtB = AtB;
    Eigen::Matrix<float, 6, 1> X = AtA.colPivHouseholderQr().solve(AtB);
    if (first) {
      Eigen::SelfAdjointEigenSolver<Eigen::Matrix<float, 6, 6> > es(AtA);
      Eigen::Matrix<float, 1, 6> E = es.eigenvalues().real();
      Eigen::Matrix<float, 6, 6> V = es.eigenvectors().real();
      Eigen::Matrix<float, 6, 6> V2 = V;
      degenerate = false;
      for (int i = 0; i < 6; i++) {
        if (E(0, i) < thr) {
          for (int j = 0; j < 6; j++) V2(i, j) = 0;
          degenerate = true;
        } else break;
      }
      P = V.inverse() * V2;
    }
    if (degenerate) {
      Eigen::Matrix<float, 6, 1> X2(X);
      X = P * X2;
    }
    return X;
  }
};

// ------------------------------------------------------------------------------------------------ odometry
struct Odom {
  float scanPeriod;
  size_t maxIter;
  float dTAbort = 0.1f, dRAbort = 0.1f;
  bool inited = false;
  long frames = 0;
  Cloud::Ptr sharp, lessSharp, flat, lessFlat, full, lastCorner, lastSurf;
  KdTree cornerTree, surfTree;
  std::vector<int> c1, c2, s1, s2, s3;
  Pose tf, sum;
  GaussNewton gn;
  size_t lastIters = 0;
  // per-iteration capture for GPU parity tests
  std::vector<float> rows, rhs;

  Odom(float sp, size_t it) : scanPeriod(sp), maxIter(it), sharp(new Cloud), lessSharp(new Cloud), flat(new Cloud),
                              lessFlat(new Cloud), full(new Cloud), lastCorner(new Cloud), lastSurf(new Cloud) {}

  // BasicLaserOdometry.cpp:40-53
  void toStart(const Pt& pi, Pt& po) const {
    float s = (1.f / scanPeriod) * (pi.intensity - int(pi.intensity));
    po.x = pi.x - s * tf.t.x;
    po.y = pi.y - s * tf.t.y;
    po.z = pi.z - s * tf.t.z;
    po.intensity = pi.intensity;
    Ang rx = -s * tf.rx.rad, ry = -s * tf.ry.rad, rz = -s * tf.rz.rad;
    zxy(po, rz, rx, ry);
  }
  // BasicLaserOdometry.cpp:57-87 (IMU terms vanish without IMU: shift 0, start/end angles 0)
  void toEnd(Cloud& c) const {
    for (auto& p : c.points) {
      float s = (1.f / scanPeriod) * (p.intensity - int(p.intensity));
      p.x -= s * tf.t.x;
      p.y -= s * tf.t.y;
      p.z -= s * tf.t.z;
      p.intensity = int(p.intensity);
      Ang rx = -s * tf.rx.rad, ry = -s * tf.ry.rad, rz = -s * tf.rz.rad;
      zxy(p, rz, rx, ry);
      yxz(p, tf.ry, tf.rx, tf.rz);
      p.x += tf.t.x - 0.f;
      p.y += tf.t.y - 0.f;
      p.z += tf.t.z - 0.f;
      Ang zero;
      zxy(p, zero, zero, zero);
      yxz(p, zero.neg(), zero.neg(), zero.neg());
    }
  }
  // BasicLaserOdometry.cpp:155-179
  static void accumulate(Ang cx, Ang cy, Ang cz, Ang lx, Ang ly, Ang lz, Ang& ox, Ang& oy, Ang& oz) {
    float srx = lx.c * cx.c * ly.s * cz.s - cx.c * cz.c * lx.s - lx.c * ly.c * cx.s;
    ox = -std::asin(srx);
    float srycrx = lx.s * (cy.c * cz.s - cz.c * cx.s * cy.s) + lx.c * ly.s * (cy.c * cz.c + cx.s * cy.s * cz.s) + lx.c * ly.c * cx.c * cy.s;
    float crycrx = lx.c * ly.c * cx.c * cy.c - lx.c * ly.s * (cz.c * cy.s - cy.c * cx.s * cz.s) - lx.s * (cy.s * cz.s + cy.c * cz.c * cx.s);
    oy = std::atan2(srycrx / ox.c, crycrx / ox.c);
    float srzcrx = cx.s * (lz.c * ly.s - ly.c * lx.s * lz.s) + cx.c * cz.s * (ly.c * lz.c + lx.s * ly.s * lz.s) + lx.c * cx.c * cz.c * lz.s;
    float crzcrx = lx.c * lz.c * cx.c * cz.c - cx.c * cz.s * (ly.c * lz.s - lz.c * lx.s * ly.s) - cx.s * (ly.s * lz.s + ly.c * lz.c * lx.s);
    oz = std::atan2(srzcrx / ox.c, crzcrx / ox.c);
  }
  // BasicLaserOdometry.cpp:91-151 with the three IMU start angles (bl*) and end angles (al*) all zero:
  // sin = 0, cos = 1 reduce every product exactly (x*1 = x, x*0 = 0, x+0 = x in IEEE arithmetic for finite x)
  static void pluginZeroImu(const Ang& bcx, const Ang& bcy, const Ang& bcz, Ang& acx, Ang& acy, Ang& acz) {
    const float sbcx = bcx.s, cbcx = bcx.c, sbcy = bcy.s, cbcy = bcy.c, sbcz = bcz.s, cbcz = bcz.c;
    const float sblx = 0.f, cblx = 1.f, sbly = 0.f, cbly = 1.f, sblz = 0.f, cblz = 1.f;
    const float salx = 0.f, calx = 1.f, saly = 0.f, caly = 1.f, salz = 0.f, calz = 1.f;
    float srx = -sbcx * (salx * sblx + calx * caly * cblx * cbly + calx * cblx * saly * sbly) -
                cbcx * cbcz * (calx * saly * (cbly * sblz - cblz * sblx * sbly) - calx * caly * (sbly * sblz + cbly * cblz * sblx) + cblx * cblz * salx) -
                cbcx * sbcz * (calx * caly * (cblz * sbly - cbly * sblx * sblz) - calx * saly * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sblz);
    acx = -std::asin(srx);
    float srycrx = (cbcy * sbcz - cbcz * sbcx * sbcy) * (calx * saly * (cbly * sblz - cblz * sblx * sbly) - calx * caly * (sbly * sblz + cbly * cblz * sblx) + cblx * cblz * salx) -
                   (cbcy * cbcz + sbcx * sbcy * sbcz) * (calx * caly * (cblz * sbly - cbly * sblx * sblz) - calx * saly * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sblz) +
                   cbcx * sbcy * (salx * sblx + calx * caly * cblx * cbly + calx * cblx * saly * sbly);
    float crycrx = (cbcz * sbcy - cbcy * sbcx * sbcz) * (calx * caly * (cblz * sbly - cbly * sblx * sblz) - calx * saly * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sblz) -
                   (sbcy * sbcz + cbcy * cbcz * sbcx) * (calx * saly * (cbly * sblz - cblz * sblx * sbly) - calx * caly * (sbly * sblz + cbly * cblz * sblx) + cblx * cblz * salx) +
                   cbcx * cbcy * (salx * sblx + calx * caly * cblx * cbly + calx * cblx * saly * sbly);
    acy = std::atan2(srycrx / acx.c, crycrx / acx.c);
    float srzcrx = sbcx * (cblx * cbly * (calz * saly - caly * salx * salz) - cblx * sbly * (caly * calz + salx * saly * salz) + calx * salz * sblx) -
                   cbcx * cbcz * ((caly * calz + salx * saly * salz) * (cbly * sblz - cblz * sblx * sbly) + (calz * saly - caly * salx * salz) * (sbly * sblz + cbly * cblz * sblx) - calx * cblx * cblz * salz) +
                   cbcx * sbcz * ((caly * calz + salx * saly * salz) * (cbly * cblz + sblx * sbly * sblz) + (calz * saly - caly * salx * salz) * (cblz * sbly - cbly * sblx * sblz) + calx * cblx * salz * sblz);
    float crzcrx = sbcx * (cblx * sbly * (caly * salz - calz * salx * saly) - cblx * cbly * (saly * salz + caly * calz * salx) + calx * calz * sblx) +
                   cbcx * cbcz * ((saly * salz + caly * calz * salx) * (sbly * sblz + cbly * cblz * sblx) + (caly * salz - calz * salx * saly) * (cbly * sblz - cblz * sblx * sbly) + calx * calz * cblx * cblz) -
                   cbcx * sbcz * ((saly * salz + caly * calz * salx) * (cblz * sbly - cbly * sblx * sblz) + (caly * salz - calz * salx * saly) * (cbly * cblz + sblx * sbly * sblz) - calx * calz * cblx * sblz);
    acz = std::atan2(srzcrx / acx.c, crzcrx / acx.c);
  }

  // one pass of the correspondence + residual loops, BasicLaserOdometry.cpp:246-482; fills rows/rhs (:497-554)
  // and returns the number of selected points.  coeffOut/selOut (optional) receive the per-query coefficient.
  int correspond(size_t iter, std::vector<float>* coeffOut, std::vector<signed char>* selOut) {
    const size_t nSharp = sharp->points.size(), nFlat = flat->points.size();
    std::vector<Pt> ori, coeffs;
    Pt sel, coeff;
    int ki[1];
    float kd[1];
    if (coeffOut) coeffOut->assign((nSharp + nFlat) * 4, 0.f);
    if (selOut) selOut->assign(nSharp + nFlat, 0);
    for (int i = 0; i < (int)nSharp; i++) {
      toStart(sharp->points[i], sel);
      if (iter % 5 == 0) {
        cornerTree.knn(sel, 1, ki, kd);
        int closest = -1, second = -1;
        if (kd[0] < 25) {
          closest = ki[0];
          int scan = int(lastCorner->points[closest].intensity);
          float dsq, best2 = 25;
          for (int j = closest + 1; j < (int)nSharp; j++) {  // bound by the CURRENT sharp count (:262)
            if (int(lastCorner->points[j].intensity) > scan + 2.5) break;
            dsq = d2(lastCorner->points[j], sel);
            if (int(lastCorner->points[j].intensity) > scan && dsq < best2) { best2 = dsq; second = j; }
          }
          for (int j = closest - 1; j >= 0; j--) {
            if (int(lastCorner->points[j].intensity) < scan - 2.5) break;
            dsq = d2(lastCorner->points[j], sel);
            if (int(lastCorner->points[j].intensity) < scan && dsq < best2) { best2 = dsq; second = j; }
          }
        }
        c1[i] = closest;
        c2[i] = second;
      }
      if (c2[i] >= 0) {
        const Pt& a = lastCorner->points[c1[i]];
        const Pt& b = lastCorner->points[c2[i]];
        float la, lb, lc, ld2;
        lineResidual(sel.x, sel.y, sel.z, a.x, a.y, a.z, b.x, b.y, b.z, la, lb, lc, ld2);
        float s = 1;
        if (iter >= 5) s = 1 - 1.8f * std::fabs(ld2);
        coeff.x = s * la; coeff.y = s * lb; coeff.z = s * lc; coeff.intensity = s * ld2;
        const bool keep = s > 0.1 && ld2 != 0;
        if (coeffOut) { float* o = &(*coeffOut)[4 * i]; o[0] = coeff.x; o[1] = coeff.y; o[2] = coeff.z; o[3] = coeff.intensity; }
        if (keep) {
          if (selOut) (*selOut)[i] = 1;
          ori.push_back(sharp->points[i]);
          coeffs.push_back(coeff);
        }
      }
    }
    for (int i = 0; i < (int)nFlat; i++) {
      toStart(flat->points[i], sel);
      if (iter % 5 == 0) {
        surfTree.knn(sel, 1, ki, kd);
        int closest = -1, second = -1, third = -1;
        if (kd[0] < 25) {
          closest = ki[0];
          int scan = int(lastSurf->points[closest].intensity);
          float dsq, best2 = 25, best3 = 25;
          for (int j = closest + 1; j < (int)nFlat; j++) {  // bound by the CURRENT flat count (:378)
            if (int(lastSurf->points[j].intensity) > scan + 2.5) break;
            dsq = d2(lastSurf->points[j], sel);
            if (int(lastSurf->points[j].intensity) <= scan) { if (dsq < best2) { best2 = dsq; second = j; } }
            else { if (dsq < best3) { best3 = dsq; third = j; } }
          }
          for (int j = closest - 1; j >= 0; j--) {
            if (int(lastSurf->points[j].intensity) < scan - 2.5) break;
            dsq = d2(lastSurf->points[j], sel);
            if (int(lastSurf->points[j].intensity) >= scan) { if (dsq < best2) { best2 = dsq; second = j; } }
            else { if (dsq < best3) { best3 = dsq; third = j; } }
          }
        }
        s1[i] = closest; s2[i] = second; s3[i] = third;
      }
      if (s2[i] >= 0 && s3[i] >= 0) {
        const Pt& t1 = lastSurf->points[s1[i]];
        const Pt& t2 = lastSurf->points[s2[i]];
        const Pt& t3 = lastSurf->points[s3[i]];
        float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
        float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
        float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
        float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
        float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
        pa /= ps; pb /= ps; pc /= ps; pd /= ps;
        float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
        float s = 1;
        if (iter >= 5) s = 1 - 1.8f * std::fabs(pd2) / std::sqrt(norm(sel));
        coeff.x = s * pa; coeff.y = s * pb; coeff.z = s * pc; coeff.intensity = s * pd2;
        const bool keep = s > 0.1 && pd2 != 0;
        if (coeffOut) { float* o = &(*coeffOut)[4 * (nSharp + i)]; o[0] = coeff.x; o[1] = coeff.y; o[2] = coeff.z; o[3] = coeff.intensity; }
        if (keep) {
          if (selOut) (*selOut)[nSharp + i] = 1;
          ori.push_back(flat->points[i]);
          coeffs.push_back(coeff);
        }
      }
    }
    // Jacobian rows, BasicLaserOdometry.cpp:497-554 (s = 1)
    const int n = (int)ori.size();
    rows.resize(6 * (size_t)n);
    rhs.resize(n);
    for (int i = 0; i < n; i++) {
      const Pt& p = ori[i];
      coeff = coeffs[i];
      float s = 1;
      float srx = std::sin(s * tf.rx.rad), crx = std::cos(s * tf.rx.rad);
      float sry = std::sin(s * tf.ry.rad), cry = std::cos(s * tf.ry.rad);
      float srz = std::sin(s * tf.rz.rad), crz = std::cos(s * tf.rz.rad);
      float tx = s * tf.t.x, ty = s * tf.t.y, tz = s * tf.t.z;
      float arx = (-s * crx * sry * srz * p.x + s * crx * crz * sry * p.y + s * srx * sry * p.z + s * tx * crx * sry * srz - s * ty * crx * crz * sry - s * tz * srx * sry) * coeff.x +
                  (s * srx * srz * p.x - s * crz * srx * p.y + s * crx * p.z + s * ty * crz * srx - s * tz * crx - s * tx * srx * srz) * coeff.y +
                  (s * crx * cry * srz * p.x - s * crx * cry * crz * p.y - s * cry * srx * p.z + s * tz * cry * srx + s * ty * crx * cry * crz - s * tx * crx * cry * srz) * coeff.z;
      float ary = ((-s * crz * sry - s * cry * srx * srz) * p.x + (s * cry * crz * srx - s * sry * srz) * p.y - s * crx * cry * p.z +
                   tx * (s * crz * sry + s * cry * srx * srz) + ty * (s * sry * srz - s * cry * crz * srx) + s * tz * crx * cry) * coeff.x +
                  ((s * cry * crz - s * srx * sry * srz) * p.x + (s * cry * srz + s * crz * srx * sry) * p.y - s * crx * sry * p.z + s * tz * crx * sry -
                   ty * (s * cry * srz + s * crz * srx * sry) - tx * (s * cry * crz - s * srx * sry * srz)) * coeff.z;
      float arz = ((-s * cry * srz - s * crz * srx * sry) * p.x + (s * cry * crz - s * srx * sry * srz) * p.y + tx * (s * cry * srz + s * crz * srx * sry) -
                   ty * (s * cry * crz - s * srx * sry * srz)) * coeff.x +
                  (-s * crx * crz * p.x - s * crx * srz * p.y + s * ty * crx * srz + s * tx * crx * crz) * coeff.y +
                  ((s * cry * crz * srx - s * sry * srz) * p.x + (s * crz * sry + s * cry * srx * srz) * p.y + tx * (s * sry * srz - s * cry * crz * srx) -
                   ty * (s * crz * sry + s * cry * srx * srz)) * coeff.z;
      float atx = -s * (cry * crz - srx * sry * srz) * coeff.x + s * crx * srz * coeff.y - s * (crz * sry + cry * srx * srz) * coeff.z;
      float aty = -s * (cry * srz + crz * srx * sry) * coeff.x - s * crx * crz * coeff.y - s * (sry * srz - cry * crz * srx) * coeff.z;
      float atz = s * crx * sry * coeff.x - s * srx * coeff.y - s * crx * cry * coeff.z;
      float* r = &rows[6 * (size_t)i];
      r[0] = arx; r[1] = ary; r[2] = arz; r[3] = atx; r[4] = aty; r[5] = atz;
      rhs[i] = -0.05 * coeff.intensity;
    }
    return n;
  }

  // BasicLaserOdometry.cpp:196-666
  void process() {
    if (!inited) {
      lessSharp.swap(lastCorner);
      lessFlat.swap(lastSurf);
      cornerTree.build(lastCorner.get());
      surfTree.build(lastSurf.get());
      inited = true;
      return;
    }
    frames++;
    lastIters = 0;
    size_t nLastCorner = lastCorner->points.size(), nLastSurf = lastSurf->points.size();
    if (nLastCorner > 10 && nLastSurf > 100) {
      const size_t nSharp = sharp->points.size(), nFlat = flat->points.size();
      c1.assign(nSharp, 0); c2.assign(nSharp, 0);
      s1.assign(nFlat, 0); s2.assign(nFlat, 0); s3.assign(nFlat, 0);
      for (size_t iter = 0; iter < maxIter; iter++) {
        lastIters = iter + 1;
        const int n = correspond(iter, nullptr, nullptr);
        if (n < 10) continue;
        Eigen::Matrix<float, 6, 1> X = gn.step(rows, rhs, iter == 0, 10.f);
        tf.rx = tf.rx.rad + X(0, 0);
        tf.ry = tf.ry.rad + X(1, 0);
        tf.rz = tf.rz.rad + X(2, 0);
        tf.t.x += X(3, 0); tf.t.y += X(4, 0); tf.t.z += X(5, 0);
        if (!std::isfinite(tf.rx.rad)) tf.rx = Ang();
        if (!std::isfinite(tf.ry.rad)) tf.ry = Ang();
        if (!std::isfinite(tf.rz.rad)) tf.rz = Ang();
        if (!std::isfinite(tf.t.x)) tf.t.x = 0.0;
        if (!std::isfinite(tf.t.y)) tf.t.y = 0.0;
        if (!std::isfinite(tf.t.z)) tf.t.z = 0.0;
        float dR = std::sqrt(std::pow(deg(X(0, 0)), 2) + std::pow(deg(X(1, 0)), 2) + std::pow(deg(X(2, 0)), 2));
        float dT = std::sqrt(std::pow(X(3, 0) * 100, 2) + std::pow(X(4, 0) * 100, 2) + std::pow(X(5, 0) * 100, 2));
        if (dR < dRAbort && dT < dTAbort) break;
      }
    }
    Ang rx, ry, rz;
    accumulate(sum.rx, sum.ry, sum.rz, tf.rx.neg(), Ang(-tf.ry.rad * 1.05), tf.rz.neg(), rx, ry, rz);
    V3 v;
    v.x = tf.t.x - 0.f;
    v.y = tf.t.y - 0.f;
    v.z = tf.t.z * 1.05 - 0.f;
    zxy(v, rz, rx, ry);
    V3 trans;
    trans.x = sum.t.x - v.x; trans.y = sum.t.y - v.y; trans.z = sum.t.z - v.z;
    pluginZeroImu(rx, ry, rz, rx, ry, rz);
    sum.rx = rx; sum.ry = ry; sum.rz = rz; sum.t = trans;
    toEnd(*lessSharp);
    toEnd(*lessFlat);
    lessSharp.swap(lastCorner);
    lessFlat.swap(lastSurf);
    nLastCorner = lastCorner->points.size();
    nLastSurf = lastSurf->points.size();
    if (nLastCorner > 10 && nLastSurf > 100) {
      cornerTree.build(lastCorner.get());
      surfTree.build(lastSurf.get());
    }
  }
};

// ------------------------------------------------------------------------------------------------ mapping
struct Mapping {
  float scanPeriod;
  size_t maxIter;
  float dTAbort = 0.05f, dRAbort = 0.05f;
  long frameCount, mapFrameCount;
  int cenW = 10, cenH = 5, cenD = 10;
  const int W = 21, H = 11, D = 21;
  Cloud::Ptr cornerLast, surfLast, fullRes, cornerStack, surfStack, cornerStackDS, surfStackDS, surround, surroundDS,
      cornerFromMap, surfFromMap;
  std::vector<Cloud::Ptr> cornerCubes, surfCubes, cornerCubesDS, surfCubesDS;
  std::vector<size_t> validInd, surroundInd;
  Pose sum, incre, tobe, bef, aft;
  pcl::VoxelGrid<Pt> fCorner, fSurf;
  bool freshMap = false;
  KdTree cornerTree, surfTree;
  GaussNewton gn;
  size_t lastIters = 0;
  std::vector<float> rows, rhs;
  Cloud scratch;

  Mapping(float sp, size_t it)
      : scanPeriod(sp), maxIter(it), cornerLast(new Cloud), surfLast(new Cloud), fullRes(new Cloud),
        cornerStack(new Cloud), surfStack(new Cloud), cornerStackDS(new Cloud), surfStackDS(new Cloud),
        surround(new Cloud), surroundDS(new Cloud), cornerFromMap(new Cloud), surfFromMap(new Cloud) {
    // BasicLaserMapping.cpp:51-100
    frameCount = 1 - 1;
    mapFrameCount = 5 - 1;
    const size_t n = (size_t)W * H * D;
    cornerCubes.resize(n); surfCubes.resize(n); cornerCubesDS.resize(n); surfCubesDS.resize(n);
    for (size_t i = 0; i < n; i++) {
      cornerCubes[i].reset(new Cloud); surfCubes[i].reset(new Cloud);
      cornerCubesDS[i].reset(new Cloud); surfCubesDS[i].reset(new Cloud);
    }
    fCorner.setLeafSize(0.2, 0.2, 0.2);
    fSurf.setLeafSize(0.4, 0.4, 0.4);
  }
  size_t at(int i, int j, int k) const { return i + (size_t)W * j + (size_t)W * H * k; }

  // BasicLaserMapping.cpp:103-167
  void predict() {
    incre.t.x = bef.t.x - sum.t.x; incre.t.y = bef.t.y - sum.t.y; incre.t.z = bef.t.z - sum.t.z;
    yxz(incre.t, sum.ry.neg(), sum.rx.neg(), sum.rz.neg());
    float sbcx = sum.rx.s, cbcx = sum.rx.c, sbcy = sum.ry.s, cbcy = sum.ry.c, sbcz = sum.rz.s, cbcz = sum.rz.c;
    float sblx = bef.rx.s, cblx = bef.rx.c, sbly = bef.ry.s, cbly = bef.ry.c, sblz = bef.rz.s, cblz = bef.rz.c;
    float salx = aft.rx.s, calx = aft.rx.c, saly = aft.ry.s, caly = aft.ry.c, salz = aft.rz.s, calz = aft.rz.c;
    float srx = -sbcx * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz) -
                cbcx * sbcy * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                cbcx * cbcy * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx);
    tobe.rx = -std::asin(srx);
    float srycrx = sbcx * (cblx * cblz * (caly * salz - calz * salx * saly) - cblx * sblz * (caly * calz + salx * saly * salz) + calx * saly * sblx) -
                   cbcx * cbcy * ((caly * calz + salx * saly * salz) * (cblz * sbly - cbly * sblx * sblz) + (caly * salz - calz * salx * saly) * (sbly * sblz + cbly * cblz * sblx) - calx * cblx * cbly * saly) +
                   cbcx * sbcy * ((caly * calz + salx * saly * salz) * (cbly * cblz + sblx * sbly * sblz) + (caly * salz - calz * salx * saly) * (cbly * sblz - cblz * sblx * sbly) + calx * cblx * saly * sbly);
    float crycrx = sbcx * (cblx * sblz * (calz * saly - caly * salx * salz) - cblx * cblz * (saly * salz + caly * calz * salx) + calx * caly * sblx) +
                   cbcx * cbcy * ((saly * salz + caly * calz * salx) * (sbly * sblz + cbly * cblz * sblx) + (calz * saly - caly * salx * salz) * (cblz * sbly - cbly * sblx * sblz) + calx * caly * cblx * cbly) -
                   cbcx * sbcy * ((saly * salz + caly * calz * salx) * (cbly * sblz - cblz * sblx * sbly) + (calz * saly - caly * salx * salz) * (cbly * cblz + sblx * sbly * sblz) - calx * caly * cblx * sbly);
    tobe.ry = std::atan2(srycrx / tobe.rx.c, crycrx / tobe.rx.c);
    float srzcrx = (cbcz * sbcy - cbcy * sbcx * sbcz) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) -
                   (cbcy * cbcz + sbcx * sbcy * sbcz) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) +
                   cbcx * sbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
    float crzcrx = (cbcy * sbcz - cbcz * sbcx * sbcy) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                   (sbcy * sbcz + cbcy * cbcz * sbcx) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) +
                   cbcx * cbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
    tobe.rz = std::atan2(srzcrx / tobe.rx.c, crzcrx / tobe.rx.c);
    V3 v = incre.t;
    zxy(v, tobe.rz, tobe.rx, tobe.ry);
    tobe.t.x = aft.t.x - v.x; tobe.t.y = aft.t.y - v.y; tobe.t.z = aft.t.z - v.z;
  }
  // BasicLaserMapping.cpp:207-219 / 223-231
  void toMap(const Pt& pi, Pt& po) const {
    po.x = pi.x; po.y = pi.y; po.z = pi.z; po.intensity = pi.intensity;
    zxy(po, tobe.rz, tobe.rx, tobe.ry);
    po.x += tobe.t.x; po.y += tobe.t.y; po.z += tobe.t.z;
  }
  void toSensor(const Pt& pi, Pt& po) const {
    po.x = pi.x - tobe.t.x; po.y = pi.y - tobe.t.y; po.z = pi.z - tobe.t.z; po.intensity = pi.intensity;
    yxz(po, tobe.ry.neg(), tobe.rx.neg(), tobe.rz.neg());
  }
  // cube index of a map-frame point, BasicLaserMapping.cpp:540-553
  bool cubeOf(const Pt& p, size_t& ind) const {
    const double SZ = 50.0, HALF = SZ / 2;
    int ci = int((p.x + HALF) / SZ) + cenW, cj = int((p.y + HALF) / SZ) + cenH, ck = int((p.z + HALF) / SZ) + cenD;
    if (p.x + HALF < 0) ci--;
    if (p.y + HALF < 0) cj--;
    if (p.z + HALF < 0) ck--;
    if (ci >= 0 && ci < W && cj >= 0 && cj < H && ck >= 0 && ck < D) { ind = at(ci, cj, ck); return true; }
    return false;
  }
  void seed(int kind, const float* p, int n) {
    for (int i = 0; i < n; i++) {
      Pt q; q.x = p[4 * i]; q.y = p[4 * i + 1]; q.z = p[4 * i + 2]; q.intensity = p[4 * i + 3];
      size_t ind;
      if (cubeOf(q, ind)) (kind == 0 ? cornerCubes : surfCubes)[ind]->push_back(q);
    }
  }
  // one cell of grid roll along an axis, BasicLaserMapping.cpp:311-441
  void roll(int axis, int dir) {
    const int dims[3] = {W, H, D};
    int c[3];
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    for (c[a1] = 0; c[a1] < dims[a1]; c[a1]++)
      for (c[a2] = 0; c[a2] < dims[a2]; c[a2]++) {
        if (dir > 0) {
          for (int t = dims[axis] - 1; t >= 1; t--) {
            c[axis] = t; size_t a = at(c[0], c[1], c[2]);
            c[axis] = t - 1; size_t b = at(c[0], c[1], c[2]);
            std::swap(cornerCubes[a], cornerCubes[b]); std::swap(surfCubes[a], surfCubes[b]);
          }
          c[axis] = 0;
        } else {
          for (int t = 0; t < dims[axis] - 1; t++) {
            c[axis] = t; size_t a = at(c[0], c[1], c[2]);
            c[axis] = t + 1; size_t b = at(c[0], c[1], c[2]);
            std::swap(cornerCubes[a], cornerCubes[b]); std::swap(surfCubes[a], surfCubes[b]);
          }
          c[axis] = dims[axis] - 1;
        }
        size_t z = at(c[0], c[1], c[2]);
        cornerCubes[z]->clear(); surfCubes[z]->clear();
      }
  }

  // correspondence pass of one iteration, BasicLaserMapping.cpp:665-817, and Jacobian rows :837-862
  int correspond(std::vector<float>* coeffOut, std::vector<signed char>* selOut) {
    const size_t nC = cornerStackDS->size(), nS = surfStackDS->size();
    std::vector<Pt> ori, coeffs;
    Pt sel, coeff, pOri;
    int ki[5];
    float kd[5];
    Eigen::Matrix<float, 5, 3> A0;
    Eigen::Matrix<float, 5, 1> B0;
    Eigen::Vector3f X0;
    Eigen::Matrix3f A1;
    Eigen::Matrix<float, 1, 3> D1;
    Eigen::Matrix3f V1;
    A0.setZero(); B0.setConstant(-1); X0.setZero(); A1.setZero(); D1.setZero(); V1.setZero();
    if (coeffOut) coeffOut->assign((nC + nS) * 4, 0.f);
    if (selOut) selOut->assign(nC + nS, 0);
    for (int i = 0; i < (int)nC; i++) {
      pOri = cornerStackDS->points[i];
      toMap(pOri, sel);
      cornerTree.knn(sel, 5, ki, kd);
      if (kd[4] < 1.0) {
        float vx = 0, vy = 0, vz = 0;
        for (int j = 0; j < 5; j++) { const Pt& q = cornerFromMap->points[ki[j]]; vx += q.x; vy += q.y; vz += q.z; }
        vx /= 5.0f; vy /= 5.0f; vz /= 5.0f;
        Eigen::Matrix3f m;
        m.setZero();
        for (int j = 0; j < 5; j++) {
          const Pt& q = cornerFromMap->points[ki[j]];
          float ax = q.x - vx, ay = q.y - vy, az = q.z - vz;
          m(0, 0) += ax * ax; m(1, 0) += ax * ay; m(2, 0) += ax * az;
          m(1, 1) += ay * ay; m(2, 1) += ay * az; m(2, 2) += az * az;
        }
        A1 = m / 5.0;
        Eigen::SelfAdjointEigenSolver<Eigen::Matrix3f> es(A1);
        D1 = es.eigenvalues().real();
        V1 = es.eigenvectors().real();
        if (D1(0, 2) > 3 * D1(0, 1)) {
          float x1 = vx + 0.1 * V1(0, 2), y1 = vy + 0.1 * V1(1, 2), z1 = vz + 0.1 * V1(2, 2);
          float x2 = vx - 0.1 * V1(0, 2), y2 = vy - 0.1 * V1(1, 2), z2 = vz - 0.1 * V1(2, 2);
          float la, lb, lc, ld2;
          lineResidual(sel.x, sel.y, sel.z, x1, y1, z1, x2, y2, z2, la, lb, lc, ld2);
          float s = 1 - 0.9f * std::fabs(ld2);
          coeff.x = s * la; coeff.y = s * lb; coeff.z = s * lc; coeff.intensity = s * ld2;
          if (coeffOut) { float* o = &(*coeffOut)[4 * i]; o[0] = coeff.x; o[1] = coeff.y; o[2] = coeff.z; o[3] = coeff.intensity; }
          if (s > 0.1) {
            if (selOut) (*selOut)[i] = 1;
            ori.push_back(pOri);
            coeffs.push_back(coeff);
          }
        }
      }
    }
    for (int i = 0; i < (int)nS; i++) {
      pOri = surfStackDS->points[i];
      toMap(pOri, sel);
      surfTree.knn(sel, 5, ki, kd);
      if (kd[4] < 1.0) {
        for (int j = 0; j < 5; j++) {
          const Pt& q = surfFromMap->points[ki[j]];
          A0(j, 0) = q.x; A0(j, 1) = q.y; A0(j, 2) = q.z;
        }
        X0 = A0.colPivHouseholderQr().solve(B0);
        float pa = X0(0, 0), pb = X0(1, 0), pc = X0(2, 0), pd = 1;
        float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
        pa /= ps; pb /= ps; pc /= ps; pd /= ps;
        bool valid = true;
        for (int j = 0; j < 5; j++) {
          const Pt& q = surfFromMap->points[ki[j]];
          if (std::fabs(pa * q.x + pb * q.y + pc * q.z + pd) > 0.2) { valid = false; break; }
        }
        if (valid) {
          float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
          float s = 1 - 0.9f * std::fabs(pd2) / std::sqrt(norm(sel));
          coeff.x = s * pa; coeff.y = s * pb; coeff.z = s * pc; coeff.intensity = s * pd2;
          if (coeffOut) { float* o = &(*coeffOut)[4 * (nC + i)]; o[0] = coeff.x; o[1] = coeff.y; o[2] = coeff.z; o[3] = coeff.intensity; }
          if (s > 0.1) {
            if (selOut) (*selOut)[nC + i] = 1;
            ori.push_back(pOri);
            coeffs.push_back(coeff);
          }
        }
      }
    }
    const float srx = tobe.rx.s, crx = tobe.rx.c, sry = tobe.ry.s, cry = tobe.ry.c, srz = tobe.rz.s, crz = tobe.rz.c;
    const int n = (int)ori.size();
    rows.resize(6 * (size_t)n);
    rhs.resize(n);
    for (int i = 0; i < n; i++) {
      const Pt& p = ori[i];
      coeff = coeffs[i];
      float arx = (crx * sry * srz * p.x + crx * crz * sry * p.y - srx * sry * p.z) * coeff.x +
                  (-srx * srz * p.x - crz * srx * p.y - crx * p.z) * coeff.y +
                  (crx * cry * srz * p.x + crx * cry * crz * p.y - cry * srx * p.z) * coeff.z;
      float ary = ((cry * srx * srz - crz * sry) * p.x + (sry * srz + cry * crz * srx) * p.y + crx * cry * p.z) * coeff.x +
                  ((-cry * crz - srx * sry * srz) * p.x + (cry * srz - crz * srx * sry) * p.y - crx * sry * p.z) * coeff.z;
      float arz = ((crz * srx * sry - cry * srz) * p.x + (-cry * crz - srx * sry * srz) * p.y) * coeff.x +
                  (crx * crz * p.x - crx * srz * p.y) * coeff.y +
                  ((sry * srz + cry * crz * srx) * p.x + (crz * sry - cry * srx * srz) * p.y) * coeff.z;
      float* r = &rows[6 * (size_t)i];
      r[0] = arx; r[1] = ary; r[2] = arz; r[3] = coeff.x; r[4] = coeff.y; r[5] = coeff.z;
      rhs[i] = -coeff.intensity;
    }
    return n;
  }

  // BasicLaserMapping.cpp:626-926 (transformUpdate :171-203 without IMU)
  void optimize() {
    lastIters = 0;
    if (cornerFromMap->size() <= 10 || surfFromMap->size() <= 100) return;
    cornerTree.build(cornerFromMap.get());
    surfTree.build(surfFromMap.get());
    for (size_t iter = 0; iter < maxIter; iter++) {
      lastIters = iter + 1;
      const int n = correspond(nullptr, nullptr);
      if (n < 50) continue;
      Eigen::Matrix<float, 6, 1> X = gn.step(rows, rhs, iter == 0, 100.f);
      tobe.rx = tobe.rx.rad + X(0, 0);
      tobe.ry = tobe.ry.rad + X(1, 0);
      tobe.rz = tobe.rz.rad + X(2, 0);
      tobe.t.x += X(3, 0); tobe.t.y += X(4, 0); tobe.t.z += X(5, 0);
      float dR = std::sqrt(std::pow(deg(X(0, 0)), 2) + std::pow(deg(X(1, 0)), 2) + std::pow(deg(X(2, 0)), 2));
      float dT = std::sqrt(std::pow(X(3, 0) * 100, 2) + std::pow(X(4, 0) * 100, 2) + std::pow(X(5, 0) * 100, 2));
      if (dR < dRAbort && dT < dTAbort) break;
    }
    bef = sum;
    aft = tobe;
  }

  // BasicLaserMapping.cpp:266-599
  bool process() {
    frameCount++;
    if (frameCount < 1) return false;
    frameCount = 0;
    Pt sel;
    predict();
    for (auto const& p : cornerLast->points) { toMap(p, sel); cornerStack->push_back(sel); }
    for (auto const& p : surfLast->points) { toMap(p, sel); surfStack->push_back(sel); }
    Pt up;
    up.x = 0.0; up.y = 10.0; up.z = 0.0;
    toMap(up, up);
    const double SZ = 50.0, HALF = SZ / 2;
    int ci = int((tobe.t.x + HALF) / SZ) + cenW, cj = int((tobe.t.y + HALF) / SZ) + cenH, ck = int((tobe.t.z + HALF) / SZ) + cenD;
    if (tobe.t.x + HALF < 0) ci--;
    if (tobe.t.y + HALF < 0) cj--;
    if (tobe.t.z + HALF < 0) ck--;
    while (ci < 3) { roll(0, +1); ci++; cenW++; }
    while (ci >= W - 3) { roll(0, -1); ci--; cenW--; }
    while (cj < 3) { roll(1, +1); cj++; cenH++; }
    while (cj >= H - 3) { roll(1, -1); cj--; cenH--; }
    while (ck < 3) { roll(2, +1); ck++; cenD++; }
    while (ck >= D - 3) { roll(2, -1); ck--; cenD--; }
    validInd.clear();
    surroundInd.clear();
    for (int i = ci - 2; i <= ci + 2; i++)
      for (int j = cj - 2; j <= cj + 2; j++)
        for (int k = ck - 2; k <= ck + 2; k++) {
          if (!(i >= 0 && i < W && j >= 0 && j < H && k >= 0 && k < D)) continue;
          float cX = 50.0f * (i - cenW), cY = 50.0f * (j - cenH), cZ = 50.0f * (k - cenD);
          Pt pos;
          pos.x = tobe.t.x; pos.y = tobe.t.y; pos.z = tobe.t.z;
          bool inFov = false;
          for (int ii = -1; ii <= 1; ii += 2)
            for (int jj = -1; jj <= 1; jj += 2)
              for (int kk = -1; kk <= 1; kk += 2) {
                Pt corner;
                corner.x = cX + 25.0f * ii; corner.y = cY + 25.0f * jj; corner.z = cZ + 25.0f * kk;
                float side1 = d2(pos, corner), side2 = d2(up, corner);
                float check1 = 100.0f + side1 - side2 - 10.0f * std::sqrt(3.0f) * std::sqrt(side1);
                float check2 = 100.0f + side1 - side2 + 10.0f * std::sqrt(3.0f) * std::sqrt(side1);
                if (check1 < 0 && check2 > 0) inFov = true;
              }
          size_t idx = at(i, j, k);
          if (inFov) validInd.push_back(idx);
          surroundInd.push_back(idx);
        }
    cornerFromMap->clear();
    surfFromMap->clear();
    for (auto ind : validInd) { *cornerFromMap += *cornerCubes[ind]; *surfFromMap += *surfCubes[ind]; }
    for (auto& p : *cornerStack) toSensor(p, p);
    for (auto& p : *surfStack) toSensor(p, p);
    cornerStackDS->clear();
    fCorner.setInputCloud(cornerStack);
    fCorner.filter(*cornerStackDS);
    size_t nC = cornerStackDS->size();
    surfStackDS->clear();
    fSurf.setInputCloud(surfStack);
    fSurf.filter(*surfStackDS);
    size_t nS = surfStackDS->size();
    cornerStack->clear();
    surfStack->clear();
    optimize();
    for (size_t i = 0; i < nC; i++) {
      toMap(cornerStackDS->points[i], sel);
      size_t ind;
      if (cubeOf(sel, ind)) cornerCubes[ind]->push_back(sel);
    }
    for (size_t i = 0; i < nS; i++) {
      toMap(surfStackDS->points[i], sel);
      size_t ind;
      if (cubeOf(sel, ind)) surfCubes[ind]->push_back(sel);
    }
    for (auto ind : validInd) {
      cornerCubesDS[ind]->clear();
      fCorner.setInputCloud(cornerCubes[ind]);
      fCorner.filter(*cornerCubesDS[ind]);
      surfCubesDS[ind]->clear();
      fSurf.setInputCloud(surfCubes[ind]);
      fSurf.filter(*surfCubesDS[ind]);
      cornerCubes[ind].swap(cornerCubesDS[ind]);
      surfCubes[ind].swap(surfCubesDS[ind]);
    }
    for (auto& p : *fullRes) toMap(p, p);  // :235-240
    // createDownsizedMap, :242-264
    mapFrameCount++;
    freshMap = false;
    if (mapFrameCount >= 5) {
      mapFrameCount = 0;
      surround->clear();
      for (auto ind : surroundInd) { *surround += *cornerCubes[ind]; *surround += *surfCubes[ind]; }
      surroundDS->clear();
      fCorner.setInputCloud(surround);
      fCorner.filter(*surroundDS);
      freshMap = true;
    }
    return true;
  }
  const Cloud& cloud(int which) {
    switch (which) {
      case 0: return *fullRes;
      case 1: return *surroundDS;
      case 2: return *cornerFromMap;
      case 3: return *surfFromMap;
      case 4: return *cornerStackDS;
      case 5: return *surfStackDS;
      default: {
        scratch.clear();
        for (auto& c : (which == 6 ? cornerCubes : surfCubes)) scratch += *c;
        return scratch;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------ glue
inline void fill(Cloud& c, const float* p, int n) {
  c.clear();
  c.points.resize(n > 0 ? n : 0);
  for (int i = 0; i < n; i++) { c.points[i].x = p[4 * i]; c.points[i].y = p[4 * i + 1]; c.points[i].z = p[4 * i + 2]; c.points[i].intensity = p[4 * i + 3]; }
  c.width = n > 0 ? n : 0;
  c.height = 1;
}
inline void dump(const Cloud& c, float* o) {
  for (size_t i = 0; i < c.points.size(); i++) { o[4 * i] = c.points[i].x; o[4 * i + 1] = c.points[i].y; o[4 * i + 2] = c.points[i].z; o[4 * i + 3] = c.points[i].intensity; }
}
inline void twist6(const Pose& t, float* o) { o[0] = t.rx.rad; o[1] = t.ry.rad; o[2] = t.rz.rad; o[3] = t.t.x; o[4] = t.t.y; o[5] = t.t.z; }
inline double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct RegH {
  ScanReg r;
  std::vector<Cloud> rings;
  const Cloud& cloud(int w) { return w == 0 ? r.full : w == 1 ? r.sharp : w == 2 ? r.lessSharp : w == 3 ? r.flat : r.lessFlat; }
  void process(const float* pts, const int* sizes, int n) {
    rings.resize(n);
    int off = 0;
    for (int i = 0; i < n; i++) { fill(rings[i], pts + 4 * (size_t)off, sizes[i]); off += sizes[i]; }
    r.process(rings);
  }
};
// Ring binning front end: MultiScanMapper (MultiScanRegistration.cpp:44-67) + MultiScanRegistration::process (:160-238).
// Plain float libm calls like the reference (std::atan / std::atan2 / std::sqrt on floats); IMU projection is the
// identity without IMU data (BasicScanRegistration.cpp:100-112).
struct MultiScan {
  float lower, upper, factor;
  int nRings;
  RegH reg;
  std::vector<Cloud> scans;
  MultiScan(float lo, float hi, int n) : lower(lo), upper(hi), factor((n - 1) / (hi - lo)), nRings(n) {}
  int ringForAngle(const float& angle) const { return int(((angle * 180 / M_PI) - lower) * factor + 0.5); }
  int process(const float* xyz, int n) {
    scans.assign(nRings, Cloud());
    if (n <= 0) return 0;
    float startOri = -std::atan2(xyz[1], xyz[0]);
    float endOri = -std::atan2(xyz[3 * (size_t)(n - 1) + 1], xyz[3 * (size_t)(n - 1) + 0]) + 2 * float(M_PI);
    if (endOri - startOri > 3 * M_PI) {
      endOri -= 2 * M_PI;
    } else if (endOri - startOri < M_PI) {
      endOri += 2 * M_PI;
    }
    bool halfPassed = false;
    Pt point;
    for (int i = 0; i < n; i++) {
      point.x = xyz[3 * (size_t)i + 1];
      point.y = xyz[3 * (size_t)i + 2];
      point.z = xyz[3 * (size_t)i + 0];
      if (!std::isfinite(point.x) || !std::isfinite(point.y) || !std::isfinite(point.z)) continue;
      if (point.x * point.x + point.y * point.y + point.z * point.z < 0.0001) continue;
      float angle = std::atan(point.y / std::sqrt(point.x * point.x + point.z * point.z));
      int scanID = ringForAngle(angle);
      if (scanID >= nRings || scanID < 0) continue;
      float ori = -std::atan2(point.x, point.z);
      if (!halfPassed) {
        if (ori < startOri - M_PI / 2) {
          ori += 2 * M_PI;
        } else if (ori > startOri + M_PI * 3 / 2) {
          ori -= 2 * M_PI;
        }
        if (ori - startOri > M_PI) halfPassed = true;
      } else {
        ori += 2 * M_PI;
        if (ori < endOri - M_PI * 3 / 2) {
          ori += 2 * M_PI;
        } else if (ori > endOri + M_PI / 2) {
          ori -= 2 * M_PI;
        }
      }
      float relTime = reg.r.cfg.scanPeriod * (ori - startOri) / (endOri - startOri);
      point.intensity = scanID + relTime;
      scans[scanID].push_back(point);
    }
    reg.r.process(scans);
    size_t kept = 0;
    for (auto& s : scans) kept += s.size();
    return (int)kept;
  }
};

struct Pipe {
  RegH reg;
  Odom odom;
  Mapping map;
  Pipe(float sp, int oi, int mi) : odom(sp, oi), map(sp, mi) {}
};

}  // namespace orc

using namespace orc;

extern "C" {

const char* loamdrv_kind(void) { return "restatement"; }

void* loamdrv_scanreg_create(void) { return new RegH(); }
void loamdrv_scanreg_destroy(void* h) { delete (RegH*)h; }
void loamdrv_scanreg_configure(void* h, float scanPeriod, int nFeatureRegions, int curvatureRegion, int maxCornerSharp,
                               int maxSurfaceFlat, float lessFlatFilterSize, float surfaceCurvatureThreshold) {
  RegCfg& c = ((RegH*)h)->r.cfg;
  c.scanPeriod = scanPeriod; c.nFeatureRegions = nFeatureRegions; c.curvatureRegion = curvatureRegion;
  c.maxCornerSharp = maxCornerSharp; c.maxCornerLessSharp = 10 * maxCornerSharp; c.maxSurfaceFlat = maxSurfaceFlat;
  c.lessFlatFilterSize = lessFlatFilterSize; c.surfaceCurvatureThreshold = surfaceCurvatureThreshold;
}
int loamdrv_scanreg_process(void* h, const float* pts, const int* ring_sizes, int n_rings) {
  ((RegH*)h)->process(pts, ring_sizes, n_rings);
  return (int)((RegH*)h)->r.full.size();
}
int loamdrv_scanreg_cloud_size(void* h, int which) { return (int)((RegH*)h)->cloud(which).size(); }
void loamdrv_scanreg_cloud_copy(void* h, int which, float* out) { dump(((RegH*)h)->cloud(which), out); }

// BasicTransformMaintenance.cpp:45-178.  transformAssociateToMap there is the same association of (sum, bef, aft) as
// BasicLaserMapping.cpp:103-167 (restated in Mapping::predict above) written on float[6] arrays: the sines / cosines are
// std::sin / std::cos of the float angles (what Ang caches), the translation steps are the same elementary rotations.
void loamdrv_transform_maintenance(const float* sum, const float* bef, const float* aft, float* out) {
  Mapping m(0.1f, 1);
  const float* src[3] = {sum, bef, aft};
  Pose* dst[3] = {&m.sum, &m.bef, &m.aft};
  for (int k = 0; k < 3; k++) {
    dst[k]->rx = Ang(src[k][0]);
    dst[k]->ry = Ang(src[k][1]);
    dst[k]->rz = Ang(src[k][2]);
    dst[k]->t.x = src[k][3];
    dst[k]->t.y = src[k][4];
    dst[k]->t.z = src[k][5];
  }
  m.predict();
  twist6(m.tobe, out);
}

void* loamdrv_multiscan_create(float lo, float hi, int n) { return new MultiScan(lo, hi, n); }
void loamdrv_multiscan_destroy(void* h) { delete (MultiScan*)h; }
int loamdrv_multiscan_process(void* h, const float* xyz, int n) { return ((MultiScan*)h)->process(xyz, n); }
void loamdrv_multiscan_binned(void* h, float* out, int* ring_sizes) {
  MultiScan* ms = (MultiScan*)h;
  size_t off = 0;
  for (size_t r = 0; r < ms->scans.size(); r++) {
    dump(ms->scans[r], out + 4 * off);
    ring_sizes[r] = (int)ms->scans[r].size();
    off += ms->scans[r].size();
  }
}
int loamdrv_multiscan_cloud_size(void* h, int which) { return (int)((MultiScan*)h)->reg.cloud(which).size(); }
void loamdrv_multiscan_cloud_copy(void* h, int which, float* out) { dump(((MultiScan*)h)->reg.cloud(which), out); }

void* loamdrv_odom_create(float scanPeriod, int maxIterations) { return new Odom(scanPeriod, maxIterations); }
void loamdrv_odom_destroy(void* h) { delete (Odom*)h; }
void loamdrv_odom_set_inputs(void* h, const float* sharp, int n_sharp, const float* less_sharp, int n_less_sharp,
                             const float* flat, int n_flat, const float* less_flat, int n_less_flat, const float* full,
                             int n_full) {
  Odom* o = (Odom*)h;
  fill(*o->sharp, sharp, n_sharp); fill(*o->lessSharp, less_sharp, n_less_sharp); fill(*o->flat, flat, n_flat);
  fill(*o->lessFlat, less_flat, n_less_flat); fill(*o->full, full, n_full);
}
void loamdrv_odom_process(void* h) { ((Odom*)h)->process(); }
void loamdrv_odom_full_to_end(void* h) { Odom* o = (Odom*)h; o->toEnd(*o->full); }
void loamdrv_odom_get_twist(void* h, int which, float* out6) { Odom* o = (Odom*)h; twist6(which == 0 ? o->tf : o->sum, out6); }
static const Cloud& odom_cloud(Odom* o, int w) { return w == 0 ? *o->lastCorner : w == 1 ? *o->lastSurf : *o->full; }
int loamdrv_odom_cloud_size(void* h, int which) { return (int)odom_cloud((Odom*)h, which).size(); }
void loamdrv_odom_cloud_copy(void* h, int which, float* out) { dump(odom_cloud((Odom*)h, which), out); }

void* loamdrv_map_create(float scanPeriod, int maxIterations) { return new Mapping(scanPeriod, maxIterations); }
void loamdrv_map_destroy(void* h) { delete (Mapping*)h; }
void loamdrv_map_seed(void* h, int kind, const float* pts, int n) { ((Mapping*)h)->seed(kind, pts, n); }
void loamdrv_map_set_inputs(void* h, const float* corner_last, int n_corner, const float* surf_last, int n_surf,
                            const float* full, int n_full) {
  Mapping* m = (Mapping*)h;
  fill(*m->cornerLast, corner_last, n_corner); fill(*m->surfLast, surf_last, n_surf); fill(*m->fullRes, full, n_full);
}
void loamdrv_map_update_odometry(void* h, const float* s) {
  Mapping* m = (Mapping*)h;
  m->sum.rx = s[0]; m->sum.ry = s[1]; m->sum.rz = s[2];
  m->sum.t.x = s[3]; m->sum.t.y = s[4]; m->sum.t.z = s[5];
}
int loamdrv_map_process(void* h) { return ((Mapping*)h)->process() ? 1 : 0; }
void loamdrv_map_get_twist(void* h, int which, float* out6) {
  Mapping* m = (Mapping*)h;
  twist6(which == 0 ? m->aft : which == 1 ? m->bef : m->tobe, out6);
}
int loamdrv_map_cloud_size(void* h, int which) { return (int)((Mapping*)h)->cloud(which).size(); }
void loamdrv_map_cloud_copy(void* h, int which, float* out) { dump(((Mapping*)h)->cloud(which), out); }

void* loamdrv_pipeline_create(float scanPeriod, int odomMaxIter, int mapMaxIter) { return new Pipe(scanPeriod, odomMaxIter, mapMaxIter); }
void loamdrv_pipeline_destroy(void* h) { delete (Pipe*)h; }
void loamdrv_pipeline_seed_map(void* h, int kind, const float* pts, int n) { ((Pipe*)h)->map.seed(kind, pts, n); }
void* loamdrv_pipeline_scanreg(void* h) { return &((Pipe*)h)->reg; }
void* loamdrv_pipeline_odom(void* h) { return &((Pipe*)h)->odom; }
void* loamdrv_pipeline_map(void* h) { return &((Pipe*)h)->map; }
int loamdrv_pipeline_sweep(void* hh, const float* pts, const int* ring_sizes, int n_rings, float* odom_sum6,
                           float* map_aft6, double* st) {
  Pipe* h = (Pipe*)hh;
  double t0 = now();
  h->reg.process(pts, ring_sizes, n_rings);
  double t1 = now();
  Odom& o = h->odom;
  *o.sharp = h->reg.r.sharp; *o.lessSharp = h->reg.r.lessSharp; *o.flat = h->reg.r.flat;
  *o.lessFlat = h->reg.r.lessFlat; *o.full = h->reg.r.full;
  o.process();
  double t2 = now();
  o.toEnd(*o.full);
  double t3 = now();
  Mapping& m = h->map;
  *m.cornerLast = *o.lastCorner; *m.surfLast = *o.lastSurf; *m.fullRes = *o.full;
  m.sum = o.sum;
  int ok = m.process() ? 1 : 0;
  double t4 = now();
  twist6(o.sum, odom_sum6);
  twist6(m.aft, map_aft6);
  if (st) { st[0] = t1 - t0; st[1] = t2 - t1; st[2] = t3 - t2; st[3] = t4 - t3; st[4] = t4 - t0; }
  return ok;
}

int loamdrv_knn(const float* pts, int m, const float* queries, int nq, int k, int* idx_out, float* d2_out) {
  Cloud c;
  fill(c, pts, m);
  KdTree t;
  t.build(&c);
  std::vector<int> ki(k);
  std::vector<float> kd(k);
  for (int q = 0; q < nq; q++) {
    Pt p; p.x = queries[4 * q]; p.y = queries[4 * q + 1]; p.z = queries[4 * q + 2];
    int found = t.knn(p, k, ki.data(), kd.data());
    for (int j = 0; j < k; j++) { idx_out[q * k + j] = j < found ? ki[j] : -1; d2_out[q * k + j] = j < found ? kd[j] : -1.f; }
  }
  return 0;
}
double loamdrv_kdtree_build_seconds(const float* pts, int m) {
  Cloud c;
  fill(c, pts, m);
  KdTree t;
  double t0 = now();
  t.build(&c);
  return now() - t0;
}
int loamdrv_voxel_grid(const float* pts, int n, float leaf, float* out) {
  Cloud::Ptr c(new Cloud);
  fill(*c, pts, n);
  pcl::VoxelGrid<Pt> f;
  f.setLeafSize(leaf, leaf, leaf);
  f.setInputCloud(c);
  Cloud o;
  f.filter(o);
  dump(o, out);
  return (int)o.size();
}
void loamdrv_qr_solve6(const float* A, const float* b, float* x) {
  Eigen::Matrix<float, 6, 6> M;
  Eigen::Matrix<float, 6, 1> B, X;
  for (int i = 0; i < 6; i++) { B(i, 0) = b[i]; for (int j = 0; j < 6; j++) M(i, j) = A[i * 6 + j]; }
  X = M.colPivHouseholderQr().solve(B);
  for (int i = 0; i < 6; i++) x[i] = X(i, 0);
}
void loamdrv_eig_sym(const float* A, int n, float* evals, float* evecs) {
  if (n == 3) {
    Eigen::Matrix3f M;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M(i, j) = A[i * 3 + j];
    Eigen::SelfAdjointEigenSolver<Eigen::Matrix3f> es(M);
    for (int i = 0; i < 3; i++) evals[i] = es.eigenvalues()(i);
    for (int j = 0; j < 3; j++) for (int i = 0; i < 3; i++) evecs[i + j * 3] = es.eigenvectors()(i, j);
  } else {
    Eigen::Matrix<float, 6, 6> M;
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) M(i, j) = A[i * 6 + j];
    Eigen::SelfAdjointEigenSolver<Eigen::Matrix<float, 6, 6> > es(M);
    for (int i = 0; i < 6; i++) evals[i] = es.eigenvalues()(i);
    for (int j = 0; j < 6; j++) for (int i = 0; i < 6; i++) evecs[i + j * 6] = es.eigenvectors()(i, j);
  }
}
void loamdrv_lsq53(const float* A, float* x) {
  Eigen::Matrix<float, 5, 3> M;
  Eigen::Matrix<float, 5, 1> B;
  B.setConstant(-1);
  for (int i = 0; i < 5; i++) for (int j = 0; j < 3; j++) M(i, j) = A[i * 3 + j];
  Eigen::Vector3f X = M.colPivHouseholderQr().solve(B);
  for (int i = 0; i < 3; i++) x[i] = X(i, 0);
}

// ---- restatement-only extras for per-iteration GPU parity (not part of driver_api.h's common subset) ----
// odometry: run the correspondence pass of iteration `iter` with the current transform; return n_selected and the
// per-query coefficient / selected flag; AtA (36, row-major) / AtB (6) from the restated dense product
int loamorc_odom_iteration(void* h, int iter, const float* transform6, float* coeff, signed char* sel, float* AtA, float* AtB) {
  Odom* o = (Odom*)h;
  o->tf.rx = transform6[0]; o->tf.ry = transform6[1]; o->tf.rz = transform6[2];
  o->tf.t.x = transform6[3]; o->tf.t.y = transform6[4]; o->tf.t.z = transform6[5];
  const size_t nSharp = o->sharp->points.size(), nFlat = o->flat->points.size();
  if (o->c1.size() != nSharp) { o->c1.assign(nSharp, -1); o->c2.assign(nSharp, -1); }
  if (o->s1.size() != nFlat) { o->s1.assign(nFlat, -1); o->s2.assign(nFlat, -1); o->s3.assign(nFlat, -1); }
  std::vector<float> c;
  std::vector<signed char> s;
  const int n = o->correspond((size_t)iter, &c, &s);
  if (coeff) std::memcpy(coeff, c.data(), c.size() * sizeof(float));
  if (sel) std::memcpy(sel, s.data(), s.size());
  if (n > 0 && AtA) {
    o->gn.step(o->rows, o->rhs, false, 10.f);
    for (int i = 0; i < 6; i++) { AtB[i] = o->gn.lastAtB(i, 0); for (int j = 0; j < 6; j++) AtA[i * 6 + j] = o->gn.lastAtA(i, j); }
  }
  return n;
}
// correspondence indices kept from the last search iteration (BasicLaserOdometry.cpp:299-301, :431-434):
// ind[3 * q + {0, 1, 2}] = closest / second / third point of query q (sharp queries first; -1 = none; corners have no third)
void loamorc_odom_indices(void* h, int* ind) {
  Odom* o = (Odom*)h;
  const size_t nSharp = o->c1.size(), nFlat = o->s1.size();
  for (size_t i = 0; i < nSharp; i++) { ind[3 * i] = o->c1[i]; ind[3 * i + 1] = o->c2[i]; ind[3 * i + 2] = -1; }
  for (size_t i = 0; i < nFlat; i++) {
    ind[3 * (nSharp + i)] = o->s1[i]; ind[3 * (nSharp + i) + 1] = o->s2[i]; ind[3 * (nSharp + i) + 2] = o->s3[i];
  }
}
// make `last` clouds + trees from explicit inputs (as if a previous sweep had been processed)
void loamorc_odom_set_last(void* h, const float* corner, int nc, const float* surf, int ns) {
  Odom* o = (Odom*)h;
  fill(*o->lastCorner, corner, nc);
  fill(*o->lastSurf, surf, ns);
  o->cornerTree.build(o->lastCorner.get());
  o->surfTree.build(o->lastSurf.get());
  o->inited = true;
}
// mapping: set map clouds / query stacks directly and run one correspondence pass at pose tobe6
void loamorc_map_set(void* h, const float* corner_map, int ncm, const float* surf_map, int nsm, const float* corner_q,
                     int ncq, const float* surf_q, int nsq) {
  Mapping* m = (Mapping*)h;
  fill(*m->cornerFromMap, corner_map, ncm);
  fill(*m->surfFromMap, surf_map, nsm);
  fill(*m->cornerStackDS, corner_q, ncq);
  fill(*m->surfStackDS, surf_q, nsq);
  m->cornerTree.build(m->cornerFromMap.get());
  m->surfTree.build(m->surfFromMap.get());
}
int loamorc_map_iteration(void* h, const float* tobe6, float* coeff, signed char* sel, float* AtA, float* AtB) {
  Mapping* m = (Mapping*)h;
  m->tobe.rx = tobe6[0]; m->tobe.ry = tobe6[1]; m->tobe.rz = tobe6[2];
  m->tobe.t.x = tobe6[3]; m->tobe.t.y = tobe6[4]; m->tobe.t.z = tobe6[5];
  std::vector<float> c;
  std::vector<signed char> s;
  const int n = m->correspond(&c, &s);
  if (coeff) std::memcpy(coeff, c.data(), c.size() * sizeof(float));
  if (sel) std::memcpy(sel, s.data(), s.size());
  if (n > 0 && AtA) {
    m->gn.step(m->rows, m->rhs, false, 100.f);
    for (int i = 0; i < 6; i++) { AtB[i] = m->gn.lastAtB(i, 0); for (int j = 0; j < 6; j++) AtA[i * 6 + j] = m->gn.lastAtA(i, j); }
  }
  return n;
}

}  // extern "C"
