// TEST INFRASTRUCTURE -- not part of the shipped product.
//
// C driver around the UNMODIFIED reference classes (loam::BasicScanRegistration / BasicLaserOdometry /
// BasicLaserMapping from /root/reference/src/lib/*.cpp) compiled against oracle/shim/.  Built only where
// /root/reference exists (see oracle/Makefile), output oracle/_ref/libloam_ref.so.
// Compiled with -fno-access-control so the driver can read private clouds and pre-seed the cube map without
// touching the reference sources.
#include <chrono>
#include <cstring>
#include <vector>

#include "loam_velodyne/BasicLaserMapping.h"
#include "loam_velodyne/BasicLaserOdometry.h"
#include "loam_velodyne/BasicScanRegistration.h"
#include "loam_velodyne/MultiScanRegistration.h"
#include "loam_velodyne/BasicTransformMaintenance.h"
#include "loam_velodyne/nanoflann_pcl.h"
#include <Eigen/Eigenvalues>
#include <Eigen/QR>

#include "driver_api.h"

namespace {

typedef pcl::PointCloud<pcl::PointXYZI> Cloud;

void fill(Cloud& c, const float* p, int n) {
  c.clear();
  c.points.resize(n);
  for (int i = 0; i < n; i++) {
    c.points[i].x = p[4 * i + 0];
    c.points[i].y = p[4 * i + 1];
    c.points[i].z = p[4 * i + 2];
    c.points[i].intensity = p[4 * i + 3];
  }
  c.width = n;
  c.height = 1;
}
void dump(const Cloud& c, float* out) {
  for (size_t i = 0; i < c.points.size(); i++) {
    out[4 * i + 0] = c.points[i].x;
    out[4 * i + 1] = c.points[i].y;
    out[4 * i + 2] = c.points[i].z;
    out[4 * i + 3] = c.points[i].intensity;
  }
}
void twist6(const loam::Twist& t, float* o) {
  o[0] = t.rot_x.rad(); o[1] = t.rot_y.rad(); o[2] = t.rot_z.rad();
  o[3] = t.pos.x(); o[4] = t.pos.y(); o[5] = t.pos.z();
}

struct MapH {
  loam::BasicLaserMapping m;
  Cloud scratch;
  MapH(float sp, int it) : m(sp, it) {}
  const Cloud& cloud(int which) {
    switch (which) {
      case 0: return *m._laserCloudFullRes;
      case 1: return *m._laserCloudSurroundDS;
      case 2: return *m._laserCloudCornerFromMap;
      case 3: return *m._laserCloudSurfFromMap;
      case 4: return *m._laserCloudCornerStackDS;
      case 5: return *m._laserCloudSurfStackDS;
      case 6:
      case 7: {
        scratch.clear();
        auto& arr = which == 6 ? m._laserCloudCornerArray : m._laserCloudSurfArray;
        for (auto& c : arr) scratch += *c;
        return scratch;
      }
    }
    scratch.clear();
    return scratch;
  }
  void seed(int kind, const float* p, int n) {
    // cube index arithmetic of BasicLaserMapping.cpp:540-553
    const double CUBE_SIZE = 50.0, CUBE_HALF = 25.0;
    for (int i = 0; i < n; i++) {
      pcl::PointXYZI pt;
      pt.x = p[4 * i]; pt.y = p[4 * i + 1]; pt.z = p[4 * i + 2]; pt.intensity = p[4 * i + 3];
      int cubeI = int((pt.x + CUBE_HALF) / CUBE_SIZE) + m._laserCloudCenWidth;
      int cubeJ = int((pt.y + CUBE_HALF) / CUBE_SIZE) + m._laserCloudCenHeight;
      int cubeK = int((pt.z + CUBE_HALF) / CUBE_SIZE) + m._laserCloudCenDepth;
      if (pt.x + CUBE_HALF < 0) cubeI--;
      if (pt.y + CUBE_HALF < 0) cubeJ--;
      if (pt.z + CUBE_HALF < 0) cubeK--;
      if (cubeI >= 0 && cubeI < (int)m._laserCloudWidth && cubeJ >= 0 && cubeJ < (int)m._laserCloudHeight &&
          cubeK >= 0 && cubeK < (int)m._laserCloudDepth) {
        size_t ind = cubeI + m._laserCloudWidth * cubeJ + m._laserCloudWidth * m._laserCloudHeight * cubeK;
        (kind == 0 ? m._laserCloudCornerArray : m._laserCloudSurfArray)[ind]->push_back(pt);
      }
    }
  }
};

struct OdomH {
  loam::BasicLaserOdometry o;
  OdomH(float sp, int it) : o(sp, it) {}
};

struct RegH {
  loam::BasicScanRegistration r;
  std::vector<Cloud> rings;
  const Cloud& cloud(int which) {
    switch (which) {
      case 0: return r.laserCloud();
      case 1: return r.cornerPointsSharp();
      case 2: return r.cornerPointsLessSharp();
      case 3: return r.surfacePointsFlat();
      default: return r.surfacePointsLessFlat();
    }
  }
};

// the reference's ring-binning adapter (ROS names resolved by oracle/shim/ros etc.; no node is ever set up)
struct MsH {
  loam::MultiScanRegistration m;
  MsH(float lo, float hi, int n) : m(loam::MultiScanMapper(lo, hi, (uint16_t)n)) {}
  loam::BasicScanRegistration& base() { return (loam::BasicScanRegistration&)m; }
  const Cloud& cloud(int which) {
    switch (which) {
      case 0: return base().laserCloud();
      case 1: return base().cornerPointsSharp();
      case 2: return base().cornerPointsLessSharp();
      case 3: return base().surfacePointsFlat();
      default: return base().surfacePointsLessFlat();
    }
  }
};

struct PipeH {
  RegH reg;
  OdomH odom;
  MapH map;
  PipeH(float sp, int oi, int mi) : odom(sp, oi), map(sp, mi) {}
};

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int reg_process(RegH* h, const float* pts, const int* ring_sizes, int n_rings) {
  h->rings.resize(n_rings);
  int off = 0;
  for (int r = 0; r < n_rings; r++) {
    fill(h->rings[r], pts + 4 * off, ring_sizes[r]);
    off += ring_sizes[r];
  }
  h->r.processScanlines(loam::Time(), h->rings);
  return off;
}

}  // namespace

extern "C" {

const char* loamdrv_kind(void) { return "reference"; }

void* loamdrv_scanreg_create(void) { return new RegH(); }
void loamdrv_scanreg_destroy(void* h) { delete (RegH*)h; }
void loamdrv_scanreg_configure(void* h, float scanPeriod, int nFeatureRegions, int curvatureRegion, int maxCornerSharp,
                               int maxSurfaceFlat, float lessFlatFilterSize, float surfaceCurvatureThreshold) {
  loam::RegistrationParams p(scanPeriod, 200, nFeatureRegions, curvatureRegion, maxCornerSharp, maxSurfaceFlat,
                             lessFlatFilterSize, surfaceCurvatureThreshold);
  ((RegH*)h)->r.configure(p);
}
int loamdrv_scanreg_process(void* h, const float* pts, const int* ring_sizes, int n_rings) {
  return reg_process((RegH*)h, pts, ring_sizes, n_rings);
}
int loamdrv_scanreg_cloud_size(void* h, int which) { return (int)((RegH*)h)->cloud(which).size(); }
void loamdrv_scanreg_cloud_copy(void* h, int which, float* out) { dump(((RegH*)h)->cloud(which), out); }

void loamdrv_transform_maintenance(const float* sum, const float* bef, const float* aft, float* out) {
  loam::BasicTransformMaintenance tm;
  tm.updateOdometry(sum[0], sum[1], sum[2], sum[3], sum[4], sum[5]);
  tm.updateMappingTransform(aft[0], aft[1], aft[2], aft[3], aft[4], aft[5], bef[0], bef[1], bef[2], bef[3], bef[4], bef[5]);
  tm.transformAssociateToMap();
  for (int i = 0; i < 6; i++) out[i] = tm.transformMapped()[i];
}

void* loamdrv_multiscan_create(float lo, float hi, int n) { return new MsH(lo, hi, n); }
void loamdrv_multiscan_destroy(void* h) { delete (MsH*)h; }
int loamdrv_multiscan_process(void* h, const float* xyz, int n) {
  pcl::PointCloud<pcl::PointXYZ> in;
  in.points.resize(n);
  for (int i = 0; i < n; i++) {
    in.points[i].x = xyz[3 * i + 0];
    in.points[i].y = xyz[3 * i + 1];
    in.points[i].z = xyz[3 * i + 2];
  }
  in.width = n;
  in.height = 1;
  MsH* ms = (MsH*)h;
  ms->m.process(in, loam::Time());
  size_t kept = 0;
  for (auto& r : ms->m._laserCloudScans) kept += r.size();
  return (int)kept;
}
void loamdrv_multiscan_binned(void* h, float* out, int* ring_sizes) {
  MsH* ms = (MsH*)h;
  size_t off = 0;
  for (size_t r = 0; r < ms->m._laserCloudScans.size(); r++) {
    dump(ms->m._laserCloudScans[r], out + 4 * off);
    ring_sizes[r] = (int)ms->m._laserCloudScans[r].size();
    off += ms->m._laserCloudScans[r].size();
  }
}
int loamdrv_multiscan_cloud_size(void* h, int which) { return (int)((MsH*)h)->cloud(which).size(); }
void loamdrv_multiscan_cloud_copy(void* h, int which, float* out) { dump(((MsH*)h)->cloud(which), out); }

void* loamdrv_odom_create(float scanPeriod, int maxIterations) { return new OdomH(scanPeriod, maxIterations); }
void loamdrv_odom_destroy(void* h) { delete (OdomH*)h; }
void loamdrv_odom_set_inputs(void* h, const float* sharp, int n_sharp, const float* less_sharp, int n_less_sharp,
                             const float* flat, int n_flat, const float* less_flat, int n_less_flat,
                             const float* full, int n_full) {
  auto& o = ((OdomH*)h)->o;
  fill(*o.cornerPointsSharp(), sharp, n_sharp);
  fill(*o.cornerPointsLessSharp(), less_sharp, n_less_sharp);
  fill(*o.surfPointsFlat(), flat, n_flat);
  fill(*o.surfPointsLessFlat(), less_flat, n_less_flat);
  fill(*o.laserCloud(), full, n_full);
}
void loamdrv_odom_process(void* h) { ((OdomH*)h)->o.process(); }
void loamdrv_odom_full_to_end(void* h) {
  auto& o = ((OdomH*)h)->o;
  o.transformToEnd(o.laserCloud());
}
void loamdrv_odom_get_twist(void* h, int which, float* out6) {
  auto& o = ((OdomH*)h)->o;
  twist6(which == 0 ? o.transform() : o.transformSum(), out6);
}
static const Cloud& odom_cloud(OdomH* h, int which) {
  return which == 0 ? *h->o.lastCornerCloud() : which == 1 ? *h->o.lastSurfaceCloud() : *h->o.laserCloud();
}
int loamdrv_odom_cloud_size(void* h, int which) { return (int)odom_cloud((OdomH*)h, which).size(); }
void loamdrv_odom_cloud_copy(void* h, int which, float* out) { dump(odom_cloud((OdomH*)h, which), out); }

void* loamdrv_map_create(float scanPeriod, int maxIterations) { return new MapH(scanPeriod, maxIterations); }
void loamdrv_map_destroy(void* h) { delete (MapH*)h; }
void loamdrv_map_seed(void* h, int kind, const float* pts, int n) { ((MapH*)h)->seed(kind, pts, n); }
void loamdrv_map_set_inputs(void* h, const float* corner_last, int n_corner, const float* surf_last, int n_surf,
                            const float* full, int n_full) {
  auto& m = ((MapH*)h)->m;
  fill(m.laserCloudCornerLast(), corner_last, n_corner);
  fill(m.laserCloudSurfLast(), surf_last, n_surf);
  fill(m.laserCloud(), full, n_full);
}
void loamdrv_map_update_odometry(void* h, const float* s) {
  loam::Twist t;
  t.rot_x = s[0]; t.rot_y = s[1]; t.rot_z = s[2];
  t.pos = loam::Vector3(s[3], s[4], s[5]);
  ((MapH*)h)->m.updateOdometry(t);
}
int loamdrv_map_process(void* h) { return ((MapH*)h)->m.process(loam::Time()) ? 1 : 0; }
void loamdrv_map_get_twist(void* h, int which, float* out6) {
  auto& m = ((MapH*)h)->m;
  twist6(which == 0 ? m._transformAftMapped : which == 1 ? m._transformBefMapped : m._transformTobeMapped, out6);
}
int loamdrv_map_cloud_size(void* h, int which) { return (int)((MapH*)h)->cloud(which).size(); }
void loamdrv_map_cloud_copy(void* h, int which, float* out) { dump(((MapH*)h)->cloud(which), out); }

void* loamdrv_pipeline_create(float scanPeriod, int odomMaxIter, int mapMaxIter) {
  return new PipeH(scanPeriod, odomMaxIter, mapMaxIter);
}
void loamdrv_pipeline_destroy(void* h) { delete (PipeH*)h; }
void loamdrv_pipeline_seed_map(void* h, int kind, const float* pts, int n) { ((PipeH*)h)->map.seed(kind, pts, n); }
void* loamdrv_pipeline_scanreg(void* h) { return &((PipeH*)h)->reg; }
void* loamdrv_pipeline_odom(void* h) { return &((PipeH*)h)->odom; }
void* loamdrv_pipeline_map(void* h) { return &((PipeH*)h)->map; }

int loamdrv_pipeline_sweep(void* hh, const float* pts, const int* ring_sizes, int n_rings, float* odom_sum6,
                           float* map_aft6, double* st) {
  PipeH* h = (PipeH*)hh;
  double t0 = now();
  reg_process(&h->reg, pts, ring_sizes, n_rings);
  double t1 = now();
  // what the ROS hop ScanRegistration::publishResult -> LaserOdometry::*Handler moves (ScanRegistration.cpp:187,
  // LaserOdometry.cpp:178-238): plain copies of the five clouds
  auto& o = h->odom.o;
  *o.cornerPointsSharp() = h->reg.r.cornerPointsSharp();
  *o.cornerPointsLessSharp() = h->reg.r.cornerPointsLessSharp();
  *o.surfPointsFlat() = h->reg.r.surfacePointsFlat();
  *o.surfPointsLessFlat() = h->reg.r.surfacePointsLessFlat();
  *o.laserCloud() = h->reg.r.laserCloud();
  o.updateIMU(h->reg.r.imuTransform());
  o.process();
  double t2 = now();
  o.transformToEnd(o.laserCloud());  // LaserOdometry.cpp:326
  double t3 = now();
  auto& m = h->map.m;
  m.laserCloudCornerLast() = *o.lastCornerCloud();
  m.laserCloudSurfLast() = *o.lastSurfaceCloud();
  m.laserCloud() = *o.laserCloud();
  m.updateOdometry(o.transformSum());
  int ok = m.process(loam::Time()) ? 1 : 0;
  double t4 = now();
  twist6(o.transformSum(), odom_sum6);
  twist6(m.transformAftMapped(), map_aft6);
  if (st) { st[0] = t1 - t0; st[1] = t2 - t1; st[2] = t3 - t2; st[3] = t4 - t3; st[4] = t4 - t0; }
  return ok;
}

int loamdrv_knn(const float* pts, int m, const float* queries, int nq, int k, int* idx_out, float* d2_out) {
  Cloud::Ptr c(new Cloud);
  fill(*c, pts, m);
  nanoflann::KdTreeFLANN<pcl::PointXYZI> tree;
  tree.setInputCloud(c);
  std::vector<int> ind(k);
  std::vector<float> d2(k);
  for (int q = 0; q < nq; q++) {
    pcl::PointXYZI p;
    p.x = queries[4 * q]; p.y = queries[4 * q + 1]; p.z = queries[4 * q + 2];
    ind.assign(k, -1);
    d2.assign(k, 0.f);
    int found = tree.nearestKSearch(p, k, ind, d2);
    for (int j = 0; j < k; j++) {
      idx_out[q * k + j] = j < found ? ind[j] : -1;
      d2_out[q * k + j] = j < found ? d2[j] : -1.f;
    }
  }
  return 0;
}
double loamdrv_kdtree_build_seconds(const float* pts, int m) {
  Cloud::Ptr c(new Cloud);
  fill(*c, pts, m);
  nanoflann::KdTreeFLANN<pcl::PointXYZI> tree;
  double t0 = now();
  tree.setInputCloud(c);
  return now() - t0;
}
int loamdrv_voxel_grid(const float* pts, int n, float leaf, float* out) {
  Cloud::Ptr c(new Cloud);
  fill(*c, pts, n);
  pcl::VoxelGrid<pcl::PointXYZI> f;
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
  for (int i = 0; i < 6; i++) {
    B(i, 0) = b[i];
    for (int j = 0; j < 6; j++) M(i, j) = A[i * 6 + j];
  }
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

}  // extern "C"
