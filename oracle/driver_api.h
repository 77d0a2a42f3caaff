/* TEST INFRASTRUCTURE -- not part of the shipped product.
 *
 * One C driver API, implemented twice:
 *   oracle/_ref/libloam_ref.so   : the UNMODIFIED reference sources (/root/reference/src/lib/Basic*.cpp +
 *                                  include/loam_velodyne/nanoflann.hpp) compiled against oracle/shim/ (ref_driver.cpp)
 *   oracle/liboracle.so          : the CPU restatement (loam_oracle.cpp), buildable without /root/reference
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load either.
 *
 * Points are packed float[4] = (x, y, z, intensity).  Twists are float[6] = (rot_x, rot_y, rot_z, pos.x, pos.y, pos.z).
 */
#ifndef LOAM_ORACLE_DRIVER_API_H
#define LOAM_ORACLE_DRIVER_API_H

#ifdef __cplusplus
extern "C" {
#endif

/* returns "reference" or "restatement" */
const char* loamdrv_kind(void);

/* ---- scan registration (BasicScanRegistration.cpp:28-46,155-254) ---- */
void* loamdrv_scanreg_create(void);
void loamdrv_scanreg_destroy(void* h);
/* params: scanPeriod, nFeatureRegions, curvatureRegion, maxCornerSharp, maxSurfaceFlat, lessFlatFilterSize, surfaceCurvatureThreshold
 * (RegistrationParams, BasicScanRegistration.h:37-44; maxCornerLessSharp = 10*maxCornerSharp) */
void loamdrv_scanreg_configure(void* h, float scanPeriod, int nFeatureRegions, int curvatureRegion, int maxCornerSharp,
                               int maxSurfaceFlat, float lessFlatFilterSize, float surfaceCurvatureThreshold);
/* ring r holds ring_sizes[r] consecutive points of pts */
int loamdrv_scanreg_process(void* h, const float* pts, const int* ring_sizes, int n_rings);
/* which: 0 laserCloud, 1 cornerPointsSharp, 2 cornerPointsLessSharp, 3 surfacePointsFlat, 4 surfacePointsLessFlat */
int loamdrv_scanreg_cloud_size(void* h, int which);
void loamdrv_scanreg_cloud_copy(void* h, int which, float* out);

/* ---- ring binning front end (MultiScanRegistration.cpp:44-67 mapper, :160-238 process): an unordered sensor-frame xyz
 * cloud -> axis swap, NaN / near-zero rejection, ring from the vertical angle, relative time from the azimuth (with the
 * half-sweep bookkeeping), intensity = ring + relTime, stable binning per ring -> then the regular registration ---- */
void* loamdrv_multiscan_create(float lower_bound_deg, float upper_bound_deg, int n_rings);
void loamdrv_multiscan_destroy(void* h);
/* xyz: n x 3 floats in the sensor frame; returns the number of points kept */
int loamdrv_multiscan_process(void* h, const float* xyz, int n);
/* ring-ordered cloud as handed to processScanlines: out_pts4 (kept x 4 floats), ring_sizes (n_rings ints) */
void loamdrv_multiscan_binned(void* h, float* out_pts4, int* ring_sizes);
/* registration results, same ids as loamdrv_scanreg_cloud_* */
int loamdrv_multiscan_cloud_size(void* h, int which);
void loamdrv_multiscan_cloud_copy(void* h, int which, float* out);

/* ---- transform maintenance (BasicTransformMaintenance.cpp:45-178): updateOdometry(sum) + updateMappingTransform(aft, bef)
 * + transformAssociateToMap -> transformMapped ---- */
void loamdrv_transform_maintenance(const float* sum6, const float* bef6, const float* aft6, float* mapped_out6);

/* ---- laser odometry (BasicLaserOdometry.cpp:196-666) ---- */
void* loamdrv_odom_create(float scanPeriod, int maxIterations);
void loamdrv_odom_destroy(void* h);
void loamdrv_odom_set_inputs(void* h, const float* sharp, int n_sharp, const float* less_sharp, int n_less_sharp,
                             const float* flat, int n_flat, const float* less_flat, int n_less_flat,
                             const float* full, int n_full);
void loamdrv_odom_process(void* h);
/* transformToEnd(laserCloud()) as LaserOdometry::publishResult does (LaserOdometry.cpp:326) */
void loamdrv_odom_full_to_end(void* h);
/* which: 0 transform, 1 transformSum */
void loamdrv_odom_get_twist(void* h, int which, float* out6);
/* which: 0 lastCornerCloud, 1 lastSurfaceCloud, 2 laserCloud */
int loamdrv_odom_cloud_size(void* h, int which);
void loamdrv_odom_cloud_copy(void* h, int which, float* out);

/* ---- laser mapping (BasicLaserMapping.cpp:266-599,626-926) ---- */
void* loamdrv_map_create(float scanPeriod, int maxIterations);
void loamdrv_map_destroy(void* h);
/* kind 0 corner / 1 surface: push points (already in map frame) into their 50 m cubes with the reference's own
 * cube-index arithmetic (BasicLaserMapping.cpp:540-553); used to pre-seed maps of a given size */
void loamdrv_map_seed(void* h, int kind, const float* pts, int n);
void loamdrv_map_set_inputs(void* h, const float* corner_last, int n_corner, const float* surf_last, int n_surf,
                            const float* full, int n_full);
void loamdrv_map_update_odometry(void* h, const float* sum6);
int loamdrv_map_process(void* h);
/* which: 0 transformAftMapped, 1 transformBefMapped, 2 transformTobeMapped */
void loamdrv_map_get_twist(void* h, int which, float* out6);
/* which: 0 laserCloud (registered full-res), 1 laserCloudSurroundDS, 2 cornerFromMap, 3 surfFromMap,
 *        4 cornerStackDS, 5 surfStackDS, 6 all corner cubes concatenated, 7 all surf cubes concatenated */
int loamdrv_map_cloud_size(void* h, int which);
void loamdrv_map_cloud_copy(void* h, int which, float* out);

/* ---- whole pipeline: registration -> odometry -> mapping on one sweep (§8b "who calls it") ---- */
void* loamdrv_pipeline_create(float scanPeriod, int odomMaxIter, int mapMaxIter);
void loamdrv_pipeline_destroy(void* h);
void loamdrv_pipeline_seed_map(void* h, int kind, const float* pts, int n);
/* stage_seconds[5]: registration, odometry, full->end, mapping, total (steady_clock) */
int loamdrv_pipeline_sweep(void* h, const float* pts, const int* ring_sizes, int n_rings, float* odom_sum6,
                           float* map_aft6, double* stage_seconds);
/* sub-object handles for the accessors above (borrowed, owned by the pipeline) */
void* loamdrv_pipeline_scanreg(void* h);
void* loamdrv_pipeline_odom(void* h);
void* loamdrv_pipeline_map(void* h);

/* ---- pieces ---- */
/* exact k-NN of each query against pts (nanoflann_pcl.h:131-152 semantics: ascending d2, L2 in float x->y->z) */
int loamdrv_knn(const float* pts, int m, const float* queries, int nq, int k, int* idx_out, float* d2_out);
/* build the tree only; returns seconds */
double loamdrv_kdtree_build_seconds(const float* pts, int m);
/* pcl::VoxelGrid<PointXYZI>; returns number of output points written to out (capacity n) */
int loamdrv_voxel_grid(const float* pts, int n, float leaf, float* out);
/* dense helpers the solve uses: x = colPivHouseholderQr(A 6x6 row-major).solve(b); eigen (ascending) of sym 6x6 / 3x3 */
void loamdrv_qr_solve6(const float* A36, const float* b6, float* x6);
void loamdrv_eig_sym(const float* A, int n, float* evals, float* evecs_colmajor);
void loamdrv_lsq53(const float* A15_rowmajor, float* x3); /* A x = -1 */

#ifdef __cplusplus
}
#endif
#endif
