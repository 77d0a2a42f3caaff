/* loam_b200_host.h -- C handles over the C++ drop-in classes loam::BasicScanRegistration / BasicLaserOdometry /
 * BasicLaserMapping (the headers under loam_velodyne_b200/csrc/host/loam_velodyne/), for callers that cannot include C++ headers
 * (the Python package, bench.py, ctypes tests).  C++ callers -- including the reference's ROS adapters, which derive
 * from the Basic* classes (ScanRegistration.h:55, LaserOdometry.h:56, LaserMapping.h:54 upstream) -- use the classes
 * directly.
 *
 * Points are packed float[4] = (x, y, z, intensity); twists are float[6] = (rot_x, rot_y, rot_z, pos.x, pos.y, pos.z).
 * Functions returning int give 0 on success and a negative value on failure; loam_b200_host_last_error() has the text
 * (per thread).  A failure here is always loud: there is no CPU fallback behind these handles.
 */
#ifndef LOAM_B200_HOST_H
#define LOAM_B200_HOST_H

#ifdef __cplusplus
extern "C" {
#endif

const char* loam_b200_host_last_error(void);
/* CUDA device used by objects created afterwards (default: $LOAM_B200_DEVICE, else $LOCAL_RANK, else 0) */
void loam_b200_host_set_device(int device);

/* The 6 x 6 Gauss-Newton step of both loops (BasicLaserOdometry.cpp:559-600, BasicLaserMapping.cpp:867-908) as the
 * library solves it on host and device (csrc/lmstep.cuh: gn_solve): x = colPivHouseholderQr(AtA).solve(AtB); on the first
 * iteration eigenvalues of AtA below eigen_threshold switch the degeneracy projection on.  Pure host arithmetic (no GPU):
 * exposed for the parity tests.  AtA row-major 36 floats; *degenerate_out receives 0 / 1. */
int loam_b200_host_gn_solve(const float* AtA, const float* AtB, int first_iteration, float eigen_threshold, float* x_out6,
                            int* degenerate_out);

/* ---- loam::BasicTransformMaintenance (host-only pose fusion, BasicTransformMaintenance.cpp:45-178) ----
 * sum / bef / aft: rot_x, rot_y, rot_z, x, y, z of the odometry pose, the odometry pose at the last mapping update and the
 * mapped pose of that update (updateOdometry + updateMappingTransform); mapped_out6 = transformMapped() after
 * transformAssociateToMap(). */
int loam_b200_transform_maintenance(const float* sum6, const float* bef6, const float* aft6, float* mapped_out6);

/* ---- loam::BasicScanRegistration ---- */
void* loam_b200_scanreg_create(void);
void loam_b200_scanreg_destroy(void* h);
int loam_b200_scanreg_configure(void* h, float scanPeriod, int nFeatureRegions, int curvatureRegion, int maxCornerSharp,
                                int maxSurfaceFlat, float lessFlatFilterSize, float surfaceCurvatureThreshold);
/* processScanlines: ring r = ring_sizes[r] consecutive points */
int loam_b200_scanreg_process(void* h, const float* pts, const int* ring_sizes, int n_rings);
/* processUnorderedSweep: MultiScanRegistration::process (MultiScanRegistration.cpp:160-238) -- n unordered sensor-frame
 * xyz points (3 floats each) binned into rings on the GPU with MultiScanMapper(lower, upper, n_rings), then extraction */
int loam_b200_scanreg_process_unordered(void* h, const float* xyz, int n, float lower_bound_deg, float upper_bound_deg,
                                        int n_rings);
/* which: 0 laserCloud, 1 cornerPointsSharp, 2 cornerPointsLessSharp, 3 surfacePointsFlat, 4 surfacePointsLessFlat */
int loam_b200_scanreg_cloud_size(void* h, int which);
int loam_b200_scanreg_cloud_copy(void* h, int which, float* out);
/* which: 1 sharp, 2 less sharp, 3 flat -> indices into laserCloud */
int loam_b200_scanreg_index_size(void* h, int which);
int loam_b200_scanreg_index_copy(void* h, int which, int* out);

/* ---- loam::BasicLaserOdometry ---- */
void* loam_b200_odom_create(float scanPeriod, int maxIterations);
void loam_b200_odom_destroy(void* h);
int loam_b200_odom_set_inputs(void* h, const float* sharp, int n_sharp, const float* less_sharp, int n_less_sharp,
                              const float* flat, int n_flat, const float* less_flat, int n_less_flat,
                              const float* full, int n_full);
int loam_b200_odom_process(void* h);
int loam_b200_odom_full_to_end(void* h); /* transformToEnd(laserCloud()), as LaserOdometry::publishResult does */
/* which: 0 transform, 1 transformSum */
int loam_b200_odom_get_twist(void* h, int which, float* out6);
/* which: 0 lastCornerCloud, 1 lastSurfaceCloud, 2 laserCloud */
int loam_b200_odom_cloud_size(void* h, int which);
int loam_b200_odom_cloud_copy(void* h, int which, float* out);
int loam_b200_odom_last_iterations(void* h);

/* ---- loam::BasicLaserMapping ---- */
void* loam_b200_map_create(float scanPeriod, int maxIterations);
void loam_b200_map_destroy(void* h);
int loam_b200_map_seed(void* h, const float* corner, int n_corner, const float* surf, int n_surf);
int loam_b200_map_set_inputs(void* h, const float* corner_last, int n_corner, const float* surf_last, int n_surf,
                             const float* full, int n_full);
int loam_b200_map_update_odometry(void* h, const float* sum6);
int loam_b200_map_process(void* h); /* 1 processed, 0 frame skipped, <0 error */
/* which: 0 transformAftMapped, 1 transformBefMapped, 2 transformTobeMapped */
int loam_b200_map_get_twist(void* h, int which, float* out6);
/* which: 0 laserCloud (registered), 1 laserCloudSurroundDS, 2 cornerFromMap, 3 surfFromMap, 4 cornerStackDS,
 *        5 surfStackDS, 6 all corner cubes, 7 all surf cubes */
int loam_b200_map_cloud_size(void* h, int which);
int loam_b200_map_cloud_copy(void* h, int which, float* out);
int loam_b200_map_last_iterations(void* h);
/* host wall seconds of the last process(): begin_sweep, LM loop, end_sweep, surround map */
int loam_b200_map_last_phase_seconds(void* h, double* out4);
/* measurement hook: the scan-to-map iteration kernel exactly as this object launches it (own context, persistent map, the
 * queries of the last process()), CUDA-event time over `reps` launches at the current pose.  out5 = average launch
 * microseconds, queries, table probes / query, candidate points / query, selected correspondences */
int loam_b200_map_kernel_profile(void* h, int reps, double* out5);
/* ... on n caller-supplied surface queries (packed float4, map frame, identity pose) against this object's persistent map:
 * the k-NN bandwidth stress (queries spread over a map far larger than L2).  Replaces the queries of the last process(). */
int loam_b200_map_kernel_profile_queries(void* h, const float* queries, int n, int reps, double* out5);
/* test hook: keep a copy of laserCloudCornerFromMap / laserCloudSurfFromMap (cloud ids 2, 3) at every process() */
int loam_b200_map_retain_from_map(void* h, int on);
/* multi-GPU, one process per GPU: rank 0 calls loam_b200_host_nccl_unique_id and distributes the 128 bytes; every rank
 * then enables sharding on its mapping object (nccl_id128 = NULL: slice the queries only, no collective) */
int loam_b200_host_nccl_unique_id(unsigned char* out128);
int loam_b200_map_enable_sharding(void* h, int rank, int world, const unsigned char* nccl_id128);

/* multi-GPU with the MAP sharded by cube slabs along x (include/loam_b200.h "peer"): every rank exports its 64-byte inbox
 * handle, the caller all-gathers them, every rank enables sharding with all handles in rank order.  A rank then stores
 * only its slabs (+ 2 m halo; loam_b200_map_seed keeps the points that belong to it) and the per-iteration normal equations
 * are all-reduced inside the iteration kernel over NVLink peer memory.  _local: objects of one process, hs[r] = rank r. */
int loam_b200_map_peer_export(void* h, unsigned char* out64);
int loam_b200_map_enable_cube_sharding(void* h, int rank, int world, const unsigned char* handles, int slab_metres);
int loam_b200_map_enable_cube_sharding_local(void** hs, int world, int slab_metres);
/* unmap the peers' inboxes: every rank calls this and the ranks synchronise BEFORE any of them destroys its object (memory
 * exported over CUDA IPC must not be freed while another process still has it mapped) */
int loam_b200_map_disable_cube_sharding(void* h);

/* ---- the three chained in-process: registration -> odometry -> mapping on one sweep ---- */
void* loam_b200_pipeline_create(float scanPeriod, int odomMaxIter, int mapMaxIter);
void loam_b200_pipeline_destroy(void* h);
int loam_b200_pipeline_seed_map(void* h, const float* corner, int n_corner, const float* surf, int n_surf);
/* stage_seconds[5]: registration, odometry, full->end, mapping, total (steady_clock, host wall time) */
int loam_b200_pipeline_sweep(void* h, const float* pts, const int* ring_sizes, int n_rings, float* odom_sum6,
                             float* map_aft6, double* stage_seconds);
/* same, the sweep already resident in GPU memory (d_pts: device pointer to the packed points) */
int loam_b200_pipeline_sweep_device(void* h, const void* d_pts, const int* ring_sizes, int n_rings, float* odom_sum6,
                                    float* map_aft6, double* stage_seconds);
/* same chain, but every hand-off goes through host pcl clouds and the reference's own entry points
 * (processScanlines(vector<PointCloud>) / cloud Ptr accessors), as separate ROS nodes would use the classes */
int loam_b200_pipeline_sweep_hostclouds(void* h, const float* pts, const int* ring_sizes, int n_rings, float* odom_sum6,
                                        float* map_aft6, double* stage_seconds);
/* Streaming form: registration, odometry and mapping run as three concurrent single-threaded stages over consecutive
 * sweeps -- the way the reference's three ROS nodes run (CMakeLists.txt:40-50) -- connected by the device-to-device
 * hand-offs; results equal the sequential chain bit for bit.  submit() queues one sweep (host `pts` or device `d_pts`,
 * the other NULL; the buffer must stay valid until its result has been collected) and blocks only while two sweeps are
 * already waiting; collect() returns the oldest finished sweep: 1 = result written, 0 = nothing pending (or, with
 * wait == 0, nothing ready yet), -1 = a stage failed (loam_b200_host_last_error).  Do not mix with the _sweep calls
 * while sweeps are in flight. */
int loam_b200_pipeline_submit(void* h, const float* pts, const void* d_pts, const int* ring_sizes, int n_rings);
int loam_b200_pipeline_collect(void* h, int wait, float* odom_sum6, float* map_aft6, int* ok);
/* out9 = seconds the registration / odometry / mapping stage threads spent [0..2] working, [3..5] waiting for a neighbour
 * stage, [6..8] in the adopt() hand-offs since streaming started (or the last reset); call while no sweep is in flight */
int loam_b200_pipeline_stage_seconds(void* h, double* out9, int reset);
/* wait until everything the three stage objects enqueued (or posted to their helper threads) has finished on the GPU */
int loam_b200_pipeline_sync(void* h);
void* loam_b200_pipeline_scanreg(void* h);
void* loam_b200_pipeline_odom(void* h);
void* loam_b200_pipeline_map(void* h);

#ifdef __cplusplus
}
#endif
#endif
