/* loam_b200.h -- C ABI of libloam_b200.so: the B200 (sm_100a) implementation of LOAM's per-sweep registration hot
 * path behind the reference's Basic* classes (laboshinl/loam_velodyne).
 *
 * Plain pointers and sizes only.  Every entry point returns 0 on success or a negative loam_b200_status; nothing
 * throws across this boundary.  A context owns one device, one stream and all device memory; it is not thread-safe
 * (the reference classes are used from one thread each, e.g. LaserOdometry.cpp:254-270) but several contexts may
 * coexist.  Unless stated otherwise host buffers are only read/written during the call (the call synchronises its
 * stream before returning), so callers may reuse them immediately.
 *
 * Points are packed float[4] = (x, y, z, intensity); indices are int32 (nanoflann_pcl.h:102).
 * "Too few points / too few correspondences" are results, not errors: they are reported through counts so the
 * caller can mirror the reference's `continue`s (BasicLaserOdometry.cpp:484-488, BasicLaserMapping.cpp:826-828).
 */
#ifndef LOAM_B200_H
#define LOAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct loam_b200_ctx loam_b200_ctx;

typedef enum {
  LOAM_B200_OK = 0,
  LOAM_B200_ERR_ARG = -1,      /* null pointer / negative size / inconsistent ranges */
  LOAM_B200_ERR_CUDA = -2,     /* a CUDA runtime call failed; loam_b200_last_error() has the text */
  LOAM_B200_ERR_NO_DEVICE = -3,/* no CUDA device / device is not sm_100 */
  LOAM_B200_ERR_STATE = -4,    /* call sequence violated (e.g. iterate before set) */
  LOAM_B200_ERR_CAPACITY = -5, /* caller-provided output capacity too small */
  LOAM_B200_ERR_COMM = -6      /* NCCL failure, or a peer of the cube-sharded map did not answer (fused all-reduce) */
} loam_b200_status;

const char* loam_b200_strerror(int status);
/* text of the last CUDA/NCCL failure seen by this context (empty string if none) */
const char* loam_b200_last_error(const loam_b200_ctx* ctx);
/* library version: major*10000 + minor*100 + patch */
int loam_b200_version(void);

/* Create a context on CUDA device `device` (>= 0).  Fails with LOAM_B200_ERR_NO_DEVICE when there is no usable GPU:
 * there is deliberately no CPU fallback. */
int loam_b200_create(loam_b200_ctx** out, int device);
int loam_b200_destroy(loam_b200_ctx* ctx);
/* Scheduling priority of the context's streams relative to other contexts on the same GPU: level 0 = default, 1 = high
 * (the latency-critical Gauss-Newton loops of the odometry stage), -1 = low.  Call right after loam_b200_create, before
 * any work has been enqueued (the streams are re-created). */
int loam_b200_set_priority(loam_b200_ctx* ctx, int level);
/* ---- multi-GPU with the map sharded by cube slabs (csrc/shard.cuh, csrc/peer.inc) --------------------
 * One process per GPU.  Rank r owns the slabs of `slab_cells` 1 m cells along x with (slab index mod world) == r and stores a
 * 2-cell halo around them; it evaluates the scan-to-map queries whose transformed position falls into a cell it owns, and the
 * 32 partial sums of every Gauss-Newton iteration are all-reduced INSIDE the iteration kernel: the folding CTA stores them
 * into every peer's inbox over NVLink (CUDA IPC mapped memory), flags them, and adds the world contributions in rank order,
 * so every rank holds bit-identical normal equations without a host round trip or a separate collective.
 *   loam_b200_peer_export   -> this rank's inbox as a 64-byte CUDA IPC handle (distribute with any all-gather)
 *   loam_b200_peer_connect  <- world x 64 bytes of handles in rank order
 *   loam_b200_peer_connect_local: the same wiring for contexts of ONE process (tests, single-process multi-GPU)
 * After connecting: loam_b200_map_pool_append / seedMap must only be given the points this rank stores
 * (loam_b200_shard_stores), inserted points are filtered by the library; loam_b200_map_iterate returns the all-reduced sums.
 * Every rank must run the same sequence of sweeps and iterations (they do: same sums, same solve). */
int loam_b200_peer_export(loam_b200_ctx* ctx, unsigned char handle_out[64]);
int loam_b200_peer_connect(loam_b200_ctx* ctx, int rank, int world, const unsigned char* handles, int slab_cells);
int loam_b200_peer_connect_local(loam_b200_ctx** ctxs, int world, int slab_cells);
int loam_b200_peer_disconnect(loam_b200_ctx* ctx);
/* host logic of the slab partition: owner of the cell with x index cell_x; does `rank` store the point with coordinate x */
int loam_b200_shard_owner(int cell_x, int slab_cells, int world);
int loam_b200_shard_stores(float x, int rank, int world, int slab_cells);
/* The library records some per-sweep enqueue sequences by stream capture and launches them as one graph.  While a stream
 * records, cudaDeviceSynchronize() from any thread of the process returns an error, so the sequence issued by the context's
 * helper thread (the asynchronous end-of-sweep map update) is only recorded after the caller has opted in here -- i.e. has
 * promised to synchronise through loam_b200_sync / loam_b200_pipeline_sync (or per stream) rather than device-wide while
 * sweeps are in flight.  The streaming pipeline (loam_b200_pipeline_submit) opts in.  LOAM_B200_NO_CAPTURE=1 disables all
 * recording. */
int loam_b200_allow_async_capture(loam_b200_ctx* ctx, int on);
/* Bind the calling host thread to a CUDA device (cudaSetDevice): every host thread other than the one that created a
 * context must call this once before using the context (helper threads of the library do so themselves). */
int loam_b200_bind_thread(int device);
int loam_b200_sync(loam_b200_ctx* ctx);
/* cudaStream_t of the context (as void*) so callers can order their own work / time with events */
void* loam_b200_stream(loam_b200_ctx* ctx);

/* ------------------------------------------------------------------------------------------------------------------
 * Scan registration: replaces BasicScanRegistration::extractFeatures with setScanBuffersFor / setRegionBuffersFor /
 * markAsPicked (BasicScanRegistration.cpp:155-254, 284-386) including the per-ring VoxelGrid of the less-flat points
 * (:246-252).  Parameter names follow RegistrationParams (BasicScanRegistration.h:37-71).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  int nFeatureRegions;              /* 6  */
  int curvatureRegion;              /* 5  */
  int maxCornerSharp;               /* 2  */
  int maxCornerLessSharp;           /* 20 */
  int maxSurfaceFlat;               /* 4  */
  float lessFlatFilterSize;         /* 0.2 */
  float surfaceCurvatureThreshold;  /* 0.1 */
} loam_b200_reg_params;

typedef struct {
  /* picks in the reference's output order (ring-major, region-major, pick order), as indices into the input cloud */
  int32_t* sharp_idx;       int sharp_cap;       int n_sharp;
  int32_t* less_sharp_idx;  int less_sharp_cap;  int n_less_sharp;
  int32_t* flat_idx;        int flat_cap;        int n_flat;
  /* per-point label, n entries: 2 sharp, 1 less sharp, 0 less flat, -1 flat (PointLabel, BasicScanRegistration.h:24-30),
   * 127 = not inside any feature region.  May be NULL. */
  int8_t* label;
  /* voxel-filtered less-flat cloud (ring-major; within a ring ascending voxel index, like pcl::VoxelGrid) */
  float* less_flat_ds;      int less_flat_cap;   int n_less_flat;
} loam_b200_features;

/* pts: n packed points, ring-ordered; ring_start/ring_end: the inclusive IndexRange of each ring exactly as
 * processScanlines builds _scanIndices (BasicScanRegistration.cpp:38-41), so empty rings are (c, c-1) or (0, 0).
 * Limits the reference does not have (a ring is staged in one CTA's shared memory): a ring longer than 5344 points
 * returns LOAM_B200_ERR_CAPACITY; curvatureRegion > 15 or nFeatureRegions > 4095 return LOAM_B200_ERR_ARG; the ring
 * binning front end (loam_b200_reg_bin) takes at most 254 rings.  (HDL-64E: 2.1 k points per ring at 10 Hz; the
 * limits are repeated in INTEGRATION.md.) */
int loam_b200_extract_features(loam_b200_ctx* ctx, const float* pts, int n, const int32_t* ring_start,
                               const int32_t* ring_end, int n_rings, const loam_b200_reg_params* params,
                               loam_b200_features* out);

/* ------------------------------------------------------------------------------------------------------------------
 * k-NN search structure: replaces nanoflann::KdTreeFLANN<PointXYZI>::setInputCloud / nearestKSearch
 * (nanoflann_pcl.h:131-152; nanoflann.hpp:1217-1262, 1354-1412) with a Morton-sorted linear BVH held in HBM.
 * Slots: the reference keeps four trees alive (last corner / last surface in odometry, corner / surface map in
 * mapping, BasicLaserOdometry.h:83-84, BasicLaserMapping.cpp:623-624).
 * ------------------------------------------------------------------------------------------------------------------ */
enum { LOAM_B200_TREE_ODOM_CORNER = 0, LOAM_B200_TREE_ODOM_SURF = 1, LOAM_B200_TREE_MAP_CORNER = 2,
       LOAM_B200_TREE_MAP_SURF = 3, LOAM_B200_NUM_TREES = 4 };

/* upload m points and (re)build the tree of `slot`; m may be 0 */
int loam_b200_tree_build(loam_b200_ctx* ctx, int slot, const float* pts, int m);
int loam_b200_tree_size(loam_b200_ctx* ctx, int slot);
/* exact k nearest neighbours (k <= 8) of nq queries (packed float[4], intensity ignored), ascending squared distance,
 * float L2 accumulated x->y->z like nanoflann's L2_Simple_Adaptor (nanoflann.hpp:372-379).  Only points with
 * d2 < max_d2 are reported (pass INFINITY for the plain search); missing neighbours have idx -1 and d2 = FLT_MAX. */
int loam_b200_tree_knn(loam_b200_ctx* ctx, int slot, const float* queries, int nq, int k, float max_d2,
                       int32_t* idx_out, float* d2_out);

/* ------------------------------------------------------------------------------------------------------------------
 * Scan-to-map Gauss-Newton iteration: replaces the body of the iteration loop of
 * BasicLaserMapping::optimizeTransformTobeMapped up to and including AtA / AtB
 * (BasicLaserMapping.cpp:665-866: pointAssociateToMap, 5-NN, line / plane fit, residual weight, Jacobian rows,
 * normal equations).  The 6x6 solve and pose update stay with the caller (:867-922).
 * ------------------------------------------------------------------------------------------------------------------ */
/* DS corner / surface stacks of this sweep (BasicLaserMapping.cpp:519-527), sensor frame */
int loam_b200_map_set_queries(loam_b200_ctx* ctx, const float* corner, int n_corner, const float* surf, int n_surf);

typedef struct {
  float rot[3];   /* rot_x, rot_y, rot_z of _transformTobeMapped */
  float sin_[3];  /* Angle::sin() of the three, as cached by the host (Angle.h:23-26) */
  float cos_[3];
  float pos[3];
} loam_b200_pose;

typedef struct {
  float AtA[36];      /* row-major 6x6, columns (rot_x, rot_y, rot_z, x, y, z) */
  float AtB[6];
  int n_selected;     /* laserCloudSelNum (BasicLaserMapping.cpp:826) */
  int n_corner_selected;
} loam_b200_normal_eq;

int loam_b200_map_iterate(loam_b200_ctx* ctx, const loam_b200_pose* pose, loam_b200_normal_eq* out);
/* debug / parity variant: additionally returns per query (corner queries first, then surface) the coefficient
 * (s*n, s*d) of _coeffSel and a selected flag; coeff: (n_corner+n_surf) x 4 floats, selected: (n_corner+n_surf) */
int loam_b200_map_iterate_debug(loam_b200_ctx* ctx, const loam_b200_pose* pose, loam_b200_normal_eq* out,
                                float* coeff, int8_t* selected);

/* instrumented variant: same result plus the number of grid-table probes (16 B each) and candidate map points (16 B
 * each) the 5-NN searches of this launch read (feeds the algorithmic-bytes figure of the roofline); slower, never
 * used on the timed path */
int loam_b200_map_iterate_stats(loam_b200_ctx* ctx, const loam_b200_pose* pose, loam_b200_normal_eq* out,
                                unsigned long long* table_probes, unsigned long long* candidate_points);

/* ------------------------------------------------------------------------------------------------------------------
 * Scan-to-scan Gauss-Newton iteration: replaces the body of the iteration loop of BasicLaserOdometry::process up to
 * AtA / AtB (BasicLaserOdometry.cpp:246-557: transformToStart, 1-NN + adjacent-ring search every 5th iteration,
 * point-to-line / point-to-plane residuals, Jacobian rows, normal equations).
 * ------------------------------------------------------------------------------------------------------------------ */
/* last less-sharp / less-flat clouds (already transformed to the sweep end, integer-ring intensities) */
int loam_b200_odom_set_last(loam_b200_ctx* ctx, const float* corner, int n_corner, const float* surf, int n_surf);
/* current sharp / flat feature clouds (intensity = ring + relTime) */
int loam_b200_odom_set_current(loam_b200_ctx* ctx, const float* sharp, int n_sharp, const float* flat, int n_flat);

typedef struct {
  float rot[3];
  float sin_[3];
  float cos_[3];
  float pos[3];
  float inv_scan_period; /* 1.f / _scanPeriod (BasicLaserOdometry.cpp:42) */
  int iter;              /* iterCount: search on iter % 5 == 0, weights on iter >= 5 */
} loam_b200_odom_pose;

int loam_b200_odom_iterate(loam_b200_ctx* ctx, const loam_b200_odom_pose* pose, loam_b200_normal_eq* out);
/* debug: coefficient + selected flag per query (sharp first, then flat) and the correspondence indices
 * (ind: (n_sharp+n_flat) x 3 ints: closest, second, third (-1 when absent)) */
int loam_b200_odom_iterate_debug(loam_b200_ctx* ctx, const loam_b200_odom_pose* pose, loam_b200_normal_eq* out,
                                 float* coeff, int8_t* selected, int32_t* ind);

/* ------------------------------------------------------------------------------------------------------------------
 * Whole Gauss-Newton loops with the pose kept on the device: replace the iteration LOOPS of
 * BasicLaserOdometry::process (BasicLaserOdometry.cpp:246-622: correspondences every 5th iteration, normal equations,
 * `continue` below 10 selected points, colPivHouseholderQr solve, degeneracy projection of the first iteration, pose
 * update, non-finite reset, deltaR / deltaT abort) and of BasicLaserMapping::optimizeTransformTobeMapped
 * (BasicLaserMapping.cpp:646-922, `continue` below 50, eigenvalue threshold 100).  Same arithmetic as the
 * per-iteration entry points + a host solve, except that the sin / cos of the updated angles come from the device
 * (double sincos rounded once instead of the host's float libm).  Queries / last clouds / map as for *_iterate.
 * loam_b200_map_solve returns LOAM_B200_ERR_STATE when a shard or communicator is set (use loam_b200_map_iterate).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
  float rot[3];     /* optimised rot_x, rot_y, rot_z */
  float pos[3];
  int iterations;   /* iterCount + 1 of the last executed iteration (0 when nothing ran) */
  int converged;    /* loop ended (abort thresholds met or iteration cap reached) */
} loam_b200_lm_result;
int loam_b200_odom_solve(loam_b200_ctx* ctx, const float rot[3], const float pos[3], float inv_scan_period,
                         int max_iterations, float delta_t_abort, float delta_r_abort, loam_b200_lm_result* out);
int loam_b200_map_solve(loam_b200_ctx* ctx, const float rot[3], const float pos[3], int max_iterations,
                        float delta_t_abort, float delta_r_abort, loam_b200_lm_result* out);

/* Test hook: the warp-parallel 6 x 6 Gauss-Newton step of the device-resident loops (csrc/lmstep_warp.cuh) on n systems
 * given as 42 floats each (AtA row-major 36, AtB 6) -> x_out (6 per system), degenerate_out (0 / 1 per system).  Must equal
 * loam_b200_host_gn_solve (include/loam_b200_host.h) bit for bit. */
int loam_b200_debug_gn_solve(loam_b200_ctx* ctx, const float* ata_atb, int n, int first_iteration, float eigen_threshold,
                             float* x_out, int* degenerate_out);

/* BasicLaserOdometry::transformToEnd (BasicLaserOdometry.cpp:57-87) without IMU terms, in place on n host points */
int loam_b200_transform_to_end(loam_b200_ctx* ctx, float* pts, int n, const loam_b200_odom_pose* pose);
/* pointAssociateToMap over n host points in place (BasicLaserMapping.cpp:207-219, 235-240) */
int loam_b200_transform_to_map(loam_b200_ctx* ctx, float* pts, int n, const loam_b200_pose* pose);

/* pcl::VoxelGrid<PointXYZI> centroid filter (call sites BasicLaserMapping.cpp:519-527,580-588): writes at most
 * cap points, ascending voxel index; *n_out receives the count. */
int loam_b200_voxel_grid(loam_b200_ctx* ctx, const float* pts, int n, float leaf, float* out, int cap, int* n_out);

/* ------------------------------------------------------------------------------------------------------------------
 * Device-resident stage API.  The Basic* drop-in classes keep their clouds in HBM between calls and only
 * materialise a pcl::PointCloud on the host when an accessor is used; these are the entry points they call.  Clouds
 * live in numbered slots of a context; each class owns a disjoint range of slots.
 * ------------------------------------------------------------------------------------------------------------------ */
enum {
  /* BasicScanRegistration outputs (BasicScanRegistration.h:156-163) */
  LOAM_B200_C_REG_FULL = 0, LOAM_B200_C_REG_SHARP, LOAM_B200_C_REG_LESS_SHARP, LOAM_B200_C_REG_FLAT, LOAM_B200_C_REG_LESS_FLAT,
  /* BasicLaserOdometry inputs / outputs (BasicLaserOdometry.h:22-31) */
  LOAM_B200_C_ODOM_SHARP, LOAM_B200_C_ODOM_LESS_SHARP, LOAM_B200_C_ODOM_FLAT, LOAM_B200_C_ODOM_LESS_FLAT, LOAM_B200_C_ODOM_FULL,
  LOAM_B200_C_ODOM_LAST_CORNER, LOAM_B200_C_ODOM_LAST_SURF,
  /* BasicLaserMapping inputs / internals / outputs (BasicLaserMapping.h:88-90,150-162) */
  LOAM_B200_C_MAP_CORNER_LAST, LOAM_B200_C_MAP_SURF_LAST, LOAM_B200_C_MAP_FULL, LOAM_B200_C_MAP_CORNER_STACK_DS,
  LOAM_B200_C_MAP_SURF_STACK_DS, LOAM_B200_C_MAP_CORNER_FROM_MAP, LOAM_B200_C_MAP_SURF_FROM_MAP, LOAM_B200_C_MAP_SURROUND_DS,
  LOAM_B200_C_MAP_CORNER_POOL, LOAM_B200_C_MAP_SURF_POOL,
  LOAM_B200_NUM_CLOUDS
};

int loam_b200_cloud_upload(loam_b200_ctx* ctx, int slot, const float* pts, int n);
/* d_pts: DEVICE pointer to n packed points (e.g. a torch tensor); device-to-device copy on the context's stream */
int loam_b200_cloud_upload_device(loam_b200_ctx* ctx, int slot, const void* d_pts, int n);
int loam_b200_cloud_download(loam_b200_ctx* ctx, int slot, float* out, int cap, int* n_out);
int loam_b200_cloud_size(loam_b200_ctx* ctx, int slot);
int loam_b200_cloud_swap(loam_b200_ctx* ctx, int slot_a, int slot_b);
/* copy a cloud between slots, possibly of two different contexts on the same device (stream-ordered, no host hop) */
int loam_b200_cloud_copy(loam_b200_ctx* dst_ctx, int dst_slot, loam_b200_ctx* src_ctx, int src_slot);
/* count copies dst[i] <- src[i] with one stream hand-shake per direction */
int loam_b200_cloud_copy_many(loam_b200_ctx* dst_ctx, const int* dst, loam_b200_ctx* src_ctx, const int* src, int count);

/* Ring-binning front end = MultiScanRegistration::process up to the processScanlines call
 * (MultiScanRegistration.cpp:160-238, mapper :44-67): n unordered sensor-frame points (xyz, 3 floats each, arrival order;
 * host memory, or device memory when on_device != 0) -> LOAM_B200_C_REG_FULL holds the ring-ordered cloud
 * (x <- y, y <- z, z <- x, intensity = ring + relTime), ring_sizes[n_rings] the points per ring; rejected points
 * (non-finite, |p|^2 < 1e-4, ring outside [0, n_rings)) are dropped.  Follow with loam_b200_reg_run. */
int loam_b200_reg_bin(loam_b200_ctx* ctx, const float* xyz, int n, int on_device, float lower_bound_deg,
                      float upper_bound_deg, int n_rings, float scan_period, int32_t* ring_sizes, int* n_kept);

/* Scan registration on the cloud in LOAM_B200_C_REG_FULL (already uploaded): fills the REG_SHARP / LESS_SHARP / FLAT /
 * LESS_FLAT slots; counts_out[4] = their sizes.  Index lists and labels of this sweep stay readable until the next run. */
int loam_b200_reg_run(loam_b200_ctx* ctx, const int32_t* ring_start, const int32_t* ring_end, int n_rings,
                      const loam_b200_reg_params* params, int counts_out[4]);
/* which: 1 sharp, 2 less sharp, 3 flat */
int loam_b200_reg_indices(loam_b200_ctx* ctx, int which, int32_t* out, int cap, int* n_out);
int loam_b200_reg_labels(loam_b200_ctx* ctx, int8_t* out, int n);

/* Odometry stage on the ODOM_* slots: prepare = take ODOM_SHARP + ODOM_FLAT as this sweep's queries;
 * rebuild_last = BVHs + ring-offset tables over ODOM_LAST_CORNER / ODOM_LAST_SURF (setInputCloud,
 * BasicLaserOdometry.cpp:203-204,662-663), built asynchronously on side streams;
 * loam_b200_odom_iterate (above) then works on them. */
int loam_b200_odom_prepare(loam_b200_ctx* ctx);
/* The registration -> odometry hand-off in one launch: the five clouds of `src_ctx` (src5 = its sharp, less sharp, flat,
 * less flat, full-resolution slots) go to the ODOM_SHARP .. ODOM_FULL slots of `ctx` like loam_b200_cloud_copy_many, and
 * this sweep's query array is filled at the same time (the next loam_b200_odom_prepare then copies nothing). */
int loam_b200_odom_adopt(loam_b200_ctx* ctx, loam_b200_ctx* src_ctx, const int* src5);
int loam_b200_odom_rebuild_last(loam_b200_ctx* ctx);
/* transformToEnd / pointAssociateToMap in place on a device cloud */
int loam_b200_cloud_transform_to_end(loam_b200_ctx* ctx, int slot, const loam_b200_odom_pose* pose);
/* the same on two clouds with one launch (slot_b < 0: only slot_a) */
int loam_b200_cloud_transform_to_end2(loam_b200_ctx* ctx, int slot_a, int slot_b, const loam_b200_odom_pose* pose);
int loam_b200_cloud_transform_to_map(loam_b200_ctx* ctx, int slot, const loam_b200_pose* pose);

/* Mapping stage.  The surrounding map is one point pool per kind (MAP_CORNER_POOL / MAP_SURF_POOL), kept sorted by 1 m
 * cell across sweeps together with its cell table (csrc/mapstore.cuh); the 21x11x21 cube grid of the reference is
 * implicit (a point's cube follows from its position and the grid centre).  Between loam_b200_map_end_sweep and the next
 * loam_b200_map_begin_sweep the host only knows an upper bound of the pool size; loam_b200_cloud_size / _download of
 * the pool slots synchronise and return the exact contents. */
typedef struct {
  int cen[3];                 /* _laserCloudCenWidth / Height / Depth after rolling (BasicLaserMapping.cpp:311-441) */
  const int32_t* valid_cubes; /* _laserCloudValidInd, cube index i + 21 j + 231 k (BasicLaserMapping.h:126-127) */
  int n_valid;                /* <= 125 */
  float corner_leaf, surf_leaf; /* VoxelGrid leaf sizes of the per-cube filters (BasicLaserMapping.cpp:98-99); >= 0.2 m, smaller
                                 * leaves are refused with LOAM_B200_ERR_ARG (8-bit per-axis voxel index inside a cube) */
} loam_b200_map_window;

/* append points (map frame) to the pool of kind 0 corner / 1 surface (unfiltered until their cube is in view at the end
 * of a sweep, like points the reference holds in a cube cloud; the pool is re-sorted at the next begin_sweep) */
int loam_b200_map_pool_append(loam_b200_ctx* ctx, int kind, const float* pts, int n);
/* stacks: MAP_CORNER_LAST / MAP_SURF_LAST -> to map -> back to sensor (predicted pose) -> VoxelGrid -> *_STACK_DS;
 * the field-of-view cube table is uploaded (the search structure itself persists from the previous sweep); queries set
 * for loam_b200_map_iterate / loam_b200_map_solve.
 * sizes_out[4] = number of map points in the valid cubes (corner, surface: the sizes of the reference's
 * laserCloudCornerFromMap / laserCloudSurfFromMap, BasicLaserMapping.cpp:503-509,628), corner_stack_ds, surf_stack_ds */
int loam_b200_map_begin_sweep(loam_b200_ctx* ctx, const loam_b200_pose* predicted, const loam_b200_map_window* win,
                              int sizes_out[4]);
/* insert the stack points with the optimised pose, voxel-filter every valid cube (BasicLaserMapping.cpp:536-593) and
 * move MAP_FULL into the map frame (:595) */
int loam_b200_map_end_sweep(loam_b200_ctx* ctx, const loam_b200_pose* optimised);
/* Same, but issued by a helper thread of the context: returns at once; the next call on this context (any entry point)
 * first waits for the update to be issued and returns its status if it failed. */
int loam_b200_map_end_sweep_async(loam_b200_ctx* ctx, const loam_b200_pose* optimised);
/* createDownsizedMap (:242-264): VoxelGrid(leaf) over the corner + surface points of the surround cubes -> MAP_SURROUND_DS */
int loam_b200_map_surround(loam_b200_ctx* ctx, const int cen[3], const int32_t* surround_cubes, int n, float leaf);
/* Same on an auxiliary context with its own stream and helper thread: returns at once, ordered behind the pending
 * end-of-sweep update; loam_b200_cloud_size / _download of MAP_SURROUND_DS wait for it. */
int loam_b200_map_surround_async(loam_b200_ctx* ctx, const int cen[3], const int32_t* surround_cubes, int n, float leaf);
/* Debug / test switch: loam_b200_map_begin_sweep additionally materialises the reference's laserCloud*FromMap
 * (BasicLaserMapping.cpp:503-509: the map points of the cubes in view) in the *_FROM_MAP cloud slots.  Off by default:
 * the persistent cell-sorted map never needs that copy, only its size (sizes_out[0..1]). */
int loam_b200_map_debug_from_map(loam_b200_ctx* ctx, int on);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU).  With a shard set, loam_b200_map_iterate evaluates this rank's contiguous slice
 * of the corner and surface queries; with a communicator the 32-float partial normal equations are all-reduced
 * (NCCL, sum, fp32) on the context's stream directly behind the kernel, so every rank returns the same totals and
 * runs the same 6x6 solve.  rank 0 creates the id, the caller distributes it (torch.distributed, MPI, a file, ...).
 * ------------------------------------------------------------------------------------------------------------------ */
int loam_b200_comm_unique_id(unsigned char out[128]);
int loam_b200_comm_init(loam_b200_ctx* ctx, int rank, int world, const unsigned char id[128]);
int loam_b200_comm_destroy(loam_b200_ctx* ctx);
/* slice only, no collective: loam_b200_map_iterate returns PARTIAL sums for the caller to reduce */
int loam_b200_map_set_shard(loam_b200_ctx* ctx, int rank, int world);

/* ------------------------------------------------------------------------------------------------------------------
 * Kernel timing (CUDA events on the context's stream) for bench.py's roofline: accumulated GPU milliseconds and
 * launch counts per kernel family since the last reset.
 * ------------------------------------------------------------------------------------------------------------------ */
enum { LOAM_B200_K_FEATURES = 0, LOAM_B200_K_TREE_BUILD = 1, LOAM_B200_K_KNN = 2, LOAM_B200_K_MAP_ITER = 3,
       LOAM_B200_K_ODOM_ITER = 4, LOAM_B200_K_TRANSFORM = 5, LOAM_B200_K_VOXEL = 6, LOAM_B200_NUM_KERNEL_FAMILIES = 7 };
int loam_b200_profile_enable(loam_b200_ctx* ctx, int on);
int loam_b200_profile_reset(loam_b200_ctx* ctx);
int loam_b200_profile_get(loam_b200_ctx* ctx, int family, double* gpu_ms, long long* launches);
/* total kernel launches issued by this context since creation */
long long loam_b200_launch_count(loam_b200_ctx* ctx);
/* ... by every context of this process */
long long loam_b200_total_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LOAM_B200_H */
