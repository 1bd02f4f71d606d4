/*
 * flimo.h — C ABI of libflimo_cuda.so: the B200 (sm_100a) registration hot path of fast_LIMO.
 *
 * The reference (fetty31/fast_LIMO) has no FFI: its boundary for this path is the C++ class
 * surface fast_limo::Mapper / fast_limo::Localizer plus the IKFoM measurement-model hook.
 * Each entry point below names the reference interface it replaces (paths relative to the
 * reference's include/ directory).  INTEGRATION.md shows the few lines a maintainer adds to
 * Mapper.cpp / use-ikfom.cpp / Localizer.cpp to bind them.
 *
 * Conventions
 *   - opaque handle; the caller owns every host buffer, the library owns all device memory;
 *   - every function returns 0 on success or a negative flimo_status; it never throws and never
 *     falls back to a CPU implementation (a missing GPU / CUDA error is an error);
 *   - one caller at a time per handle (the reference calls this path under Localizer::mtx_ikfom,
 *     fast_limo/Modules/Localizer.cpp:326-353); different handles are independent;
 *   - quaternions are (x, y, z, w); matrices are row-major; points are float32 xyz with a caller
 *     supplied stride in BYTES (16 for float4, 32 for fast_limo::Point, fast_limo/Common.hpp:100-113).
 *   - state26 = state_ikfom flattened in declaration order (IKFoM/use-ikfom.hpp:12-21):
 *       pos[3] rot[4] offset_R_L_I[4] offset_T_L_I[3] vel[3] bg[3] ba[3] grav[3]
 *     state14 = its first 14 doubles (all the measurement model reads).
 *     Covariances are 23x23 row-major in the error-state order pos rot offR offT vel bg ba grav(2).
 */
#ifndef FLIMO_H_
#define FLIMO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flimo_ctx* flimo_handle;

typedef enum {
  FLIMO_OK = 0,
  FLIMO_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
  FLIMO_ERR_CUDA = -2,         /* CUDA runtime error (see flimo_last_error) */
  FLIMO_ERR_NO_DEVICE = -3,    /* no usable sm_100 device */
  FLIMO_ERR_STATE = -4,        /* call sequence error (e.g. match before scan_set) */
  FLIMO_ERR_NOMEM = -5
} flimo_status;

/* Mirrors fast_limo::Config::iKFoM::Mapping (fast_limo/Utils/Config.hpp:59-72) and the two
 * iKFoM flags the hot path reads (Config.hpp:74-77).  Set by flimo_cfg_default() to the values of
 * Mapper::Mapper() (fast_limo/Modules/Mapper.cpp:23-32) / src/main.cpp:148-156. */
typedef struct {
  int32_t NUM_MATCH_POINTS;     /* only 5 is compiled in (every shipped YAML uses 5) */
  int32_t MAX_NUM_MATCHES;      /* first-N cap on Jacobian rows (Localizer.cpp:539) */
  int32_t MAX_NUM_PC2MATCH;     /* first-N cap on queried points (Mapper.cpp:63-69) */
  int32_t estimate_extrinsics;  /* Localizer.cpp:569 */
  double MAX_DIST_PLANE;        /* compared with the 5th SQUARED distance (Plane.cpp:45-48) */
  double PLANE_THRESHOLD;       /* Plane.cpp:107-114 */
  /* Octree block (Config.hpp:65-69).  bucket_size is accepted and IGNORED exactly as the
   * reference ignores it (Octree.hpp:178-180 is a self-assignment; effective bucket = 32). */
  int32_t octree_bucket_size;
  int32_t octree_downsampling;
  float octree_min_extent;
  /* Extensions (defaults keep reference behaviour): */
  float knn_cell;               /* side of the device search grid in metres; 0 = choose automatically */
  int32_t sort_scan;            /* 1 = Morton-sort the scan on upload; 0 (default) = the kernel scatters the query order itself */
  float knn_level_ratio;        /* cell growth between index levels; 0 = default (1.5) */
  int32_t knn_tau;              /* a query starts on the finest level whose 3x3x3 block holds >= knn_tau points; 0 = default (8) */
} flimo_cfg;

void flimo_cfg_default(flimo_cfg* cfg);

/* Mapper::getInstance()/set_config (Mapper.cpp:38-45).  `device` is a CUDA ordinal. */
int flimo_create(const flimo_cfg* cfg, int device, flimo_handle* out);
void flimo_destroy(flimo_handle h);
const char* flimo_last_error(flimo_handle h);   /* valid until the next call on h; h may be NULL */
const char* flimo_version(void);

/* ---- map (fast_limo::Mapper) -------------------------------------------------------------- */

/* Mapper::add (Mapper.cpp:88-96) -> Octree::initialize / Octree::update (Octree.hpp:282-432):
 * world-frame points; the first non-empty call builds the map and never down-samples, later calls
 * apply the reference's leaf-split / drop-batch rule.  NaN points are skipped (Octree.hpp:243). */
int flimo_map_add(flimo_handle h, const float* xyz, size_t n, size_t stride_bytes, double stamp);
/* Same, points already in device memory (bench / pipelines that keep the scan on the GPU). */
int flimo_map_add_device(flimo_handle h, const void* d_xyz, size_t n, size_t stride_bytes, double stamp);
/* Mapper::size / exists / last_time (Mapper.cpp:47-57). */
int flimo_map_size(flimo_handle h, size_t* n_points);
int flimo_map_exists(flimo_handle h);
double flimo_map_last_time(flimo_handle h);
/* Copy the map points out (any order); for tests and for saving benchmark maps. */
int flimo_map_get_points(flimo_handle h, float* out_xyz, size_t cap_points, size_t* n_points);

/* ---- scan + one measurement pass ------------------------------------------------------------ */

/* Binds Localizer::pc2match (use-ikfom.cpp:18): body-frame points of the current scan.  Only the
 * first min(n, MAX_NUM_PC2MATCH) points are kept (Mapper.cpp:63-69).  One H2D copy per scan. */
int flimo_scan_set(flimo_handle h, const float* xyz_body, size_t n, size_t stride_bytes);
/* Same for a scan that already lives in device memory.  The array is read IN PLACE by the measurement
 * passes (no copy): it must stay valid and unchanged until the next scan is bound. */
int flimo_scan_set_device(flimo_handle h, const void* d_xyz_body, size_t n, size_t stride_bytes);
/* Optional: starts the host-to-device copy of the NEXT scan on a separate copy stream and returns at
 * once, so that the copy overlaps the registration of the current scan (a LiDAR driver delivers scan
 * k+1 while scan k is being registered).  A later flimo_scan_set with the same (pointer, n, stride)
 * binds the prefetched copy instead of copying again; any other flimo_scan_set simply ignores it.
 * The host buffer must stay unchanged until that flimo_scan_set returns. */
int flimo_scan_prefetch(flimo_handle h, const float* xyz_body, size_t n, size_t stride_bytes);
/* Multi-GPU: restrict this handle to the contiguous slice [begin, end) of the (capped) scan. */
int flimo_scan_shard(flimo_handle h, size_t begin, size_t end);

/* One IKFoM::h_share_model evaluation (use-ikfom.cpp:10-31 = Mapper::match, Mapper.cpp:59-86 +
 * Localizer::calculate_H, Localizer.cpp:537-577) fused with the two products that consume its
 * output in esekf::update_iterated_dyn_share_modified (esekfom.hpp:1723 HTH = H^T H, :1727 H^T h).
 * H itself is never materialised.  Blocking.
 *   HTH[144] row-major 12x12, HTh[12];
 *   n_valid  = accepted matches (Mapper::match result size),
 *   n_rows   = min(n_valid, MAX_NUM_MATCHES) rows that contributed (calculate_H's N),
 *   sum_sq_res = sum of dist^2 over those rows.                                              */
int flimo_match_reduce(flimo_handle h, const double state14[14], double HTH[144], double HTh[12],
                       int64_t* n_valid, int64_t* n_rows, double* sum_sq_res);

/* Asynchronous form for pipelines / multi-GPU: launches on `cuda_stream` (a cudaStream_t; NULL =
 * the handle's own stream, flimo_stream(h) — the legacy default stream cannot be selected) and leaves 96 doubles in DEVICE memory at d_out:
 *   [0..77]  upper triangle of HTH, row by row (i<=j)   [78..89] HTh
 *   [90] n_rows  [91] sum_sq_res  [92] n_valid  [93..95] reserved (0)
 * With FLIMO sharding, summing d_out over ranks (ncclAllReduce, ncclDouble, 96) gives the
 * whole-scan normal equations.  MAX_NUM_MATCHES truncation is NOT applied in this form unless
 * n_valid <= MAX_NUM_MATCHES on every shard (the blocking form handles the general case). */
int flimo_match_reduce_async(flimo_handle h, const double state14[14], double* d_out96, void* cuda_stream);
/* Fused exchange for ranks that are processes of ONE node (alternative to the NCCL all-reduce above):
 * `shared_host_mem` is a host segment mapped by every rank (e.g. POSIX shared memory), at least
 * world * FLIMO_EXCHANGE_BYTES_PER_RANK bytes, zero-initialised by its creator.  The library pins and
 * maps it (cudaHostRegister) so that the LAST CTA of the measurement kernel stores this rank's 96 packed
 * doubles, each tagged with a sequence number, straight into its slot; flimo_match_reduce_exchange then spins until the
 * slots of all ranks carry the same sequence number and sums them in rank order (deterministic).  No
 * D2H copy, no stream synchronise, no collective launch: ~2 us after the slowest rank's kernel ends.
 * Every rank must make the same sequence of flimo_match_reduce_exchange calls. */
#define FLIMO_EXCHANGE_BYTES_PER_RANK 4096
int flimo_exchange_attach(flimo_handle h, void* shared_host_mem, size_t bytes, int rank, int world);
int flimo_match_reduce_exchange(flimo_handle h, const double state14[14], double HTH[144], double HTh[12],
                                int64_t* n_valid, int64_t* n_rows, double* sum_sq_res);

/* flimo_update with every measurement pass summed over ranks through the exchange segment.  The host-segment
 * exchange cannot order accepted matches across shards: when the summed n_valid exceeds MAX_NUM_MATCHES
 * (Localizer.cpp:539) both exchange calls FAIL with FLIMO_ERR_STATE instead of returning normal equations that
 * differ from the single-GPU ones -- flimo_update_peer below applies the first-N rule across shards. */
int flimo_update_exchange(flimo_handle h, double state26[26], double P529[529], int max_iter,
                          const double limit23[23], double R_noise, double D_degeneracy, int* passes_out);

/* ---- multi-GPU over NVLink peer memory (one process per GPU, one node) ------------------------
 * The scan is sharded over the ranks (flimo_scan_shard), the map is replicated.  Every rank owns an INBOX in
 * device memory; flimo_peer_export returns its CUDA IPC handle (64 bytes), the ranks exchange the handles by
 * any means (torch.distributed.all_gather in fast_limo_b200/dist.py) and flimo_peer_attach maps all of them.
 * flimo_update_peer then runs the WHOLE iterated update (esekfom.hpp:1620-1823) inside one kernel per rank:
 * the CTA that completes a pass stores its 96 partial sums into every rank's inbox with peer stores over
 * NVLink, sums the records of all ranks in rank order and runs the filter step on the device -- identical on
 * every rank, no host, no collective launch, no pose exchange.  The MAX_NUM_MATCHES first-N rule
 * (Localizer.cpp:539,547-548) is applied across shards (the accepted-match bits travel the same way), so the
 * result equals the single-GPU one up to the float64 summation order. */
#define FLIMO_PEER_HANDLE_BYTES 64
int flimo_peer_export(flimo_handle h, void* ipc_handle_64);
int flimo_peer_attach(flimo_handle h, int rank, int world, const void* ipc_handles /* world x 64 bytes, rank order */);
int flimo_update_peer(flimo_handle h, double state26[26], double P529[529], int max_iter,
                      const double limit23[23], double R_noise, double D_degeneracy, int* passes_out);

/* Per-pass records of the last update that ran on the device (flimo_update, flimo_update_peer): for each pass
 * 32 doubles = state26 after the pass, n_valid, n_rows, device time of the pass in ns, row limit, 0, 0.
 * Returns the number of passes recorded (at most 16). */
int flimo_update_trace(flimo_handle h, double* out32, size_t cap_passes, size_t* n_passes);

/* Expands the 96-double packed form into HTH/HTh/... on the host. */
void flimo_unpack96(const double packed[96], double HTH[144], double HTh[12], int64_t* n_valid,
                    int64_t* n_rows, double* sum_sq_res);

/* Per-point results of the LAST pass, for Localizer::get_matches (Localizer.cpp:139-141,575-576),
 * RViz markers (ROSutils.hpp:216-252) and the parity tests.  Record layout (16 floats per queried
 * point, original scan order): world xyz[3], plane ABCD[4], dist, valid (0/1), nn_d2[5], pad[2].
 * Re-runs the pass with the per-point output enabled (debug path, not timed). */
int flimo_match_debug(flimo_handle h, const double state14[14], float* out16, size_t cap_points,
                      size_t* n_points);

/* ---- scan preparation: what Localizer::updatePointCloud does before the hot path ---------------- */

/* Config::filters + the sensor flags deskewPointCloud reads (fast_limo/Utils/Config.hpp:43-55,83-91). */
typedef struct {
  int32_t crop_active, dist_active, rate_active, fov_active, voxel_active;
  float cropBoxMin[3], cropBoxMax[3];   /* negative crop box (Localizer.cpp:57-59,268-271) */
  double min_dist;                      /* cast to float once, as Localizer.cpp:274 does */
  int32_t rate_value;
  float fov_angle;                      /* radians, compared with |atan2(y, x)| (Localizer.cpp:866-869) */
  float leafSize;                       /* leafSize[0]: the reference passes it for all three axes (Localizer.cpp:61) */
  int32_t sensor_type;                  /* 0 OUSTER, 1 VELODYNE, 2 HESAI, 3 LIVOX (Localizer.cpp:747-777) */
  int32_t end_of_sweep;
} flimo_prep_cfg;

/* fast_limo::State — the members State::update(t) and get_RT() read (fast_limo/Objects/State.hpp). */
typedef struct {
  double time;
  float q[4];                           /* x y z w */
  float p[3], v[3], w[3], a[3], bg[3], ba[3], g[3];
} flimo_frame;

/* Localizer.cpp:262-302 + the time sort of deskewPointCloud (:744-789) on a raw LiDAR message of
 * 32-byte fast_limo::Point records (Common.hpp:100-113: xyz at 0, intensity at 16, time union at 24).
 * One H2D copy; returns the size of the filtered cloud and the time of its last point
 * (extract_point_time of the last sorted point, :797,:802) — the host needs it to choose the IMU frames. */
int flimo_prep_filter_sort(flimo_handle h, const void* raw_points, size_t n, double sweep_ref_time,
                           const flimo_prep_cfg* cfg, size_t* n_kept, double* t_last);
/* Same for a sensor_msgs/PointCloud2 payload as it arrives on the wire (what src/main.cpp:22-23 feeds to
 * pcl::fromROSMsg): `data` holds n points of `point_step` bytes, the layout names the byte offsets of the
 * fields fast_limo::Point registers (Common.hpp:156-163).  The decode to the 32-byte record runs on the
 * device; offsets < 0 mean "field absent" (left 0, as pcl::fromROSMsg leaves unmatched fields). */
typedef struct {
  int32_t off_x, off_y, off_z;          /* FLOAT32 */
  int32_t off_intensity;                /* FLOAT32, or -1 */
  int32_t off_time;                     /* the per-point time field of the sensor, or -1 */
  int32_t time_datatype;                /* sensor_msgs/PointField datatype of that field: 6 UINT32 ("t"), 7 FLOAT32 ("time"), 8 FLOAT64 ("timestamp") */
} flimo_msg_layout;
int flimo_prep_filter_sort_msg(flimo_handle h, const void* data, size_t n, size_t point_step,
                               const flimo_msg_layout* layout, double sweep_ref_time, const flimo_prep_cfg* cfg,
                               size_t* n_kept, double* t_last);
/* The per-point loop of deskewPointCloud (Localizer.cpp:822-843) on the cloud left by
 * flimo_prep_filter_sort: frames = integrateImu(prev_scan_stamp, scan_stamp) (:805), last_q/last_p =
 * pose of State(_iKFoM.get_x()) (:820), T_lidar2baselink = extr.lidar2baselink_T row-major, offset as
 * computed at :797-801.  If cfg.voxel_active the voxel grid of :313-321 follows.  The result IS
 * pc2match: it stays in device memory and is bound as the current scan (as flimo_scan_set_device). */
int flimo_prep_deskew(flimo_handle h, const flimo_frame* frames, int n_frames, const float last_q[4],
                      const float last_p[3], const float T_lidar2baselink[16], double offset, size_t* n_pc2match);
/* Copies an intermediate cloud out (tests, debug clouds of Localizer.cpp:303-305,850-851):
 * what = 0: indices (uint32) of the filtered + time-sorted points in the raw message,
 *        1: deskewed points in the world frame (xyz1 float4), 2: in the body frame of the last state
 *        (deskewed_Xt2), 3: pc2match (after the voxel grid). */
int flimo_prep_get(flimo_handle h, int what, void* out, size_t cap_items, size_t* n_items);
/* pcl::VoxelGrid with leaf (leaf, leaf, leaf) on n points given as xyz1 float4 (host); centroids in
 * ascending voxel index.  Utility / test entry for the voxel stage alone. */
int flimo_voxel_grid(flimo_handle h, const float* xyz4, size_t n, float leaf, float* out_xyz4, size_t cap_points,
                     size_t* n_out);

/* ---- iterated update (IKFoM) ---------------------------------------------------------------- */

/* esekf::update_iterated_dyn_share_modified (esekfom.hpp:1620-1823) with the measurement pass
 * above as h_dyn_share.  state26 / P529 in: x_ and P_ after prediction; out: updated.
 * max_iter = Config::iKFoM::MAX_NUM_ITERS (passes <= max_iter + 1), limit23 = LIMITS,
 * R_noise / D_degeneracy = the literals at Localizer.cpp:333 (0.001, 5.0).
 * passes_out (may be NULL) receives the number of measurement passes executed. */
int flimo_update(flimo_handle h, double state26[26], double P529[529], int max_iter,
                 const double limit23[23], double R_noise, double D_degeneracy, int* passes_out);

/* The same update as a pass-wise state machine, so a driver can put a collective between the
 * measurement pass and the filter algebra (multi-GPU) or supply its own normal equations:
 *   begin -> { ekf_state -> [measurement pass] -> ekf_step } until *done -> ekf_end            */
int flimo_ekf_begin(flimo_handle h, const double state26[26], const double P529[529], int max_iter,
                    const double limit23[23], double R_noise, double D_degeneracy);
int flimo_ekf_state(flimo_handle h, double state26[26]);
int flimo_ekf_step(flimo_handle h, const double HTH[144], const double HTh[12], int64_t n_rows, int* done);
int flimo_ekf_end(flimo_handle h, double state26[26], double P529[529]);

/* ---- IMU rate: prediction and the propagated states the deskew stage reads (host algebra) ------- */

/* fast_limo::IMUmeas (Common.hpp:126-132) — the members Localizer::propagateImu reads: already moved to
 * the base-link frame and bias-corrected by Localizer::updateIMU (:512-520), which stays with the caller. */
typedef struct {
  double stamp, dt;
  float ang_vel[3], lin_accel[3];
} flimo_imu;

/* Localizer::propagateImu(const IMUmeas&) (Localizer.cpp:583-608): esekf::predict (esekfom.hpp:279-384)
 * with the process model of use-ikfom.cpp:46-91 and Q = diag(cov4[0] x3, cov4[1] x3, cov4[2] x3, cov4[3] x3),
 * cov4 = Config::iKFoM {cov_gyro, cov_acc, cov_bias_gyro, cov_bias_acc} (:588-592); then
 * State(x, imu.stamp, imu.lin_accel, imu.ang_vel) is pushed on the handle's ring of propagated states
 * (propagated_buffer, capacity 2000, :54).  state26 / P529: x_ and P_, updated in place. */
int flimo_ekf_predict(flimo_handle h, double state26[26], double P529[529], const flimo_imu* imu,
                      const double cov4[4]);
/* Localizer::integrateImu(start_time, end_time) (Localizer.cpp:855-871) over that ring with the selection
 * rule of propagatedFromTimeRange (:878-913): frames oldest first, from the last state before start_time up
 * to the first one at or after end_time — what flimo_prep_deskew takes.  *n_frames = 0 with FLIMO_OK is the
 * reference's "not enough propagated states" (e.g. the first scan, prev_scan_stamp = 0); FLIMO_ERR_STATE when
 * the newest state is older than end_time, where the reference blocks until the IMU thread catches up.
 * out may be NULL with cap = 0 to query the count. */
int flimo_propagated_frames(flimo_handle h, double start_time, double end_time, flimo_frame* out, size_t cap,
                            size_t* n_frames);
int flimo_propagated_clear(flimo_handle h);

/* pcl::transformPointCloud(pc2match -> world) of Localizer.cpp:361-374 on the device copy of the bound
 * cloud, using State::get_RT() of state14.  ALL points of the cloud passed to flimo_scan_set* are
 * transformed, in their original order -- the MAX_NUM_PC2MATCH cap only limits what Mapper::match
 * queries (Mapper.cpp:63-69), the reference maps the whole pc2match (Localizer.cpp:361,377). */
int flimo_scan_to_world(flimo_handle h, const double state14[14], float* out_xyz, size_t cap_points,
                        size_t* n_points);

/* Localizer.cpp:361 + :377 without leaving the device: transformPointCloud(pc2match, state.get_RT())
 * followed by Mapper::add(world cloud, stamp) (Mapper.cpp:88-96). */
int flimo_map_add_scan(flimo_handle h, const double state14[14], double stamp);

/* ---- introspection (bench / profiles) ------------------------------------------------------- */
typedef struct {
  uint64_t kernel_launches;     /* CUDA kernels launched by this handle so far */
  uint64_t match_launches;      /* measurement passes executed (one kernel launch each, or one iteration of the persistent kernel) */
  float last_match_ms;          /* device time of the last timed match launch (CUDA events) */
  double match_ms_total;        /* sum of device times of all match launches (events on the launch stream) */
  uint64_t match_timed;         /* number of launches in that sum */
  float knn_cell;               /* finest grid cell in use */
  int32_t grid_nx, grid_ny, grid_nz;
  int32_t n_levels;
  uint64_t table_bytes, map_bytes;   /* device memory of the prefix tables / of the canonical points + super-row entries (all levels, incl. head-room) */
  double persist_ms_total;      /* in-kernel device time (%globaltimer) of the passes run by the persistent kernel */
  uint64_t persist_passes;      /* number of such passes */
  uint64_t index_builds, index_updates;   /* Mapper::add calls that rebuilt the search index / merged the batch into the touched rows */
  uint64_t index_rows_moved;    /* rows that outgrew their segment during those merges and moved to the free tail */
  double exchange_ms_total;     /* device-resident updates: time from "own tiles complete" to "pass sums of all ranks in hand" (collect + NVLink peer exchange + wait for the slowest rank), summed over the persist_passes */
  uint64_t update_stalls;       /* flimo_update calls whose resident kernels stopped answering (watchdog) and that were redone with one launch per pass */
} flimo_stats;
int flimo_get_stats(flimo_handle h, flimo_stats* out);
void* flimo_stream(flimo_handle h);   /* the handle's cudaStream_t */

#ifdef __cplusplus
}
#endif
#endif /* FLIMO_H_ */
