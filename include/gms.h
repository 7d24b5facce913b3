/*
 * gms.h — C-ABI of the grid-map SLAM hot path (libgms.so = CUDA/sm_100a, libgms_ref.so = CPU oracle).
 *
 * The reference (antbern/gridmap-slam-robot, java/GridMapGL, package com.fmsz.gridmapgl.slam) has no
 * FFI of its own: its boundary is the public Java surface of `SLAM` and `GridMap` as used by
 * `GridMapApp`.  Every entry point below names the Java member it replaces
 * (paths relative to java/GridMapGL/src/main/java/com/fmsz/gridmapgl/).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no C++/torch types.
 *   - every call returns GMS_OK (0) or a negative gms_status; nothing throws or aborts.
 *     gms_last_error(h) returns a human-readable message for the last failure on that handle
 *     (h == NULL: the last gms_create failure of this thread).
 *   - the caller owns every host buffer; the library owns all device memory.
 *   - all calls on one handle come from one thread (the reference is single threaded:
 *     GridMapApp.java:217 -> onHandleData -> slam.update on the GL render thread).
 *   - "host" entry points are synchronous w.r.t. the outputs they return.  "*_dev" entry points take
 *     DEVICE pointers, only enqueue work on the handle's stream and return immediately.
 *   - libgms.so has NO CPU fallback: without a usable CUDA device gms_create fails with GMS_ERR_CUDA.
 *
 * Layouts
 *   - maps are row-major, idx = x + y*W (GridMap.java:135), one map per "slot".
 *   - a beam is Observation.Measurement (Observation.java:37-51): localX, localY, distance, wasHit.
 *     beam_xy is interleaved {localX0, localY0, localX1, localY1, ...} (f64, as in Java).
 *   - a pose is Pose.java:21-34: three f32 {x, y, theta}.
 */
#ifndef GMS_H_
#define GMS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMS_ABI_VERSION 2

typedef enum gms_status {
    GMS_OK = 0,
    GMS_ERR_INVALID_ARG = -1,   /* NULL pointer, bad size, particle/kind out of range            */
    GMS_ERR_CUDA = -2,          /* CUDA runtime failure (message in gms_last_error)               */
    GMS_ERR_OOM = -3,           /* host or device allocation failed                               */
    GMS_ERR_STATE = -4,         /* call not valid in this state / mode                            */
    GMS_ERR_UNSUPPORTED = -5    /* feature not available in this build (e.g. *_dev on the oracle) */
} gms_status;

/* map_mode */
#define GMS_MAP_PER_PARTICLE 0  /* the reference's mode: every particle owns a map (SLAM.java:30-38) */
#define GMS_MAP_SHARED 1        /* extension: one map, updated from the strongest pose */

/* resample_mode: how the CDF of SLAM.resample (SLAM.java:137-145) is accumulated */
#define GMS_RESAMPLE_AUTO 0     /* LITERAL when P <= 2048, FIXED above                             */
#define GMS_RESAMPLE_LITERAL 1  /* sequential f64 running sum in particle order: Java's own order  */
#define GMS_RESAMPLE_FIXED 2    /* u64 fixed point (w * 2^60, truncated): associative, so a block-
                                   wide / multi-rank scan gives identical indices                  */

/* update_mode: how the ray cells of one scan are accumulated into the shared map */
#define GMS_UPDATE_ATOMIC 0     /* integer atomics on the counters (default; order-independent => deterministic) */
#define GMS_UPDATE_SORTED 1     /* atomic-free: key sort of the ray cells + run-length accumulation, one writer
                                   per cell (the scatter north_star sketches; kept as a measured alternative) */

/* kinds for gms_get_map */
#define GMS_MAP_LOG 0           /* f64[W*H]: nFree*L_free + nOcc*L_occ   (GridMapData.logData)       */
#define GMS_MAP_LIKELIHOOD 1    /* f64[W*H]: thresholded + blurred field (GridMapData.likelihoodData) */
#define GMS_MAP_FREE_COUNT 2    /* u32[W*H]: number of "+= logOdds(P_FREE)" increments              */
#define GMS_MAP_OCC_COUNT 3     /* u32[W*H]: number of "+= logOdds(P_OCCUPPIED)" increments         */

/* Every constant the reference bakes into its arithmetic; gms_config_default() fills in the
 * reference values.  Fields are only ever appended (struct_size versions the struct). */
typedef struct gms_config {
    uint32_t struct_size;       /* = sizeof(gms_config)                                            */
    int32_t num_particles;      /* SLAM.java:50 (500) — GLOBAL particle count                      */
    float map_width_m;          /* GridMap ctor args, SLAM.java:57: 6.0f, 6.0f, 0.05f, (-3,-3)     */
    float map_height_m;
    float resolution;
    float origin_x;
    float origin_y;
    float sensor_max_range;     /* SensorModel.java:20 (10.0f): uniform / random terms of scoring  */
    double z_hit;               /* GridMap.java:259 (0.9); zRandom = 1 - z_hit                     */
    float hit_tolerance;        /* GridMap.java:223 (2): occupied band is measured +- tol/2 cells  */
    int32_t extra_steps;        /* GridMap.java:210 (2): cells traced past the end point           */
    float p_free;               /* SensorModel.java:23-24: 0.30f, 0.9f (P_PRIOR = 0.5 is structural)*/
    float p_occ;
    double noise_center_base;   /* Odometry.java:63: sd_c = (base + |dCenter| * gain) / 2          */
    double noise_center_gain;
    double noise_theta_base_deg;/* Odometry.java:64: sd_t = base_deg*pi/180 + gain * |dTheta|      */
    double noise_theta_gain;
    double skip_update_deg;     /* SLAM.java:82 (30): no map integration above this |dTheta|       */
    double likelihood_sigma_num;/* GridMap.java:94 (0.05): sigma = sqrt(num / resolution) cells    */
    int32_t map_mode;           /* GMS_MAP_*                                                       */
    int32_t resample_mode;      /* GMS_RESAMPLE_*                                                  */
    int32_t device;             /* CUDA device ordinal (ignored by the oracle)                     */
    int32_t rank;               /* this process' rank / number of ranks sharing the particle set   */
    int32_t nranks;
    int32_t update_mode;        /* GMS_UPDATE_* (shared map only)                                  */
    uint64_t seed;              /* Philox key for device-generated motion noise / resample draws   */
} gms_config;

typedef struct gms_info {
    int32_t abi_version;
    int32_t is_cuda;            /* 1 = libgms (CUDA), 0 = libgms_ref (CPU oracle)                  */
    int32_t grid_w, grid_h;     /* GridMap.java:85: ceil(width / resolution)                       */
    int32_t num_particles;      /* global                                                          */
    int32_t local_begin;        /* particles [local_begin, local_begin + local_count) live here    */
    int32_t local_count;
    int32_t num_slots;          /* maps held by this handle (1 in shared mode)                     */
    int32_t kernel_taps;        /* 2*ceil(3 sigma)+1 (GridMap.java:95)                             */
    int32_t resample_mode;      /* resolved (never AUTO)                                           */
    double kernel[32];          /* Util.generateGaussianKernel (Util.java:428-455)                 */
    double l_free, l_occ;       /* Util.logOdds(double) of p_free / p_occ (Util.java:35-37)        */
    double world_w, world_h;    /* GridMap.java:88                                                 */
} gms_info;

typedef struct gms_handle gms_handle;

/* ---- lifecycle ------------------------------------------------------------------------------ */
int gms_config_default(gms_config* cfg);                    /* SLAM.java:50,57 + constants above   */
int gms_create(const gms_config* cfg, gms_handle** out);    /* new SLAM() SLAM.java:56-62          */
int gms_destroy(gms_handle* h);
const char* gms_last_error(const gms_handle* h);
int gms_get_info(const gms_handle* h, gms_info* info);
int gms_reset(gms_handle* h);                               /* SLAM.reset SLAM.java:65-77          */

/* ---- the SLAM step -------------------------------------------------------------------------- */
/* Largest scan any entry point accepts (the hit-beam table of the scoring kernels is staged in one
 * CTA's shared memory: 12800 * 16 B = 200 KB); more beams -> GMS_ERR_INVALID_ARG, no state change.
 * The reference's sweeps have 360 (ConnectionThread.java) .. 720 beams. */
#define GMS_MAX_BEAMS 12800

/* SLAM.update(Observation, Odometry) SLAM.java:80-131.  `normals` = 2*local_count standard normal
 * draws, particle-major {z_d, z_theta} (the two NormalDistribution.sample() calls of
 * Odometry.apply, Odometry.java:80-81, d first), or NULL to draw them on the device (Philox keyed
 * by cfg.seed, global particle index and step).  Returns Neff (SLAM.java:124) in *neff_out. */
int gms_update(gms_handle* h, const double* beam_xy, const double* beam_dist, const uint8_t* beam_hit,
               int32_t num_beams, double d_center, double d_theta, const double* normals,
               double* neff_out);

/* SLAM.resample() SLAM.java:133-153.  u01 in [0,1) replaces Math.random() (SLAM.java:136);
 * u01 < 0 draws it from the handle's Philox stream.  Like the Java method it returns nothing to wait for:
 * the work is enqueued on the handle's stream and observed through the getters (which synchronise). */
int gms_resample(gms_handle* h, double u01);

/* A4 — GridMap.findBestPoseOptim GridMap.java:348-369 (called per particle between the motion sample and the
 * weight, SLAM.java:97) as an optional CPU hook.  The reference's own optimiser is the identity (its objective
 * is multiplied by Odometry.probabiliyOf == 0, Odometry.java:99-103), so the default is NO hook and nothing of
 * this runs.  With a hook installed every update calls it once, after the motion update and the likelihood
 * refresh and before the scoring, with the poses of the handle's local particles [first, first + count) in a
 * host array it may modify in place; beam_* are host copies of the scan.  The hook may call the GridMap
 * operators of this header on `h` (gms_map_probability_of is the reference's objective) but no step entry
 * point.  A non-zero return aborts the update with GMS_ERR_STATE. */
typedef int (*gms_pose_optimizer_fn)(void* user, gms_handle* h, int32_t first_particle, int32_t count,
                                     float* poses_xyt /* 3*count, in/out */, const double* beam_xy,
                                     const double* beam_dist, const uint8_t* beam_hit, int32_t num_beams,
                                     double d_center, double d_theta);
int gms_set_pose_optimizer(gms_handle* h, gms_pose_optimizer_fn fn /* NULL: identity */, void* user);

int gms_calculate_neff(gms_handle* h, double* neff_out);    /* SLAM.calculateNeff SLAM.java:180-190 */
int gms_get_weighted_pose(gms_handle* h, float pose_xyt[3]);/* SLAM.getWeightedPose SLAM.java:165-178 */
/* SLAM.getStrongestParticle SLAM.java:196-198 (+ public fields weight, pose SLAM.java:31-32): the particle
 * that was strongest at the last update.  pose / weight are its values at that update (Java keeps referencing
 * that Particle object, which a later resample() does not modify).  index = where it lives NOW: its global
 * index after the update; after a resampling the index of its first child, which inherits its map slot
 * (GridMapApp.java:376-393 keeps drawing strongestParticle.m); -1 before any update or if it left no child. */
int gms_get_strongest(gms_handle* h, int32_t* index, float pose_xyt[3], double* weight);
/* Multi-rank handles on the peer exchange hold only their own block of particles (poses; with a shared map also
 * weights, log-weights and parent indices: every rank normalises and resamples its own block).  The getters of
 * per-particle arrays copy the other blocks out of their owners' arrays through the peer mappings, so they are
 * collective in spirit: call them only when EVERY rank has finished the step (gms_sync / gms_read_neff on each rank,
 * then a barrier) and do not let a rank start its next step before the others have read (a barrier after the reads).
 * gms_read_neff, gms_get_strongest and the map getters need no such care. */
int gms_get_poses(gms_handle* h, float* xyt /* 3*P */);     /* getParticles().get(i).pose          */
int gms_get_weights(gms_handle* h, double* w /* P */);      /* getParticles().get(i).weight        */
int gms_get_log_weights(gms_handle* h, double* lw /* P */); /* ln of the un-normalised products of
                                                               the last update (GridMap.java:261-294);
                                                               finite where Java's product underflows */
int gms_get_parents(gms_handle* h, int32_t* parents /* P */);/* index i chosen for each m, SLAM.java:147 */
/* getParticles().get(particle).m.{logData,likelihoodData} GridMap.java:72-74 (particle ignored in
 * shared mode).  bytes must equal W*H*sizeof(element of kind).
 * Per-particle maps hold only the counter pairs in device memory.  Two things the reference does eagerly are
 * evaluated on demand, with results identical to the eager arrays at every observable point:
 *  - likelihoodData: the step scores the scan against the field evaluated at the looked-up cells straight from the
 *    counters; GMS_MAP_LIKELIHOOD (and gms_render_map / gms_map_probability_of) materialise the array Java would
 *    hold at that moment — rebuilt by computeLikelihoodMap before the scan is scored, i.e. from the counters as
 *    they were BEFORE the last update integrated its scan (SLAM.java:93-105);
 *  - integrateObservation of an update: left pending until the next call that reads or writes a map (or the next
 *    update).  If gms_resample / the resampling of gms_update_end_dev comes first, the scan is integrated only into
 *    the maps of the particles it selects as parents; the others are dropped with their maps (SLAM.java:133-153).
 * GMS_DEFER_INTEGRATION=0 (environment) integrates inside every update instead. */
int gms_get_map(gms_handle* h, int32_t particle, int32_t kind, void* dst, size_t bytes);

/* ---- state injection (tests, replay of a saved state) ---------------------------------------- */
int gms_set_poses(gms_handle* h, const float* xyt /* 3*P */);
int gms_set_weights(gms_handle* h, const double* w /* P */);
int gms_set_map_counts(gms_handle* h, int32_t particle, const uint32_t* n_free, const uint32_t* n_occ);

/* ---- GridMap operators on one particle's map ------------------------------------------------- */
/* GridMap.applyMeasurement GridMap.java:194-228 (grid-coordinate floats). */
int gms_map_apply_measurement(gms_handle* h, int32_t particle, float start_x, float start_y,
                              float end_x, float end_y, float measured_distance, int32_t was_hit);
/* GridMap.integrateObservation GridMap.java:173-191. */
int gms_map_integrate_observation(gms_handle* h, int32_t particle, const float pose_xyt[3],
                                  const double* beam_xy, const double* beam_dist,
                                  const uint8_t* beam_hit, int32_t num_beams);
/* GridMap.computeLikelihoodMap GridMap.java:233-250 (whole map). */
int gms_map_compute_likelihood(gms_handle* h, int32_t particle);
/* GridMap.probabilityOf GridMap.java:261-294 against the map's CURRENT likelihood field.
 * *log_prob = ln(product); *prob = the product itself (may underflow to 0). Either may be NULL. */
int gms_map_probability_of(gms_handle* h, int32_t particle, const float pose_xyt[3],
                           const double* beam_xy, const uint8_t* beam_hit, int32_t num_beams,
                           double* log_prob, double* prob);
/* RayIterator.init/hasNext/next RayIterator.java:65-130 for num_rays rays given as
 * {x0,y0,x1,y1} f32 quadruples.  Ray r writes its visited cells {x,y} to cells_xy[2*cap*r ..] and its
 * count to counts[r] (cells beyond cap are counted but not stored). */
int gms_trace_rays(gms_handle* h, const float* rays_x0y0x1y1, int32_t num_rays, int32_t extra_steps,
                   int32_t* cells_xy, int32_t cap, int32_t* counts);
/* Odometry(int,int) Odometry.java:41-55: encoder counts -> dCenter, dTheta (host arithmetic). */
int gms_odometry_from_counts(int32_t left, int32_t right, double* d_center, double* d_theta);

/* ---- device-resident / multi-rank entry points (libgms only) --------------------------------- */
/* policy for gms_step_dev */
#define GMS_RESAMPLE_NEVER 0
#define GMS_RESAMPLE_IF_NEFF_LOW 1  /* GridMapApp.java:185: neff < P/2, decided on the device */
#define GMS_RESAMPLE_ALWAYS 2

/* One SLAM step with every input already on the device: update + (policy) resample, enqueued on
 * the handle's stream without any host synchronisation.  d_normals may be NULL (device Philox);
 * u01 < 0 draws the resampling uniform on the device. */
int gms_step_dev(gms_handle* h, const double* d_beam_xy, const double* d_beam_dist,
                 const uint8_t* d_beam_hit, int32_t num_beams, double d_center, double d_theta,
                 const double* d_normals, int32_t resample_policy, double u01);
int gms_sync(gms_handle* h);
/* Stream-level join (no host synchronisation): everything the last step enqueued on the handle's internal side
 * streams (the shared-map integration runs next to the resampling) is ordered before whatever is enqueued on
 * the handle's stream next — e.g. an event that closes a timed region. */
int gms_join_streams(gms_handle* h);
/* Use an existing cudaStream_t for all work of this handle (NULL = the handle's own stream). */
int gms_set_stream(gms_handle* h, void* cuda_stream);
/* Per-phase device timing (CUDA events on the handle's stream).  Phases: see GMS_PHASE_*. */
#define GMS_PHASE_MOTION 0
#define GMS_PHASE_LIKELIHOOD 1
#define GMS_PHASE_SCORE 2
#define GMS_PHASE_NORMALISE 3
#define GMS_PHASE_MAP_UPDATE 4
#define GMS_PHASE_RESAMPLE 5
#define GMS_PHASE_MAP_COPY 6
#define GMS_PHASE_EXCHANGE 7    /* multi-rank: push of the log-weights (GMS_PULL=0) / import of all-gathered records /
                                   the maps-final round of per-particle maps; the default pull is inside NORMALISE */
#define GMS_PHASE_COUNT 9       /* the last slot collects everything else (getters, resets, rows of §8f) */
int gms_profile_enable(gms_handle* h, int32_t on);
/* ms[GMS_PHASE_COUNT], launches[GMS_PHASE_COUNT]: accumulated since the last gms_profile_reset. */
int gms_profile_read(gms_handle* h, double* ms, int64_t* launches);
int gms_profile_reset(gms_handle* h);
/* Total kernels launched by this handle since creation. */
int gms_launch_count(gms_handle* h, int64_t* launches);

/* Multi-rank split of one update (one process per GPU; the exchange itself is the caller's
 * collective, e.g. ncclAllGather via torch.distributed, on the handle's stream):
 *   gms_update_begin_dev : motion + likelihood + scoring of the LOCAL particles; fills the local
 *                          exchange block.
 *   <all-gather of gms_exchange_local -> gms_exchange_global>
 *   gms_update_end_dev   : normalise over ALL particles, strongest, Neff, map integration,
 *                          (policy) resample — computed redundantly and identically on every rank.
 * Exchange record = 24 bytes per particle {f64 log_weight; f32 x, y, theta; u32 pad}. */
int gms_exchange_buffers(gms_handle* h, void** d_local, size_t* local_bytes, void** d_global,
                         size_t* global_bytes);
int gms_update_begin_dev(gms_handle* h, const double* d_beam_xy, const double* d_beam_dist,
                         const uint8_t* d_beam_hit, int32_t num_beams, double d_center,
                         double d_theta, const double* d_normals);
int gms_update_end_dev(gms_handle* h, int32_t resample_policy, double u01);
int gms_read_neff(gms_handle* h, double* neff_out);  /* sync + read Neff of the last *_dev update */

/* Per-particle maps (GMS_MAP_PER_PARTICLE) across ranks — the reference's Particle(Particle) deep copy
 * (SLAM.java:41-45, GridMap.java:118-124) when parent and child live on different GPUs of one node.
 * Each rank exports IPC handles of its map arenas, the caller all-gathers them (rank-major) and every
 * rank imports the lot; resampling then pulls remote parents' maps over NVLink inside
 * gms_update_end_dev / gms_resample.  The caller must run a barrier across ranks after the resampling
 * call before the next update (old slots are only released then).  Handles hold 2x the slots.
 * The same import also switches the exchange of every multi-rank handle to the PEER path: the scoring of
 * gms_update_begin_dev writes the rank's block of f64 log-weights (8 bytes per particle) into the rank's own
 * exchange buffer, and the normalise kernel of gms_update_end_dev raises a flag per sender on every rank and reads
 * every other rank's block through its peer mapping over NVLink (the caller skips its all-gather between begin
 * and end; GMS_PULL=0 in the environment: each rank stores its block into every rank's receive buffer instead);
 * the resampling reads a remote parent's pose through the peer mapping of that rank's pose array.  Between begin
 * and end the log-weights of the running step are not readable through the getters (they are filed in the
 * handle's arrays by the normalise). */
#define GMS_IPC_HANDLE_BYTES 64
#define GMS_IPC_NUM_HANDLES 18
int gms_ipc_export(gms_handle* h, void* handles /* GMS_IPC_NUM_HANDLES * GMS_IPC_HANDLE_BYTES */);
int gms_ipc_import(gms_handle* h, const void* all_handles /* nranks * the above, rank-major */);

/* ---- rows adjacent to the path --------------------------------------------------- */
/* Scan de-skew + beam-table build, GridMapApp.onHandleData GridMapApp.java:140-175 followed by
 * Measurement(x, y, wasHit, 0) Observation.java:69-76: beam i of n is rotated / shifted back by the
 * fraction -(n-i)/n of the odometry.  out_xy = {localX, localY} pairs, out_dist = sqrt(x^2+y^2). */
int gms_deskew(gms_handle* h, const double* angle, const double* dist, int32_t num_beams, double d_center,
               double d_theta, double* out_xy, double* out_dist);
/* SLAM.update on a RAW sweep (angle, distance, wasHit as received, ConnectionThread.java:73-81): the
 * de-skew runs on the device, fused with the beam upload, then the step of gms_update. */
int gms_update_raw(gms_handle* h, const double* angle, const double* dist, const uint8_t* beam_hit,
                   int32_t num_beams, double d_center, double d_theta, const double* normals, double* neff_out);
/* Map hand-off to the renderer, GridMap.render GridMap.java:371-388 + Util.getColorBitsGrayscale
 * Util.java:106-108 + Color.colorToFloatBits Color.java:62-66: one packed ABGR word per cell
 * (likelihood != 0: the likelihood field, else 1 - p(occupied)), 4 bytes/cell D2H instead of 8-16. */
int gms_render_map(gms_handle* h, int32_t particle, int32_t likelihood, uint32_t* abgr_out /* W*H */);
/* Combined-map fusion, GridMapApp.calculateCombined GridMapApp.java:439-458: per cell
 * logOdds(1 - prod_p (1 - invLogOdds(logData_p))) over all particles, then computeLikelihoodMap of it.
 * Either output may be NULL.  Per-particle maps. */
int gms_combined_map(gms_handle* h, double* log_out /* W*H */, double* likelihood_out /* W*H */);
/* The same fusion when the particles (and their maps) are sharded over ranks: begin leaves the product over the
 * handle's LOCAL particles, prod_p (1 - invLogOdds(logData_p)), in a device buffer of W*H doubles and returns its
 * address; the caller multiplies the ranks' buffers element-wise (one PRODUCT all-reduce, e.g. ncclAllReduce with
 * ncclProd on the handle's stream); end turns the product into the combined log-odds and its likelihood field.
 * gms_combined_map() is begin + end on a single-rank handle. */
int gms_combined_map_begin_dev(gms_handle* h, void** d_product, size_t* bytes);
int gms_combined_map_end(gms_handle* h, double* log_out /* W*H */, double* likelihood_out /* W*H */);

#ifdef __cplusplus
}
#endif
#endif /* GMS_H_ */
