/*
 * nsb.h -- C ABI of the B200-native NICE-SLAM ray-rendering hot path (libnsb.so).
 *
 * The reference (cjpurackal/nice-slam-cpp) has no plugin/FFI layer: its boundary is the C++ class API of
 * include/Renderer.h, include/Mapper.h, include/Tracker.h and the free functions of
 * include/torchlib/utils.h.  Every entry point below names the reference interface it replaces
 * (file:line relative to the reference tree).  The same-named C++ wrapper classes that keep the
 * reference's signatures live in include/nsb/ (Renderer, Mapper, Tracker, NICE) and call only this ABI.
 *
 * Conventions
 *   - plain pointers and sizes, no torch / STL types; every call returns 0 on success, non-zero on error,
 *     nsb_last_error(ctx) gives the message (the reference throws c10::Error instead);
 *   - "host" entry points take HOST buffers and include the host<->device copies (this is what the
 *     reference's callers do: Mapper.cpp:430 uploads the rays inside the call expression);
 *     "_dev" entry points take DEVICE pointers, enqueue on the context's stream and do not synchronise;
 *   - one context per GPU, not thread-safe, no allocation after nsb_create on the hot path;
 *   - there is NO CPU fallback: nsb_create fails if no sm_100 device is present.
 *
 * Layouts
 *   - grids cross the ABI in the reference's layout (1, C, Z, Y, X) fp32 channel-first (main.cpp:39-43);
 *     on the device they are stored channel-last [Z][Y][X][C] so that one voxel corner is one 128-byte line;
 *   - decoders cross the ABI as ONE flat fp32 vector per decoder (PyTorch Linear layout, weight[out][in]):
 *       MLP (middle/fine/color; MLP.cpp:14-46):
 *         B[3][E] | for i<5: W_i[H][K_i], b_i[H] | for i<5: Fc_i[H][C], bc_i[H] | Wo[O][H], bo[O]
 *         E=93, H=32, K={E,H,H,E+H,H} (skip input is cat(e,h)), C=c_dim (fine: 2*c_dim = cat(fine,middle)),
 *         O=1 (middle, fine) or 4 (color)
 *       MLP_no_xyz (coarse; MLP.cpp:104-138): for i<5: W_i[H][K_i], b_i[H] | Wo[1][H], bo[1],
 *         K={C,H,H,C+H,H} (skip input is cat(c,h))
 *   - a camera is either a row-major 4x4 c2w (16 floats) or the 7-vector (qw,qx,qy,qz,tx,ty,tz)
 *     consumed by quad2rotation (utils.h:174-195).
 */
#ifndef NSB_H
#define NSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSB_ABI_VERSION 2

enum { NSB_COARSE = 0, NSB_MIDDLE = 1, NSB_FINE = 2, NSB_COLOR = 3 }; /* grid level / decoder / stage id */
/* decoder MMA precision.  0 (default): fp32-grade -- every operand is split into fp16 hi + lo and a product is three tensor-core
 * instructions (a_lo.b_hi + a_hi.b_lo + a_hi.b_hi, fp32 accumulate, ~2^-22 per product).
 * 1: one fp16 product (~2^-11), fast mode, OUTSIDE the 1e-4 parity target -- never the benchmarked mode. */
enum { NSB_PREC_FP32_GRADE = 0, NSB_PREC_FP16_SINGLE = 1 };
enum { NSB_RAYDIR_REFERENCE = 0, NSB_RAYDIR_PINHOLE = 1 }; /* utils.h:44-47 as written / upstream pinhole */
enum { NSB_DISTNORM_PER_RAY = 0, NSB_DISTNORM_REFERENCE = 1 }; /* upstream intent / utils.h:153 as written */

typedef struct nsb_ctx nsb_ctx;

/* All hyper-parameters the hot path consumes.  Key names follow config/nice_slam.yaml and
 * config/cofusion.yaml; nsb_config_default() fills the reference's values. */
typedef struct nsb_config {
    /* cam.* (cofusion.yaml:23-29; Mapper.cpp:22-27, Tracker.cpp:25-30) */
    int H, W;
    float fx, fy, cx, cy;
    /* scene bound, hard-coded in the reference (Renderer.cpp:15, Mapper.cpp:29, Tracker.cpp:23, main.cpp:33) */
    float bound[3][2];
    /* grid_len.* and model.* (main.cpp:22-30); grid_dim[level] = {Z,Y,X}; 0 = derive from bound/grid_len as main.cpp:34-78 */
    float grid_len[4];
    int coarse_bound_enlarge;
    int grid_dim[4][3];
    int c_dim;                       /* model.c_dim = 32 (only value this build supports) */
    /* rendering.* -- the reference hard-codes them in Renderer::Renderer (Renderer.cpp:5-15) */
    int n_samples, n_surface;        /* 32, 16 */
    int occupancy;                   /* 0: density branch, what raw2outputs_nerf_color computes (utils.h:155); 1: sigmoid(10 raw) */
    int dist_norm;                   /* NSB_DISTNORM_* */
    int raydir;                      /* NSB_RAYDIR_* */
    /* mapping.* (Mapper.cpp:11-19,28-33,351-368,502-503,521-522) */
    int mapping_pixels, mapping_iters, mapping_iters_first, mapping_window_size, keyframe_every;
    float middle_iter_ratio, fine_iter_ratio;
    int second_stage;                /* stage used for middle_ratio < it <= fine_ratio: NSB_MIDDLE (Mapper.cpp:355-356 as written) or NSB_FINE (upstream) */
    float lr_factor, lr_first_factor;
    float stage_lr[4][5];            /* [stage][group: decoders, coarse, middle, fine, color] (nice_slam.yaml:102-126) */
    float mapping_w_color_loss;      /* Mapper.cpp:33 reads tracking.w_color_loss */
    int fix_fine, fix_color, frustum_feature_selection, BA;
    float BA_cam_lr;
    /* tracking.* (Tracker.cpp:14-21; lr / iters are hard-coded at Tracker.cpp:103,107 to 1e-2 / 10) */
    float tracking_lr;
    int tracking_iters, tracking_pixels, ignore_edge_W, ignore_edge_H, handle_dynamic, use_color_in_tracking;
    float w_color_loss;
    /* engine */
    int precision;                   /* NSB_PREC_* */
    int max_rays;                    /* capacity of the per-call ray batch */
    int max_frames;                  /* resident frame slots (keyframes + current) */
    int color_refine;                /* mapping.color_refine (nice_slam.yaml:86): the last frame is mapped once more with the settings of Mapper.cpp:505-513 */
} nsb_config;

/* ---- configuration ------------------------------------------------------------------------------ */
/* Reference defaults: config/nice_slam.yaml + config/cofusion.yaml + the literals of Renderer.cpp:5-15.
 * ONE rule for the two places where the transliteration is degenerate: the default is the upstream formula the
 * reference transliterates -- raydir = NSB_RAYDIR_PINHOLE (utils.h:45 uses the column index for the y direction, so every pixel
 * of a column would cast the same ray) and dist_norm = NSB_DISTNORM_PER_RAY (utils.h:153 passes the dim as the norm order).
 * nsb_config_reference_literal() switches both to the reference's literal arithmetic; the parity tests pin that mode bit-for-bit
 * (rays) / to 1e-4 (render) against the reference's own compiled code. */
void nsb_config_default(nsb_config* cfg);
void nsb_config_reference_literal(nsb_config* cfg);
/* Replaces YAML::LoadFile + the .as<T>() lookups of Mapper.cpp:6-34, Tracker.cpp:5-34, main.cpp:7-30.
 * Either path may be NULL.  Reads the YAML subset those files use (nested maps, scalars, comments). */
int nsb_config_load_yaml(nsb_config* cfg, const char* nice_slam_yaml, const char* dataset_yaml, char* err, int err_len);
/* main.cpp:34-78: grid dims from bound / grid_len with the reference's fp32 arithmetic + truncation. */
void nsb_grid_dims(const nsb_config* cfg, int level, int* Z, int* Y, int* X);
int64_t nsb_decoder_count(int which, int c_dim);

/* ---- context ------------------------------------------------------------------------------------ */
int nsb_create(const nsb_config* cfg, int device, nsb_ctx** out);   /* fails without an sm_100 GPU */
void nsb_destroy(nsb_ctx* ctx);
const char* nsb_last_error(const nsb_ctx* ctx);
int nsb_abi_version(void);
const char* nsb_build_info(void);
int nsb_synchronize(nsb_ctx* ctx);
void* nsb_stream(nsb_ctx* ctx);                                       /* cudaStream_t the context enqueues on */

/* ---- state: grids (main.cpp:33-78 c10::Dict "grid_*"), decoders (NICE.h:9-11), t-tables ------------ */
int nsb_set_grid(nsb_ctx* ctx, int level, const float* host_ncdhw);
int nsb_get_grid(nsb_ctx* ctx, int level, float* host_ncdhw);
int nsb_get_grid_grad(nsb_ctx* ctx, int level, float* host_ncdhw);   /* gradient left by the last backward */
int nsb_set_decoder(nsb_ctx* ctx, int which, const float* host_flat, int64_t n);
int nsb_get_decoder(nsb_ctx* ctx, int which, float* host_flat, int64_t n);
int nsb_get_decoder_grad(nsb_ctx* ctx, int which, float* host_flat, int64_t n);
/* torch::linspace(0,1,32) / (0,1,16) of Renderer.cpp:86,101 are ISA dependent in the last bit, so the
 * tables are inputs; the default is the symmetric fp32 formula. */
int nsb_set_ttables(nsb_ctx* ctx, const float* t_samples, const float* t_surface);
/* Optional frustum voxel mask per level (Z*Y*X bytes, NULL clears): Adam touches masked voxels only
 * (intent of Mapper.cpp:260-290,333-350,448-464). */
int nsb_set_voxel_mask(nsb_ctx* ctx, int level, const uint8_t* host_mask_zyx);
/* Mapper::get_mask_from_c2w (Mapper.h:23, Mapper.cpp:42-130): frustum voxel mask of grid `level` for the depth frame in
 * `slot` seen from c2w16 (NULL = the slot's pose), computed on the GPU.  host_mask_zyx may be NULL; install != 0 makes it
 * the level's Adam mask.  nsb_mapping_begin does this itself for the current frame when frustum_feature_selection is on. */
int nsb_frustum_mask(nsb_ctx* ctx, int slot, const float* c2w16, int level, uint8_t* host_mask_zyx, int install);

/* ---- frames (KeyFrame, Mapper.h:11-15) ------------------------------------------------------------ */
/* Upload one RGB-D frame into resident slot `slot`: depth (H,W), colour (H,W,3), c2w row-major 4x4. */
int nsb_set_frame(nsb_ctx* ctx, int slot, const float* host_depth, const float* host_color, const float* c2w16);
int nsb_set_frame_pose(nsb_ctx* ctx, int slot, const float* c2w16);
/* Asynchronous frame ingest: the H2D copy runs on its own stream while the optimiser keeps iterating; the next call that reads
 * frames waits for it on the device.  Pass buffers from nsb_host_alloc (pinned) for a truly asynchronous copy and keep them alive
 * until nsb_frames_ready returns (host-side wait) or another frame-reading call has been made and synchronised. */
int nsb_set_frame_async(nsb_ctx* ctx, int slot, const float* host_depth, const float* host_color, const float* c2w16);
int nsb_frames_ready(nsb_ctx* ctx);
int nsb_host_alloc(void** p, size_t bytes);
int nsb_host_free(void* p);
/* Checkpoint of the map (nice_slam.yaml mapping.ckpt_freq): the four grids in the reference's (1,C,Z,Y,X) layout and the four
 * flat decoder vectors, in one little-endian file; load validates the grid dimensions / decoder sizes against the context. */
int nsb_save_checkpoint(nsb_ctx* ctx, const char* path);
int nsb_load_checkpoint(nsb_ctx* ctx, const char* path);

/* ---- camera utilities (utils.h:174-231) ------------------------------------------------------------- */
void nsb_quad2rotation(const float* q4, float* R9);                    /* utils.h:174-195 */
void nsb_get_camera_from_tensor(const float* cam7, float* RT12);      /* utils.h:198-210 */
void nsb_get_tensor_from_camera(const float* c2w16, float* cam7);     /* utils.h:212-231, fixed: uses R, emits (w,x,y,z) */

/* ---- ray sampling: get_samples / raySampler (utils.h:13-55,141-146) + inside filter (Mapper.cpp:416-427) ---- */
/* Draws n pixels in the crop [H0,H1)x[W0,W1) of frame `slot`.  idx (host, n int64, flat index into the crop) may
 * be NULL: then indices are std::mt19937(seed)() % crop, i.e. exactly what torch::randint draws on the CPU
 * generator (utils.h:32), the stream continuing across calls until nsb_seed() reseeds it.
 * c2w16 NULL = the slot's pose.  Outputs are host arrays and may be NULL. inside[i] = AABB exit t >= gt_depth. */
int nsb_seed(nsb_ctx* ctx, uint64_t seed);
int nsb_get_samples(nsb_ctx* ctx, int slot, const float* c2w16, int H0, int H1, int W0, int W1, int n,
                    const int64_t* idx, float* rays_o, float* rays_d, float* gt_depth, float* gt_color,
                    uint8_t* inside, int64_t* idx_out);

/* ---- rendering: Renderer::render_batch_ray (Renderer.h:13, Renderer.cpp:44-125) ---------------------- */
/* stage in {NSB_COARSE..NSB_COLOR} replaces the std::string stage.  gt_depth NULL = the no-depth path
 * (Renderer.cpp:54-58: N_surface = 0, near = 0.01).  weights is (n, n_samples + n_surface) or (n, n_samples). */
int nsb_render_batch_ray(nsb_ctx* ctx, int stage, int n, const float* rays_d, const float* rays_o,
                         const float* gt_depth, float* rgb, float* depth, float* var, float* weights);
int nsb_render_batch_ray_dev(nsb_ctx* ctx, int stage, int n, const float* d_rays_d, const float* d_rays_o,
                             const float* d_gt_depth, float* d_rgb, float* d_depth, float* d_var, float* d_weights);
/* Renderer::eval_points (Renderer.h:12, Renderer.cpp:19-42): raw (P,4) for P points.  The stage assembly (NICE.cpp:16-51) and the
 * out-of-bound occupancy 100 (Renderer.cpp:26-36) are applied on the device; host form: one synchronisation at the end. */
int nsb_eval_points(nsb_ctx* ctx, int stage, int P, const float* pts, float* raw);
int nsb_eval_points_dev(nsb_ctx* ctx, int stage, int P, const float* d_pts, float* d_raw);
/* eval_points over a regular nx x ny x nz lattice between lo3 and hi3 (NULL: the scene bound) -- the mesh-extraction query of
 * nice_slam.yaml meshing.resolution.  The points are generated on the device; outputs are host arrays, either may be NULL:
 * raw4 (n,4), occ (n) = channel 3.  Point order q = (j*nx + i)*nz + k <-> (x_i, y_j, z_k) (numpy.meshgrid(x,y,z).ravel()). */
int nsb_eval_lattice(nsb_ctx* ctx, int stage, int nx, int ny, int nz, const float* lo3, const float* hi3, float* raw4, float* occ);
/* Dense render of all H*W pixels of the frame in `slot` seen from c2w16 (NULL: the slot's pose) -- upstream's
 * Renderer::render_img, of which the reference keeps only ray_batch_size (Renderer.cpp:5); equals ONE render_batch_ray over
 * the H*W rays in row-major pixel order.  use_gt_depth != 0: depth-guided with the frame's depth image (48 samples/ray);
 * 0: the no-depth path (32 samples/ray, e.g. the coarse level).  Outputs are host arrays (H*W*3, H*W, H*W), any may be NULL. */
int nsb_render_img(nsb_ctx* ctx, int slot, const float* c2w16, int stage, int use_gt_depth, float* rgb, float* depth, float* var);
/* z_vals of the last render (n, S) -- the "sample placement" intermediate of Renderer.cpp:61-119. */
int nsb_get_last_zvals(nsb_ctx* ctx, int n, int S, float* z_vals);

/* Vector-Jacobian product of render_batch_ray: what loss.backward() (Mapper.cpp:444, Tracker.cpp:84) does
 * for L = sum(g_rgb*rgb + g_depth*depth + g_var*var).  Leaves grid / decoder gradients in the context
 * (nsb_get_grid_grad / nsb_get_decoder_grad) and returns the ray gradients (may be NULL).
 * flags: bit0 grid grads, bit1 colour-decoder weight grads, bit2 ray (pose) grads. */
int nsb_render_vjp(nsb_ctx* ctx, int stage, int n, const float* rays_d, const float* rays_o, const float* gt_depth,
                   const float* g_rgb, const float* g_depth, const float* g_var, int flags,
                   float* d_rays_d, float* d_rays_o);

/* Mapper::keyframe_selection_overlap (Mapper.cpp:132-196): samples `pixels` (reference: 100) pixels of the current frame
 * (slot cur_slot, pose cur_c2w16 or the slot's own when NULL; pixel indices from `idx` or the context's mt19937 stream),
 * places n_samples (16) vertices per ray between 0.8 d and d + 0.5, projects them into each of the n_kf keyframe poses
 * (kf_c2w16: [n_kf][16] row-major 4x4) and returns the indices of the k_overlap keyframes that see the largest fraction
 * (> 0) of them, best first.  percent_out ([n_kf], may be NULL) receives every keyframe's fraction. */
int nsb_keyframe_selection_overlap(nsb_ctx* ctx, int cur_slot, const float* cur_c2w16, int n_kf, const float* kf_c2w16, int k_overlap,
                                   const int64_t* idx, int pixels, int n_samples, int* selected, int* n_selected, float* percent_out);

/* ---- mapping: Mapper::optimize_map inner loop (Mapper.cpp:330-465) ------------------------------------ */
/* begin = the per-call setup of Mapper.cpp:198-330: chooses the frame slots to optimise (optimize_frame),
 * resets Adam (a fresh torch::optim::Adam is built at :330).  lr_factor as Mapper::run passes it. */
int nsb_mapping_begin(nsb_ctx* ctx, int n_frames, const int* slots, int n_iters, float lr_factor);
/* Bundle adjustment (Mapper.cpp:305-329,366-368,382-399,467-489): like nsb_mapping_begin, and the poses of the frames whose
 * bit is set in ba_mask (bit f <-> slots[f]; the reference: every frame of optimize_frame except the oldest keyframe) are
 * optimised together with the map as 7-vectors (quaternion w,x,y,z + translation), lr = mapping.BA_cam_lr in the colour
 * stage and 0 before.  nsb_mapping_end writes the optimised poses back into the frame slots (est_c2w) and optionally
 * returns the 7-vectors ([n_frames][7]); nsb_get_frame_pose reads a slot's current [R|t] (3x4 row-major). */
int nsb_mapping_begin_ba(nsb_ctx* ctx, int n_frames, const int* slots, int n_iters, float lr_factor, uint32_t ba_mask);
/* General form.  flags:
 *   NSB_MAP_COARSE      the coarse mapper's optimize_map (Mapper(ns, cf, coarse_mapper = true); Mapper.cpp:335-338,351-352,450-453):
 *                       stage "coarse" for every iteration, render_batch_ray("coarse"), only grid_coarse is optimised (coarse_lr);
 *   NSB_MAP_FIX_COLOR   the colour decoder stays fixed for this call (color_refine sets fix_color, Mapper.cpp:505-513);
 *   NSB_MAP_NO_FRUSTUM  no frustum feature selection for this call (color_refine: frustum_feature_selection = false);
 *   NSB_MAP_ZERO_RATIOS middle_iter_ratio = fine_iter_ratio = 0 for this call (color_refine, Mapper.cpp:508-509);
 *   NSB_MAP_COLOR_REFINE = the three color_refine settings of Mapper.cpp:505-513 together.
 * The coarse mapper follows upstream's intent: the transliteration renders "color" at Mapper.cpp:430 whatever the stage, which
 * never touches grid_coarse, so its coarse mapper would optimise nothing; here the coarse mapper renders stage "coarse" without
 * depth guidance (upstream: gt_depth = None) and back-propagates the depth loss into grid_coarse. */
enum { NSB_MAP_COARSE = 1, NSB_MAP_FIX_COLOR = 2, NSB_MAP_NO_FRUSTUM = 4, NSB_MAP_ZERO_RATIOS = 8, NSB_MAP_COLOR_REFINE = 2 | 4 | 8 };
int nsb_mapping_begin_ex(nsb_ctx* ctx, int n_frames, const int* slots, int n_iters, float lr_factor, uint32_t ba_mask, int flags);
int nsb_mapping_end(nsb_ctx* ctx, float* cam7s_out);
/* d L / d (q, t) per frame at the last BA iteration, [n_frames][7] (parity checks of the pose-gradient chain). */
int nsb_mapping_cam_grads(nsb_ctx* ctx, float* g7s);
int nsb_get_frame_pose(nsb_ctx* ctx, int slot, float* c2w12);
/* One joint iteration.  idx: host int64 [n_frames * (mapping_pixels / n_frames)] flat pixel indices, or NULL to
 * draw them from the context's mt19937 stream.  loss (host, may be NULL; reading it synchronises). */
int nsb_mapping_iter(nsb_ctx* ctx, int iter, const int64_t* idx, float* loss);
/* Enqueue only (no host sync).  The iteration is ONE cudaGraphLaunch: the kernel sequence of each variant (geometry / colour /
 * bundle adjustment) is captured once, and everything that changes from iteration to iteration (statistics slot, index-pool row,
 * Adam step count, peer-barrier epoch) is read by the kernels from a device-resident iteration state (NSB_GRAPH=0 enqueues the
 * kernels one by one instead).  Losses are later read with nsb_mapping_losses, indexed by STEP: the k-th nsb_mapping_iter* call
 * since nsb_mapping_begin is step k - 1, whatever `iter` it was given (the ring keeps max(4096, n_iters + 1) steps). */
int nsb_mapping_iter_async(nsb_ctx* ctx, int iter, const int64_t* idx);
int nsb_mapping_losses(nsb_ctx* ctx, int first_step, int n, float* losses, int* n_inside);
/* Optional: park the pixel indices of n_iters iterations ([n_iters][n] int64) in device memory; iterations called
 * with idx = NULL then consume the rows in order (wrapping) instead of drawing/uploading (device-resident benchmarking). */
int nsb_mapping_set_index_pool(nsb_ctx* ctx, const int64_t* host_idx, int n_iters, int n);
/* Whole optimize_map: n_iters iterations, indices from the mt19937 stream; losses may be NULL. */
int nsb_optimize_map(nsb_ctx* ctx, int n_frames, const int* slots, int n_iters, float lr_factor, float* losses);
/* Parity aid: with capture on, every iteration keeps a copy of its gradient arena (after loss.backward() of Mapper.cpp:444, before
 * the exchange / optimiser step); the getters return it in the layouts of nsb_get_grid_grad / nsb_get_decoder_grad. */
int nsb_mapping_capture_grads(nsb_ctx* ctx, int on);
int nsb_get_captured_grid_grad(nsb_ctx* ctx, int level, float* host_ncdhw);
int nsb_get_captured_decoder_grad(nsb_ctx* ctx, int which, float* host_flat, int64_t n);

/* ---- tracking: Tracker::run / optimize_cam_in_batch (Tracker.cpp:41-113) ------------------------------ */
int nsb_tracking_begin(nsb_ctx* ctx, int slot, const float* cam7);       /* Tracker.cpp:96-103 */
int nsb_tracking_iter(nsb_ctx* ctx, const int64_t* idx, float* loss, float* cam_grad7);   /* Tracker.cpp:41-89 */
int nsb_tracking_get_camera(nsb_ctx* ctx, float* cam7);

/* ---- multi-GPU: rays sharded over ranks, one fp32 SUM all-reduce of the gradient arena per iteration ---- */
int nsb_comm_unique_id(char* id128);                                       /* ncclGetUniqueId */
int nsb_comm_init(nsb_ctx* ctx, const char* id128, int rank, int world);  /* ncclCommInitRank on the ctx device */
/* Peer-memory mode (NVLink / NVSwitch, one process per GPU): every rank exports the CUDA IPC handles of its gradient arena,
 * parameter arena and flag block (192 bytes), the host gathers them in rank order ([world][192]) and every rank imports them.
 * From then on a mapping iteration replaces "ncclAllReduce(gradients) + Adam" by ONE kernel: each rank sums its slice of all
 * ranks' gradients with P2P loads, applies Adam to that slice and stores the updated parameters into every rank's arena, with
 * two flag barriers in peer memory.  nsb_comm_init is still required first (rank / world; NCCL stays the fallback path).
 * Put a host barrier between the imports and the first nsb_mapping_iter.  all_handles == NULL switches back to the NCCL path. */
int nsb_comm_p2p_export(nsb_ctx* ctx, char* handles192);
int nsb_comm_p2p_import(nsb_ctx* ctx, const char* all_handles, int rank, int world);
/* Multi-GPU ray order (rank-major interleave of the reference's frame-major batch, see RayOrder in ray_kernels.cuh): returns
 * the reference batch element that ray i of the rendered order is, and its frame.  Rank r renders rays [r*n/world, (r+1)*n/world). */
int nsb_ray_order_source(int world, int n_rays, int pix_per_frame, int i, int* frame);
int nsb_comm_rank_world(nsb_ctx* ctx, int* rank, int* world);
/* Timing of the fused exchange kernel, stamped on the device with %globaltimer: out8 = {last wait-for-peers us, last kernel us,
 * mean wait us, mean kernel us, exchanges averaged, mean exchanged range in bytes, NVLink bytes per direction and rank under the
 * (W-1)/W model, GB/s per direction over kernel - wait}; reset != 0 clears the sums.  wait = barrier 1 (the slowest rank's backward),
 * kernel - wait = reduce-scatter + Adam + all-gather + barrier 2.  All ranks must issue the same call sequence: a barrier that
 * waits longer than NSB_P2P_TIMEOUT_MS (default 20000) gives up and the next synchronising mapping call returns an error. */
int nsb_comm_p2p_stats(nsb_ctx* ctx, double* out8, int reset);

/* ---- instrumentation ----------------------------------------------------------------------------------- */
/* Number of kernels this library launched since the last reset (bench.py's gpu_launches). */
int64_t nsb_launch_count(nsb_ctx* ctx, int reset);
/* Device time of the last instrumented kernels in ms (CUDA events on the ctx stream): [0] sample+zvals,
 * [1] decode fwd, [2] composite/loss, [3] decode bwd, [4] wgrad, [5] adam, [6] allreduce. Enable with nsb_set_profiling. */
int nsb_set_profiling(nsb_ctx* ctx, int on);
int nsb_get_kernel_ms(nsb_ctx* ctx, float* ms7);
/* Times the trilinear grid sampling alone (three 32-channel grids, the rays and z values of the last forward): average
 * device ms of `reps` launches.  Algorithmic bytes per launch = n * S * 3 * 8 * 128. */
int nsb_bench_gather(nsb_ctx* ctx, int reps, float* ms_out);
/* Development aid: cycle counters of the tcgen05 forward kernel (filled only by the NSB_TC_TIMING build variant). */
int nsb_debug_counters(nsb_ctx* ctx, unsigned long long* out32);

#ifdef __cplusplus
}
#endif
#endif /* NSB_H */
