// Host side of libnsb.so: context, device arena, the C ABI of include/nsb.h.
// Plain C++ + CUDA runtime; no libtorch, no CPU fallback (nsb_create fails without an sm_100 device).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>
#include <random>
#include <string>
#include <vector>

#include "../../include/nsb.h"
#include "misc_kernels.cuh"
#include "p2p_kernels.cuh"
#include "params.h"
#include "ray_kernels.cuh"

namespace nsb {
cudaError_t launch_decode_fwd(const DecodeParams& P, int precision, int grid, cudaStream_t st);
cudaError_t launch_decode_bwd_0(const DecodeParams& P, int precision, int grid, cudaStream_t st);
cudaError_t launch_decode_bwd_1(const DecodeParams& P, int precision, int grid, cudaStream_t st);
cudaError_t launch_decode_bwd_2(const DecodeParams& P, int precision, int grid, cudaStream_t st);
cudaError_t launch_decode_bwd_3(const DecodeParams& P, int precision, int grid, cudaStream_t st);
cudaError_t launch_decode_bwd_4(const DecodeParams& P, int precision, int grid, cudaStream_t st);
cudaError_t launch_decode_bwd_5(const DecodeParams& P, int precision, int grid, cudaStream_t st);
// flags: bit0 grid gradients, bit1 colour-decoder weight-gradient stash, bit2 ray (pose) gradients
inline cudaError_t launch_decode_bwd(const DecodeParams& P, int precision, int grid, cudaStream_t st) {
    switch (P.flags & 7) {
        case 1: return launch_decode_bwd_0(P, precision, grid, st);
        case 3: return launch_decode_bwd_1(P, precision, grid, st);
        case 4: return launch_decode_bwd_2(P, precision, grid, st);
        case 5: return launch_decode_bwd_3(P, precision, grid, st);
        case 7: return launch_decode_bwd_4(P, precision, grid, st);
        case 2: return launch_decode_bwd_5(P, precision, grid, st);
        default: return cudaErrorInvalidValue;   // 6 = RAY|WG is not instantiated
    }
}
cudaError_t wgrad_init();
size_t wgrad_fused_img_bytes();
int wgrad_fused_scratch_floats();
cudaError_t wgrad_fused_init();
cudaError_t launch_build_wgimg(const float* flat, uint8_t* img, cudaStream_t st);
cudaError_t launch_wgrad_fused(const DecodeParams& P, const uint8_t* img, const float* flat, float* dflat, float* scratch, int n_sm, cudaStream_t st);
cudaError_t launch_wgrad(const float* stash_buf, const uint8_t* valid, int P, int S, const float* flat, float* dflat, float* mscr, int h_is_u, int precision, int grid, cudaStream_t st);
int decode_fwd_occupancy(int precision);
cudaError_t launch_gather_only(const DecodeParams& P, float* out, int grid, cudaStream_t st);
cudaError_t launch_coarse_bwd(const DecodeParams& P, int precision, int grid, cudaStream_t st);
cudaError_t launch_build_wimg(const float* const flat[4], const float* const comp[4], float* const img_fwd[4], float* const img_bwd[4], float* const img_cmp[4],
                              int mask, int cmp_mask, cudaStream_t st);
size_t wimg_floats(int which);
cudaError_t launch_decode_fwd_tc(const DecodeParams& P, int grid, cudaStream_t st);
cudaError_t launch_decode_fwd_tc16(const DecodeParams& P, int grid, cudaStream_t st);
cudaError_t launch_decode_fwd_t5(const DecodeParams& P, int grid, cudaStream_t st);
size_t t5_img_bytes(int which);
size_t t5b_img_bytes();
cudaError_t launch_build_t5bimg(const float* const flat[4], const float* const comp[4], uint8_t* const img[4], int mask, cudaStream_t st);
cudaError_t launch_decode_bwd_t5(const DecodeParams& P, int grid, cudaStream_t st);
cudaError_t launch_build_t5img(const float* const flat[4], const float* const comp[4], uint8_t* const img[4], int mask, cudaStream_t st);
cudaError_t launch_compose(const float* const flat[4], float* const comp[4], int mask, cudaStream_t st);
int compose_floats(int which);
}  // namespace nsb

using namespace nsb;

// ---- NCCL through dlopen: the library must not pull a second NCCL into a process that already has one ----------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (h) break; }   // reuse the one already mapped (torch's)
        if (!h) for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) return false;
        GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
        CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
        GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
        GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
        GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
        return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
    }
};
static NcclApi g_nccl;
enum { NCCL_FLOAT32 = 7, NCCL_SUM = 0 };

// ---- context ---------------------------------------------------------------------------------------------------
enum { T_SAMPLE = 0, T_FWD, T_COMP, T_BWD, T_WGRAD, T_ADAM, T_COMM, T_N };
constexpr int LOSS_RING_MIN = 4096;   // slots of the statistics / loss ring (grown to n_iters by nsb_mapping_begin)

// One captured joint iteration (Mapper.cpp:331-465) per variant; replayed with a single cudaGraphLaunch.
struct IterGraph { cudaGraphExec_t exec = nullptr; int launches = 0; };

struct nsb_ctx {
    nsb_config cfg;
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    Bound bnd;
    // arena: [grid0 | dec0 | dec1 | grid1 | grid2 | grid3 | dec2 | dec3 | cams | tail]
    size_t arena_n = 0;
    float *param = nullptr, *grad = nullptr, *m = nullptr, *v = nullptr;
    size_t off_grid[4], off_dec[4], off_cam = 0, off_tail = 0, off_train = 0;
    int gdim[4][3];
    size_t nvox[4];
    int64_t dec_n[4];
    uint8_t* vmask[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t* vmask_tmp[4] = {nullptr, nullptr, nullptr, nullptr};
    float *t_samples = nullptr, *t_surface = nullptr;
    // frames
    float *f_depth = nullptr, *f_color = nullptr, *f_pose = nullptr;
    // per-ray buffers
    int cap = 0;
    float *rays_o = nullptr, *rays_d = nullptr, *gt_depth = nullptr, *gt_color = nullptr, *z = nullptr;
    float *raw_rgb = nullptr, *occ[3] = {nullptr, nullptr, nullptr}, *g_raw = nullptr;
    float *o_rgb = nullptr, *o_depth = nullptr, *o_var = nullptr, *o_w = nullptr;
    float *g_rgb = nullptr, *g_depth = nullptr, *g_var = nullptr, *d_rays = nullptr, *absdiff = nullptr;
    uint8_t* valid = nullptr;
    int64_t* idx = nullptr;
    int64_t* idx_pool = nullptr; int pool_iters = 0, pool_n = 0, pool_cursor = 0;   // optional device-resident pixel indices for many iterations
    float* pts = nullptr;
    float* h_stats = nullptr;     // pinned host landing zone of nsb_mapping_losses (H_STATS_STEPS slots): no staged pageable copy on the per-step path
    float* stats = nullptr;       // mapping loop: [ring][4]: max gt depth, n inside, sum 1/|d|, loss -- one slot per step
    int ring = 0;                 // slots of `stats` (>= the n_iters of the current optimize_map: no loss is lost to a wrap)
    float* rstats = nullptr;      // [2][4] scratch of the render / sampling / keyframe entry points (never the mapping ring)
    int* it_state = nullptr;      // device iteration state of the mapping loop (IterRef, common.cuh)
    double* bc1_tab = nullptr; float* bc2s_tab = nullptr; int tab_n = 0;   // Adam bias corrections per step (k_adam reads row state[0])
    std::unordered_map<uint64_t, IterGraph> graphs;   // captured iteration per variant key
    uint64_t graph_sig = 0;       // signature of everything baked into the captured launches; a change drops the cache
    int use_graph = 1;            // NSB_GRAPH=0: enqueue every kernel from the host (also used while profiling)
    bool capturing = false;
    bool capture_grads = false; float* grad_snap = nullptr;   // nsb_mapping_capture_grads: copy of the gradient arena before the optimiser step
    float* median = nullptr; int* count = nullptr;
    float* trk_scratch = nullptr;   // [32] per-iteration tracking scratch (see nsb_tracking_iter)
    bool trk_hook = false; int* trk_count = nullptr;   // set around the tracking forward: the composite compacts |gt - depth|
    float* stash = nullptr; size_t stash_rows = 0;   // weight-gradient stash of the split-K k_wgrad path (wg_stash = 1)
    uint8_t* wg_img = nullptr; float* wg_scratch = nullptr;   // fused weight-gradient kernel: plane image of the colour decoder, M_i scratch
    int wg_stash = 0;
    uint32_t* masks = nullptr;   // relu masks of the last training forward
    int mask_layout = 0, mask_stride = 0;
    float* comp[4] = {nullptr, nullptr, nullptr, nullptr};   // composed weights for the tcgen05 forward
    uint8_t* wimg_t5b[4] = {nullptr, nullptr, nullptr, nullptr};  // images of the tcgen05 backward (decoders 1, 2)
    int bwd_t5 = 0;              // NSB_BWD_T5=1: geometry iterations run the data gradient of decoders 1, 2 on the tcgen05 backward (decode_bwd_t5.cu);
                                 // measured equal to the warp-MMA kernel (both are dominated by the scatter walk), so the older kernel stays default
    uint8_t* wimg_t5[4] = {nullptr, nullptr, nullptr, nullptr};   // images of the tcgen05 forward (NSB_TCGEN05=3), rebuilt with the composed images
    bool pending_join = false;   // a rebuild of the colour decoder's images is in flight on aux_stream (fork_rebuild / join_rebuild)
    int comp_dirty = 0xE;        // bit d: decoder d's composed weights are stale
    float* wimg_fwd[4] = {nullptr, nullptr, nullptr, nullptr};   // pre-split shared-memory images of the decoders (k_build_wimg)
    float* wimg_bwd[4] = {nullptr, nullptr, nullptr, nullptr};
    float* wimg_cmp[4] = {nullptr, nullptr, nullptr, nullptr};
    int wimg_dirty = 0xE;        // bit d: decoder d's plain forward / backward images are stale
    int wimg_cmp_dirty = 0xE;    // bit d: decoder d's composed forward image is stale (rebuilt lazily: a colour decoder that is being
                                 // trained changes every iteration but runs on the plain image while its stash is needed)
    float* wg_mscr = nullptr;    // [5][32][32] scratch of k_wgrad (sums of g_u (x) c), consumed and cleared by k_wgrad_finish
    int* ray_list = nullptr; int* ray_count = nullptr;   // valid-ray compaction for the tcgen05 forward (k_zvals fills, the forward consumes and clears)
    unsigned long long* tile_ctr = nullptr;          // [8] ticket counters of the decoder kernels' tile scheduler: [0..3] forward,
                                                     // [4..7] backward; cleared on the device by the kernel preceding each decoder launch
    int use_tc = 3;              // forward decoder kernel (NSB_TCGEN05 env; fp32-grade precision only): 3 = decode_fwd_t5.cu (default), 0 = warp MMA
    unsigned long long* dbg = nullptr;   // 32 cycle counters (NSB_TC_TIMING builds)
    float* scratch_ncdhw = nullptr; size_t scratch_n = 0;
    int last_n = 0, last_S = 0;
    // host RNG (std::mt19937 == the CPU generator torch::randint uses, utils.h:32)
    std::mt19937 rng{5489u};
    std::vector<int64_t> h_idx;
    // mapping state
    int map_frames = 0, map_slots[MAX_OPT_FRAMES], map_iters = 0, map_step = 0; float map_lr_factor = 1.0f;
    float* cam_grad_last = nullptr;   // [MAX_OPT_FRAMES][8] camera gradients of the last BA iteration (Adam zeroes the arena)
    cudaStream_t upload_stream = nullptr; cudaEvent_t ev_upload = nullptr; bool upload_pending = false;   // nsb_set_frame_async
    bool map_color_touched = false;   // a colour iteration has run since nsb_mapping_begin (see run_adam)
    bool coarse_map = false;          // this optimize_map is the coarse mapper's (Mapper.cpp:335-338,351-352): stage "coarse", grid_coarse only
    bool map_fix_color = false;       // colour decoder fixed for this optimize_map (mapping.fix_color, or color_refine: Mapper.cpp:505-513)
    bool map_no_mask = false;         // frustum feature selection off for this optimize_map (color_refine)
    bool map_zero_ratios = false;     // middle_iter_ratio = fine_iter_ratio = 0 for this optimize_map (color_refine)
    unsigned long long p2p_timeout_ns = 20000000000ull;   // peer-barrier time-out (NSB_P2P_TIMEOUT_MS)
    uint32_t map_ba_mask = 0;      // bundle adjustment: frames (bit f) whose 7-vector pose is optimised with the map (Mapper.cpp:305-329)
    // tracking state
    int trk_slot = 0, trk_step = 0;
    // comm
    ncclComm_t comm = nullptr; int rank = 0, world = 1;
    cudaStream_t comm_stream = nullptr; cudaEvent_t ev_bwd = nullptr, ev_comm = nullptr;   // grid all-reduce overlapped with the wgrad kernel
    cudaStream_t aux_stream = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // colour iterations: the stashing colour decoder runs beside the tcgen05 forward
    int t5_stash = 1;            // colour iterations with a stash: ONE tcgen05 forward launch writes it (NSB_T5_STASH=0: warp-MMA colour decoder beside the tcgen05 one)
    bool stash_is_u = false;     // the stash's H slots hold relu outputs (written by the tcgen05 forward): k_wgrad_finish adds the Fc c + bc part
    int compact_rays = 1;        // tcgen05 forward walks the compacted list of valid rays (NSB_COMPACT_RAYS=0: all rays, filtered rows idle)
    int split_color_sms = 65;    // SMs given to the colour decoder's warp-MMA forward in that split (NSB_SPLIT_COLOR_SMS; 0 = one warp-MMA launch)
    bool ar_request = false, ar_overlapped = false;
    // peer-memory optimiser step (nsb_comm_p2p_import): every rank's gradient / parameter arena and flag block, opened through CUDA IPC
    bool p2p = false; uint32_t* p2p_flags = nullptr;
    float* peer_grad[P2P_MAX_WORLD] = {nullptr}; float* peer_param[P2P_MAX_WORLD] = {nullptr}; uint32_t* peer_flags[P2P_MAX_WORLD] = {nullptr};
    int ar_mode = 1;               // NSB_AR_MODE: 0 = one full all-reduce per iteration; 1 (default) = short prefix in geometry iterations;
                                   // 2 = additionally the grid all-reduce of colour iterations on a second stream under the wgrad kernel
                                   // (measured: fine on 2 GPUs, pathological on 8 -- NCCL's NVLS kernel and k_wgrad fight for SM residency)
    // instrumentation
    int64_t launches = 0; bool profiling = false;
    struct EvRec { int id; cudaEvent_t a, b; };
    std::vector<EvRec> ev_pool; size_t ev_used = 0;   // profiling: one event pair per timed region since nsb_set_profiling(1)
    int occ_blocks[2] = {0, 0};
};

static int fail(nsb_ctx* c, const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf;
    return -1;
}
#define CK(call)                                                                                         \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)

struct Timer {
    nsb_ctx* c; int id; size_t slot;
    Timer(nsb_ctx* c_, int id_) : c(c_), id(id_), slot(0) {
        if (!c->profiling) return;
        if (c->ev_used == c->ev_pool.size()) { nsb_ctx::EvRec r; r.id = id; cudaEventCreate(&r.a); cudaEventCreate(&r.b); c->ev_pool.push_back(r); }
        slot = c->ev_used++;
        c->ev_pool[slot].id = id;
        cudaEventRecord(c->ev_pool[slot].a, c->stream);
    }
    ~Timer() { if (c->profiling) cudaEventRecord(c->ev_pool[slot].b, c->stream); }
};

// Makes the compute stream wait (on the device) for pending asynchronous frame uploads; called by every sampling entry point.
static int wait_uploads(nsb_ctx* ctx) {
    if (ctx->upload_pending) { CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_upload, 0)); ctx->upload_pending = false; }
    return 0;
}
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static size_t pad32(size_t n) { return (n + 31) / 32 * 32; }

// ---- configuration -----------------------------------------------------------------------------------------------
extern "C" void nsb_config_default(nsb_config* c) {
    memset(c, 0, sizeof *c);
    c->H = 480; c->W = 640; c->fx = 360.f; c->fy = 360.f; c->cx = 320.f; c->cy = 240.f;        // cofusion.yaml:23-29
    const float b[3][2] = {{-4.5f, 3.82f}, {-1.5f, 2.02f}, {-3.0f, 2.76f}};                     // Renderer.cpp:15
    memcpy(c->bound, b, sizeof b);
    c->grid_len[0] = 2.f; c->grid_len[1] = 0.32f; c->grid_len[2] = 0.16f; c->grid_len[3] = 0.16f;   // nice_slam.yaml:7-11
    c->coarse_bound_enlarge = 2; c->c_dim = 32;
    c->n_samples = 32; c->n_surface = 16; c->occupancy = 0;                                      // Renderer.cpp:9-10, utils.h:155
    // one rule for the two places where the transliteration is degenerate (SURVEY.md 8-A.3 decision table): the DEFAULT is the
    // upstream formula the reference set out to transliterate (pinhole directions, per-ray L2 norm); nsb_config_reference_literal()
    // switches both to the reference's literal arithmetic, which the parity tests pin bit-for-bit against the reference's own compiled code
    c->dist_norm = NSB_DISTNORM_PER_RAY; c->raydir = NSB_RAYDIR_PINHOLE;
    c->mapping_pixels = 1000; c->mapping_iters = 60; c->mapping_iters_first = 1500;              // cofusion.yaml:20-22
    c->mapping_window_size = 5; c->keyframe_every = 50;
    c->middle_iter_ratio = 0.4f; c->fine_iter_ratio = 0.6f; c->second_stage = NSB_MIDDLE;        // Mapper.cpp:353-356
    c->lr_factor = 1.f; c->lr_first_factor = 5.f;
    const float lr[4][5] = {{0.f, 0.001f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.1f, 0.f, 0.f},
                            {0.f, 0.f, 0.005f, 0.005f, 0.f}, {0.005f, 0.f, 0.005f, 0.005f, 0.005f}};   // nice_slam.yaml:102-126
    memcpy(c->stage_lr, lr, sizeof lr);
    c->mapping_w_color_loss = 0.5f;                                                              // Mapper.cpp:33
    c->fix_fine = 1; c->fix_color = 0; c->frustum_feature_selection = 1; c->BA = 1; c->BA_cam_lr = 0.001f;
    c->tracking_lr = 0.01f; c->tracking_iters = 10;                                              // Tracker.cpp:103,107
    c->tracking_pixels = 200; c->ignore_edge_W = 20; c->ignore_edge_H = 20;
    c->handle_dynamic = 1; c->use_color_in_tracking = 1; c->w_color_loss = 0.5f;
    c->precision = NSB_PREC_FP32_GRADE; c->max_rays = 8192; c->max_frames = 8; c->color_refine = 1;                         // nice_slam.yaml:86
}

// Minimal YAML subset: nested maps by indentation, "key: scalar", comments, quotes.  Flattened to "a.b.c" -> value.
static bool yaml_flat(const char* path, std::map<std::string, std::string>& out, std::string& err) {
    FILE* f = fopen(path, "r");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::vector<std::pair<int, std::string>> stack;
    char line[1024];
    while (fgets(line, sizeof line, f)) {
        std::string s(line);
        bool inq = false; char qc = 0;
        for (size_t i = 0; i < s.size(); ++i) {
            if (inq) { if (s[i] == qc) inq = false; }
            else if (s[i] == '\'' || s[i] == '"') { inq = true; qc = s[i]; }
            else if (s[i] == '#') { s.erase(i); break; }
        }
        while (!s.empty() && (s.back() == '\n' || s.back() == '\r' || s.back() == ' ' || s.back() == '\t')) s.pop_back();
        size_t ind = 0; while (ind < s.size() && s[ind] == ' ') ++ind;
        if (ind >= s.size()) continue;
        const size_t colon = s.find(':', ind);
        if (colon == std::string::npos) continue;
        std::string key = s.substr(ind, colon - ind), val = s.substr(colon + 1);
        while (!val.empty() && val.front() == ' ') val.erase(0, 1);
        if (val.size() >= 2 && (val.front() == '\'' || val.front() == '"')) val = val.substr(1, val.size() - 2);
        while (!stack.empty() && stack.back().first >= (int)ind) stack.pop_back();
        std::string full;
        for (auto& p : stack) full += p.second + ".";
        full += key;
        if (val.empty()) stack.emplace_back((int)ind, key); else out[full] = val;
    }
    fclose(f);
    return true;
}
static bool ybool(const std::string& v) { return v == "True" || v == "true" || v == "1" || v == "yes"; }

extern "C" int nsb_config_load_yaml(nsb_config* c, const char* ns_yaml, const char* ds_yaml, char* errbuf, int errlen) {
    std::string err;
    std::map<std::string, std::string> ns, ds;
    if (ns_yaml && !yaml_flat(ns_yaml, ns, err)) { if (errbuf) snprintf(errbuf, errlen, "%s", err.c_str()); return -1; }
    if (ds_yaml && !yaml_flat(ds_yaml, ds, err)) { if (errbuf) snprintf(errbuf, errlen, "%s", err.c_str()); return -1; }
    auto F = [](std::map<std::string, std::string>& m, const char* k, float& dst) { auto it = m.find(k); if (it != m.end()) dst = strtof(it->second.c_str(), nullptr); };
    auto I = [](std::map<std::string, std::string>& m, const char* k, int& dst) { auto it = m.find(k); if (it != m.end()) dst = (int)strtol(it->second.c_str(), nullptr, 10); };
    auto Bo = [](std::map<std::string, std::string>& m, const char* k, int& dst) { auto it = m.find(k); if (it != m.end()) dst = ybool(it->second) ? 1 : 0; };
    // nice_slam.yaml first, the dataset file overrides it (upstream's config inheritance; the reference reads
    // cam.* and mapping.pixels from cofusion.yaml: Mapper.cpp:22-28)
    for (auto* m : {&ns, &ds}) {
        I(*m, "cam.H", c->H); I(*m, "cam.W", c->W); F(*m, "cam.fx", c->fx); F(*m, "cam.fy", c->fy); F(*m, "cam.cx", c->cx); F(*m, "cam.cy", c->cy);
        F(*m, "grid_len.coarse", c->grid_len[0]); F(*m, "grid_len.middle", c->grid_len[1]); F(*m, "grid_len.fine", c->grid_len[2]); F(*m, "grid_len.color", c->grid_len[3]);
        I(*m, "model.c_dim", c->c_dim); I(*m, "model.coarse_bound_enlarge", c->coarse_bound_enlarge);
        I(*m, "rendering.N_samples", c->n_samples); I(*m, "rendering.N_surface", c->n_surface);
        I(*m, "mapping.pixels", c->mapping_pixels); I(*m, "mapping.iters", c->mapping_iters); I(*m, "mapping.iters_first", c->mapping_iters_first);
        I(*m, "mapping.mapping_window_size", c->mapping_window_size); I(*m, "mapping.keyframe_every", c->keyframe_every);
        F(*m, "mapping.middle_iter_ratio", c->middle_iter_ratio); F(*m, "mapping.fine_iter_ratio", c->fine_iter_ratio);
        F(*m, "mapping.lr_factor", c->lr_factor); F(*m, "mapping.lr_first_factor", c->lr_first_factor);
        Bo(*m, "mapping.fix_fine", c->fix_fine); Bo(*m, "mapping.fix_color", c->fix_color);
        Bo(*m, "mapping.frustum_feature_selection", c->frustum_feature_selection); Bo(*m, "mapping.BA", c->BA);
        F(*m, "mapping.BA_cam_lr", c->BA_cam_lr); Bo(*m, "mapping.color_refine", c->color_refine);
        const char* st[4] = {"coarse", "middle", "fine", "color"};
        const char* gr[5] = {"decoders_lr", "coarse_lr", "middle_lr", "fine_lr", "color_lr"};
        for (int s = 0; s < 4; ++s) for (int g = 0; g < 5; ++g) {
            const std::string k = std::string("mapping.stage.") + st[s] + "." + gr[g];
            F(*m, k.c_str(), c->stage_lr[s][g]);
        }
        F(*m, "tracking.w_color_loss", c->w_color_loss); F(*m, "tracking.w_color_loss", c->mapping_w_color_loss);   // Mapper.cpp:33
        F(*m, "tracking.lr", c->tracking_lr); I(*m, "tracking.iters", c->tracking_iters); I(*m, "tracking.pixels", c->tracking_pixels);
        I(*m, "tracking.ignore_edge_W", c->ignore_edge_W); I(*m, "tracking.ignore_edge_H", c->ignore_edge_H);
        Bo(*m, "tracking.handle_dynamic", c->handle_dynamic); Bo(*m, "tracking.use_color_in_tracking", c->use_color_in_tracking);
    }
    return 0;
}

extern "C" void nsb_config_reference_literal(nsb_config* c) { c->raydir = NSB_RAYDIR_REFERENCE; c->dist_norm = NSB_DISTNORM_REFERENCE; }

extern "C" void nsb_grid_dims(const nsb_config* c, int level, int* Z, int* Y, int* X) {
    if (c->grid_dim[level][0] > 0) { *Z = c->grid_dim[level][0]; *Y = c->grid_dim[level][1]; *X = c->grid_dim[level][2]; return; }
    int d[3];
    for (int a = 0; a < 3; ++a) {   // main.cpp:34,38,48,59,70: fp32 arithmetic, then .item<int>() truncation
        const float len = c->bound[a][1] - c->bound[a][0];
        float v = level == 0 ? len * (float)c->coarse_bound_enlarge / c->grid_len[0] : len / c->grid_len[level];
        d[a] = (int)v;
    }
    *X = d[0]; *Y = d[1]; *Z = d[2];
}

extern "C" int64_t nsb_decoder_count(int which, int c_dim) {
    const int E = EMB, H = HID;
    if (which == 0) { const int K[5] = {c_dim, H, H, c_dim + H, H}; int64_t n = 0; for (int i = 0; i < 5; ++i) n += H * K[i] + H; return n + H + 1; }
    const int C = which == 2 ? 2 * c_dim : c_dim, O = which == 3 ? 4 : 1;
    const int K[5] = {E, H, H, E + H, H};
    int64_t n = 3 * E;
    for (int i = 0; i < 5; ++i) n += H * K[i] + H;
    return n + 5 * (H * C + H) + O * H + O;
}

extern "C" int nsb_abi_version(void) { return NSB_ABI_VERSION; }
extern "C" const char* nsb_build_info(void) {
    return "libnsb sm_100a"
#ifdef NSB_PRECISE_SIN
           " precise-sin"
#endif
        ;
}
extern "C" const char* nsb_last_error(const nsb_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" void* nsb_stream(nsb_ctx* c) { return c->stream; }
extern "C" int nsb_synchronize(nsb_ctx* ctx) { CK(cudaStreamSynchronize(ctx->stream)); return 0; }

// ---- camera utilities (host) ------------------------------------------------------------------------------------
extern "C" void nsb_quad2rotation(const float* q4, float* R9) { quad2rotation(q4, R9); }
extern "C" void nsb_get_camera_from_tensor(const float* cam7, float* RT12) {
    float R[9]; quad2rotation(cam7, R);
    for (int r = 0; r < 3; ++r) { RT12[4 * r] = R[3 * r]; RT12[4 * r + 1] = R[3 * r + 1]; RT12[4 * r + 2] = R[3 * r + 2]; RT12[4 * r + 3] = cam7[4 + r]; }
}
extern "C" void nsb_get_tensor_from_camera(const float* c2w, float* cam7) {
    const double m00 = c2w[0], m01 = c2w[1], m02 = c2w[2], m10 = c2w[4], m11 = c2w[5], m12 = c2w[6], m20 = c2w[8], m21 = c2w[9], m22 = c2w[10];
    double w, x, y, z; const double tr = m00 + m11 + m22;
    if (tr > 0) { double s = std::sqrt(tr + 1.0) * 2; w = s / 4; x = (m21 - m12) / s; y = (m02 - m20) / s; z = (m10 - m01) / s; }
    else if (m00 > m11 && m00 > m22) { double s = std::sqrt(1.0 + m00 - m11 - m22) * 2; w = (m21 - m12) / s; x = s / 4; y = (m01 + m10) / s; z = (m02 + m20) / s; }
    else if (m11 > m22) { double s = std::sqrt(1.0 + m11 - m00 - m22) * 2; w = (m02 - m20) / s; x = (m01 + m10) / s; y = s / 4; z = (m12 + m21) / s; }
    else { double s = std::sqrt(1.0 + m22 - m00 - m11) * 2; w = (m10 - m01) / s; x = (m02 + m20) / s; y = (m12 + m21) / s; z = s / 4; }
    cam7[0] = (float)w; cam7[1] = (float)x; cam7[2] = (float)y; cam7[3] = (float)z; cam7[4] = c2w[3]; cam7[5] = c2w[7]; cam7[6] = c2w[11];
}

// ---- create / destroy ------------------------------------------------------------------------------------------
template <typename T> static cudaError_t dalloc(T** p, size_t n) {
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e == cudaSuccess && getenv("NSB_DEBUG_FILL")) e = cudaMemset(*p, atoi(getenv("NSB_DEBUG_FILL")), bytes);   // debug: every buffer starts from a known byte pattern
    return e;
}

static void drop_graphs(nsb_ctx* ctx) {
    for (auto& kv : ctx->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    ctx->graphs.clear();
}

// Statistics ring and the Adam bias-correction tables, sized for `slots` steps (libtorch: bias_correction = 1 - beta^step in double).
static int ensure_ring(nsb_ctx* ctx, int slots) {
    if (ctx->ring >= slots) return 0;
    if (ctx->stream) CK(cudaStreamSynchronize(ctx->stream));
    drop_graphs(ctx);                                   // the ring pointers are baked into the captured launches
    if (ctx->stats) cudaFree(ctx->stats);
    if (ctx->bc1_tab) cudaFree(ctx->bc1_tab);
    if (ctx->bc2s_tab) cudaFree(ctx->bc2s_tab);
    ctx->stats = nullptr; ctx->bc1_tab = nullptr; ctx->bc2s_tab = nullptr; ctx->ring = 0;
    CK(dalloc(&ctx->stats, 4 * (size_t)slots)); CK(dalloc(&ctx->bc1_tab, slots)); CK(dalloc(&ctx->bc2s_tab, slots));
    std::vector<double> b1(slots); std::vector<float> b2(slots);
    for (int t = 1; t <= slots; ++t) { b1[t - 1] = 1.0 - std::pow(0.9, (double)t); b2[t - 1] = (float)std::sqrt(1.0 - std::pow(0.999, (double)t)); }
    CK(cudaMemcpy(ctx->bc1_tab, b1.data(), slots * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->bc2s_tab, b2.data(), slots * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemset(ctx->stats, 0, 4 * (size_t)slots * 4));
    ctx->ring = slots; ctx->tab_n = slots;
    return 0;
}

extern "C" int nsb_create(const nsb_config* cfg, int device, nsb_ctx** out) {
    *out = nullptr;
    nsb_ctx* ctx = new nsb_ctx();
    *out = ctx;   // returned even on failure so that nsb_last_error works; caller destroys it
    ctx->cfg = *cfg; ctx->device = device;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(ctx, "no CUDA device: libnsb has no CPU fallback");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(ctx, "device %s is sm_%d%d; libnsb is built for sm_100a only", prop.name, prop.major, prop.minor);
    ctx->n_sm = prop.multiProcessorCount;
    if (cfg->c_dim != CDIM) return fail(ctx, "c_dim %d unsupported (this build: 32)", cfg->c_dim);
    if (cfg->n_samples != 32 || (cfg->n_surface != 16 && cfg->n_surface != 0)) return fail(ctx, "n_samples/n_surface %d/%d unsupported (32 / 16|0)", cfg->n_samples, cfg->n_surface);
    if (cfg->max_frames < 1 || cfg->max_rays < 16) return fail(ctx, "max_frames / max_rays too small");
    if (!cfg->fix_fine) return fail(ctx, "mapping.fix_fine = False is not supported: this build has no weight-gradient kernel for the fine decoder "
                                         "(the reference adds fine_decoder.parameters() to the optimiser at Mapper.cpp:292-296); refusing instead of silently not training it");
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&ctx->ev_upload, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_bwd, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    { const char* e = getenv("NSB_SPLIT_COLOR_SMS"); if (e) ctx->split_color_sms = atoi(e); }
    { const char* e = getenv("NSB_COMPACT_RAYS"); if (e) ctx->compact_rays = atoi(e); }
    { const char* e = getenv("NSB_T5_STASH"); if (e) ctx->t5_stash = atoi(e); }
    for (int a = 0; a < 3; ++a) { ctx->bnd.lo[a] = cfg->bound[a][0]; ctx->bnd.hi[a] = cfg->bound[a][1]; ctx->bnd.len[a] = cfg->bound[a][1] - cfg->bound[a][0]; ctx->bnd.inv_len[a] = 1.0f / ctx->bnd.len[a]; }
    size_t off = 0;
    auto seg = [&](size_t n) { size_t o = off; off += pad32(n); return o; };
    for (int l = 0; l < 4; ++l) {
        nsb_grid_dims(cfg, l, &ctx->gdim[l][0], &ctx->gdim[l][1], &ctx->gdim[l][2]);
        ctx->nvox[l] = (size_t)ctx->gdim[l][0] * ctx->gdim[l][1] * ctx->gdim[l][2];
        if (ctx->nvox[l] == 0) return fail(ctx, "grid level %d has zero voxels", l);
        ctx->dec_n[l] = nsb_decoder_count(l, cfg->c_dim);
    }
    ctx->off_grid[0] = seg(ctx->nvox[0] * CDIM);
    ctx->off_dec[0] = seg(ctx->dec_n[0]); ctx->off_dec[1] = seg(ctx->dec_n[1]);
    ctx->off_train = off;
    ctx->off_tail = seg(32);          // the loss scalar rides at the HEAD of the all-reduced range, so that a geometry-stage
                                      // iteration (no colour / decoder / camera gradients) reduces one short contiguous prefix
    for (int l = 1; l < 4; ++l) ctx->off_grid[l] = seg(ctx->nvox[l] * CDIM);
    ctx->off_dec[2] = seg(ctx->dec_n[2]); ctx->off_dec[3] = seg(ctx->dec_n[3]);
    ctx->off_cam = seg(8 * (size_t)cfg->max_frames);
    ctx->arena_n = off;
    CK(dalloc(&ctx->param, off)); CK(dalloc(&ctx->grad, off)); CK(dalloc(&ctx->m, off)); CK(dalloc(&ctx->v, off));
    CK(cudaMemsetAsync(ctx->param, 0, off * 4, ctx->stream)); CK(cudaMemsetAsync(ctx->grad, 0, off * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->m, 0, off * 4, ctx->stream)); CK(cudaMemsetAsync(ctx->v, 0, off * 4, ctx->stream));
    CK(dalloc(&ctx->t_samples, 32)); CK(dalloc(&ctx->t_surface, 16));
    {
        float t32[32], t16[16];   // torch::linspace symmetric formula
        for (int n : {32, 16}) {
            float* t = n == 32 ? t32 : t16; const float step = 1.0f / (float)(n - 1);
            for (int i = 0; i < n; ++i) t[i] = i < n / 2 ? (float)i * step : 1.0f - (float)(n - 1 - i) * step;
        }
        CK(cudaMemcpyAsync(ctx->t_samples, t32, sizeof t32, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->t_surface, t16, sizeof t16, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    const size_t hw = (size_t)cfg->H * cfg->W;
    CK(dalloc(&ctx->f_depth, hw * cfg->max_frames)); CK(dalloc(&ctx->f_color, 3 * hw * cfg->max_frames)); CK(dalloc(&ctx->f_pose, 12 * (size_t)cfg->max_frames));
    const int cap = ctx->cap = (cfg->max_rays + 15) / 16 * 16;
    const int S = cfg->n_samples + cfg->n_surface;
    const size_t PS = (size_t)cap * S;
    CK(dalloc(&ctx->rays_o, 3 * (size_t)cap)); CK(dalloc(&ctx->rays_d, 3 * (size_t)cap)); CK(dalloc(&ctx->gt_depth, cap)); CK(dalloc(&ctx->gt_color, 3 * (size_t)cap));
    CK(dalloc(&ctx->z, PS)); CK(dalloc(&ctx->raw_rgb, 4 * PS)); CK(dalloc(&ctx->g_raw, 4 * PS));
    for (int k = 0; k < 3; ++k) CK(dalloc(&ctx->occ[k], PS));
    CK(dalloc(&ctx->o_rgb, 3 * (size_t)cap)); CK(dalloc(&ctx->o_depth, cap)); CK(dalloc(&ctx->o_var, cap)); CK(dalloc(&ctx->o_w, PS));
    CK(dalloc(&ctx->g_rgb, 3 * (size_t)cap)); CK(dalloc(&ctx->g_depth, cap)); CK(dalloc(&ctx->g_var, cap)); CK(dalloc(&ctx->d_rays, 6 * (size_t)cap));
    CK(dalloc(&ctx->absdiff, cap)); CK(dalloc(&ctx->valid, cap)); CK(dalloc(&ctx->idx, cap)); CK(dalloc(&ctx->pts, 3 * PS));
    CK(dalloc(&ctx->masks, 3 * (PS / TILE) * 96));
    for (int d = 1; d < 4; ++d) CK(dalloc(&ctx->comp[d], (size_t)compose_floats(d)));
    for (int d = 1; d < 4; ++d) { CK(dalloc(&ctx->wimg_fwd[d], wimg_floats(d))); CK(dalloc(&ctx->wimg_bwd[d], wimg_floats(d))); CK(dalloc(&ctx->wimg_cmp[d], wimg_floats(d))); CK(dalloc(&ctx->wimg_t5[d], t5_img_bytes(d))); if (d < 3) CK(dalloc(&ctx->wimg_t5b[d], t5b_img_bytes())); }
    { const char* e = getenv("NSB_BWD_T5"); if (e) ctx->bwd_t5 = atoi(e); }
    { const char* e = getenv("NSB_TCGEN05"); ctx->use_tc = e ? atoi(e) : 3; }   // forward decoders: 3 = tcgen05, operands in tensor memory (default); 0 = warp MMA; 1, 2 = earlier tcgen05 generations
    { const char* e = getenv("NSB_AR_MODE"); ctx->ar_mode = e ? atoi(e) : 1; }
    CK(dalloc(&ctx->dbg, 32)); CK(cudaMemsetAsync(ctx->dbg, 0, 32 * 8, ctx->stream));
    CK(dalloc(&ctx->wg_mscr, 5 * 1024)); CK(cudaMemsetAsync(ctx->wg_mscr, 0, 5 * 1024 * 4, ctx->stream));
    CK(dalloc(&ctx->ray_list, cap)); CK(dalloc(&ctx->ray_count, 4)); CK(cudaMemsetAsync(ctx->ray_count, 0, 16, ctx->stream));
    CK(dalloc(&ctx->tile_ctr, 8)); CK(cudaMemsetAsync(ctx->tile_ctr, 0, 8 * sizeof(unsigned long long), ctx->stream));
    CK(dalloc(&ctx->it_state, 8)); CK(cudaMemsetAsync(ctx->it_state, 0, 8 * sizeof(int), ctx->stream));
    CK(dalloc(&ctx->rstats, 8)); CK(cudaMemsetAsync(ctx->rstats, 0, 8 * 4, ctx->stream));
    { const char* e = getenv("NSB_GRAPH"); ctx->use_graph = e ? atoi(e) : 1; }
    { const char* e = getenv("NSB_P2P_TIMEOUT_MS"); if (e) ctx->p2p_timeout_ns = 1000000ull * strtoull(e, nullptr, 10); }
    CK(wgrad_init()); CK(wgrad_fused_init());
    CK(cudaMalloc((void**)&ctx->wg_img, wgrad_fused_img_bytes()));
    CK(dalloc(&ctx->wg_scratch, (size_t)wgrad_fused_scratch_floats())); CK(cudaMemsetAsync(ctx->wg_scratch, 0, wgrad_fused_scratch_floats() * 4, ctx->stream));
    // colour-decoder weight gradient: 1 (default) = activation stash + split-K k_wgrad (fastest measured: 0.19 ms + ~0.07 ms of stash
    // writes per colour iteration at 5000 rays, but 1.2 GB of HBM traffic); 0 = k_wgrad_fused, no stash (13 MB of HBM traffic, 0.43 ms:
    // the recomputation costs more issue slots than the traffic it removes costs bandwidth -- DESIGN.md section 4)
    { const char* e = getenv("NSB_WGRAD_STASH"); ctx->wg_stash = e ? atoi(e) : 1; }
    CK(dalloc(&ctx->p2p_flags, 32)); CK(cudaMemsetAsync(ctx->p2p_flags, 0, 32 * 4, ctx->stream));
    CK(dalloc(&ctx->cam_grad_last, 8 * MAX_OPT_FRAMES)); CK(cudaMemsetAsync(ctx->cam_grad_last, 0, 8 * MAX_OPT_FRAMES * 4, ctx->stream));
    if (ensure_ring(ctx, LOSS_RING_MIN)) return -1;
    CK(dalloc(&ctx->median, 4)); CK(dalloc(&ctx->count, 4));
    CK(dalloc(&ctx->trk_scratch, 32)); CK(cudaMemsetAsync(ctx->trk_scratch, 0, 32 * 4, ctx->stream));
    ctx->occ_blocks[0] = decode_fwd_occupancy(0); ctx->occ_blocks[1] = decode_fwd_occupancy(1);
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" void nsb_destroy(nsb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->p2p) for (int w = 0; w < c->world; ++w) if (w != c->rank) { cudaIpcCloseMemHandle(c->peer_grad[w]); cudaIpcCloseMemHandle(c->peer_param[w]); cudaIpcCloseMemHandle(c->peer_flags[w]); }
    if (c->p2p_flags) cudaFree(c->p2p_flags);
    drop_graphs(c);
    void* ptrs[] = {c->wg_img, c->wg_scratch, c->it_state, c->rstats, c->bc1_tab, c->bc2s_tab, c->grad_snap, c->param, c->grad, c->m, c->v, c->t_samples, c->t_surface, c->f_depth, c->f_color, c->f_pose, c->rays_o, c->rays_d, c->gt_depth,
                    c->gt_color, c->z, c->raw_rgb, c->occ[0], c->occ[1], c->occ[2], c->g_raw, c->o_rgb, c->o_depth, c->o_var, c->o_w, c->g_rgb, c->g_depth,
                    c->g_var, c->d_rays, c->absdiff, c->valid, c->idx, c->idx_pool, c->pts, c->stats, c->median, c->count, c->stash, c->masks, c->comp[1], c->comp[2], c->comp[3], c->dbg, c->scratch_ncdhw,
                    c->cam_grad_last, c->trk_scratch, c->tile_ctr, c->ray_list, c->ray_count, c->wg_mscr, c->wimg_fwd[1], c->wimg_fwd[2], c->wimg_fwd[3], c->wimg_bwd[1], c->wimg_bwd[2], c->wimg_bwd[3], c->wimg_cmp[1], c->wimg_cmp[2], c->wimg_cmp[3], c->wimg_t5[1], c->wimg_t5[2], c->wimg_t5[3], c->wimg_t5b[1], c->wimg_t5b[2], c->vmask[0], c->vmask[1], c->vmask[2], c->vmask[3], c->vmask_tmp[0], c->vmask_tmp[1], c->vmask_tmp[2], c->vmask_tmp[3]};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    for (auto& r : c->ev_pool) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (c->ev_upload) cudaEventDestroy(c->ev_upload);
    if (c->upload_stream) cudaStreamDestroy(c->upload_stream);
    if (c->ev_bwd) cudaEventDestroy(c->ev_bwd);
    if (c->ev_comm) cudaEventDestroy(c->ev_comm);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// ---- state ---------------------------------------------------------------------------------------------------------
static int ensure_scratch(nsb_ctx* ctx, size_t n) {
    if (ctx->scratch_n >= n) return 0;
    if (ctx->scratch_ncdhw) cudaFree(ctx->scratch_ncdhw);
    ctx->scratch_ncdhw = nullptr; ctx->scratch_n = 0;
    CK(dalloc(&ctx->scratch_ncdhw, n));
    ctx->scratch_n = n;
    return 0;
}

extern "C" int nsb_set_grid(nsb_ctx* ctx, int level, const float* host) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (level < 0 || level > 3) return fail(ctx, "bad level %d", level);
    const size_t n = ctx->nvox[level] * CDIM;
    if (ensure_scratch(ctx, n)) return -1;
    CK(cudaMemcpyAsync(ctx->scratch_ncdhw, host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_ncdhw_to_cl<<<cdiv((int)n, 256), 256, 0, ctx->stream>>>(ctx->scratch_ncdhw, ctx->param + ctx->off_grid[level], CDIM, (int)ctx->nvox[level]);
    ctx->launches++;
    CK(cudaGetLastError()); CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
static int get_cl(nsb_ctx* ctx, const float* src, int level, float* host) {
    const size_t n = ctx->nvox[level] * CDIM;
    if (ensure_scratch(ctx, n)) return -1;
    k_cl_to_ncdhw<<<cdiv((int)n, 256), 256, 0, ctx->stream>>>(src, ctx->scratch_ncdhw, CDIM, (int)ctx->nvox[level]);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host, ctx->scratch_ncdhw, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int nsb_get_grid(nsb_ctx* ctx, int level, float* host) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (level < 0 || level > 3) return fail(ctx, "bad level %d", level);
    return get_cl(ctx, ctx->param + ctx->off_grid[level], level, host);
}
extern "C" int nsb_get_grid_grad(nsb_ctx* ctx, int level, float* host) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (level < 0 || level > 3) return fail(ctx, "bad level %d", level);
    return get_cl(ctx, ctx->grad + ctx->off_grid[level], level, host);
}
extern "C" int nsb_set_decoder(nsb_ctx* ctx, int which, const float* host, int64_t n) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (which < 0 || which > 3 || n != ctx->dec_n[which]) return fail(ctx, "decoder %d: expected %lld floats, got %lld", which, (long long)ctx->dec_n[which], (long long)n);
    CK(cudaMemcpyAsync(ctx->param + ctx->off_dec[which], host, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->comp_dirty |= 1 << which; ctx->wimg_dirty |= 1 << which; ctx->wimg_cmp_dirty |= 1 << which;
    return 0;
}
static int get_dec(nsb_ctx* ctx, const float* base, int which, float* host, int64_t n) {
    if (which < 0 || which > 3 || n != ctx->dec_n[which]) return fail(ctx, "decoder %d: expected %lld floats, got %lld", which, (long long)ctx->dec_n[which], (long long)n);
    CK(cudaMemcpyAsync(host, base + ctx->off_dec[which], n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int nsb_get_decoder(nsb_ctx* ctx, int which, float* host, int64_t n) { return get_dec(ctx, ctx->param, which, host, n); }
extern "C" int nsb_get_decoder_grad(nsb_ctx* ctx, int which, float* host, int64_t n) { return get_dec(ctx, ctx->grad, which, host, n); }
extern "C" int nsb_set_ttables(nsb_ctx* ctx, const float* t32, const float* t16) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    CK(cudaMemcpyAsync(ctx->t_samples, t32, 32 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->t_surface, t16, 16 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int nsb_set_voxel_mask(nsb_ctx* ctx, int level, const uint8_t* host) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (level < 0 || level > 3) return fail(ctx, "bad level %d", level);
    if (!host) { if (ctx->vmask[level]) { cudaFree(ctx->vmask[level]); ctx->vmask[level] = nullptr; } return 0; }
    if (!ctx->vmask[level]) CK(dalloc(&ctx->vmask[level], ctx->nvox[level]));
    CK(cudaMemcpyAsync(ctx->vmask[level], host, ctx->nvox[level], cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// Mapper::get_mask_from_c2w (Mapper.cpp:42-130): frustum voxel mask of grid `level` for the depth frame in `slot` seen from
// c2w16 (NULL = the slot's pose).  host_mask_zyx (Z*Y*X bytes) may be NULL; install != 0 makes it the level's Adam mask.
extern "C" int nsb_frustum_mask(nsb_ctx* ctx, int slot, const float* c2w16, int level, uint8_t* host_mask_zyx, int install) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (wait_uploads(ctx)) return -1;
    if (level < 0 || level > 3) return fail(ctx, "bad level %d", level);
    if (slot < 0 || slot >= ctx->cfg.max_frames) return fail(ctx, "bad frame slot %d", slot);
    float c2w[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (c2w16) memcpy(c2w, c2w16, 12 * sizeof(float));
    else { CK(cudaMemcpyAsync(c2w, ctx->f_pose + 12 * slot, 12 * 4, cudaMemcpyDeviceToHost, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); }
    const size_t nv = ctx->nvox[level];
    if (!ctx->vmask_tmp[level]) CK(dalloc(&ctx->vmask_tmp[level], nv));
    if (ensure_scratch(ctx, nv)) return -1;
    FrustumParams P; memset(&P, 0, sizeof P);
    // general 4x4 inverse of [R|t; 0 0 0 1] in double (np.linalg.inv upstream): w2c = [R^-1 | -R^-1 t]
    {
        const double a[9] = {c2w[0], c2w[1], c2w[2], c2w[4], c2w[5], c2w[6], c2w[8], c2w[9], c2w[10]};
        const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
        double inv[9] = {(a[4] * a[8] - a[5] * a[7]) / det, (a[2] * a[7] - a[1] * a[8]) / det, (a[1] * a[5] - a[2] * a[4]) / det,
                         (a[5] * a[6] - a[3] * a[8]) / det, (a[0] * a[8] - a[2] * a[6]) / det, (a[2] * a[3] - a[0] * a[5]) / det,
                         (a[3] * a[7] - a[4] * a[6]) / det, (a[1] * a[6] - a[0] * a[7]) / det, (a[0] * a[4] - a[1] * a[3]) / det};
        const double t[3] = {c2w[3], c2w[7], c2w[11]};
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) P.w2c[4 * r + c] = (float)inv[3 * r + c];
            P.w2c[4 * r + 3] = (float)(-(inv[3 * r] * t[0] + inv[3 * r + 1] * t[1] + inv[3 * r + 2] * t[2]));
            P.cam_o[r] = (float)t[r];
        }
    }
    P.depth = ctx->f_depth + (size_t)ctx->cfg.H * ctx->cfg.W * slot;
    P.bnd = ctx->bnd; P.H = ctx->cfg.H; P.W = ctx->cfg.W; P.Z = ctx->gdim[level][0]; P.Y = ctx->gdim[level][1]; P.X = ctx->gdim[level][2];
    P.fx = ctx->cfg.fx; P.fy = ctx->cfg.fy; P.cx = ctx->cfg.cx; P.cy = ctx->cfg.cy;
    P.vdepth = ctx->scratch_ncdhw; P.stats = ctx->median; P.mask = ctx->vmask_tmp[level];
    CK(cudaMemsetAsync(ctx->median, 0, 16, ctx->stream));
    if (level == NSB_COARSE) {   // Mapper.cpp:54-59: the coarse grid is always fully selected
        CK(cudaMemsetAsync(ctx->vmask_tmp[level], 1, nv, ctx->stream));
    } else {
        k_frustum_depth<<<cdiv((int)nv, 256), 256, 0, ctx->stream>>>(P);
        k_frustum_mask<<<cdiv((int)nv, 256), 256, 0, ctx->stream>>>(P);
        ctx->launches += 2;
        CK(cudaGetLastError());
    }
    if (host_mask_zyx) CK(cudaMemcpyAsync(host_mask_zyx, ctx->vmask_tmp[level], nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (install) {
        if (!ctx->vmask[level]) CK(dalloc(&ctx->vmask[level], nv));
        CK(cudaMemcpyAsync(ctx->vmask[level], ctx->vmask_tmp[level], nv, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (host_mask_zyx) CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int nsb_set_frame_pose(nsb_ctx* ctx, int slot, const float* c2w16) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (slot < 0 || slot >= ctx->cfg.max_frames) return fail(ctx, "bad frame slot %d", slot);
    CK(cudaMemcpyAsync(ctx->f_pose + 12 * slot, c2w16, 12 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
extern "C" int nsb_set_frame(nsb_ctx* ctx, int slot, const float* depth, const float* color, const float* c2w16) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (slot < 0 || slot >= ctx->cfg.max_frames) return fail(ctx, "bad frame slot %d", slot);
    const size_t hw = (size_t)ctx->cfg.H * ctx->cfg.W;
    CK(cudaMemcpyAsync(ctx->f_depth + hw * slot, depth, hw * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->f_color + 3 * hw * slot, color, 3 * hw * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (c2w16) CK(cudaMemcpyAsync(ctx->f_pose + 12 * slot, c2w16, 12 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// Frame ingest without stalling the optimiser: the copy runs on a separate stream from (ideally pinned, see
// nsb_host_alloc) host buffers while the compute stream keeps iterating; the next call that samples frames waits for it on the
// device (event), not on the host.  The caller keeps the buffers alive until nsb_frames_ready or the next synchronising call.
extern "C" int nsb_set_frame_async(nsb_ctx* ctx, int slot, const float* depth, const float* color, const float* c2w16) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (slot < 0 || slot >= ctx->cfg.max_frames) return fail(ctx, "bad frame slot %d", slot);
    const size_t hw = (size_t)ctx->cfg.H * ctx->cfg.W;
    CK(cudaMemcpyAsync(ctx->f_depth + hw * slot, depth, hw * 4, cudaMemcpyHostToDevice, ctx->upload_stream));
    CK(cudaMemcpyAsync(ctx->f_color + 3 * hw * slot, color, 3 * hw * 4, cudaMemcpyHostToDevice, ctx->upload_stream));
    if (c2w16) CK(cudaMemcpyAsync(ctx->f_pose + 12 * slot, c2w16, 12 * 4, cudaMemcpyHostToDevice, ctx->upload_stream));
    CK(cudaEventRecord(ctx->ev_upload, ctx->upload_stream));
    ctx->upload_pending = true;
    return 0;
}
extern "C" int nsb_frames_ready(nsb_ctx* ctx) { CK(cudaStreamSynchronize(ctx->upload_stream)); return 0; }
extern "C" int nsb_host_alloc(void** p, size_t bytes) { return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? 0 : -1; }
extern "C" int nsb_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? 0 : -1; }

// ---- checkpoint (nice_slam.yaml mapping.ckpt_freq; upstream saves the grids and the decoders): flat little-endian file,
// grids in the reference's (1,C,Z,Y,X) layout, decoders as the flat vectors of nsb_set_decoder -------------------------------------
struct CkptHeader { char magic[8]; int32_t abi, c_dim, gdim[4][3]; int64_t dec_n[4]; };
extern "C" int nsb_save_checkpoint(nsb_ctx* ctx, const char* path) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    FILE* f = fopen(path, "wb");
    if (!f) return fail(ctx, "cannot open %s for writing", path);
    CkptHeader h; memset(&h, 0, sizeof h); memcpy(h.magic, "NSBCKPT1", 8); h.abi = NSB_ABI_VERSION; h.c_dim = CDIM;
    for (int l = 0; l < 4; ++l) { for (int a = 0; a < 3; ++a) h.gdim[l][a] = ctx->gdim[l][a]; h.dec_n[l] = ctx->dec_n[l]; }
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    std::vector<float> buf;
    for (int l = 0; l < 4 && ok; ++l) {
        buf.resize(ctx->nvox[l] * CDIM);
        if (nsb_get_grid(ctx, l, buf.data())) { fclose(f); return -1; }
        ok = fwrite(buf.data(), 4, buf.size(), f) == buf.size();
    }
    for (int d = 0; d < 4 && ok; ++d) {
        buf.resize((size_t)ctx->dec_n[d]);
        if (nsb_get_decoder(ctx, d, buf.data(), ctx->dec_n[d])) { fclose(f); return -1; }
        ok = fwrite(buf.data(), 4, buf.size(), f) == buf.size();
    }
    fclose(f);
    return ok ? 0 : fail(ctx, "short write to %s", path);
}
extern "C" int nsb_load_checkpoint(nsb_ctx* ctx, const char* path) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    FILE* f = fopen(path, "rb");
    if (!f) return fail(ctx, "cannot open %s", path);
    CkptHeader h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "NSBCKPT1", 8) != 0) { fclose(f); return fail(ctx, "%s is not an nsb checkpoint", path); }
    for (int l = 0; l < 4; ++l) {
        if (h.dec_n[l] != ctx->dec_n[l] || h.gdim[l][0] != ctx->gdim[l][0] || h.gdim[l][1] != ctx->gdim[l][1] || h.gdim[l][2] != ctx->gdim[l][2] || h.c_dim != CDIM) {
            fclose(f); return fail(ctx, "checkpoint %s was written for another configuration (level %d)", path, l);
        }
    }
    std::vector<float> buf;
    for (int l = 0; l < 4; ++l) {
        buf.resize(ctx->nvox[l] * CDIM);
        if (fread(buf.data(), 4, buf.size(), f) != buf.size()) { fclose(f); return fail(ctx, "checkpoint %s is truncated", path); }
        if (nsb_set_grid(ctx, l, buf.data())) { fclose(f); return -1; }
    }
    for (int d = 0; d < 4; ++d) {
        buf.resize((size_t)ctx->dec_n[d]);
        if (fread(buf.data(), 4, buf.size(), f) != buf.size()) { fclose(f); return fail(ctx, "checkpoint %s is truncated", path); }
        if (nsb_set_decoder(ctx, d, buf.data(), ctx->dec_n[d])) { fclose(f); return -1; }
    }
    fclose(f);
    return 0;
}

extern "C" int nsb_seed(nsb_ctx* ctx, uint64_t seed) { ctx->rng.seed((uint32_t)seed); return 0; }

// ---- shared pipeline pieces ---------------------------------------------------------------------------------------
static void draw_indices(nsb_ctx* ctx, int n, int64_t range, int64_t* dst) {
    for (int i = 0; i < n; ++i) dst[i] = (int64_t)((uint64_t)ctx->rng() % (uint64_t)range);   // torch::randint on the CPU generator
}

static GridView grid_view(nsb_ctx* ctx, int l) {
    GridView g; g.data = ctx->param + ctx->off_grid[l]; g.grad = ctx->grad + ctx->off_grid[l];
    g.Z = ctx->gdim[l][0]; g.Y = ctx->gdim[l][1]; g.X = ctx->gdim[l][2];
    return g;
}

// Split `grid` CTAs between the decoders in proportion to their per-tile cost.
static void partition(int grid, const float w[4], int cta_begin[5]) {
    float tot = 0; int nact = 0;
    for (int d = 0; d < 4; ++d) { tot += w[d]; nact += w[d] > 0; }
    grid = std::max(grid, nact);
    int n[4], used = 0;
    for (int d = 0; d < 4; ++d) { n[d] = w[d] > 0 ? std::max(1, (int)std::floor(grid * w[d] / tot)) : 0; used += n[d]; }
    for (int d = 0; used < grid; d = (d + 1) % 4) if (w[d] > 0) { n[d]++; used++; }
    for (int d = 3; used > grid; d = (d + 3) % 4) if (n[d] > 1) { n[d]--; used--; }
    cta_begin[0] = 0;
    for (int d = 0; d < 4; ++d) cta_begin[d + 1] = cta_begin[d] + n[d];
}

static void env_weights(const char* name, float w[4]) {
    const char* e = getenv(name);
    if (!e) return;
    float a, b, c, d;
    if (sscanf(e, "%f,%f,%f,%f", &a, &b, &c, &d) == 4) { if (w[0] > 0) w[0] = a; if (w[1] > 0) w[1] = b; if (w[2] > 0) w[2] = c; if (w[3] > 0) w[3] = d; }
}

// Rebuilds the pre-split shared-memory images of the decoders whose weights changed (set_decoder / an Adam step with a decoder
// learning rate).  cmp_mask: decoders whose COMPOSED forward image the coming launches read.  force: rebuild these decoders
// whatever the dirty bits say (the captured colour iteration ends with the rebuild of the colour decoder it has just stepped).
static int refresh_images(nsb_ctx* ctx, int cmp_mask, int force = 0, int lazy = 0, int skip = 0, cudaStream_t on = nullptr) {
    // lazy (the stash path's colour iterations, `force` = the colour decoder, stepped by the previous iteration): only the images the
    // iteration itself reads are rebuilt -- at its START, on the auxiliary stream beside k_sample / k_zvals (enqueue_iteration); the host
    // marks all of the decoder's images stale after every such iteration (mark_color_stale) and the next API entry that needs them
    // rebuilds them once.
    //   lazy 1 (warp-MMA stash forward): only the plain forward / backward images now; composed weights, composed image, tcgen05 image later
    //   lazy 2 (tcgen05 stash forward):  plain images + composed weights + tcgen05 image now; the warp-MMA composed image later
    // skip: decoders left alone whatever their dirty bits say (a stash colour iteration rebuilds the colour decoder's images itself).
    // on: the stream the launches go to (default: the context's main stream).
    cudaStream_t st = on ? on : ctx->stream;
    const int plain = (ctx->wimg_dirty & ~skip) | force;
    const int cmp_need = (ctx->wimg_cmp_dirty & cmp_mask & ~skip) | (lazy ? 0 : force);
    const int t5_need = cmp_need | (lazy == 2 ? force : 0);
    if (!plain && !cmp_need && !t5_need) return 0;
    const float* flat[4]; for (int d = 0; d < 4; ++d) flat[d] = ctx->param + ctx->off_dec[d];
    const int comp_need = (ctx->comp_dirty & (cmp_need | t5_need)) | (lazy == 1 ? 0 : force);
    if (comp_need) {   // the composed images are built from k_compose's output
        CK(launch_compose(flat, ctx->comp, comp_need, st)); ctx->launches++;
        ctx->comp_dirty &= ~comp_need;
    }
    if (plain || cmp_need) { CK(launch_build_wimg(flat, ctx->comp, ctx->wimg_fwd, ctx->wimg_bwd, ctx->wimg_cmp, plain, cmp_need, st)); ctx->launches++; }
    if (ctx->use_tc == 3 && t5_need) { CK(launch_build_t5img(flat, ctx->comp, ctx->wimg_t5, t5_need, st)); ctx->launches++; }
    if (ctx->use_tc == 3 && (cmp_need & 0x6)) { CK(launch_build_t5bimg(flat, ctx->comp, ctx->wimg_t5b, cmp_need, st)); ctx->launches++; }
    if ((plain & 8) && !ctx->wg_stash) { CK(launch_build_wgimg(flat[3], ctx->wg_img, st)); ctx->launches++; }   // plane image of the colour decoder (stash-free weight gradient only)
    ctx->wimg_dirty &= ~plain; ctx->wimg_cmp_dirty &= ~cmp_need;
    return 0;
}

// Colour iterations of the stash path rebuild the stepped colour decoder's images lazily (see refresh_images): 0 = no, 1 / 2 = which set.
static int lazy_color_images(const nsb_ctx* ctx) {
    if (!ctx->wg_stash || ctx->map_fix_color || ctx->coarse_map || ctx->use_tc != 3) return 0;
    return ctx->t5_stash ? 2 : 1;
}

static void fill_decode_params(nsb_ctx* ctx, DecodeParams& P, int n, int S, const uint8_t* valid) {
    memset(&P, 0, sizeof P);
    for (int d = 0; d < 4; ++d) { P.dec_flat[d] = ctx->param + ctx->off_dec[d]; P.grid[d] = grid_view(ctx, d); P.wimg_fwd[d] = ctx->wimg_fwd[d]; P.wimg_bwd[d] = ctx->wimg_bwd[d]; P.wimg_cmp[d] = ctx->wimg_cmp[d]; P.wimg_t5[d] = ctx->wimg_t5[d]; P.wimg_t5b[d] = ctx->wimg_t5b[d]; }
    P.bnd = ctx->bnd;
    P.rays_o = ctx->rays_o; P.rays_d = ctx->rays_d; P.z = ctx->z; P.valid = valid; P.pts = nullptr;
    P.S = S; P.P = n * S;
    P.out_rgb = ctx->raw_rgb; P.out_occ[0] = ctx->occ[0]; P.out_occ[1] = ctx->occ[1]; P.out_occ[2] = ctx->occ[2];
    P.g_raw = ctx->g_raw; P.d_rays = ctx->d_rays; P.stash = nullptr; P.masks = nullptr;
}

static const IterRef NO_ITER = {nullptr, 1};

// Stash colour iterations rebuild the colour decoder's images on the auxiliary stream at their start (see refresh_images): fork / join.
static int fork_rebuild(nsb_ctx* ctx, int lazy) {
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    if (refresh_images(ctx, 0, 1 << 3, lazy, 0, ctx->aux_stream)) return -1;
    CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    ctx->pending_join = true;
    return 0;
}
static int join_rebuild(nsb_ctx* ctx) {
    if (!ctx->pending_join) return 0;
    ctx->pending_join = false;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    return 0;
}

static int decode_grid_size(nsb_ctx* ctx, int P) {
    const int occ = std::max(1, ctx->occ_blocks[ctx->cfg.precision ? 1 : 0]);
    const int tiles = cdiv(P, TILE);
    return std::max(1, std::min(ctx->n_sm * occ, cdiv(tiles, std::min(FWD_WARPS, BWD_WARPS))));
}

// Decoders a stage evaluates (NICE.cpp:16-51).
static void stage_decoders(int stage, float w[4]) {
    w[0] = w[1] = w[2] = w[3] = 0;
    if (stage == NSB_COARSE) w[0] = 300;
    else if (stage == NSB_MIDDLE) w[1] = 732;
    else if (stage == NSB_FINE) { w[1] = 732; w[2] = 972; }
    else { w[1] = 732; w[2] = 972; w[3] = 732; }
}

// rays (ctx->rays_*, ctx->gt_depth or null, ctx->valid or null) -> z, raw, composite outputs in ctx buffers.
// `off`/`n` select the slice of the ray batch this rank renders (ranks share the batch-global statistics).
static int ensure_stash(nsb_ctx* ctx, size_t rows) {
    if (ctx->stash_rows >= rows) return 0;
    if (ctx->stash) cudaFree(ctx->stash);
    ctx->stash = nullptr; ctx->stash_rows = 0;
    CK(dalloc(&ctx->stash, rows * stash::W + 64));
    ctx->stash_rows = rows;
    return 0;
}

// train: keep the relu masks for run_backward; stash_fwd: also write the colour decoder's activations to the wgrad stash.
// stats: base of a statistics ring, `it` selects the slot on the device ({nullptr, 1}: slot 0).  The decoder images must be fresh
// (refresh_images) before this is called.
static int run_forward(nsb_ctx* ctx, int stage, int off, int n, bool have_depth, const uint8_t* valid, float* stats, IterRef it, bool want_weights,
                       bool train = false, bool stash_fwd = false, bool skip_composite = false) {
    const nsb_config& c = ctx->cfg;
    const int S = have_depth ? c.n_samples + c.n_surface : c.n_samples;
    ctx->last_n = n; ctx->last_S = S;
    // will the tcgen05 forward run (alone, or beside the stashing colour decoder)?  Then k_zvals also compacts the rays that pass the inside filter.
    const bool want_stash = train && stash_fwd && stage == NSB_COLOR && ctx->wg_stash;
    const bool t5_base = ctx->use_tc == 3 && c.precision == NSB_PREC_FP32_GRADE && stage != NSB_COARSE && !(train && stash_fwd && stage == NSB_COLOR && !ctx->wg_stash);
    const bool t5_all_stash = t5_base && want_stash && ctx->t5_stash && !((ctx->comp_dirty >> 1) & 7);     // one tcgen05 launch also writes the stash
    const bool t5_split = t5_base && want_stash && !t5_all_stash && ctx->split_color_sms > 0 && ctx->split_color_sms < ctx->n_sm - 8 && !((ctx->comp_dirty >> 1) & 3);
    const bool compact = valid != nullptr && ctx->compact_rays && (t5_split || t5_all_stash || (t5_base && !want_stash));
    {
        Timer t(ctx, T_SAMPLE);
        ZParams Z; Z.rays_o = ctx->rays_o + 3 * off; Z.rays_d = ctx->rays_d + 3 * off; Z.gt_depth = have_depth ? ctx->gt_depth + off : nullptr;
        Z.valid = valid ? valid + off : nullptr; Z.stats = stats; Z.it = it; Z.zero_ctr = ctx->tile_ctr; Z.t_samples = ctx->t_samples; Z.t_surface = ctx->t_surface; Z.bnd = ctx->bnd;
        Z.ray_list = compact ? ctx->ray_list : nullptr; Z.ray_count = ctx->ray_count;
        Z.n = n; Z.n_samples = c.n_samples; Z.n_surface = have_depth ? c.n_surface : 0; Z.z = ctx->z + (size_t)off * S;
        k_zvals<<<cdiv(n * 32, 256), 256, 0, ctx->stream>>>(Z); ctx->launches++;
        CK(cudaGetLastError());
    }
    if (join_rebuild(ctx)) return -1;   // the colour decoder's images, rebuilt beside the two kernels above
    {
        Timer t(ctx, T_FWD);
        // the forward runs on the composed images, except for a colour decoder whose activations are stashed for the weight gradient
        const bool will_stash = train && stash_fwd && stage == NSB_COLOR;
        (void)will_stash;
        DecodeParams P; fill_decode_params(ctx, P, n, S, valid ? valid + off : nullptr);
        P.rays_o += 3 * off; P.rays_d += 3 * off; P.z += (size_t)off * S;
        P.out_rgb += 4 * (size_t)off * S; for (int k = 0; k < 3; ++k) P.out_occ[k] += (size_t)off * S;
        if (train) P.masks = ctx->masks;
        if (compact) { P.ray_list = ctx->ray_list; P.ray_count = ctx->ray_count; }
        if (train && stash_fwd && stage == NSB_COLOR && ctx->wg_stash) { if (ensure_stash(ctx, (size_t)n * S)) return -1; P.stash = ctx->stash; }
        float w[4]; stage_decoders(stage, w);
        if (P.stash) { w[1] = 700; w[2] = 972; w[3] = 860; }   // the colour decoder also writes its activations to the wgrad stash (measured split, tools/sweep_split.sh)
        env_weights("NSB_SPLIT_FWD", w);
        // the stash-free weight-gradient kernel reads the warp-MMA forward's fragment-packed relu masks
        const bool fused_wg = train && stash_fwd && stage == NSB_COLOR && !ctx->wg_stash;
        const bool tc_ok = ctx->use_tc && c.precision == NSB_PREC_FP32_GRADE && stage != NSB_COARSE && P.stash == nullptr && !fused_wg;
        // Colour iteration with a weight-gradient stash: decoders 1 and 2 run on the tcgen05 kernel while the stashing colour
        // decoder runs on the warp-MMA kernel, side by side on disjoint SMs (both kernels take a whole SM per CTA), forked
        // onto a second stream and joined before the composite -- capturable, so it lives inside the iteration's graph.
        const bool split_fwd = t5_split && P.stash != nullptr;
        if (P.stash) ctx->stash_is_u = false;
        if (t5_all_stash && P.stash) {
            // all three decoders on the tcgen05 kernel; the colour decoder's threads write the stash rows of their samples
            for (int d = 0; d < 4; ++d) P.comp[d] = ctx->comp[d];
            P.mask_layout = 0xE; P.mask_stride = n * S;
            ctx->mask_layout = 0xE; ctx->mask_stride = n * S; ctx->stash_is_u = true;
            float w5[4] = {0, 700.f, 1300.f, 1300.f}; env_weights("NSB_SPLIT_FWD_T5S", w5);     // measured (the colour decoder also writes 288 floats per sample)
            partition(std::min(ctx->n_sm, std::max(1, cdiv(n * S, 512))), w5, P.cta_begin);
            P.tile_ctr = ctx->tile_ctr;
            CK(launch_decode_fwd_t5(P, P.cta_begin[4], ctx->stream)); ctx->launches++;
        } else if (split_fwd) {
            DecodeParams PA = P, PB = P;
            PA.stash = nullptr; PB.ray_list = nullptr; PB.ray_count = nullptr;
            for (int d = 0; d < 4; ++d) PA.comp[d] = ctx->comp[d];
            PA.mask_layout = 0x6; PA.mask_stride = n * S; PA.tile_ctr = ctx->tile_ctr;
            float wa[4] = {0, 700.f, 1300.f, 0}; env_weights("NSB_SPLIT_FWD_T5", wa); wa[3] = 0;
            partition(ctx->n_sm - ctx->split_color_sms, wa, PA.cta_begin);
            const float wb[4] = {0, 0, 0, 1};
            partition(ctx->split_color_sms, wb, PB.cta_begin);
            PB.tile_ctr = ctx->tile_ctr;
            ctx->mask_layout = 0x6; ctx->mask_stride = n * S;     // decoders 1, 2: per-sample words; decoder 3: fragment-packed
            CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
            CK(launch_decode_fwd(PB, c.precision, PB.cta_begin[4], ctx->aux_stream)); ctx->launches++;
            CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
            CK(launch_decode_fwd_t5(PA, PA.cta_begin[4], ctx->stream)); ctx->launches++;
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        } else if (tc_ok) {
            // tcgen05 path: one 320-thread CTA per SM, per-sample mask words, composed weights refreshed when stale
            int need = 0;
            for (int d = 1; d < 4; ++d) if (w[d] > 0 && ((ctx->comp_dirty >> d) & 1)) need |= 1 << d;
            if (need) {
                if (ctx->capturing) return fail(ctx, "tcgen05 forward: composed weights must be fresh before graph capture");
                const float* flat[4]; for (int d = 0; d < 4; ++d) flat[d] = ctx->param + ctx->off_dec[d];
                CK(launch_compose(flat, ctx->comp, need, ctx->stream)); ctx->launches++;
                ctx->comp_dirty &= ~need;
            }
            for (int d = 0; d < 4; ++d) P.comp[d] = ctx->comp[d];
            P.dbg = ctx->dbg;
            P.mask_layout = 0xE; P.mask_stride = n * S;     // bit d: decoder d's relu masks are per-sample words
            if (train) { ctx->mask_layout = 0xE; ctx->mask_stride = n * S; }
            float wt[4] = {0, w[1], w[2], w[3]}; env_weights("NSB_SPLIT_FWD_TC", wt);
            if (ctx->use_tc >= 2) {   // kind::f16 kernels: three 128-sample tiles per CTA (3 = every A operand in tensor memory)
                partition(std::min(ctx->n_sm, std::max(1, cdiv(n * S, 384))), wt, P.cta_begin);
                if (ctx->use_tc == 3) {   // four tile groups per CTA (three for the fine decoder), tiles drawn from the ticket counters
                    float w5[4] = {0, w[1] > 0 ? 700.f : 0.f, w[2] > 0 ? 1300.f : 0.f, w[3] > 0 ? 760.f : 0.f}; env_weights("NSB_SPLIT_FWD_T5", w5);
                    partition(std::min(ctx->n_sm, std::max(1, cdiv(n * S, 512))), w5, P.cta_begin);
                    P.tile_ctr = ctx->tile_ctr;
                    CK(launch_decode_fwd_t5(P, P.cta_begin[4], ctx->stream));
                }
                else CK(launch_decode_fwd_tc16(P, P.cta_begin[4], ctx->stream));
                ctx->launches++;
            } else {
                partition(std::min(ctx->n_sm, std::max(1, cdiv(n * S, 256))), wt, P.cta_begin);
                CK(launch_decode_fwd_tc(P, P.cta_begin[4], ctx->stream)); ctx->launches++;
            }
        } else {
            if (train) { ctx->mask_layout = 0; ctx->mask_stride = 0; }
            const int grid = decode_grid_size(ctx, n * S);
            partition(grid, w, P.cta_begin);
            P.tile_ctr = ctx->tile_ctr;
            CK(launch_decode_fwd(P, c.precision, P.cta_begin[4], ctx->stream)); ctx->launches++;
        }
    }
    if (!skip_composite) {
        Timer t(ctx, T_COMP);
        CompositeParams Q; memset(&Q, 0, sizeof Q);
        Q.rays_o = ctx->rays_o + 3 * off; Q.rays_d = ctx->rays_d + 3 * off; Q.z = ctx->z + (size_t)off * S; Q.valid = valid ? valid + off : nullptr;
        Q.raw_rgb = ctx->raw_rgb + 4 * (size_t)off * S; for (int k = 0; k < 3; ++k) Q.occ[k] = ctx->occ[k] + (size_t)off * S;
        Q.stats = stats; Q.it = it; Q.bnd = ctx->bnd; Q.n = n; Q.S = S; Q.stage = stage; Q.occupancy = c.occupancy; Q.dist_norm = c.dist_norm;
        Q.rgb = ctx->o_rgb + 3 * off; Q.depth = ctx->o_depth + off; Q.var = ctx->o_var + off; Q.weights = want_weights ? ctx->o_w + (size_t)off * S : nullptr;
        if (ctx->trk_hook) { Q.trk_gt_depth = ctx->gt_depth + off; Q.trk_absdiff = ctx->absdiff; Q.trk_count = ctx->trk_count; }
        k_composite_fwd<<<cdiv(n * 32, 256), 256, 0, ctx->stream>>>(Q); ctx->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

// Sum all-reduce of the gradient-arena floats [begin, end) over the ranks, on stream st.
static int allreduce_range(nsb_ctx* ctx, size_t begin, size_t end, cudaStream_t st) {
    if (end <= begin) return 0;
    const int rc = g_nccl.AllReduce(ctx->grad + begin, ctx->grad + begin, end - begin, NCCL_FLOAT32, NCCL_SUM, ctx->comm, st);
    if (rc != 0) return fail(ctx, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return 0;
}

// cotangents in ctx->g_rgb / g_depth / g_var -> g_raw -> decoder backward (+ wgrad).  flags: F_GRID=1, F_WGRAD=2, F_RAY=4.
static int run_backward(nsb_ctx* ctx, int stage, int off, int n, const uint8_t* valid, float* stats, IterRef it, int flags, bool color_active,
                        bool skip_composite = false) {
    const nsb_config& c = ctx->cfg;
    const int S = ctx->last_S;
    if (!skip_composite) {
        Timer t(ctx, T_COMP);
        CompositeParams Q; memset(&Q, 0, sizeof Q);
        Q.rays_o = ctx->rays_o + 3 * off; Q.rays_d = ctx->rays_d + 3 * off; Q.z = ctx->z + (size_t)off * S; Q.valid = valid ? valid + off : nullptr;
        Q.raw_rgb = ctx->raw_rgb + 4 * (size_t)off * S; for (int k = 0; k < 3; ++k) Q.occ[k] = ctx->occ[k] + (size_t)off * S;
        Q.stats = stats; Q.it = it; Q.zero_ctr = ctx->tile_ctr + 4; Q.bnd = ctx->bnd; Q.n = n; Q.S = S; Q.stage = stage; Q.occupancy = c.occupancy; Q.dist_norm = c.dist_norm;
        Q.g_rgb = ctx->g_rgb + 3 * off; Q.g_depth = ctx->g_depth + off; Q.g_var = ctx->g_var + off; Q.g_raw = ctx->g_raw + 4 * (size_t)off * S;
        if (flags & 4) { CK(cudaMemsetAsync(ctx->d_rays + 6 * (size_t)off, 0, 6 * (size_t)n * 4, ctx->stream)); Q.d_rays = ctx->d_rays + 6 * (size_t)off; }
        k_composite_bwd<<<cdiv(n * 32, 256), 256, 0, ctx->stream>>>(Q); ctx->launches++;
        CK(cudaGetLastError());
    }
    const bool wg = (flags & 2) && stage == NSB_COLOR && color_active;
    const bool wg_stash = wg && ctx->wg_stash;      // stash + split-K k_wgrad; otherwise the stash-free fused kernel
    if (wg_stash && ctx->stash_rows < (size_t)n * S) return fail(ctx, "wgrad stash missing: the forward must run with stash_fwd");
    DecodeParams Pw; memset(&Pw, 0, sizeof Pw);
    {
        Timer t(ctx, T_BWD);
        DecodeParams P; fill_decode_params(ctx, P, n, S, valid ? valid + off : nullptr);
        P.rays_o += 3 * off; P.rays_d += 3 * off; P.z += (size_t)off * S; P.g_raw += 4 * (size_t)off * S; P.d_rays += 6 * (size_t)off;
        P.stash = ctx->stash; P.masks = ctx->masks; P.mask_layout = ctx->mask_layout; P.mask_stride = ctx->mask_stride;
        P.flags = (flags & 1) | (wg_stash ? 2 : 0) | (flags & 4);
        Pw = P;
        if (P.flags == 0 && !wg) return 0;
        float w[4] = {0, 0, 0, 0};
        const float ge = (flags & 4) ? 288.f : 0.f;
        if (stage == NSB_MIDDLE) w[1] = 480 + ge;
        else if (stage == NSB_FINE) { w[1] = 480 + ge; w[2] = 480 + ge; }
        else if (stage == NSB_COLOR) {
            w[1] = 480 + ge; w[2] = 480 + ge; if (color_active) w[3] = 480 + (wg_stash ? 288.f + 300.f : ge);
            if (!ge && !color_active) { w[1] = 460; w[2] = 500; }               // geometry iterations: the fine grid's scatter touches more vertices (measured)
            if (wg_stash && !ge) { w[1] = 460; w[2] = 500; w[3] = 1100; }       // measured with the 552-float stash (tools/bench_short.sh sweeps)
        }
        else {   // coarse stage (the coarse mapper): MLP_no_xyz data gradient -> grid_coarse; no embedding, so no ray gradient path here
            if (flags & 4) return fail(ctx, "ray gradients through the coarse stage are not supported (the coarse mapper runs without bundle adjustment, Mapper.cpp:530)");
            if (ctx->mask_layout != 0) return fail(ctx, "coarse backward needs the warp-MMA forward's mask layout");
            P.tile_ctr = nullptr;
            CK(launch_coarse_bwd(P, c.precision, std::max(1, std::min(ctx->n_sm * 4, cdiv(cdiv(n * S, TILE), 8))), ctx->stream)); ctx->launches++;
            return 0;
        }
        env_weights("NSB_SPLIT_BWD", w);
        P.tile_ctr = ctx->tile_ctr + 4;
        // geometry iterations (grid gradients of the two occupancy decoders only) run on the tcgen05 backward when the forward
        // left its per-sample relu masks for both decoders
        const bool t5_bwd = ctx->use_tc == 3 && ctx->bwd_t5 && c.precision == NSB_PREC_FP32_GRADE && P.flags == 1 && !wg && w[1] > 0 && w[2] > 0 && w[3] == 0 &&
                            (ctx->mask_layout & 0x6) == 0x6;
        if (t5_bwd) {
            float wb[4] = {0, 1000.f, 1000.f, 0}; env_weights("NSB_SPLIT_BWD_T5", wb);
            partition(std::min(ctx->n_sm, std::max(2, cdiv(n * S, 512))), wb, P.cta_begin);
            CK(launch_decode_bwd_t5(P, P.cta_begin[4], ctx->stream)); ctx->launches++;
        } else {
        const int grid = decode_grid_size(ctx, n * S);
        partition(grid, w, P.cta_begin);
        P.cta_begin[1] = 0;   // no coarse CTAs: decoder 1 starts at block 0
        if (P.flags != 0) { CK(launch_decode_bwd(P, c.precision, P.cta_begin[4], ctx->stream)); ctx->launches++; }
        }
    }
    ctx->ar_overlapped = false;
    if (wg && ctx->ar_request && ctx->world > 1) {
        // multi-GPU colour iteration: the grid gradients are final once k_decode_bwd has run, so their all-reduce goes to the
        // communication stream now and overlaps the weight-gradient kernel; the decoder / camera gradients follow after it
        CK(cudaEventRecord(ctx->ev_bwd, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_bwd, 0));
        if (allreduce_range(ctx, ctx->off_train, ctx->off_dec[2], ctx->comm_stream)) return -1;
        CK(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
        ctx->ar_overlapped = true;
    }
    if (wg_stash) {
        Timer t(ctx, T_WGRAD);
        CK(launch_wgrad(ctx->stash, valid ? valid + off : nullptr, n * S, S, ctx->param + ctx->off_dec[3], ctx->grad + ctx->off_dec[3], ctx->wg_mscr, ctx->stash_is_u ? 1 : 0, c.precision, ctx->n_sm, ctx->stream)); ctx->launches += 2;
    } else if (wg) {
        // colour-decoder weight gradient without a stash: recomputed per tile, contracted through shared memory (wgrad_fused.cu)
        Timer t(ctx, T_WGRAD);
        if ((ctx->mask_layout >> 3) & 1) return fail(ctx, "the fused weight-gradient kernel needs the warp-MMA forward's relu-mask layout (NSB_TCGEN05 must be 0)");
        if (c.precision != NSB_PREC_FP32_GRADE) return fail(ctx, "the fused weight-gradient kernel is fp32-grade only");
        CK(launch_wgrad_fused(Pw, ctx->wg_img, ctx->param + ctx->off_dec[3], ctx->grad + ctx->off_dec[3], ctx->wg_scratch, ctx->n_sm, ctx->stream)); ctx->launches += 2;
    }
    return 0;
}

static int zero_grads(nsb_ctx* ctx) {
    CK(cudaMemsetAsync(ctx->grad, 0, ctx->arena_n * 4, ctx->stream));
    return 0;
}

// ---- sampling API ----------------------------------------------------------------------------------------------------
static void fill_sample_params(nsb_ctx* ctx, SampleParams& P, int n, int H0, int H1, int W0, int W1, float* stats, int apply_filter) {
    const nsb_config& c = ctx->cfg;
    memset(&P, 0, sizeof P);
    P.depth = ctx->f_depth; P.color = ctx->f_color; P.poses = ctx->f_pose; P.cams = ctx->param + ctx->off_cam; P.cam_mask = 0; P.idx = ctx->idx;
    P.H = c.H; P.W = c.W; P.H0 = H0; P.W0 = W0; P.Wc = W1 - W0;
    P.fx = c.fx; P.fy = c.fy; P.cx = c.cx; P.cy = c.cy; P.raydir = c.raydir; P.bnd = ctx->bnd; P.n = n;
    P.rays_o = ctx->rays_o; P.rays_d = ctx->rays_d; P.gt_depth = ctx->gt_depth; P.gt_color = ctx->gt_color; P.valid = ctx->valid;
    P.stats = stats; P.apply_filter = apply_filter; P.it = NO_ITER; P.pool = nullptr; P.pool_iters = 1;
}

extern "C" int nsb_get_samples(nsb_ctx* ctx, int slot, const float* c2w16, int H0, int H1, int W0, int W1, int n, const int64_t* idx,
                               float* rays_o, float* rays_d, float* gt_depth, float* gt_color, uint8_t* inside, int64_t* idx_out) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (wait_uploads(ctx)) return -1;
    if (n > ctx->cap) return fail(ctx, "n %d exceeds max_rays %d", n, ctx->cap);
    if (slot < 0 || slot >= ctx->cfg.max_frames) return fail(ctx, "bad frame slot %d", slot);
    ctx->h_idx.resize(n);
    if (idx) memcpy(ctx->h_idx.data(), idx, n * sizeof(int64_t)); else draw_indices(ctx, n, (int64_t)(H1 - H0) * (W1 - W0), ctx->h_idx.data());
    if (idx_out) memcpy(idx_out, ctx->h_idx.data(), n * sizeof(int64_t));
    CK(cudaMemcpyAsync(ctx->idx, ctx->h_idx.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (c2w16) CK(cudaMemcpyAsync(ctx->f_pose + 12 * slot, c2w16, 12 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->rstats, 0, 16, ctx->stream));
    SampleParams P; fill_sample_params(ctx, P, n, H0, H1, W0, W1, ctx->rstats, 1);
    P.slots[0] = slot; P.n_frames = 1; P.pix_per_frame = n;
    k_sample<<<cdiv(n, 128), 128, 0, ctx->stream>>>(P); ctx->launches++;
    CK(cudaGetLastError());
    if (rays_o) CK(cudaMemcpyAsync(rays_o, ctx->rays_o, 12 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (rays_d) CK(cudaMemcpyAsync(rays_d, ctx->rays_d, 12 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (gt_depth) CK(cudaMemcpyAsync(gt_depth, ctx->gt_depth, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (gt_color) CK(cudaMemcpyAsync(gt_color, ctx->gt_color, 12 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (inside) CK(cudaMemcpyAsync(inside, ctx->valid, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- render API -------------------------------------------------------------------------------------------------------
// the render / sampling entry points keep their batch statistics in ctx->rstats, never in the mapping loop's ring
static int render_prepare_stats(nsb_ctx* ctx, int n, bool have_depth) {
    CK(cudaMemsetAsync(ctx->rstats, 0, 16, ctx->stream));
    if (have_depth) { k_depth_max<<<cdiv(n, 256), 256, 0, ctx->stream>>>(ctx->gt_depth, n, ctx->rstats); ctx->launches++; }
    if (ctx->cfg.dist_norm == NSB_DISTNORM_REFERENCE) { k_dirnorm_ref<<<cdiv(n, 256), 256, 0, ctx->stream>>>(ctx->rays_d, nullptr, n, ctx->rstats, NO_ITER); ctx->launches++; }
    CK(cudaGetLastError());
    return 0;
}

extern "C" int nsb_render_batch_ray_dev(nsb_ctx* ctx, int stage, int n, const float* d_rays_d, const float* d_rays_o, const float* d_gt_depth,
                                        float* d_rgb, float* d_depth, float* d_var, float* d_weights) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (n > ctx->cap) return fail(ctx, "n %d exceeds max_rays %d", n, ctx->cap);
    if (stage < 0 || stage > 3) return fail(ctx, "bad stage %d", stage);
    const bool hd = d_gt_depth != nullptr;
    const int S = hd ? ctx->cfg.n_samples + ctx->cfg.n_surface : ctx->cfg.n_samples;
    CK(cudaMemcpyAsync(ctx->rays_d, d_rays_d, 12 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->rays_o, d_rays_o, 12 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (hd) CK(cudaMemcpyAsync(ctx->gt_depth, d_gt_depth, 4 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (render_prepare_stats(ctx, n, hd)) return -1;
    if (refresh_images(ctx, 0xE)) return -1;
    if (run_forward(ctx, stage, 0, n, hd, nullptr, ctx->rstats, NO_ITER, d_weights != nullptr)) return -1;
    if (d_rgb) CK(cudaMemcpyAsync(d_rgb, ctx->o_rgb, 12 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_depth) CK(cudaMemcpyAsync(d_depth, ctx->o_depth, 4 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_var) CK(cudaMemcpyAsync(d_var, ctx->o_var, 4 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_weights) CK(cudaMemcpyAsync(d_weights, ctx->o_w, 4 * (size_t)n * S, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

static int upload_rays(nsb_ctx* ctx, int n, const float* rays_d, const float* rays_o, const float* gt_depth) {
    CK(cudaMemcpyAsync(ctx->rays_d, rays_d, 12 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->rays_o, rays_o, 12 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if (gt_depth) CK(cudaMemcpyAsync(ctx->gt_depth, gt_depth, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

extern "C" int nsb_render_batch_ray(nsb_ctx* ctx, int stage, int n, const float* rays_d, const float* rays_o, const float* gt_depth,
                                    float* rgb, float* depth, float* var, float* weights) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (stage < 0 || stage > 3) return fail(ctx, "bad stage %d", stage);
    if (n <= 0) return 0;
    const bool hd = gt_depth != nullptr;
    const int S = hd ? ctx->cfg.n_samples + ctx->cfg.n_surface : ctx->cfg.n_samples;
    // the batch-global scalars of Renderer.cpp:76,93 (and utils.h:153 in reference mode) span the WHOLE call, so
    // they are computed over all n rays before the batch is cut into max_rays chunks
    float gmax = 0.f; double inv = 0.0;
    if (hd) for (int i = 0; i < n; ++i) gmax = std::max(gmax, gt_depth[i]);
    if (ctx->cfg.dist_norm == NSB_DISTNORM_REFERENCE) for (int i = 0; i < 3 * n; ++i) inv += 1.0 / std::fabs((double)rays_d[i]);
    const float st[4] = {gmax, 0.f, (float)inv, 0.f};
    if (refresh_images(ctx, 0xE)) return -1;
    for (int o = 0; o < n; o += ctx->cap) {
        const int m = std::min(ctx->cap, n - o);
        if (upload_rays(ctx, m, rays_d + 3 * (size_t)o, rays_o + 3 * (size_t)o, hd ? gt_depth + o : nullptr)) return -1;
        CK(cudaMemcpyAsync(ctx->rstats, st, sizeof st, cudaMemcpyHostToDevice, ctx->stream));
        if (run_forward(ctx, stage, 0, m, hd, nullptr, ctx->rstats, NO_ITER, weights != nullptr)) return -1;
        if (rgb) CK(cudaMemcpyAsync(rgb + 3 * (size_t)o, ctx->o_rgb, 12 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        if (depth) CK(cudaMemcpyAsync(depth + o, ctx->o_depth, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        if (var) CK(cudaMemcpyAsync(var + o, ctx->o_var, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        if (weights) CK(cudaMemcpyAsync(weights + (size_t)o * S, ctx->o_w, 4 * (size_t)m * S, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

// Dense render of every pixel of a resident frame (upstream Renderer.render_img; the reference keeps only its chunk constant,
// Renderer.cpp:5): rays of all H*W pixels in row-major order, optionally depth-guided by the frame's own depth image,
// rendered in chunks of max_rays.  The batch-global scalars of Renderer.cpp:76,93 (and utils.h:153 in reference mode) are
// taken over the whole image in a first pass, exactly as one render_batch_ray call over all pixels would.
extern "C" int nsb_render_img(nsb_ctx* ctx, int slot, const float* c2w16, int stage, int use_gt_depth, float* rgb, float* depth, float* var) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (wait_uploads(ctx)) return -1;
    if (stage < 0 || stage > 3) return fail(ctx, "bad stage %d", stage);
    if (slot < 0 || slot >= ctx->cfg.max_frames) return fail(ctx, "bad frame slot %d", slot);
    const nsb_config& c = ctx->cfg;
    const int HW = c.H * c.W;
    if (c2w16) CK(cudaMemcpyAsync(ctx->f_pose + 12 * slot, c2w16, 12 * 4, cudaMemcpyHostToDevice, ctx->stream));
    float* stats = ctx->rstats + 4;     // whole-image scalars (kept apart from the per-chunk scratch in rstats[0..3])
    CK(cudaMemsetAsync(stats, 0, 16, ctx->stream));
    if (refresh_images(ctx, 0xE)) return -1;
    const bool ref_norm = c.dist_norm == NSB_DISTNORM_REFERENCE;
    auto sample_chunk = [&](int o, int m) -> int {
        k_iota<<<cdiv(m, 256), 256, 0, ctx->stream>>>(ctx->idx, (int64_t)o, m);
        CK(cudaMemsetAsync(ctx->rstats, 0, 16, ctx->stream));
        SampleParams P; fill_sample_params(ctx, P, m, 0, c.H, 0, c.W, ctx->rstats, 0);
        P.slots[0] = slot; P.n_frames = 1; P.pix_per_frame = m;
        k_sample<<<cdiv(m, 128), 128, 0, ctx->stream>>>(P); ctx->launches += 2;
        CK(cudaGetLastError());
        return 0;
    };
    if (use_gt_depth) { k_depth_max<<<cdiv(HW, 256), 256, 0, ctx->stream>>>(ctx->f_depth + (size_t)HW * slot, HW, stats); ctx->launches++; }
    if (ref_norm) {
        for (int o = 0; o < HW; o += ctx->cap) {
            const int m = std::min(ctx->cap, HW - o);
            if (sample_chunk(o, m)) return -1;
            k_dirnorm_ref<<<cdiv(m, 256), 256, 0, ctx->stream>>>(ctx->rays_d, nullptr, m, stats, NO_ITER); ctx->launches++;
        }
    }
    for (int o = 0; o < HW; o += ctx->cap) {
        const int m = std::min(ctx->cap, HW - o);
        if (sample_chunk(o, m)) return -1;
        if (run_forward(ctx, stage, 0, m, use_gt_depth != 0, nullptr, stats, NO_ITER, false)) return -1;
        if (rgb) CK(cudaMemcpyAsync(rgb + 3 * (size_t)o, ctx->o_rgb, 12 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        if (depth) CK(cudaMemcpyAsync(depth + o, ctx->o_depth, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        if (var) CK(cudaMemcpyAsync(var + o, ctx->o_var, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int nsb_get_last_zvals(nsb_ctx* ctx, int n, int S, float* z) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (n != ctx->last_n || S != ctx->last_S) return fail(ctx, "last render was %d x %d, asked %d x %d", ctx->last_n, ctx->last_S, n, S);
    CK(cudaMemcpyAsync(z, ctx->z, 4 * (size_t)n * S, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Decoders + stage assembly + bound mask for m points already in ctx->pts: raw (m,4) to d_raw4 and / or the occupancy channel to
// d_occ (device pointers, either may be null).  Everything stays on the device; no synchronisation.
static int eval_chunk(nsb_ctx* ctx, int stage, int m, float* d_raw4, float* d_occ) {
    const nsb_config& c = ctx->cfg;
    DecodeParams P; fill_decode_params(ctx, P, 0, 16, nullptr);
    P.pts = ctx->pts; P.P = m;
    float w[4]; stage_decoders(stage, w);
    partition(decode_grid_size(ctx, m), w, P.cta_begin);
    P.tile_ctr = ctx->tile_ctr;
    CK(cudaMemsetAsync(ctx->tile_ctr, 0, 4 * sizeof(unsigned long long), ctx->stream));
    CK(launch_decode_fwd(P, c.precision, P.cta_begin[4], ctx->stream)); ctx->launches++;
    k_assemble_raw<<<cdiv(m, 256), 256, 0, ctx->stream>>>(ctx->pts, ctx->raw_rgb, ctx->occ[0], ctx->occ[1], ctx->occ[2], stage, ctx->bnd, m, d_raw4, d_occ);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// Renderer::eval_points (Renderer.cpp:19-42), host buffers: chunks of the context's point capacity, one synchronisation at the end.
extern "C" int nsb_eval_points(nsb_ctx* ctx, int stage, int Pn, const float* pts, float* raw) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (stage < 0 || stage > 3) return fail(ctx, "bad stage %d", stage);
    const nsb_config& c = ctx->cfg;
    const size_t cap_pts = (size_t)ctx->cap * (c.n_samples + c.n_surface);
    if (refresh_images(ctx, 0xE)) return -1;
    for (size_t o = 0; o < (size_t)Pn; o += cap_pts) {
        const int m = (int)std::min(cap_pts, (size_t)Pn - o);
        CK(cudaMemcpyAsync(ctx->pts, pts + 3 * o, 12 * (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        if (eval_chunk(ctx, stage, m, ctx->g_raw, nullptr)) return -1;           // g_raw doubles as the (m,4) staging buffer
        CK(cudaMemcpyAsync(raw + 4 * o, ctx->g_raw, 16 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
// Device-pointer form: d_pts (P,3) -> d_raw (P,4), enqueued on the context's stream, no host synchronisation.
extern "C" int nsb_eval_points_dev(nsb_ctx* ctx, int stage, int Pn, const float* d_pts, float* d_raw) {
    cudaSetDevice(ctx->device);
    if (stage < 0 || stage > 3) return fail(ctx, "bad stage %d", stage);
    const nsb_config& c = ctx->cfg;
    const size_t cap_pts = (size_t)ctx->cap * (c.n_samples + c.n_surface);
    if (refresh_images(ctx, 0xE)) return -1;
    for (size_t o = 0; o < (size_t)Pn; o += cap_pts) {
        const int m = (int)std::min(cap_pts, (size_t)Pn - o);
        CK(cudaMemcpyAsync(ctx->pts, d_pts + 3 * o, 12 * (size_t)m, cudaMemcpyDeviceToDevice, ctx->stream));
        if (eval_chunk(ctx, stage, m, d_raw + 4 * o, nullptr)) return -1;
    }
    return 0;
}
// Mesh-extraction query (nice_slam.yaml meshing: eval_points over a regular lattice): the nx*ny*nz lattice points between lo and hi
// are generated on the device, evaluated chunk by chunk and masked by the bound; the outputs are HOST arrays (either may be NULL):
// raw4 (n,4) and / or occ (n) = the occupancy channel, point order q = (j*nx + i)*nz + k (numpy.meshgrid(x,y,z) ravelled).
extern "C" int nsb_eval_lattice(nsb_ctx* ctx, int stage, int nx, int ny, int nz, const float* lo3, const float* hi3, float* raw4, float* occ) {
    cudaSetDevice(ctx->device);
    if (stage < 0 || stage > 3) return fail(ctx, "bad stage %d", stage);
    if (nx < 1 || ny < 1 || nz < 1) return fail(ctx, "bad lattice %d x %d x %d", nx, ny, nz);
    const nsb_config& c = ctx->cfg;
    const size_t cap_pts = (size_t)ctx->cap * (c.n_samples + c.n_surface);
    const long long total = (long long)nx * ny * nz;
    LatticeParams L; L.nx = nx; L.ny = ny; L.nz = nz;
    for (int a = 0; a < 3; ++a) { L.lo[a] = lo3 ? lo3[a] : c.bound[a][0]; L.hi[a] = hi3 ? hi3[a] : c.bound[a][1]; }
    if (refresh_images(ctx, 0xE)) return -1;
    for (long long o = 0; o < total; o += (long long)cap_pts) {
        const int m = (int)std::min<long long>((long long)cap_pts, total - o);
        k_lattice_points<<<cdiv(m, 256), 256, 0, ctx->stream>>>(L, o, m, ctx->pts); ctx->launches++;
        if (eval_chunk(ctx, stage, m, raw4 ? ctx->g_raw : nullptr, occ ? ctx->o_w : nullptr)) return -1;      // o_w: (cap, S) scratch
        if (raw4) CK(cudaMemcpyAsync(raw4 + 4 * o, ctx->g_raw, 16 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        if (occ) CK(cudaMemcpyAsync(occ + o, ctx->o_w, 4 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int nsb_render_vjp(nsb_ctx* ctx, int stage, int n, const float* rays_d, const float* rays_o, const float* gt_depth,
                              const float* g_rgb, const float* g_depth, const float* g_var, int flags, float* d_rays_d, float* d_rays_o) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (n > ctx->cap) return fail(ctx, "n %d exceeds max_rays %d", n, ctx->cap);
    if (stage < 1 || stage > 3) return fail(ctx, "vjp supports stages middle/fine/color");
    const bool hd = gt_depth != nullptr;
    if (upload_rays(ctx, n, rays_d, rays_o, gt_depth)) return -1;
    if (render_prepare_stats(ctx, n, hd)) return -1;
    if (zero_grads(ctx)) return -1;
    if (refresh_images(ctx, 0xE)) return -1;
    if (run_forward(ctx, stage, 0, n, hd, nullptr, ctx->rstats, NO_ITER, false, true, (flags & 2) != 0)) return -1;
    CK(cudaMemcpyAsync(ctx->g_rgb, g_rgb, 12 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->g_depth, g_depth, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->g_var, g_var, 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    if (run_backward(ctx, stage, 0, n, nullptr, ctx->rstats, NO_ITER, flags, true)) return -1;
    if (flags & 4) {
        std::vector<float> h(6 * (size_t)n);
        CK(cudaMemcpyAsync(h.data(), ctx->d_rays, 24 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) {
            if (d_rays_o) d_rays_o[3 * i + a] = h[6 * (size_t)i + a];
            if (d_rays_d) d_rays_d[3 * i + a] = h[6 * (size_t)i + 3 + a];
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Peer-memory mode: did a barrier of k_reduce_adam give up (a peer never launched its kernel)?  Called at the synchronising
// entry points of the mapping loop.
static int p2p_check(nsb_ctx* ctx) {
    if (!ctx->p2p) return 0;
    uint32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, ctx->p2p_flags + 18, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (flag) return fail(ctx, "peer-memory optimiser step: a barrier timed out after %llu ms (a rank did not launch its iteration; "
                               "all ranks must issue the same call sequence)", ctx->p2p_timeout_ns / 1000000ull);
    return 0;
}

// ---- Adam -----------------------------------------------------------------------------------------------------------
// color_pristine: no colour gradient has been produced since the optimiser was created (geometry iterations before the first
// colour iteration): gradient, m and v of the colour grid / colour decoder are all exactly zero, the Adam update is the identity
// (p - step * 0 / (0 + eps) = p), so those segments are left out of the launch.
// The step count is NOT a launch parameter: the kernel reads it from the device iteration state (IterRef) and looks the bias
// corrections up in the tables ensure_ring() wrote, so the same launch serves every iteration (CUDA-graph replay).
static void build_adam(nsb_ctx* ctx, AdamParams& A, const float lr_group[6], bool dec_fine, bool dec_color, int n_cam_floats, bool color_pristine) {
    memset(&A, 0, sizeof A);
    A.param = ctx->param; A.grad = ctx->grad; A.m = ctx->m; A.v = ctx->v;
    const double b1 = 0.9, b2 = 0.999;
    A.beta1 = (float)b1; A.beta2 = (float)b2; A.om_beta1 = (float)(1.0 - b1); A.om_beta2 = (float)(1.0 - b2); A.eps = 1e-8f;
    A.bc1 = 1.0; A.bc2_sqrt = 1.0f; A.bc1_tab = ctx->bc1_tab; A.bc2s_tab = ctx->bc2s_tab; A.tab_n = ctx->tab_n;
    A.it.state = ctx->it_state; A.it.ring = ctx->ring;
    A.grad_scale = 1.0f;
    int k = 0;
    auto add = [&](size_t begin, size_t n, float lr, const uint8_t* mask, int active) {
        AdamSegment& s = A.seg[k++]; s.begin = (int)begin; s.end = (int)(begin + pad32(n)); s.lr = lr; s.mask = mask; s.active = active;
    };
    if (ctx->coarse_map) add(ctx->off_grid[0], ctx->nvox[0] * CDIM, lr_group[1], ctx->map_no_mask ? nullptr : ctx->vmask[0], 1);                                     // group 1 = coarse (coarse mapper only)
    else for (int l = 1; l < (color_pristine ? 3 : 4); ++l) add(ctx->off_grid[l], ctx->nvox[l] * CDIM, lr_group[1 + l], ctx->map_no_mask ? nullptr : ctx->vmask[l], 1);   // groups 2,3,4 = middle, fine, color
    if (!ctx->coarse_map) {
        add(ctx->off_dec[2], ctx->dec_n[2], lr_group[0], nullptr, dec_fine ? 1 : 0);
        if (!color_pristine) add(ctx->off_dec[3], ctx->dec_n[3], lr_group[0], nullptr, dec_color ? 1 : 0);
    }
    if (n_cam_floats > 0) add(ctx->off_cam, n_cam_floats, lr_group[5], nullptr, 1);
    add(ctx->off_tail, 32, 0.f, nullptr, 0);
    A.n_seg = k;
    A.cum4[0] = 0;
    for (int s = 0; s < k; ++s) A.cum4[s + 1] = A.cum4[s] + (A.seg[s].end - A.seg[s].begin) / 4;
    A.loss_dst = ctx->stats; A.loss_idx4 = (int)(ctx->off_tail / 4);
}

static int run_adam(nsb_ctx* ctx, const float lr_group[6], bool dec_fine, bool dec_color, int n_cam_floats, bool color_pristine) {
    Timer t(ctx, T_ADAM);
    AdamParams A; build_adam(ctx, A, lr_group, dec_fine, dec_color, n_cam_floats, color_pristine);
    k_adam<<<cdiv(A.cum4[A.n_seg], 256 * ADAM_VEC), 256, 0, ctx->stream>>>(A); ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// Multi-GPU optimiser step over peer memory (p2p_kernels.cuh): floats [begin, end) of the arena are exchanged and stepped.
static int run_reduce_adam(nsb_ctx* ctx, const float lr_group[6], bool dec_fine, bool dec_color, int n_cam_floats, bool color_pristine,
                           size_t begin, size_t end) {
    Timer t(ctx, T_COMM);
    P2PParams P; memset(&P, 0, sizeof P);
    build_adam(ctx, P.A, lr_group, dec_fine, dec_color, n_cam_floats, color_pristine);
    for (int w = 0; w < ctx->world; ++w) { P.peer_grad[w] = ctx->peer_grad[w]; P.peer_param[w] = ctx->peer_param[w]; P.peer_flags[w] = ctx->peer_flags[w]; }
    P.rank = ctx->rank; P.world = ctx->world; P.lo4 = (int)(begin / 4); P.hi4 = (int)(end / 4); P.loss4 = (int)(ctx->off_tail / 4);
    P.stats_base = ctx->stats; P.timeout_ns = ctx->p2p_timeout_ns;
    k_reduce_adam<<<ctx->n_sm * 2, 256, 0, ctx->stream>>>(P); ctx->launches++;
    CK(cudaGetLastError());
    // every peer has finished reading this rank's gradients when the kernel retires: clear them for the next iteration
    CK(cudaMemsetAsync(ctx->grad + ctx->off_train, 0, (ctx->arena_n - ctx->off_train) * 4, ctx->stream));
    return 0;
}

// ---- mapping ----------------------------------------------------------------------------------------------------------
static int stage_of_iter(const nsb_ctx* ctx, int it, int n_iters) {   // Mapper.cpp:351-358
    const nsb_config& c = ctx->cfg;
    if (ctx->coarse_map) return NSB_COARSE;                                         // Mapper.cpp:351-352
    const float mr = ctx->map_zero_ratios ? 0.f : c.middle_iter_ratio, fr = ctx->map_zero_ratios ? 0.f : c.fine_iter_ratio;   // color_refine: Mapper.cpp:508-509
    if (it <= (int)((float)n_iters * mr)) return NSB_MIDDLE;
    if (it <= (int)((float)n_iters * fr)) return c.second_stage;
    return NSB_COLOR;
}

extern "C" int nsb_mapping_begin(nsb_ctx* ctx, int n_frames, const int* slots, int n_iters, float lr_factor) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    return nsb_mapping_begin_ex(ctx, n_frames, slots, n_iters, lr_factor, 0u, 0);
}
extern "C" int nsb_mapping_begin_ba(nsb_ctx* ctx, int n_frames, const int* slots, int n_iters, float lr_factor, uint32_t ba_mask) {
    cudaSetDevice(ctx->device);
    return nsb_mapping_begin_ex(ctx, n_frames, slots, n_iters, lr_factor, ba_mask, 0);
}

static uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
// Everything a captured iteration bakes into its launch parameters besides nsb_config (fixed per context).
static uint64_t graph_signature(const nsb_ctx* ctx) {
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(h, &ctx->map_frames, sizeof ctx->map_frames); h = fnv1a(h, ctx->map_slots, sizeof(int) * ctx->map_frames);
    h = fnv1a(h, &ctx->map_lr_factor, sizeof(float)); h = fnv1a(h, &ctx->map_ba_mask, sizeof(uint32_t));
    h = fnv1a(h, &ctx->coarse_map, sizeof ctx->coarse_map); h = fnv1a(h, &ctx->map_fix_color, sizeof ctx->map_fix_color);
    h = fnv1a(h, &ctx->map_no_mask, sizeof ctx->map_no_mask);
    h = fnv1a(h, &ctx->rank, sizeof(int)); h = fnv1a(h, &ctx->world, sizeof(int)); h = fnv1a(h, &ctx->p2p, sizeof(bool));
    h = fnv1a(h, ctx->vmask, sizeof ctx->vmask); h = fnv1a(h, &ctx->idx_pool, sizeof ctx->idx_pool);
    h = fnv1a(h, &ctx->pool_iters, sizeof(int)); h = fnv1a(h, &ctx->pool_n, sizeof(int));
    h = fnv1a(h, &ctx->stash, sizeof ctx->stash); h = fnv1a(h, &ctx->grad_snap, sizeof ctx->grad_snap);
    h = fnv1a(h, &ctx->cfg.mapping_pixels, sizeof(int)); h = fnv1a(h, &ctx->ar_mode, sizeof(int));
    return h;
}

// ba_mask bit f: frame f's pose joins the optimisation as a 7-vector (Mapper.cpp:305-329: every frame of optimize_frame but
// the oldest one when BA is on); its lr is BA_cam_lr in the colour stage and 0 before (Mapper.cpp:366-368).
// flags: NSB_MAP_COARSE = the coarse mapper (Mapper.cpp:335-338,351-352,450-453: stage "coarse", only grid_coarse is optimised);
// NSB_MAP_FIX_COLOR = keep the colour decoder fixed for this call (color_refine, Mapper.cpp:505-513 sets fix_color);
// NSB_MAP_NO_FRUSTUM = no frustum feature selection for this call (color_refine: frustum_feature_selection = false);
// NSB_MAP_ZERO_RATIOS = middle_iter_ratio = fine_iter_ratio = 0 for this call (color_refine, Mapper.cpp:508-509).
extern "C" int nsb_mapping_begin_ex(nsb_ctx* ctx, int n_frames, const int* slots, int n_iters, float lr_factor, uint32_t ba_mask, int flags) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (wait_uploads(ctx)) return -1;
    if (n_frames < 1 || n_frames > MAX_OPT_FRAMES) return fail(ctx, "n_frames %d out of range", n_frames);
    if (n_frames > ctx->cfg.max_frames) return fail(ctx, "n_frames %d exceeds max_frames %d", n_frames, ctx->cfg.max_frames);
    const int pix = ctx->cfg.mapping_pixels / n_frames;   // Mapper.cpp:223
    if (pix * n_frames > ctx->cap) return fail(ctx, "mapping_pixels %d exceeds max_rays %d", ctx->cfg.mapping_pixels, ctx->cap);
    if ((flags & NSB_MAP_COARSE) && ctx->world > 1) return fail(ctx, "the coarse mapper runs on one GPU (its gradient is 15 KB)");
    for (int f = 0; f < n_frames; ++f) { if (slots[f] < 0 || slots[f] >= ctx->cfg.max_frames) return fail(ctx, "bad slot %d", slots[f]); ctx->map_slots[f] = slots[f]; }
    ctx->map_frames = n_frames; ctx->map_iters = n_iters; ctx->map_step = 0; ctx->map_lr_factor = lr_factor; ctx->map_color_touched = false;
    ctx->coarse_map = (flags & NSB_MAP_COARSE) != 0;
    ctx->map_fix_color = ctx->cfg.fix_color || (flags & NSB_MAP_FIX_COLOR) != 0;
    ctx->map_no_mask = (flags & NSB_MAP_NO_FRUSTUM) != 0;   // explicit masks (nsb_set_voxel_mask) are ignored too
    ctx->map_zero_ratios = (flags & NSB_MAP_ZERO_RATIOS) != 0;
    if (ensure_ring(ctx, std::max(LOSS_RING_MIN, n_iters + 1))) return -1;   // every step of this optimize_map keeps its loss slot
    // a fresh torch::optim::Adam is constructed per optimize_map (Mapper.cpp:330): state starts at zero
    CK(cudaMemsetAsync(ctx->m, 0, ctx->arena_n * 4, ctx->stream)); CK(cudaMemsetAsync(ctx->v, 0, ctx->arena_n * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->grad, 0, ctx->arena_n * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->stats, 0, 4 * (size_t)ctx->ring * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->it_state, 0, sizeof(int), ctx->stream));             // step count; the index-pool cursor [1] carries on
    CK(cudaMemsetAsync(ctx->it_state + 2, 0, sizeof(int), ctx->stream));
    ctx->map_ba_mask = ba_mask & ((n_frames >= 32) ? 0xffffffffu : ((1u << n_frames) - 1u));
    if (ctx->coarse_map) ctx->map_ba_mask = 0;
    if (ctx->map_ba_mask) {   // camera_tensor = get_tensor_from_camera(c2w) per optimised frame (Mapper.cpp:322-325)
        std::vector<float> hp(12 * (size_t)ctx->cfg.max_frames), cams(8 * (size_t)n_frames, 0.f);
        CK(cudaMemcpyAsync(hp.data(), ctx->f_pose, hp.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int f = 0; f < n_frames; ++f) {
            float m16[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1};
            memcpy(m16, hp.data() + 12 * slots[f], 12 * sizeof(float));
            nsb_get_tensor_from_camera(m16, cams.data() + 8 * f);
        }
        CK(cudaMemcpyAsync(ctx->param + ctx->off_cam, cams.data(), cams.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (!ctx->map_no_mask && ctx->cfg.frustum_feature_selection) {   // Mapper.cpp:226-290: masks from the current frame (the last of optimize_frame)
        for (int l = ctx->coarse_map ? 0 : 1; l < (ctx->coarse_map ? 1 : 4); ++l) if (nsb_frustum_mask(ctx, slots[n_frames - 1], nullptr, l, nullptr, 1)) return -1;
    }
    if (!ctx->map_fix_color && !ctx->coarse_map && ctx->wg_stash) {   // make sure the stash exists before the hot loop
        if (ensure_stash(ctx, (size_t)cdiv(pix * n_frames, ctx->world) * (ctx->cfg.n_samples + ctx->cfg.n_surface) + 64)) return -1;
    }
    const uint64_t sig = graph_signature(ctx);
    if (sig != ctx->graph_sig) { CK(cudaStreamSynchronize(ctx->stream)); drop_graphs(ctx); ctx->graph_sig = sig; }
    return 0;
}

// Ray order of a mapping iteration: identity on one GPU, rank-major interleave when the frames divide evenly over the ranks.
static RayOrder map_order(const nsb_ctx* ctx, int n, int pix) {
    RayOrder o; o.world = 1; o.per = n; o.pix = pix; o.ppr = pix;
    if (ctx->world > 1 && pix % ctx->world == 0 && n % ctx->world == 0) { o.world = ctx->world; o.per = n / ctx->world; o.ppr = pix / ctx->world; }
    return o;
}

// Host-visible copy of the ray-order rule (tests): which reference batch element, and which frame, ray i of the rendered order is.
extern "C" int nsb_ray_order_source(int world, int n_rays, int pix_per_frame, int i, int* frame) {
    RayOrder o; o.world = 1; o.per = n_rays; o.pix = pix_per_frame; o.ppr = pix_per_frame;
    if (world > 1 && pix_per_frame % world == 0 && n_rays % world == 0) { o.world = world; o.per = n_rays / world; o.ppr = pix_per_frame / world; }
    if (frame) *frame = o.frame(i, pix_per_frame);
    return o.source(i);
}

// One joint iteration in its launch-parameter-free form: nothing below depends on the iteration number except through the
// device iteration state, so the same sequence is either enqueued kernel by kernel or captured once and replayed as a graph.
struct IterPlan { int stage; bool use_color, pristine, use_pool; };

// Which lazy set (refresh_images) this iteration rebuilds at its start: colour iterations of the stash path that step the colour decoder.
static int iter_lazy(const nsb_ctx* ctx, const IterPlan& pl) {
    if (pl.stage != NSB_COLOR || pl.pristine) return 0;
    if (ctx->cfg.stage_lr[NSB_COLOR][0] * ctx->map_lr_factor == 0.f) return 0;
    return lazy_color_images(ctx);
}

static int enqueue_iteration(nsb_ctx* ctx, const IterPlan& pl) {
    const nsb_config& c = ctx->cfg;
    const int pix = c.mapping_pixels / ctx->map_frames, n = pix * ctx->map_frames;
    const int stage = pl.stage;
    const bool coarse = ctx->coarse_map;
    const int render_stage = coarse ? NSB_COARSE : NSB_COLOR;   // Mapper.cpp:430 renders "color" whatever the stage; the coarse mapper "coarse"
    IterRef it; it.state = ctx->it_state; it.ring = ctx->ring;
    float* stats = ctx->stats;
    const bool pool = pl.use_pool;
    // the previous colour iteration stepped the colour decoder: its images are rebuilt now, beside the sampling kernels (joined in run_forward)
    const int lz = iter_lazy(ctx, pl);
    if (lz && fork_rebuild(ctx, lz)) return -1;
    {
        Timer t(ctx, T_SAMPLE);
        SampleParams P; fill_sample_params(ctx, P, n, 0, c.H, 0, c.W, stats, 1);
        P.it = it; P.cam_mask = ctx->map_ba_mask;
        if (pool) { P.pool = ctx->idx_pool; P.pool_iters = ctx->pool_iters; }
        P.order = map_order(ctx, n, pix);
        for (int f = 0; f < ctx->map_frames; ++f) P.slots[f] = ctx->map_slots[f];
        P.n_frames = ctx->map_frames; P.pix_per_frame = pix;
        k_sample<<<cdiv(n, 128), 128, 0, ctx->stream>>>(P); ctx->launches++;
        if (c.dist_norm == NSB_DISTNORM_REFERENCE) { k_dirnorm_ref<<<cdiv(n, 256), 256, 0, ctx->stream>>>(ctx->rays_d, ctx->valid, n, stats, it); ctx->launches++; }
        CK(cudaGetLastError());
    }
    // this rank's slice of the (already filtered, batch-global) ray list (rank-major order, see RayOrder)
    const int per = cdiv(n, ctx->world), off = std::min(n, ctx->rank * per), nl = std::max(0, std::min(per, n - off));
    const bool use_color = pl.use_color;
    const bool train_color = use_color && !ctx->map_fix_color && !coarse;
    if (nl > 0) {
        // the coarse mapper renders without depth guidance (upstream passes gt_depth = None for it: 32 stratified samples)
        if (run_forward(ctx, render_stage, off, nl, !coarse, ctx->valid, stats, it, false, true, train_color, true)) return -1;
        {
            // composite + loss (Mapper.cpp:435-442) + composite backward fused: the cotangents are local to each ray
            Timer t(ctx, T_COMP);
            const int S = ctx->last_S;
            CompositeParams Q; memset(&Q, 0, sizeof Q);
            Q.rays_o = ctx->rays_o + 3 * off; Q.rays_d = ctx->rays_d + 3 * off; Q.z = ctx->z + (size_t)off * S; Q.valid = ctx->valid + off;
            Q.raw_rgb = ctx->raw_rgb + 4 * (size_t)off * S; for (int k = 0; k < 3; ++k) Q.occ[k] = ctx->occ[k] + (size_t)off * S;
            Q.stats = stats; Q.it = it; Q.zero_ctr = ctx->tile_ctr + 4;
            Q.bnd = ctx->bnd; Q.n = nl; Q.S = S; Q.stage = render_stage; Q.occupancy = c.occupancy; Q.dist_norm = c.dist_norm;
            Q.rgb = ctx->o_rgb + 3 * off; Q.depth = ctx->o_depth + off; Q.var = ctx->o_var + off; Q.weights = nullptr;
            Q.g_raw = ctx->g_raw + 4 * (size_t)off * S;
            if (ctx->map_ba_mask) {   // bundle adjustment: ray gradients (d L / d rays_o, rays_d) are accumulated for the pose chain
                CK(cudaMemsetAsync(ctx->d_rays + 6 * (size_t)off, 0, 6 * (size_t)nl * 4, ctx->stream));
                Q.d_rays = ctx->d_rays + 6 * (size_t)off;
            }
            // the loss rides in the tail of the gradient arena so that the all-reduce sums it too
            k_composite_map<<<cdiv(nl * 32, 256), 256, 0, ctx->stream>>>(Q, ctx->gt_depth + off, ctx->gt_color + 3 * off, use_color ? 1 : 0,
                                                                       c.mapping_w_color_loss, ctx->grad + ctx->off_tail);
            ctx->launches++;
            CK(cudaGetLastError());
        }
        const int flags = coarse ? 1 : (1 | (ctx->map_fix_color ? 0 : 2) | (ctx->map_ba_mask ? 4 : 0));
        ctx->ar_request = ctx->world > 1 && ctx->ar_mode == 2 && !ctx->p2p;
        const int rb = run_backward(ctx, render_stage, off, nl, ctx->valid, stats, it, flags, use_color, true);
        ctx->ar_request = false;
        if (rb) return -1;
        if (ctx->map_ba_mask) {   // chain to (q, t) of every optimised frame; other ranks' partial sums arrive through the all-reduce
            Timer t(ctx, T_COMP);
            PoseGradParams G; memset(&G, 0, sizeof G);
            G.d_rays = ctx->d_rays; G.idx = ctx->idx; G.it = it; G.n_total = n;
            if (pool) { G.pool = ctx->idx_pool; G.pool_iters = ctx->pool_iters; }
            G.valid = ctx->valid; G.cams = ctx->param + ctx->off_cam; G.cam_mask = ctx->map_ba_mask;
            G.pix_per_frame = pix; G.n_frames = ctx->map_frames; G.lo = off; G.hi = off + nl; G.order = map_order(ctx, n, pix);
            G.H0 = 0; G.W0 = 0; G.Wc = c.W; G.raydir = c.raydir; G.fx = c.fx; G.fy = c.fy; G.cx = c.cx; G.cy = c.cy;
            G.g_cams = ctx->grad + ctx->off_cam;
            k_pose_grad<<<ctx->map_frames, 1024, 0, ctx->stream>>>(G); ctx->launches++;
            CK(cudaGetLastError());
        }
    }
    if (join_rebuild(ctx)) return -1;   // a rank without rays has not joined in run_forward
    float lr[6];
    for (int g = 0; g < 5; ++g) lr[g] = c.stage_lr[stage][g] * ctx->map_lr_factor;   // Mapper.cpp:360-364
    lr[5] = (ctx->map_ba_mask && stage == NSB_COLOR) ? c.BA_cam_lr : 0.f;   // Mapper.cpp:366-368 (not scaled by lr_factor)
    const bool pristine = pl.pristine;   // colour grid / decoder: gradient, m, v all exactly zero so far
    const int n_cam = ctx->map_ba_mask ? 8 * ctx->map_frames : 0;
    const bool dec_color = !ctx->map_fix_color;
    // the gradient of this iteration for the parity tests (nsb_mapping_capture_grads), before the optimiser clears it
    if (ctx->capture_grads) CK(cudaMemcpyAsync(ctx->grad_snap, ctx->grad, ctx->arena_n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    // geometry iteration before any colour iteration, no BA: colour-grid, decoder and camera gradients are exact zeros on every
    // rank (Mapper.cpp:435-442 adds the colour term only in stage "color"), so only [loss | grid_middle | grid_fine] is exchanged
    const bool prefix_only = pristine && !ctx->map_ba_mask && ctx->ar_mode != 0;
    if (ctx->world > 1 && ctx->p2p) {
        // peer-memory mode: reduce-scatter + Adam + all-gather in one kernel (p2p_kernels.cuh); the summed loss lands in the parameter arena
        if (ctx->map_ba_mask) CK(cudaMemcpyAsync(ctx->cam_grad_last, ctx->grad + ctx->off_cam, 8 * ctx->map_frames * 4, cudaMemcpyDeviceToDevice, ctx->stream));   // this rank's partial sums
        if (run_reduce_adam(ctx, lr, !c.fix_fine, dec_color, n_cam, pristine, ctx->off_train, prefix_only ? ctx->off_grid[3] : ctx->arena_n)) return -1;
    } else {
        if (ctx->world > 1) {
            Timer t(ctx, T_COMM);
            if (nl > 0 && ctx->ar_overlapped) {   // colour iteration: grids went out under the wgrad kernel, the small remainder follows
                CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
                if (allreduce_range(ctx, ctx->off_dec[2], ctx->arena_n, ctx->stream)) return -1;
            } else if (prefix_only) {
                if (allreduce_range(ctx, ctx->off_train, ctx->off_grid[3], ctx->stream)) return -1;
            } else {
                if (allreduce_range(ctx, ctx->off_train, ctx->arena_n, ctx->stream)) return -1;
            }
            ctx->ar_overlapped = false;
        }
        if (ctx->map_ba_mask) CK(cudaMemcpyAsync(ctx->cam_grad_last, ctx->grad + ctx->off_cam, 8 * ctx->map_frames * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        // every parameter in the optimiser receives a (possibly zero) gradient each iteration, so one step counter serves all groups;
        // the Adam kernel also moves the (all-reduced) loss out of the gradient arena into the step's statistics slot
        if (run_adam(ctx, lr, !c.fix_fine, dec_color, n_cam, pristine)) return -1;
    }
    // the colour decoder has just been stepped.  Stash colour iterations (lz) leave its images stale: the next one rebuilds what it reads at
    // its start and the host marks the rest (mark_color_stale).  Any other iteration that moves it (a geometry stage configured with a
    // decoder learning rate, the stash-free path) rebuilds all of its images right away.
    if (!lz && dec_color && !coarse && !pristine && lr[0] != 0.f) { if (refresh_images(ctx, 0, 1 << 3)) return -1; }
    return 0;
}

// Host side of the lazy rebuild: an iteration that steps the colour decoder leaves its composed / tcgen05 images stale (this runs per
// iteration on the host, also when the iteration itself is a graph replay).
static void mark_color_stale(nsb_ctx* ctx, const IterPlan& pl) {
    if (!iter_lazy(ctx, pl)) return;
    ctx->wimg_dirty |= 8; ctx->wimg_cmp_dirty |= 8; ctx->comp_dirty |= 8;
}

extern "C" int nsb_mapping_iter_async(nsb_ctx* ctx, int iter, const int64_t* idx) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    const nsb_config& c = ctx->cfg;
    if (ctx->map_frames < 1) return fail(ctx, "nsb_mapping_begin was not called");
    if (wait_uploads(ctx)) return -1;
    const int pix = c.mapping_pixels / ctx->map_frames, n = pix * ctx->map_frames;
    IterPlan pl;
    pl.stage = stage_of_iter(ctx, iter, ctx->map_iters);
    pl.use_color = pl.stage == NSB_COLOR;
    if (pl.use_color) ctx->map_color_touched = true;
    pl.pristine = !ctx->map_color_touched;
    pl.use_pool = !idx && ctx->idx_pool && ctx->pool_n == n;
    // pixel indices of this iteration: caller's, the resident pool's next row (selected on the device), or the mt19937 stream
    if (idx) CK(cudaMemcpyAsync(ctx->idx, idx, n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    else if (!pl.use_pool) {
        ctx->h_idx.resize(n);
        for (int f = 0; f < ctx->map_frames; ++f) draw_indices(ctx, pix, (int64_t)c.H * c.W, ctx->h_idx.data() + (size_t)f * pix);   // one randint per frame (Mapper.cpp:404)
        CK(cudaMemcpyAsync(ctx->idx, ctx->h_idx.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    // no-op unless a decoder was replaced since the last iteration; a stash colour iteration rebuilds the colour decoder's images itself
    // (the previous one left them stale on purpose), so they are skipped here
    {
        const int lz = iter_lazy(ctx, pl);
        if (refresh_images(ctx, lz ? 0x6 : 0xE, 0, 0, lz ? 8 : 0)) return -1;
    }
    ctx->map_step++;
    const bool graph_ok = ctx->use_graph && !ctx->profiling && (ctx->world == 1 || ctx->p2p);
    if (!graph_ok) { const int rc = enqueue_iteration(ctx, pl); mark_color_stale(ctx, pl); return rc; }
    const uint64_t key = (uint64_t)pl.stage | (pl.use_color ? 8u : 0u) | (pl.pristine ? 16u : 0u) | (pl.use_pool ? 32u : 0u) | (ctx->capture_grads ? 64u : 0u);
    auto g = ctx->graphs.find(key);
    if (g == ctx->graphs.end()) {
        const int64_t l0 = ctx->launches;
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
        ctx->capturing = true;
        const int rc = enqueue_iteration(ctx, pl);
        ctx->capturing = false;
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        if (rc != 0 || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return rc != 0 ? -1 : fail(ctx, "graph capture of the mapping iteration failed: %s", cudaGetErrorString(e));
        }
        IterGraph ig; ig.launches = (int)(ctx->launches - l0); ctx->launches = l0;
        const cudaError_t ei = cudaGraphInstantiate(&ig.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return fail(ctx, "cudaGraphInstantiate: %s", cudaGetErrorString(ei));
        g = ctx->graphs.emplace(key, ig).first;
    }
    CK(cudaGraphLaunch(g->second.exec, ctx->stream));
    ctx->launches += g->second.launches;
    mark_color_stale(ctx, pl);
    return 0;
}

// Bundle-adjustment write-back (Mapper.cpp:467-489): est_c2w of every optimised frame <- get_camera_from_tensor(camera_tensor).
// cam7s_out ([n_frames][7], may be NULL) receives the optimised 7-vectors (frames outside the mask: their initial pose).
extern "C" int nsb_mapping_end(nsb_ctx* ctx, float* cam7s_out) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (!ctx->map_ba_mask) { CK(cudaStreamSynchronize(ctx->stream)); return p2p_check(ctx); }
    const int nf = ctx->map_frames;
    std::vector<float> cams(8 * (size_t)nf);
    CK(cudaMemcpyAsync(cams.data(), ctx->param + ctx->off_cam, cams.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<float> RT(12 * (size_t)nf);
    for (int f = 0; f < nf; ++f) {
        if (cam7s_out) memcpy(cam7s_out + 7 * f, cams.data() + 8 * f, 7 * sizeof(float));
        if (!((ctx->map_ba_mask >> f) & 1u)) continue;
        nsb_get_camera_from_tensor(cams.data() + 8 * f, RT.data() + 12 * f);
        CK(cudaMemcpyAsync(ctx->f_pose + 12 * ctx->map_slots[f], RT.data() + 12 * f, 12 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return p2p_check(ctx);
}
// d L / d (q, t) of the optimised frames at the last bundle-adjustment iteration ([n_frames][7]; zeros outside the mask).
extern "C" int nsb_mapping_cam_grads(nsb_ctx* ctx, float* g7s) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    std::vector<float> h(8 * (size_t)ctx->map_frames);
    CK(cudaMemcpyAsync(h.data(), ctx->cam_grad_last, h.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int f = 0; f < ctx->map_frames; ++f) for (int k = 0; k < 7; ++k) g7s[7 * f + k] = ((ctx->map_ba_mask >> f) & 1u) ? h[8 * f + k] : 0.f;
    return 0;
}
extern "C" int nsb_get_frame_pose(nsb_ctx* ctx, int slot, float* c2w12) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (slot < 0 || slot >= ctx->cfg.max_frames) return fail(ctx, "bad frame slot %d", slot);
    CK(cudaMemcpyAsync(c2w12, ctx->f_pose + 12 * slot, 12 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Pre-load pixel indices for n_iters iterations ([n_iters][n] int64, host) so that nsb_mapping_iter(idx = NULL) runs
// with every input already resident in HBM; rows are consumed in order, wrapping around (the cursor lives on the device and is
// advanced by the optimiser kernel).  NULL clears the pool.
extern "C" int nsb_mapping_set_index_pool(nsb_ctx* ctx, const int64_t* host_idx, int n_iters, int n) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->idx_pool) { cudaFree(ctx->idx_pool); ctx->idx_pool = nullptr; ctx->pool_iters = ctx->pool_n = 0; }
    drop_graphs(ctx); ctx->graph_sig = 0;
    if (!host_idx) return 0;
    CK(dalloc(&ctx->idx_pool, (size_t)n_iters * n));
    CK(cudaMemcpyAsync(ctx->idx_pool, host_idx, (size_t)n_iters * n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->it_state + 1, 0, sizeof(int), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->pool_iters = n_iters; ctx->pool_n = n;
    return 0;
}

// Losses / inside counts of steps [first, first + n): at most two contiguous copies out of the ring.
extern "C" int nsb_mapping_losses(nsb_ctx* ctx, int first, int n, float* losses, int* n_inside) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (n <= 0) return 0;
    if (n > ctx->ring) return fail(ctx, "the loss ring keeps the last %d steps, asked for %d", ctx->ring, n);
    constexpr int H_STATS_STEPS = 256;
    std::vector<float> big;
    float* h;
    if (n <= H_STATS_STEPS) {
        if (!ctx->h_stats) CK(cudaHostAlloc((void**)&ctx->h_stats, 16 * H_STATS_STEPS, cudaHostAllocDefault));
        h = ctx->h_stats;
    } else { big.resize(4 * (size_t)n); h = big.data(); }
    const int s0 = ((first % ctx->ring) + ctx->ring) % ctx->ring, n0 = std::min(n, ctx->ring - s0);
    CK(cudaMemcpyAsync(h, ctx->stats + 4 * (size_t)s0, 16 * (size_t)n0, cudaMemcpyDeviceToHost, ctx->stream));
    if (n0 < n) CK(cudaMemcpyAsync(h + 4 * (size_t)n0, ctx->stats, 16 * (size_t)(n - n0), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; ++i) {
        if (losses) losses[i] = h[4 * i + 3];
        if (n_inside) { int v; memcpy(&v, &h[4 * i + 1], 4); n_inside[i] = v; }
    }
    return p2p_check(ctx);
}

extern "C" int nsb_mapping_iter(nsb_ctx* ctx, int iter, const int64_t* idx, float* loss) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (nsb_mapping_iter_async(ctx, iter, idx)) return -1;
    if (loss) return nsb_mapping_losses(ctx, ctx->map_step - 1, 1, loss, nullptr);
    return 0;
}

extern "C" int nsb_optimize_map(nsb_ctx* ctx, int n_frames, const int* slots, int n_iters, float lr_factor, float* losses) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (nsb_mapping_begin(ctx, n_frames, slots, n_iters, lr_factor)) return -1;
    for (int it = 0; it < n_iters; ++it) if (nsb_mapping_iter_async(ctx, it, nullptr)) return -1;
    if (losses) return nsb_mapping_losses(ctx, 0, n_iters, losses, nullptr);   // the ring was sized for n_iters by nsb_mapping_begin
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Parity aid: keep a copy of every iteration's gradient arena (taken after the backward, before the exchange / optimiser step).
extern "C" int nsb_mapping_capture_grads(nsb_ctx* ctx, int on) {
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    if (on && !ctx->grad_snap) { CK(dalloc(&ctx->grad_snap, ctx->arena_n)); CK(cudaMemset(ctx->grad_snap, 0, ctx->arena_n * 4)); drop_graphs(ctx); ctx->graph_sig = 0; }
    ctx->capture_grads = on != 0;
    return 0;
}
extern "C" int nsb_get_captured_grid_grad(nsb_ctx* ctx, int level, float* host) {
    cudaSetDevice(ctx->device);
    if (level < 0 || level > 3) return fail(ctx, "bad level %d", level);
    if (!ctx->grad_snap) return fail(ctx, "nsb_mapping_capture_grads was not enabled");
    return get_cl(ctx, ctx->grad_snap + ctx->off_grid[level], level, host);
}
extern "C" int nsb_get_captured_decoder_grad(nsb_ctx* ctx, int which, float* host, int64_t n) {
    cudaSetDevice(ctx->device);
    if (!ctx->grad_snap) return fail(ctx, "nsb_mapping_capture_grads was not enabled");
    return get_dec(ctx, ctx->grad_snap, which, host, n);
}

// ---- keyframe selection (Mapper.cpp:132-196) ------------------------------------------------------------------------------
// 3x4 inverse of a rigid-or-not [A|t; 0 0 0 1] in double (Eigen's Matrix4f::inverse() upstream): [A^-1 | -A^-1 t]
static void invert_pose(const float* c2w12, float* w2c12) {
    const double a = c2w12[0], b = c2w12[1], cc = c2w12[2], d = c2w12[4], e = c2w12[5], f = c2w12[6], g = c2w12[8], h = c2w12[9], i = c2w12[10];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + cc * (d * h - e * g);
    const double inv[9] = {(e * i - f * h) / det, (cc * h - b * i) / det, (b * f - cc * e) / det,
                           (f * g - d * i) / det, (a * i - cc * g) / det, (cc * d - a * f) / det,
                           (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det};
    const double t[3] = {c2w12[3], c2w12[7], c2w12[11]};
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) w2c12[4 * r + k] = (float)inv[3 * r + k];
        w2c12[4 * r + 3] = (float)-(inv[3 * r] * t[0] + inv[3 * r + 1] * t[1] + inv[3 * r + 2] * t[2]);
    }
}

extern "C" int nsb_keyframe_selection_overlap(nsb_ctx* ctx, int cur_slot, const float* cur_c2w16, int n_kf, const float* kf_c2w16, int k_overlap,
                                              const int64_t* idx, int pixels, int n_samples, int* selected, int* n_selected, float* percent_out) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    *n_selected = 0;
    if (n_kf <= 0) return 0;
    if (wait_uploads(ctx)) return -1;
    if (cur_slot < 0 || cur_slot >= ctx->cfg.max_frames) return fail(ctx, "bad frame slot %d", cur_slot);
    if (pixels > ctx->cap) return fail(ctx, "pixels %d exceeds max_rays %d", pixels, ctx->cap);
    if (n_samples != 16) return fail(ctx, "n_samples %d unsupported (the reference hard-codes 16, Mapper.cpp:136)", n_samples);
    const nsb_config& c = ctx->cfg;
    // get_samples(0, H, 0, W, pixels, ...) on the current frame (:137): consumes the host RNG stream like the reference
    ctx->h_idx.resize(pixels);
    if (idx) memcpy(ctx->h_idx.data(), idx, pixels * sizeof(int64_t)); else draw_indices(ctx, pixels, (int64_t)c.H * c.W, ctx->h_idx.data());
    CK(cudaMemcpyAsync(ctx->idx, ctx->h_idx.data(), pixels * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (cur_c2w16) CK(cudaMemcpyAsync(ctx->f_pose + 12 * cur_slot, cur_c2w16, 12 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->rstats, 0, 16, ctx->stream));
    SampleParams S; fill_sample_params(ctx, S, pixels, 0, c.H, 0, c.W, ctx->rstats, 0);
    S.slots[0] = cur_slot; S.n_frames = 1; S.pix_per_frame = pixels;
    k_sample<<<cdiv(pixels, 128), 128, 0, ctx->stream>>>(S); ctx->launches++;
    std::vector<float> w2c(12 * (size_t)n_kf);
    for (int k = 0; k < n_kf; ++k) invert_pose(kf_c2w16 + 16 * k, w2c.data() + 12 * k);
    if (ensure_scratch(ctx, 13 * (size_t)n_kf)) return -1;
    CK(cudaMemcpyAsync(ctx->scratch_ncdhw, w2c.data(), w2c.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    OverlapParams P; memset(&P, 0, sizeof P);
    P.rays_o = ctx->rays_o; P.rays_d = ctx->rays_d; P.gt_depth = ctx->gt_depth; P.t_vals = ctx->t_surface; P.w2c = ctx->scratch_ncdhw;
    P.pixels = pixels; P.n_samples = n_samples; P.n_kf = n_kf; P.H = c.H; P.W = c.W; P.edge = 20;
    P.fx = c.fx; P.fy = c.fy; P.cx = c.cx; P.cy = c.cy; P.percent = ctx->scratch_ncdhw + 12 * (size_t)n_kf;
    k_kf_overlap<<<n_kf, 256, 0, ctx->stream>>>(P); ctx->launches++;
    CK(cudaGetLastError());
    std::vector<float> pct(n_kf);
    CK(cudaMemcpyAsync(pct.data(), P.percent, n_kf * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (percent_out) memcpy(percent_out, pct.data(), n_kf * sizeof(float));
    std::vector<int> order;
    for (int k = 0; k < n_kf; ++k) if (pct[k] > 0.f) order.push_back(k);                        // :178-179
    std::stable_sort(order.begin(), order.end(), [&](int l, int r) { return pct[l] > pct[r]; });   // :182-191 (ties: lower index first)
    const int take = std::min((int)order.size(), std::max(0, k_overlap));                       // :193-194 (intent: the first k_overlap)
    for (int k = 0; k < take; ++k) selected[k] = order[k];
    *n_selected = take;
    return 0;
}

// ---- tracking -----------------------------------------------------------------------------------------------------------
extern "C" int nsb_tracking_begin(nsb_ctx* ctx, int slot, const float* cam7) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (wait_uploads(ctx)) return -1;
    if (slot < 0 || slot >= ctx->cfg.max_frames) return fail(ctx, "bad frame slot %d", slot);
    if (ctx->cfg.tracking_pixels > ctx->cap) return fail(ctx, "tracking_pixels %d exceeds max_rays %d", ctx->cfg.tracking_pixels, ctx->cap);
    ctx->trk_slot = slot; ctx->trk_step = 0;
    CK(cudaMemcpyAsync(ctx->param + ctx->off_cam, cam7, 28, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->m + ctx->off_cam, 0, 32, ctx->stream)); CK(cudaMemsetAsync(ctx->v + ctx->off_cam, 0, 32, ctx->stream));   // fresh Adam (Tracker.cpp:103)
    CK(cudaMemsetAsync(ctx->grad + ctx->off_cam, 0, 32, ctx->stream));
    return 0;
}

extern "C" int nsb_tracking_iter(nsb_ctx* ctx, const int64_t* idx, float* loss, float* cam_grad7) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    const nsb_config& c = ctx->cfg;
    const int n = c.tracking_pixels;
    const int H0 = c.ignore_edge_H, H1 = c.H - c.ignore_edge_H, W0 = c.ignore_edge_W, W1 = c.W - c.ignore_edge_W;   // Tracker.cpp:46
    // per-iteration scratch, cleared with one memset: [0..3] batch statistics + loss, [4] survivor count, [8..20] pose-gradient partial sums
    float* stats = ctx->trk_scratch;
    int* count = reinterpret_cast<int*>(ctx->trk_scratch + 4);
    float* cam = ctx->param + ctx->off_cam;
    if (wait_uploads(ctx)) return -1;
    {
        Timer t(ctx, T_SAMPLE);
        if (idx) CK(cudaMemcpyAsync(ctx->idx, idx, n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        else {
            ctx->h_idx.resize(n);
            draw_indices(ctx, n, (int64_t)(H1 - H0) * (W1 - W0), ctx->h_idx.data());
            CK(cudaMemcpyAsync(ctx->idx, ctx->h_idx.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
        }
        CK(cudaMemsetAsync(ctx->trk_scratch, 0, 32 * 4, ctx->stream));
        SampleParams P; fill_sample_params(ctx, P, n, H0, H1, W0, W1, stats, 1);
        P.slots[0] = ctx->trk_slot; P.n_frames = 1; P.pix_per_frame = n; P.cam_mask = 1u;
        k_sample<<<cdiv(n, 128), 128, 0, ctx->stream>>>(P); ctx->launches++;
        if (c.dist_norm == NSB_DISTNORM_REFERENCE) { k_dirnorm_ref<<<cdiv(n, 256), 256, 0, ctx->stream>>>(ctx->rays_d, ctx->valid, n, stats, NO_ITER); ctx->launches++; }
        CK(cudaGetLastError());
    }
    if (refresh_images(ctx, 0xE)) return -1;
    // render (Tracker.cpp:61); with handle_dynamic the composite also compacts |gt - depth| of the surviving rays for the median (:69)
    ctx->trk_hook = c.handle_dynamic != 0; ctx->trk_count = count;
    const int rf = run_forward(ctx, NSB_COLOR, 0, n, true, ctx->valid, stats, NO_ITER, false, true, false);
    ctx->trk_hook = false;
    if (rf) return -1;
    const int S = ctx->last_S;
    {
        Timer t(ctx, T_COMP);
        if (c.handle_dynamic) { k_median<<<1, 1024, 0, ctx->stream>>>(ctx->absdiff, count, ctx->median); ctx->launches++; }
        // loss (Tracker.cpp:67-82) + composite backward in one pass over the rays
        CK(cudaMemsetAsync(ctx->d_rays, 0, 6 * (size_t)n * 4, ctx->stream));
        CompositeParams Q; memset(&Q, 0, sizeof Q);
        Q.rays_o = ctx->rays_o; Q.rays_d = ctx->rays_d; Q.z = ctx->z; Q.valid = ctx->valid;
        Q.raw_rgb = ctx->raw_rgb; for (int k = 0; k < 3; ++k) Q.occ[k] = ctx->occ[k];
        Q.stats = stats; Q.it = NO_ITER; Q.zero_ctr = ctx->tile_ctr + 4;
        Q.bnd = ctx->bnd; Q.n = n; Q.S = S; Q.stage = NSB_COLOR; Q.occupancy = c.occupancy; Q.dist_norm = c.dist_norm;
        Q.g_raw = ctx->g_raw; Q.d_rays = ctx->d_rays;
        k_composite_track<<<cdiv(n * 32, 256), 256, 0, ctx->stream>>>(Q, ctx->gt_depth, ctx->gt_color, ctx->median, c.handle_dynamic, c.use_color_in_tracking,
                                                                     c.w_color_loss, stats + 3);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    if (run_backward(ctx, NSB_COLOR, 0, n, ctx->valid, stats, NO_ITER, 4, true, true)) return -1;
    ctx->trk_step++;
    {
        // chain to (q, t) (utils.h:174-210) and the Adam step on the 7-vector (Tracker.cpp:85), fused: the last block applies it
        Timer t(ctx, T_ADAM);
        TrackStepParams T; memset(&T, 0, sizeof T);
        PoseGradParams& G = T.G;
        G.d_rays = ctx->d_rays; G.idx = ctx->idx; G.valid = ctx->valid; G.cams = cam; G.cam_mask = 1u; G.pix_per_frame = n; G.n_frames = 1; G.lo = 0; G.hi = n;
        G.H0 = H0; G.W0 = W0; G.Wc = W1 - W0; G.raydir = c.raydir; G.fx = c.fx; G.fy = c.fy; G.cx = c.cx; G.cy = c.cy;
        T.partial = ctx->trk_scratch + 8; T.cam = cam; T.m = ctx->m + ctx->off_cam; T.v = ctx->v + ctx->off_cam; T.g_out = ctx->cam_grad_last;
        const double b1 = 0.9, b2 = 0.999, bc1 = 1.0 - std::pow(b1, (double)ctx->trk_step), bc2 = 1.0 - std::pow(b2, (double)ctx->trk_step);
        T.beta1 = (float)b1; T.beta2 = (float)b2; T.om_beta1 = (float)(1.0 - b1); T.om_beta2 = (float)(1.0 - b2); T.eps = 1e-8f;
        T.bc2_sqrt = (float)std::sqrt(bc2); T.step = (float)((double)c.tracking_lr / bc1);
        k_track_step<<<std::max(1, std::min(32, cdiv(n, 256))), 256, 0, ctx->stream>>>(T); ctx->launches++;
        CK(cudaGetLastError());
    }
    if (cam_grad7) CK(cudaMemcpyAsync(cam_grad7, ctx->cam_grad_last, 28, cudaMemcpyDeviceToHost, ctx->stream));
    if (loss) CK(cudaMemcpyAsync(loss, stats + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (loss || cam_grad7) CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int nsb_tracking_get_camera(nsb_ctx* ctx, float* cam7) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    CK(cudaMemcpyAsync(cam7, ctx->param + ctx->off_cam, 28, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- multi-GPU -----------------------------------------------------------------------------------------------------------
extern "C" int nsb_comm_unique_id(char* id128) {
    if (!g_nccl.load()) return -1;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return -1;
    memcpy(id128, id.internal, 128);
    return 0;
}
extern "C" int nsb_comm_init(nsb_ctx* ctx, const char* id128, int rank, int world) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (!g_nccl.load()) return fail(ctx, "libnccl.so.2 not found");
    ncclUniqueId id; memcpy(id.internal, id128, 128);
    CK(cudaSetDevice(ctx->device));
    const int rc = g_nccl.CommInitRank(&ctx->comm, world, id, rank);
    if (rc != 0) return fail(ctx, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    ctx->rank = rank; ctx->world = world;
    return 0;
}
// ---- peer-memory mode: CUDA IPC handles of {gradient arena, parameter arena, flag block} ----------------------------------------
extern "C" int nsb_comm_p2p_export(nsb_ctx* ctx, char* handles192) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    cudaIpcMemHandle_t h[3];
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcGetMemHandle(&h[0], ctx->grad)); CK(cudaIpcGetMemHandle(&h[1], ctx->param)); CK(cudaIpcGetMemHandle(&h[2], ctx->p2p_flags));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handles192, h, sizeof h);
    return 0;
}
// all_handles: [world][192] gathered from every rank's nsb_comm_p2p_export (rank order).  After this call the mapping iteration
// replaces ncclAllReduce + Adam by k_reduce_adam.  The caller must put a host barrier between the imports and the first iteration.
extern "C" int nsb_comm_p2p_import(nsb_ctx* ctx, const char* all_handles, int rank, int world) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    if (!all_handles) { CK(cudaStreamSynchronize(ctx->stream)); drop_graphs(ctx); ctx->graph_sig = 0; ctx->p2p = false; return 0; }   // back to the NCCL path (e.g. another rank could not open the handles)
    if (world < 2 || world > P2P_MAX_WORLD) return fail(ctx, "peer-memory mode supports 2..%d ranks, got %d", P2P_MAX_WORLD, world);
    if (ctx->comm && (rank != ctx->rank || world != ctx->world)) return fail(ctx, "rank/world differ from nsb_comm_init");
    CK(cudaSetDevice(ctx->device));
    for (int w = 0; w < world; ++w) {
        if (w == rank) { ctx->peer_grad[w] = ctx->grad; ctx->peer_param[w] = ctx->param; ctx->peer_flags[w] = ctx->p2p_flags; continue; }
        cudaIpcMemHandle_t h[3]; memcpy(h, all_handles + 192 * (size_t)w, sizeof h);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h[0], cudaIpcMemLazyEnablePeerAccess)); ctx->peer_grad[w] = (float*)p;
        CK(cudaIpcOpenMemHandle(&p, h[1], cudaIpcMemLazyEnablePeerAccess)); ctx->peer_param[w] = (float*)p;
        CK(cudaIpcOpenMemHandle(&p, h[2], cudaIpcMemLazyEnablePeerAccess)); ctx->peer_flags[w] = (uint32_t*)p;
    }
    // the flag block starts from zero (epochs, time-out flag): a re-import on a live context must not see old epochs.  Every rank
    // does this BEFORE the host barrier the caller places between the imports and the first iteration.
    CK(cudaMemsetAsync(ctx->p2p_flags, 0, 32 * 4, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    drop_graphs(ctx); ctx->graph_sig = 0;
    ctx->rank = rank; ctx->world = world; ctx->p2p = true;
    return 0;
}
// Instrumentation of the fused exchange (k_reduce_adam stamps %globaltimer): out8 = {last barrier-1 wait us, last kernel total us,
// mean wait us, mean total us, exchanges averaged, mean exchanged range bytes, NVLink bytes per direction ((W-1)/W model), GB/s per
// direction}.  The wait is the time this rank spent waiting for the slowest rank's backward;
// total - wait is the reduce-scatter + Adam + all-gather + barrier 2 itself.  reset != 0 clears the sums.
extern "C" int nsb_comm_p2p_stats(nsb_ctx* ctx, double* out8, int reset) {
    cudaSetDevice(ctx->device);
    for (int i = 0; i < 8; ++i) out8[i] = 0.0;
    if (!ctx->p2p) return 0;
    uint32_t h[32];
    CK(cudaMemcpyAsync(h, ctx->p2p_flags, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    unsigned long long ts[4], bytes; memcpy(ts, h + 20, sizeof ts); memcpy(&bytes, h + 30, sizeof bytes);
    const double n = (double)h[28];
    out8[0] = ts[0] * 1e-3; out8[1] = ts[1] * 1e-3; out8[2] = n > 0 ? ts[2] * 1e-3 / n : 0.0; out8[3] = n > 0 ? ts[3] * 1e-3 / n : 0.0; out8[4] = n;
    // NVLink traffic model of the fused exchange: a rank reads its 1/W slice of the range from each of the W-1 peers and writes
    // the updated slice to each of them: (W-1)/W of the range per direction and rank
    out8[5] = n > 0 ? (double)bytes / n : 0.0;
    out8[6] = out8[5] * (ctx->world - 1) / std::max(1, ctx->world);
    const double ex_us = out8[3] - out8[2];
    out8[7] = ex_us > 0 ? out8[6] / ex_us * 1e-3 : 0.0;     // GB/s per direction over (kernel - barrier-1 wait)
    if (reset) { CK(cudaMemsetAsync(ctx->p2p_flags + 24, 0, 8 * 4, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); }
    return 0;
}
extern "C" int nsb_comm_rank_world(nsb_ctx* ctx, int* rank, int* world) { *rank = ctx->rank; *world = ctx->world; return 0; }

// Grid sampling alone on the rays / z values of the last forward (n rays, S samples): returns the average device time of
// `reps` launches in ms.  Algorithmic traffic per launch = n * S * 3 lookups * 8 corners * 128 B.
extern "C" int nsb_bench_gather(nsb_ctx* ctx, int reps, float* ms_out) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    const int n = ctx->last_n, S = ctx->last_S;
    if (n <= 0) return fail(ctx, "no forward has run yet");
    DecodeParams P; fill_decode_params(ctx, P, n, S, nullptr);
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int grid = ctx->n_sm * 8;
    CK(launch_gather_only(P, ctx->occ[0], grid, ctx->stream));
    CK(cudaEventRecord(a, ctx->stream));
    for (int r = 0; r < reps; ++r) CK(launch_gather_only(P, ctx->occ[0], grid, ctx->stream));
    CK(cudaEventRecord(b, ctx->stream));
    CK(cudaEventSynchronize(b));
    float ms = 0.f; CK(cudaEventElapsedTime(&ms, a, b));
    ctx->launches += reps + 1;
    *ms_out = ms / (float)reps;
    cudaEventDestroy(a); cudaEventDestroy(b);
    return 0;
}

// ---- instrumentation ----------------------------------------------------------------------------------------------------
// Cycle counters of the tcgen05 forward (only filled by the NSB_TC_TIMING build variant): 32 values, reset on read.
extern "C" int nsb_debug_counters(nsb_ctx* ctx, unsigned long long* out32) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    CK(cudaMemcpyAsync(out32, ctx->dbg, 32 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemsetAsync(ctx->dbg, 0, 32 * 8, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int64_t nsb_launch_count(nsb_ctx* ctx, int reset) { const int64_t v = ctx->launches; if (reset) ctx->launches = 0; return v; }
extern "C" int nsb_set_profiling(nsb_ctx* ctx, int on) { ctx->profiling = on != 0; ctx->ev_used = 0; return 0; }
// Sum of the device times of every region timed since nsb_set_profiling(1), per category.
extern "C" int nsb_get_kernel_ms(nsb_ctx* ctx, float* ms) {
    cudaSetDevice(ctx->device);   // one context = one GPU; the caller's current device may differ
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < T_N; ++i) ms[i] = 0.f;
    for (size_t k = 0; k < ctx->ev_used; ++k) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, ctx->ev_pool[k].a, ctx->ev_pool[k].b));
        ms[ctx->ev_pool[k].id] += t;
    }
    return 0;
}
