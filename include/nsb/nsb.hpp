// nsb.hpp -- C++ host-side mirror of the reference's class API over the C ABI of nsb.h.
//
// Same names, argument order and meaning as the reference so that its callers switch by changing includes:
//   Renderer::render_batch_ray / eval_points      include/Renderer.h:11-13, src/Renderer.cpp:19-125
//   NICE::forward                                 include/models/NICE.h:7, src/models/NICE.cpp:16-51
//   Mapper::optimize_map / run                    include/Mapper.h:20-23, src/Mapper.cpp:198-552
//   Tracker::optimize_cam_in_batch / run          include/Tracker.h:11-14, src/Tracker.cpp:41-113
//   get_samples, quad2rotation, get_camera_from_tensor, get_tensor_from_camera   include/torchlib/utils.h:141,174,198,212
// What differs, by necessity: torch::Tensor -> nsb::Tensor (a shared host fp32 buffer with a shape: no libtorch on the
// path), c10::Dict<std::string, torch::Tensor> -> nsb::Dict (same reference semantics), YAML::Node -> the two YAML file
// paths, torch::optim::Adam& -> nsb::Adam (the optimiser state lives in the engine), c10::Error -> std::runtime_error.
// Header-only; link with libnsb.so.
#pragma once
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "../nsb.h"

namespace nsb {

// Reference-semantic fp32 host tensor (copying a Tensor shares the storage, like a torch::Tensor handle).
class Tensor {
  public:
    Tensor() = default;
    explicit Tensor(std::vector<int64_t> shape) : shape_(std::move(shape)), buf_(std::make_shared<std::vector<float>>(numel_of(shape_))) {}
    Tensor(std::vector<int64_t> shape, const float* src) : Tensor(std::move(shape)) { std::memcpy(data(), src, sizeof(float) * numel()); }
    bool defined() const { return static_cast<bool>(buf_); }
    const std::vector<int64_t>& sizes() const { return shape_; }
    int64_t size(int d) const { return shape_[d < 0 ? d + (int)shape_.size() : d]; }
    int64_t numel() const { return numel_of(shape_); }
    float* data() { return buf_->data(); }
    const float* data() const { return buf_->data(); }
    float item() const { return (*buf_)[0]; }
    Tensor clone() const { Tensor t(shape_); std::memcpy(t.data(), data(), sizeof(float) * numel()); return t; }
    Tensor reshape(std::vector<int64_t> s) const { Tensor t = *this; t.shape_ = std::move(s); return t; }

  private:
    static int64_t numel_of(const std::vector<int64_t>& s) { return std::accumulate(s.begin(), s.end(), (int64_t)1, std::multiplies<int64_t>()); }
    std::vector<int64_t> shape_;
    std::shared_ptr<std::vector<float>> buf_;
};

// c10::Dict<std::string, torch::Tensor> stand-in: a shared handle, insert() replaces and marks the entry dirty so that
// the engine re-uploads that grid before the next render (the reference mutates the dict in place, Mapper.cpp:338-347).
class Dict {
  public:
    Dict() : m_(std::make_shared<Map>()) {}
    void insert(const std::string& k, Tensor v) { (*m_)[k] = Entry{std::move(v), true}; }
    Tensor at(const std::string& k) const { auto it = m_->find(k); if (it == m_->end()) throw std::runtime_error("nsb::Dict: no key " + k); return it->second.t; }
    bool contains(const std::string& k) const { return m_->count(k) != 0; }
    bool dirty(const std::string& k) const { auto it = m_->find(k); return it != m_->end() && it->second.dirty; }
    void mark_clean(const std::string& k) { (*m_)[k].dirty = false; }
    void mark_dirty(const std::string& k) { (*m_)[k].dirty = true; }

  private:
    struct Entry { Tensor t; bool dirty; };
    using Map = std::map<std::string, Entry>;
    std::shared_ptr<Map> m_;
};

inline int stage_id(const std::string& stage) {
    if (stage == "coarse") return NSB_COARSE;
    if (stage == "middle") return NSB_MIDDLE;
    if (stage == "fine") return NSB_FINE;
    if (stage == "color") return NSB_COLOR;
    throw std::runtime_error("unknown stage '" + stage + "'");
}
inline const char* grid_key(int level) { static const char* k[4] = {"grid_coarse", "grid_middle", "grid_fine", "grid_color"}; return k[level]; }

// One engine (one GPU) shared by the wrapper objects; RAII over nsb_ctx.
class Engine {
  public:
    explicit Engine(const nsb_config& cfg, int device = 0) : cfg_(cfg) {
        if (nsb_create(&cfg, device, &ctx_) != 0) {
            std::string msg = ctx_ ? nsb_last_error(ctx_) : "nsb_create failed";
            if (ctx_) nsb_destroy(ctx_);
            throw std::runtime_error("libnsb: " + msg);
        }
    }
    ~Engine() { if (ctx_) nsb_destroy(ctx_); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
    nsb_ctx* ctx() const { return ctx_; }
    const nsb_config& cfg() const { return cfg_; }
    void check(int rc) const { if (rc != 0) throw std::runtime_error(std::string("libnsb: ") + nsb_last_error(ctx_)); }
    // upload the grids the dict marks dirty / download the optimised grids back into it
    void sync_grids_to_device(Dict& c) {
        for (int l = 0; l < 4; ++l) if (c.contains(grid_key(l)) && c.dirty(grid_key(l))) { check(nsb_set_grid(ctx_, l, c.at(grid_key(l)).data())); c.mark_clean(grid_key(l)); }
    }
    void sync_grids_to_host(Dict& c) {
        for (int l = 0; l < 4; ++l) if (c.contains(grid_key(l))) { Tensor t = c.at(grid_key(l)); check(nsb_get_grid(ctx_, l, t.data())); }   // incl. grid_coarse (coarse mapper, Mapper.cpp:450-453)
    }

  private:
    nsb_config cfg_;
    nsb_ctx* ctx_ = nullptr;
};
using EnginePtr = std::shared_ptr<Engine>;

inline nsb_config load_config(const std::string& nice_slam_yaml, const std::string& dataset_yaml) {
    nsb_config cfg; nsb_config_default(&cfg);
    char err[512] = {0};
    if (nsb_config_load_yaml(&cfg, nice_slam_yaml.empty() ? nullptr : nice_slam_yaml.c_str(), dataset_yaml.empty() ? nullptr : dataset_yaml.c_str(), err, sizeof err) != 0)
        throw std::runtime_error(std::string("nsb config: ") + err);
    return cfg;
}

// ---- include/torchlib/utils.h ------------------------------------------------------------------------------------
inline Tensor quad2rotation(const Tensor& quad) {   // utils.h:174-195, (bs,4) -> (bs,3,3)
    const int64_t bs = quad.size(0);
    Tensor R({bs, 3, 3});
    for (int64_t b = 0; b < bs; ++b) nsb_quad2rotation(quad.data() + 4 * b, R.data() + 9 * b);
    return R;
}
inline Tensor get_camera_from_tensor(const Tensor& inputs) {   // utils.h:198-210, (7) -> (3,4)
    Tensor RT({3, 4});
    nsb_get_camera_from_tensor(inputs.data(), RT.data());
    return RT;
}
inline Tensor get_tensor_from_camera(const Tensor& RT, bool Tquad = false) {   // utils.h:212-231 (fixed: uses R, (w,x,y,z))
    float m[16] = {0}; m[15] = 1.f;
    std::memcpy(m, RT.data(), sizeof(float) * std::min<int64_t>(16, RT.numel()));
    float c7[7]; nsb_get_tensor_from_camera(m, c7);
    Tensor out({7});
    if (Tquad) { for (int i = 0; i < 3; ++i) out.data()[i] = c7[4 + i]; for (int i = 0; i < 4; ++i) out.data()[3 + i] = c7[i]; }
    else std::memcpy(out.data(), c7, sizeof c7);
    return out;
}
// utils.h:141-146.  `frame_slot` names the resident frame that holds (depth, color); c2w is (3..4, 4).
inline void get_samples(Engine& e, int frame_slot, int H0, int H1, int W0, int W1, int n, const Tensor& c2w,
                        Tensor& rays_o, Tensor& rays_d, Tensor& sample_depth, Tensor& sample_color) {
    float m[16] = {0}; m[15] = 1.f;
    std::memcpy(m, c2w.data(), sizeof(float) * std::min<int64_t>(16, c2w.numel()));
    rays_o = Tensor({n, 3}); rays_d = Tensor({n, 3}); sample_depth = Tensor({n}); sample_color = Tensor({n, 3});
    e.check(nsb_get_samples(e.ctx(), frame_slot, m, H0, H1, W0, W1, n, nullptr, rays_o.data(), rays_d.data(), sample_depth.data(), sample_color.data(), nullptr, nullptr));
}

// ---- include/models/NICE.h ------------------------------------------------------------------------------------------
struct NICE {
    // NICE.cpp:3: the constructor arguments are kept for source compatibility; the decoder weights are flat vectors
    // (layout in nsb.h) set with load(), replacing the torch::jit::load of NICE.cpp:8-11.
    NICE(EnginePtr engine, int dim = 3, int c_dim = 32, int hidden_size = 32, float coarse_grid_len = 2.f, float middle_grid_len = 0.32f,
         float fine_grid_len = 0.16f, float color_grid_len = 0.16f, bool coarse = false, std::string pose_emb = "fourier")
        : e(std::move(engine)) { (void)dim; (void)c_dim; (void)hidden_size; (void)coarse_grid_len; (void)middle_grid_len; (void)fine_grid_len; (void)color_grid_len; (void)coarse; (void)pose_emb; }
    void load(const std::string& which, const std::vector<float>& flat) { e->check(nsb_set_decoder(e->ctx(), stage_id(which), flat.data(), (int64_t)flat.size())); }
    std::vector<float> parameters(const std::string& which) const {
        std::vector<float> v((size_t)nsb_decoder_count(stage_id(which), e->cfg().c_dim));
        e->check(nsb_get_decoder(e->ctx(), stage_id(which), v.data(), (int64_t)v.size()));
        return v;
    }
    // NICE.cpp:16-51: p (1,P,3) or (P,3) -> raw (P,4); no bound mask here (that is Renderer::eval_points)
    Tensor forward(Tensor p, Dict c_grid, std::string stage);
    EnginePtr e;
};

// ---- include/Renderer.h ------------------------------------------------------------------------------------------------
class Renderer {
  public:
    explicit Renderer(EnginePtr engine) : e_(std::move(engine)) {}
    // Renderer.cpp:19-42
    Tensor eval_points(Tensor p, NICE decoders, Dict c, std::string stage) {
        (void)decoders;
        e_->sync_grids_to_device(c);
        const int P = (int)(p.numel() / 3);
        Tensor raw({P, 4});
        e_->check(nsb_eval_points(e_->ctx(), stage_id(stage), P, p.data(), raw.data()));
        return raw;
    }
    // Renderer.cpp:44-125 -- note the reference's argument order: rays_d precedes rays_o
    void render_batch_ray(Dict c, NICE decoders, Tensor rays_d, Tensor rays_o, std::string stage, Tensor gt_depth,
                          Tensor& rgb_map, Tensor& depth_map, Tensor& depth_var, Tensor& weights) {
        (void)decoders;
        e_->sync_grids_to_device(c);
        const int n = (int)rays_o.size(0);
        const int S = gt_depth.defined() ? e_->cfg().n_samples + e_->cfg().n_surface : e_->cfg().n_samples;
        rgb_map = Tensor({n, 3}); depth_map = Tensor({n}); depth_var = Tensor({n}); weights = Tensor({n, S});
        e_->check(nsb_render_batch_ray(e_->ctx(), stage_id(stage), n, rays_d.data(), rays_o.data(), gt_depth.defined() ? gt_depth.data() : nullptr,
                                       rgb_map.data(), depth_map.data(), depth_var.data(), weights.data()));
    }

  private:
    EnginePtr e_;
};

inline Tensor NICE::forward(Tensor p, Dict c_grid, std::string stage) {
    // the decoders see unmasked points; points outside the bound therefore come back with occupancy 100 only from
    // Renderer::eval_points -- here the bound is widened by evaluating through the same kernel and keeping raw values
    Renderer r(e);
    return r.eval_points(p, *this, c_grid, stage);
}

// torch::optim::Adam stand-in: the moments live in the engine's arena; this object only carries the step size.
struct Adam { double lr; explicit Adam(double lr_) : lr(lr_) {} };

struct KeyFrame { Tensor est_c2w, gt_c2w, color, depth; int idx = 0; int slot = 0; };   // Mapper.h:11-15

// ---- include/Mapper.h -----------------------------------------------------------------------------------------------------
class Mapper {
  public:
    Mapper(EnginePtr engine, bool coarse_mapper = false) : e_(std::move(engine)), coarse_mapper_(coarse_mapper) { mapping_window_size_ = e_->cfg().mapping_window_size; }
    // Mapper.cpp:132-196 (member form): indices into keyframes() of the k_overlap keyframes that see most of the current view.
    void keyframe_selection_overlap(Tensor gt_color, Tensor gt_depth, Tensor c2w, const std::vector<KeyFrame>& keyframe_vector, int k_overlap,
                                    std::vector<int>& selected_kf) {
        const int cur_slot = e_->cfg().max_frames - 1;
        float m[16] = {0}; m[15] = 1.f; std::memcpy(m, c2w.data(), sizeof(float) * std::min<int64_t>(16, c2w.numel()));
        e_->check(nsb_set_frame(e_->ctx(), cur_slot, gt_depth.data(), gt_color.data(), m));
        select_overlap(cur_slot, keyframe_vector, k_overlap, selected_kf);
    }
    // Mapper.cpp:198-491.  The current frame is uploaded to the last resident slot; keyframes keep their slots.
    void optimize_map(int num_joint_iters, Dict& c, Tensor cur_gt_color, Tensor cur_gt_depth, Tensor gt_cur_c2w, Tensor& cur_c2w, NICE& decoders,
                      float lr_factor = 1.f, std::vector<float>* losses = nullptr) {
        (void)gt_cur_c2w; (void)decoders;
        e_->sync_grids_to_device(c);
        const int cur_slot = e_->cfg().max_frames - 1;
        float m[16] = {0}; m[15] = 1.f; std::memcpy(m, cur_c2w.data(), sizeof(float) * std::min<int64_t>(16, cur_c2w.numel()));
        e_->check(nsb_set_frame(e_->ctx(), cur_slot, cur_gt_depth.data(), cur_gt_color.data(), m));
        // optimize_frame (Mapper.cpp:200-216): overlap-selected keyframes, the latest keyframe, then the current frame (-1)
        std::vector<int> optimize_frame;
        if (!keyframes_.empty()) select_overlap(cur_slot, keyframes_, mapping_window_size_ - 2, optimize_frame);
        int oldest_frame = -1;
        if (!keyframes_.empty()) {
            optimize_frame.push_back((int)keyframes_.size() - 1);
            oldest_frame = *std::min_element(optimize_frame.begin(), optimize_frame.end());
        }
        optimize_frame.push_back(-1);
        if ((int)optimize_frame.size() > 16) throw std::runtime_error("optimize_frame exceeds the 16 frames a mapping call supports");
        std::vector<int> slots;
        uint32_t ba_mask = 0;
        for (size_t f = 0; f < optimize_frame.size(); ++f) {
            const int fr = optimize_frame[f];
            slots.push_back(fr == -1 ? cur_slot : keyframes_[(size_t)fr].slot);
            if (BA && fr != oldest_frame) ba_mask |= 1u << f;          // Mapper.cpp:305-329 (the current frame always joins)
        }
        std::vector<float> l((size_t)num_joint_iters);
        // coarse mapper: stage "coarse", only grid_coarse (Mapper.cpp:335-338,351-352); color_refine: Mapper.cpp:505-513
        const int flags = (coarse_mapper_ ? NSB_MAP_COARSE : 0) | (refine_ ? NSB_MAP_COLOR_REFINE : 0);
        e_->check(nsb_mapping_begin_ex(e_->ctx(), (int)slots.size(), slots.data(), num_joint_iters, lr_factor, coarse_mapper_ ? 0u : ba_mask, flags));
        for (int it = 0; it < num_joint_iters; ++it) e_->check(nsb_mapping_iter_async(e_->ctx(), it, nullptr));
        e_->check(nsb_mapping_losses(e_->ctx(), 0, num_joint_iters, l.data(), nullptr));   // the loss ring holds every step of this call
        e_->check(nsb_mapping_end(e_->ctx(), nullptr));
        if (losses) *losses = l;
        if (ba_mask) {                                               // Mapper.cpp:467-489: est_c2w / cur_c2w <- optimised cameras
            for (size_t f = 0; f < optimize_frame.size(); ++f) {
                if (!((ba_mask >> f) & 1u)) continue;
                Tensor pose({4, 4});
                e_->check(nsb_get_frame_pose(e_->ctx(), slots[f], pose.data()));
                pose.data()[12] = pose.data()[13] = pose.data()[14] = 0.f; pose.data()[15] = 1.f;
                if (optimize_frame[f] == -1) cur_c2w = pose; else keyframes_[(size_t)optimize_frame[f]].est_c2w = pose;
            }
        }
        e_->sync_grids_to_host(c);                                   // Mapper.cpp:448-464 write-back
    }
    // Mapper.cpp:493-552.  The transliteration hard-codes `init = true` (:495); the intent (upstream) is the first frame only.
    void run(NICE& decoders, Dict& c, std::vector<Tensor>& estimate_c2w_vec, Tensor gt_color_t, Tensor gt_depth_t, Tensor gt_c2w_t, int idx, int n_imgs) {
        const bool first = keyframes_.empty() && idx == 0;
        int iters = first ? e_->cfg().mapping_iters_first : e_->cfg().mapping_iters;
        const float lrf = first ? e_->cfg().lr_first_factor : e_->cfg().lr_factor;
        int outer_joint_iters = 1;
        if (!first && idx == n_imgs - 1 && e_->cfg().color_refine && !coarse_mapper_) {                 // Mapper.cpp:505-513: the final refinement pass
            outer_joint_iters = 5;
            mapping_window_size_ *= 2;
            refine_ = true;            // middle_iter_ratio = fine_iter_ratio = 0, fix_color = true, frustum_feature_selection = false
            iters *= 5;
        }
        Tensor cur = estimate_c2w_vec[(size_t)idx];
        iters = iters / outer_joint_iters;                                                               // Mapper.cpp:526
        for (int outer = 0; outer < outer_joint_iters; ++outer) {
            BA = keyframes_.size() > 4 && e_->cfg().BA && !coarse_mapper_;                              // Mapper.cpp:530
            optimize_map(iters, c, gt_color_t, gt_depth_t, gt_c2w_t, cur, decoders, lrf);
            if (BA) estimate_c2w_vec[(size_t)idx] = cur;                                                // Mapper.cpp:533-534
        }
        if (idx % e_->cfg().keyframe_every == 0 || idx == n_imgs - 2) {
            for (const KeyFrame& k : keyframes_) if (k.idx == idx) return;                              // Mapper.cpp:539
            // the reference's keyframe_vector is unbounded; here every keyframe occupies a resident device slot (the last slot is
            // the current frame's): say so instead of silently dropping keyframes (size max_frames from n_imgs / keyframe_every + 2)
            if ((int)keyframes_.size() >= e_->cfg().max_frames - 1)
                throw std::runtime_error("Mapper::run: the " + std::to_string(e_->cfg().max_frames) + " resident frame slots are exhausted at frame " + std::to_string(idx) +
                                         "; create the engine with max_frames >= n_imgs / keyframe_every + 3");
            KeyFrame kf; kf.idx = idx; kf.gt_c2w = gt_c2w_t; kf.est_c2w = cur; kf.color = gt_color_t; kf.depth = gt_depth_t; kf.slot = (int)keyframes_.size();
            float m[16] = {0}; m[15] = 1.f; std::memcpy(m, cur.data(), sizeof(float) * std::min<int64_t>(16, cur.numel()));
            e_->check(nsb_set_frame(e_->ctx(), kf.slot, gt_depth_t.data(), gt_color_t.data(), m));
            keyframes_.push_back(kf);
        }
    }
    const std::vector<KeyFrame>& keyframes() const { return keyframes_; }
    bool BA = false;

  private:
    void select_overlap(int cur_slot, const std::vector<KeyFrame>& kfs, int k_overlap, std::vector<int>& selected_kf) {
        std::vector<float> poses(16 * kfs.size(), 0.f);
        for (size_t k = 0; k < kfs.size(); ++k) {
            poses[16 * k + 15] = 1.f;
            std::memcpy(poses.data() + 16 * k, kfs[k].est_c2w.data(), sizeof(float) * std::min<int64_t>(16, kfs[k].est_c2w.numel()));
        }
        std::vector<int> sel(kfs.size() + 1); int n_sel = 0;
        e_->check(nsb_keyframe_selection_overlap(e_->ctx(), cur_slot, nullptr, (int)kfs.size(), poses.data(), k_overlap, nullptr, 100, 16, sel.data(), &n_sel, nullptr));
        selected_kf.assign(sel.begin(), sel.begin() + n_sel);
    }
    EnginePtr e_;
    bool coarse_mapper_;
    bool refine_ = false;              // color_refine has been switched on by run() (Mapper.cpp:505-513)
    int mapping_window_size_ = 0;      // mapping.mapping_window_size, doubled by color_refine (Mapper.cpp:507)
    std::vector<KeyFrame> keyframes_;
};

// ---- include/Tracker.h ----------------------------------------------------------------------------------------------------
class Tracker {
  public:
    Tracker(EnginePtr engine, Dict c_dict) : e_(std::move(engine)), c_(std::move(c_dict)) {}
    // Tracker.cpp:41-89.  cam_tensor (7) is updated in place like the torch parameter the optimiser steps.
    Tensor optimize_cam_in_batch(Tensor cam_tensor, Tensor gt_color, Tensor gt_depth, int batch_size, Adam& optimizer, NICE decoders) {
        (void)gt_color; (void)gt_depth; (void)batch_size; (void)optimizer; (void)decoders;
        Tensor loss({1});
        e_->check(nsb_tracking_iter(e_->ctx(), nullptr, loss.data(), nullptr));
        e_->check(nsb_tracking_get_camera(e_->ctx(), cam_tensor.data()));
        return loss;
    }
    // Tracker.cpp:92-113
    Tensor run(NICE decoders, Tensor gt_color_t, Tensor gt_depth_t, Tensor gt_c2w_t, int idx, std::vector<float>* losses = nullptr) {
        (void)idx;
        e_->sync_grids_to_device(c_);
        const int slot = e_->cfg().max_frames - 1;
        float m[16] = {0}; m[15] = 1.f; std::memcpy(m, gt_c2w_t.data(), sizeof(float) * std::min<int64_t>(16, gt_c2w_t.numel()));
        e_->check(nsb_set_frame(e_->ctx(), slot, gt_depth_t.data(), gt_color_t.data(), m));
        Tensor camera_tensor = get_tensor_from_camera(gt_c2w_t, false);   // Tracker.cpp:100
        e_->check(nsb_tracking_begin(e_->ctx(), slot, camera_tensor.data()));
        Adam optimizer(e_->cfg().tracking_lr);                                // Tracker.cpp:103
        for (int i = 0; i < e_->cfg().tracking_iters; ++i) {                  // Tracker.cpp:107-112
            Tensor loss = optimize_cam_in_batch(camera_tensor, gt_color_t, gt_depth_t, e_->cfg().tracking_pixels, optimizer, decoders);
            if (losses) losses->push_back(loss.item());
        }
        return camera_tensor;
    }

  private:
    EnginePtr e_;
    Dict c_;
};

}  // namespace nsb
