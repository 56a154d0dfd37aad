// TEST INFRASTRUCTURE -- oracle/_ref harness.  Not part of the product path; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the
// library this file builds (oracle/_ref/libnsbref*.so).
//
// What it is: the reference's OWN src/Renderer.cpp and include/torchlib/utils.h, compiled where they
// lie under /root/reference (see oracle/Makefile), with
//   * `#define kCUDA kCPU` after the torch headers so the hard-coded torch::Device(torch::kCUDA, 0)
//     (Renderer.cpp:31,69,76,87,90,94,102; utils.h:156,161,168) resolves to the CPU,
//   * shims for the headers that do not exist in this image (oracle/shims/Eigen, opencv2),
//   * oracle/shims/models/NICE.h restating the decoders (the reference loads them from absent
//     TorchScript files, NICE.cpp:8-11).
// Built twice: NSB_REF_VERBATIM (Renderer.cpp untouched: forward only, autograd is off because of the
// function-scoped NoGradGuard at Renderer.cpp:66) and the "grad" build from a sed-patched copy
// (guard scoped to lines 67-73; reshape at :124 uses the member N_surface) -- SURVEY.md 8-A.3.
//
// Mapper.cpp / Tracker.cpp cannot be compiled here (OpenCV + yaml-cpp; main.cpp:96 type error), so the
// iteration loops below restate Mapper.cpp:330-465 and Tracker.cpp:41-113 around the reference's own
// get_samples / render_batch_ray / get_camera_from_tensor, with torch::optim::Adam as the reference uses.
#include <torch/torch.h>
#include <torch/script.h>
#define kCUDA kCPU
#include "Renderer.cpp"  // -I selects /root/reference/src (verbatim) or oracle/_ref/gen (patched copy)

#include <cstring>
#include <chrono>

namespace {

struct RefCtx {
    c10::Dict<std::string, torch::Tensor> grids;
    torch::Tensor flat[4];  // coarse, middle, fine, color decoders (flat leaves)
    NICE nice;
    Renderer renderer;
    std::string err;
    // optional: host buffers that receive the gradients of chosen mapping iterations (ref_mapping_capture_grads)
    struct Cap { int iter; float* grid[4]; float* dec_color; };
    std::vector<Cap> caps;
};

const char* kGridName[4] = {"grid_coarse", "grid_middle", "grid_fine", "grid_color"};

torch::Tensor from_host(const float* p, std::vector<int64_t> shape) {
    return torch::from_blob(const_cast<float*>(p), shape, torch::kFloat32).clone();
}
void to_host(const torch::Tensor& t, float* out) {
    if (!out) return;
    auto c = t.detach().to(torch::kFloat32).contiguous();
    std::memcpy(out, c.data_ptr<float>(), sizeof(float) * c.numel());
}
void grad_to_host(const torch::Tensor& t, float* out) {
    if (!out) return;
    if (t.grad().defined()) to_host(t.grad(), out);
    else std::memset(out, 0, sizeof(float) * t.numel());
}
void zero_grads(RefCtx* c) {
    for (int l = 0; l < 4; ++l) {
        if (c->grids.contains(kGridName[l])) { auto g = c->grids.at(kGridName[l]); if (g.grad().defined()) g.mutable_grad() = torch::Tensor(); }
        if (c->flat[l].defined() && c->flat[l].grad().defined()) c->flat[l].mutable_grad() = torch::Tensor();
    }
}

}  // namespace

#define REF_TRY try {
#define REF_CATCH(c) } catch (const std::exception& e) { (c)->err = e.what(); return -1; } return 0;

extern "C" {

int ref_is_verbatim() {
#ifdef NSB_REF_VERBATIM
    return 1;
#else
    return 0;
#endif
}
void ref_set_threads(int n) { at::set_num_threads(n); }
int ref_get_threads() { return at::get_num_threads(); }

void* ref_create() { return new RefCtx(); }
void ref_destroy(void* h) { delete static_cast<RefCtx*>(h); }
const char* ref_last_error(void* h) { return static_cast<RefCtx*>(h)->err.c_str(); }

// grid: (1, C, Z, Y, X) channel-first fp32, the reference's layout (main.cpp:39-43).
int ref_set_grid(void* h, int level, const float* ncdhw, int C, int Z, int Y, int X) {
    auto c = static_cast<RefCtx*>(h);
    REF_TRY
    c->grids.insert_or_assign(kGridName[level], from_host(ncdhw, {1, C, Z, Y, X}));
    REF_CATCH(c)
}
int ref_get_grid(void* h, int level, float* out) {
    auto c = static_cast<RefCtx*>(h);
    REF_TRY
    to_host(c->grids.at(kGridName[level]), out);
    REF_CATCH(c)
}
int64_t ref_decoder_count(int which, int E, int H, int C, int O) {
    return NsbRefDecoder::count(which == 0, E, H, C, O);
}
int ref_set_decoder(void* h, int which, const float* flat, int64_t n, int E, int H, int C, int O) {
    auto c = static_cast<RefCtx*>(h);
    REF_TRY
    c->flat[which] = from_host(flat, {n});
    NsbRefDecoder* d[4] = {&c->nice.coarse, &c->nice.middle, &c->nice.fine, &c->nice.color};
    d[which]->bind(c->flat[which], which == 0, E, H, C, O);
    REF_CATCH(c)
}
int ref_get_decoder(void* h, int which, float* out) {
    auto c = static_cast<RefCtx*>(h);
    REF_TRY
    to_host(c->flat[which], out);
    REF_CATCH(c)
}

static void rebind(RefCtx* c, int which) {
    NsbRefDecoder* d[4] = {&c->nice.coarse, &c->nice.middle, &c->nice.fine, &c->nice.color};
    auto& dd = *d[which];
    dd.bind(c->flat[which], dd.no_xyz, dd.E, dd.H, dd.C, dd.O);
}

// Renderer::render_batch_ray (Renderer.cpp:44-125), forward.  gt_depth may be NULL (no-depth path).
int ref_render_batch_ray(void* h, const char* stage, int n, const float* rays_d, const float* rays_o,
                         const float* gt_depth, float* rgb, float* depth, float* var, float* weights) {
    auto c = static_cast<RefCtx*>(h);
    REF_TRY
    torch::NoGradGuard ng;
    auto rd = from_host(rays_d, {n, 3}), ro = from_host(rays_o, {n, 3});
    torch::Tensor gd; if (gt_depth) gd = from_host(gt_depth, {n});
    torch::Tensor o_rgb, o_depth, o_var, o_w;
    Renderer r;  // fresh: the no-depth path mutates the member N_surface (Renderer.cpp:56)
    r.render_batch_ray(c->grids, c->nice, rd, ro, stage, gd, o_rgb, o_depth, o_var, o_w);
    to_host(o_rgb, rgb); to_host(o_depth, depth); to_host(o_var, var); to_host(o_w, weights);
    REF_CATCH(c)
}

// Renderer::eval_points (Renderer.cpp:19-42): raw (P,4).
int ref_eval_points(void* h, const char* stage, int P, const float* pts, float* raw) {
    auto c = static_cast<RefCtx*>(h);
    REF_TRY
    torch::NoGradGuard ng;
    Renderer r;
    auto out = r.eval_points(from_host(pts, {P, 3}), c->nice, c->grids, stage);
    to_host(out, raw);
    REF_CATCH(c)
}

// raw2outputs_nerf_color (utils.h:148-172) on its own.
int ref_raw2outputs(int n, int S, const float* raw, const float* z_vals, const float* rays_d,
                    float* rgb, float* depth, float* var, float* weights) {
    try {
        torch::NoGradGuard ng;
        torch::Tensor o_rgb, o_depth, o_var, o_w;
        raw2outputs_nerf_color(from_host(raw, {n, S, 4}), from_host(z_vals, {n, S}), false,
                               from_host(rays_d, {n, 3}), o_rgb, o_depth, o_var, o_w);
        to_host(o_rgb, rgb); to_host(o_depth, depth); to_host(o_var, var); to_host(o_w, weights);
    } catch (...) { return -1; }
    return 0;
}

// quad2rotation (utils.h:174-195) and get_camera_from_tensor (utils.h:198-210).
int ref_quad2rotation(const float* q4, float* R9) {
    try { to_host(quad2rotation(from_host(q4, {1, 4}))[0], R9); } catch (...) { return -1; }
    return 0;
}
int ref_get_camera_from_tensor(const float* cam7, float* RT12) {
    try { to_host(get_camera_from_tensor(from_host(cam7, {7})), RT12); } catch (...) { return -1; }
    return 0;
}

// raySampler / get_samples (utils.h:13-55, 141-146).  Seeds the CPU generator, returns the reference's
// outputs and the pixel indices it drew (recovered by replaying torch::randint with the same seed).
int ref_get_samples(int H0, int H1, int W0, int W1, int n, int H, int W, float fx, float fy, float cx, float cy,
                    const float* c2w16, const float* depth, const float* color, uint64_t seed,
                    float* rays_o, float* rays_d, float* gt_depth, float* gt_color, int64_t* idx) {
    try {
        torch::NoGradGuard ng;
        torch::manual_seed(seed);
        torch::Tensor ro, rd, sd, sc;
        get_samples(H0, H1, W0, W1, n, H, W, (int)fx, (int)fy, (int)cx, (int)cy, from_host(c2w16, {4, 4}),
                    from_host(depth, {H, W}), from_host(color, {H, W, 3}), ro, rd, sd, sc);
        to_host(ro.contiguous(), rays_o); to_host(rd, rays_d); to_host(sd, gt_depth); to_host(sc, gt_color);
        if (idx) {
            torch::manual_seed(seed);
            auto ind = torch::randint((int64_t)(H1 - H0) * (W1 - W0), {n}).to(torch::kLong).contiguous();
            std::memcpy(idx, ind.data_ptr<int64_t>(), sizeof(int64_t) * n);
        }
    } catch (...) { return -1; }
    return 0;
}

#ifndef NSB_REF_VERBATIM
// Gradients of  L = sum(g_rgb*rgb) + sum(g_depth*depth) + sum(g_var*var)  through render_batch_ray,
// w.r.t. the grids, the decoder flats and the rays.  Any output pointer may be NULL.
int ref_render_vjp(void* h, const char* stage, int n, const float* rays_d, const float* rays_o,
                   const float* gt_depth, const float* g_rgb, const float* g_depth, const float* g_var,
                   float* d_grid_coarse, float* d_grid_middle, float* d_grid_fine, float* d_grid_color,
                   float* d_dec_coarse, float* d_dec_middle, float* d_dec_fine, float* d_dec_color,
                   float* d_rays_d, float* d_rays_o) {
    auto c = static_cast<RefCtx*>(h);
    REF_TRY
    float* dg[4] = {d_grid_coarse, d_grid_middle, d_grid_fine, d_grid_color};
    float* dd[4] = {d_dec_coarse, d_dec_middle, d_dec_fine, d_dec_color};
    for (int l = 0; l < 4; ++l) {
        if (c->grids.contains(kGridName[l])) c->grids.at(kGridName[l]).requires_grad_(true);
        if (c->flat[l].defined()) { c->flat[l].requires_grad_(true); rebind(c, l); }
    }
    zero_grads(c);
    auto rd = from_host(rays_d, {n, 3}).requires_grad_(true), ro = from_host(rays_o, {n, 3}).requires_grad_(true);
    torch::Tensor gd; if (gt_depth) gd = from_host(gt_depth, {n});
    torch::Tensor o_rgb, o_depth, o_var, o_w;
    Renderer r;
    r.render_batch_ray(c->grids, c->nice, rd, ro, stage, gd, o_rgb, o_depth, o_var, o_w);
    auto L = (o_rgb * from_host(g_rgb, {n, 3})).sum() + (o_depth * from_host(g_depth, {n})).sum() +
             (o_var * from_host(g_var, {n})).sum();
    L.backward();
    for (int l = 0; l < 4; ++l) {
        if (dg[l] && c->grids.contains(kGridName[l])) grad_to_host(c->grids.at(kGridName[l]), dg[l]);
        if (dd[l] && c->flat[l].defined()) grad_to_host(c->flat[l], dd[l]);
    }
    grad_to_host(rd, d_rays_d); grad_to_host(ro, d_rays_o);
    zero_grads(c);
    for (int l = 0; l < 4; ++l) {
        if (c->grids.contains(kGridName[l])) c->grids.at(kGridName[l]).requires_grad_(false);
        if (c->flat[l].defined()) { c->flat[l].requires_grad_(false); rebind(c, l); }
    }
    REF_CATCH(c)
}

// The next ref_mapping_iters call copies the gradients loss.backward() left at iteration `iter` (Mapper.cpp:444, before
// optimizer.step) into these host buffers: grids middle / fine / color in (1,C,Z,Y,X), the colour decoder flat.  Any may be NULL.
int ref_mapping_capture_grads(void* h, int iter, float* g_middle, float* g_fine, float* g_color, float* d_dec_color) {
    auto c = static_cast<RefCtx*>(h);
    if (iter < 0) { c->caps.clear(); return 0; }
    RefCtx::Cap cap; cap.iter = iter; cap.grid[0] = nullptr; cap.grid[1] = g_middle; cap.grid[2] = g_fine; cap.grid[3] = g_color; cap.dec_color = d_dec_color;
    c->caps.push_back(cap);
    return 0;
}

// Mapping iterations, Mapper.cpp:330-465 (non-coarse mapper, no BA), frames = optimize_frame list.
//   lr[stage(4: coarse,middle,fine,color)][group(5: decoders,coarse,middle,fine,color)] already times lr_factor.
//   voxel masks (Z*Y*X bytes per level, NULL = frustum_feature_selection off) follow the intent of
//   Mapper.cpp:264-288,333-350: Adam touches masked voxels only.
//   stage_of_iter: 1 = middle, 2 = fine, 3 = color per Mapper.cpp:351-358; chosen by the caller.
int ref_mapping_iters(void* h, int n_frames, int H, int W, float fx, float fy, float cx, float cy,
                      const float* depths, const float* colors, const float* c2ws16,
                      int mapping_pixels, int n_iters, const int* stage_of_iter, const float* lr,
                      float w_color_loss, int fix_fine, int fix_color, uint64_t seed,
                      const uint8_t* mask_middle, const uint8_t* mask_fine, const uint8_t* mask_color,
                      float* losses, int* n_inside, double* seconds) {
    auto c = static_cast<RefCtx*>(h);
    REF_TRY
    auto bound = torch::tensor({{-4.5, 3.82}, {-1.5, 2.02}, {-3.0, 2.76}});  // Mapper.cpp:29
    std::vector<torch::Tensor> dlist;  // Mapper.cpp:292-301
    if (!fix_fine) dlist.push_back(c->flat[2]);
    if (!fix_color) dlist.push_back(c->flat[3]);
    for (auto& t : dlist) t.requires_grad_(true);
    rebind(c, 2); rebind(c, 3);
    const uint8_t* masks[4] = {nullptr, mask_middle, mask_fine, mask_color};
    torch::Tensor mask_t[4];
    std::vector<torch::Tensor> gp[4];
    for (int l = 1; l < 4; ++l) {
        auto g = c->grids.at(kGridName[l]);
        g.requires_grad_(true);
        gp[l].push_back(g);
        if (masks[l]) {
            auto m = torch::from_blob(const_cast<uint8_t*>(masks[l]), {1, 1, g.size(2), g.size(3), g.size(4)}, torch::kUInt8).clone();
            mask_t[l] = m.to(torch::kFloat32).expand_as(g);
        }
    }
    std::vector<torch::optim::OptimizerParamGroup> groups;
    groups.emplace_back(dlist); groups.emplace_back(std::vector<torch::Tensor>{});
    groups.emplace_back(gp[1]); groups.emplace_back(gp[2]); groups.emplace_back(gp[3]);
    torch::optim::Adam opt(groups, torch::optim::AdamOptions(0));  // Mapper.cpp:330
    std::vector<torch::Tensor> fdepth, fcolor, fc2w;
    for (int f = 0; f < n_frames; ++f) {
        fdepth.push_back(from_host(depths + (size_t)f * H * W, {H, W}));
        fcolor.push_back(from_host(colors + (size_t)f * H * W * 3, {H, W, 3}));
        fc2w.push_back(from_host(c2ws16 + f * 16, {4, 4}));
    }
    int pix = mapping_pixels / n_frames;  // Mapper.cpp:223
    torch::manual_seed(seed);
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < n_iters; ++it) {
        int st = stage_of_iter[it];
        for (int gidx = 0; gidx < 5; ++gidx)
            static_cast<torch::optim::AdamOptions&>(opt.param_groups()[gidx].options()).lr(lr[st * 5 + gidx]);  // :360-364
        opt.zero_grad();
        std::vector<torch::Tensor> vo, vd, vdep, vcol;
        for (int f = 0; f < n_frames; ++f) {
            torch::Tensor ro, rd, sd, sc;
            get_samples(0, H, 0, W, pix, H, W, (int)fx, (int)fy, (int)cx, (int)cy, fc2w[f], fdepth[f], fcolor[f], ro, rd, sd, sc);  // :404
            vo.push_back(ro); vd.push_back(rd); vdep.push_back(sd); vcol.push_back(sc);
        }
        auto b_d = torch::cat(vd), b_o = torch::cat(vo), b_dep = torch::cat(vdep), b_col = torch::cat(vcol);
        {
            torch::NoGradGuard ng;  // Mapper.cpp:416-427, guard scoped to the filter
            auto t_ = (bound.unsqueeze(0) - b_o.unsqueeze(-1)) / b_d.unsqueeze(-1);
            auto t = std::get<0>(torch::min(std::get<0>(torch::max(t_, 2)), 1));
            auto inside = t >= b_dep;
            b_d = b_d.index({inside}); b_o = b_o.index({inside}); b_dep = b_dep.index({inside}); b_col = b_col.index({inside});
        }
        if (n_inside) n_inside[it] = (int)b_d.size(0);
        torch::Tensor color, depth, unc, weights;
        c->renderer.render_batch_ray(c->grids, c->nice, b_d, b_o, "color", b_dep, color, depth, unc, weights);  // :430
        auto dmask = b_dep > 0;
        auto loss = torch::abs(b_dep.index({dmask}) - depth.index({dmask})).sum();  // :435-436
        if (st == 3) loss = loss + w_color_loss * torch::abs(b_col - color).sum();  // :438-442
        loss.backward();
        for (auto& cap : c->caps) if (cap.iter == it) {
            for (int l = 1; l < 4; ++l) grad_to_host(gp[l][0], cap.grid[l]);
            if (!fix_color) grad_to_host(c->flat[3], cap.dec_color);
        }
        for (int l = 1; l < 4; ++l)
            if (mask_t[l].defined() && gp[l][0].grad().defined()) gp[l][0].mutable_grad().mul_(mask_t[l]);
        opt.step();
        opt.zero_grad();
        if (losses) losses[it] = loss.item<float>();
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& t : dlist) t.requires_grad_(false);
    for (int l = 1; l < 4; ++l) c->grids.at(kGridName[l]).requires_grad_(false);
    rebind(c, 2); rebind(c, 3);
    zero_grads(c);
    c->caps.clear();
    REF_CATCH(c)
}

// Tracking iterations, Tracker.cpp:41-113 (guard at :48 scoped to the filter; lr / iters from the caller,
// the reference hard-codes 1e-2 and 10 at :103,:107).  cam7 = (qw,qx,qy,qz,tx,ty,tz), updated in place.
int ref_tracking_iters(void* h, int H, int W, float fx, float fy, float cx, float cy, int edge_h, int edge_w,
                       const float* depth, const float* color, float* cam7, int pixels, int n_iters, float lr,
                       int handle_dynamic, int use_color, float w_color_loss, uint64_t seed,
                       float* losses, float* cam_grad_first, int* n_inside, double* seconds) {
    auto c = static_cast<RefCtx*>(h);
    REF_TRY
    auto bound = torch::tensor({{-4.5, 3.82}, {-1.5, 2.02}, {-3.0, 2.76}});  // Tracker.cpp:23
    auto cam = from_host(cam7, {7}).requires_grad_(true);
    torch::optim::Adam opt(std::vector<torch::Tensor>{cam}, torch::optim::AdamOptions(lr));  // :103
    auto gdepth = from_host(depth, {H, W}), gcolor = from_host(color, {H, W, 3});
    torch::manual_seed(seed);
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < n_iters; ++it) {
        opt.zero_grad();
        auto c2w = get_camera_from_tensor(cam);  // :44
        torch::Tensor b_o, b_d, b_dep, b_col;
        get_samples(edge_h, H - edge_h, edge_w, W - edge_w, pixels, H, W, (int)fx, (int)fy, (int)cx, (int)cy, c2w, gdepth, gcolor, b_o, b_d, b_dep, b_col);  // :46
        torch::Tensor inside;
        {
            torch::NoGradGuard ng;  // :48-54, guard scoped to the mask
            auto t_ = (bound.unsqueeze(0) - b_o.unsqueeze(-1)) / b_d.unsqueeze(-1);
            auto t = std::get<0>(torch::min(std::get<0>(torch::max(t_, 2)), 1));
            inside = t >= b_dep;
        }
        b_d = b_d.index({inside}); b_o = b_o.index({inside});  // :55-58, differentiable w.r.t. the pose
        b_dep = b_dep.index({inside}); b_col = b_col.index({inside});
        if (n_inside) n_inside[it] = (int)b_dep.size(0);
        torch::Tensor color_o, depth_o, unc, weights;
        c->renderer.render_batch_ray(c->grids, c->nice, b_d, b_o, "color", b_dep, color_o, depth_o, unc, weights);  // :61
        torch::Tensor mask;
        if (handle_dynamic) {  // :67-71
            auto tmp = torch::abs(b_dep - depth_o);
            mask = (tmp < 10 * tmp.median()) & (b_dep > 0);
        } else mask = b_dep > 0;
        auto loss = (torch::abs(b_dep - depth_o) / torch::sqrt(unc + 1e-10)).index({mask}).sum();  // :75
        if (use_color) loss = loss + w_color_loss * torch::abs(b_col - color_o).index({mask}).sum();  // :77-82
        loss.backward();
        if (it == 0 && cam_grad_first) grad_to_host(cam, cam_grad_first);
        opt.step();
        opt.zero_grad();
        if (losses) losses[it] = loss.item<float>();
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    to_host(cam, cam7);
    REF_CATCH(c)
}
#endif  // !NSB_REF_VERBATIM

}  // extern "C"
