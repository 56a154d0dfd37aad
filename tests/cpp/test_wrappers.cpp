// Exercises the C++ wrapper classes of include/nsb/nsb.hpp exactly the way a caller of the reference would use
// Renderer / Mapper / Tracker (same method names and argument order).  Inputs are raw fp32 files written by
// tests/test_cpp_wrappers.py; outputs go back the same way.  usage: test_wrappers <dir>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include "nsb/nsb.hpp"

static std::vector<float> rd(const std::string& p) {
    std::ifstream f(p, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + p);
    const size_t n = (size_t)f.tellg() / sizeof(float);
    std::vector<float> v(n);
    f.seekg(0); f.read(reinterpret_cast<char*>(v.data()), n * sizeof(float));
    return v;
}
static void wr(const std::string& p, const float* d, size_t n) {
    std::ofstream f(p, std::ios::binary);
    f.write(reinterpret_cast<const char*>(d), n * sizeof(float));
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s <dir>\n", argv[0]); return 2; }
    const std::string d = std::string(argv[1]) + "/";
    try {
        nsb_config cfg; nsb_config_default(&cfg);
        cfg.mapping_pixels = 400; cfg.tracking_pixels = 300; cfg.tracking_lr = 1e-3f; cfg.tracking_iters = 2; cfg.frustum_feature_selection = 0;
        auto engine = std::make_shared<nsb::Engine>(cfg, 0);
        const char* lv[4] = {"coarse", "middle", "fine", "color"};
        nsb::Dict c;
        nsb::NICE decoders(engine, 3, 32, 32, 2.f, 0.32f, 0.16f, 0.16f, false, "fourier");
        for (int l = 0; l < 4; ++l) {
            int Z, Y, X; nsb_grid_dims(&cfg, l, &Z, &Y, &X);
            auto g = rd(d + "grid_" + lv[l] + ".bin");
            c.insert(nsb::grid_key(l), nsb::Tensor({1, 32, Z, Y, X}, g.data()));
            decoders.load(lv[l], rd(d + "dec_" + lv[l] + ".bin"));
        }
        auto tt = rd(d + "t_samples.bin"), ts = rd(d + "t_surface.bin");
        engine->check(nsb_set_ttables(engine->ctx(), tt.data(), ts.data()));
        auto ro = rd(d + "rays_o.bin"), rdv = rd(d + "rays_d.bin"), gd = rd(d + "gt_depth.bin");
        const int64_t n = (int64_t)gd.size();
        nsb::Renderer renderer(engine);
        nsb::Tensor rgb, depth, var, w;
        renderer.render_batch_ray(c, decoders, nsb::Tensor({n, 3}, rdv.data()), nsb::Tensor({n, 3}, ro.data()), "color", nsb::Tensor({n}, gd.data()), rgb, depth, var, w);
        wr(d + "out_rgb.bin", rgb.data(), rgb.numel()); wr(d + "out_depth.bin", depth.data(), depth.numel());
        wr(d + "out_var.bin", var.data(), var.numel()); wr(d + "out_weights.bin", w.data(), w.numel());
        // mapping: one frame, 3 joint iterations (Mapper::optimize_map), then the grids come back through the dict
        auto fdepth = rd(d + "frame_depth.bin"), fcolor = rd(d + "frame_color.bin"), c2w = rd(d + "c2w.bin");
        nsb::Tensor depth_t({cfg.H, cfg.W}, fdepth.data()), color_t({cfg.H, cfg.W, 3}, fcolor.data()), c2w_t({4, 4}, c2w.data());
        nsb::Mapper mapper(engine, false);
        engine->check(nsb_seed(engine->ctx(), 3));
        std::vector<float> losses;
        mapper.optimize_map(3, c, color_t, depth_t, c2w_t, c2w_t, decoders, 1.f, &losses);
        wr(d + "out_map_losses.bin", losses.data(), losses.size());
        wr(d + "out_grid_middle.bin", c.at("grid_middle").data(), (size_t)c.at("grid_middle").numel());
        // the coarse mapper (Mapper(ns, cf, coarse_mapper = true), Mapper.cpp:335-338,351-352): two iterations move grid_coarse only
        {
            const size_t ncoarse = (size_t)c.at("grid_coarse").numel(), nmid = (size_t)c.at("grid_middle").numel();
            std::vector<float> coarse0(c.at("grid_coarse").data(), c.at("grid_coarse").data() + ncoarse), mid0(c.at("grid_middle").data(), c.at("grid_middle").data() + nmid);
            nsb::Mapper coarse_mapper(engine, true);
            std::vector<float> cl;
            coarse_mapper.optimize_map(2, c, color_t, depth_t, c2w_t, c2w_t, decoders, 1.f, &cl);
            float dc = 0.f, dm = 0.f;
            for (size_t i = 0; i < ncoarse; ++i) dc = std::max(dc, std::fabs(c.at("grid_coarse").data()[i] - coarse0[i]));
            for (size_t i = 0; i < nmid; ++i) dm = std::max(dm, std::fabs(c.at("grid_middle").data()[i] - mid0[i]));
            const float cm[4] = {cl[0], cl[1], dc, dm};
            wr(d + "out_coarse_mapper.bin", cm, 4);
        }
        // tracking: Tracker::run (2 iterations)
        nsb::Tracker tracker(engine, c);
        engine->check(nsb_seed(engine->ctx(), 5));
        std::vector<float> tl;
        nsb::Tensor cam = tracker.run(decoders, color_t, depth_t, c2w_t, 0, &tl);
        wr(d + "out_trk_losses.bin", tl.data(), tl.size());
        wr(d + "out_cam.bin", cam.data(), 7);
        // Mapper::run over a short sequence (Mapper.cpp:493-552): every frame becomes a keyframe, so from the sixth frame on
        // bundle adjustment is active (keyframes > 4, :530), the overlap selection picks the window and poses are written back
        {
            nsb_config cfg2 = cfg; cfg2.mapping_iters = 6; cfg2.mapping_iters_first = 6; cfg2.keyframe_every = 1; cfg2.max_frames = 8; cfg2.mapping_window_size = 5;
            cfg2.middle_iter_ratio = 0.2f; cfg2.fine_iter_ratio = 0.2f; cfg2.BA_cam_lr = 0.002f;
            auto engine2 = std::make_shared<nsb::Engine>(cfg2, 0);
            nsb::Dict c2; nsb::NICE dec2(engine2);
            for (int l = 0; l < 4; ++l) {
                int Z, Y, X; nsb_grid_dims(&cfg2, l, &Z, &Y, &X);
                auto g = rd(d + "grid_" + lv[l] + ".bin");
                c2.insert(nsb::grid_key(l), nsb::Tensor({1, 32, Z, Y, X}, g.data()));
                dec2.load(lv[l], rd(d + "dec_" + lv[l] + ".bin"));
            }
            engine2->check(nsb_set_ttables(engine2->ctx(), tt.data(), ts.data()));
            engine2->check(nsb_seed(engine2->ctx(), 9));
            nsb::Mapper mapper2(engine2, false);
            const int n_imgs = 7;
            std::vector<nsb::Tensor> est;
            for (int i = 0; i < n_imgs; ++i) { nsb::Tensor p = c2w_t.clone(); p.data()[3] += 0.01f * (float)i; est.push_back(p); }
            std::vector<float> ba_flags, pose_delta;
            for (int i = 0; i < n_imgs; ++i) {
                nsb::Tensor before = est[(size_t)i].clone();
                mapper2.run(dec2, c2, est, color_t, depth_t, est[(size_t)i], i, n_imgs);
                float dmax = 0.f;
                for (int k = 0; k < 12; ++k) dmax = std::max(dmax, std::fabs(est[(size_t)i].data()[k] - before.data()[k]));
                ba_flags.push_back(mapper2.BA ? 1.f : 0.f); pose_delta.push_back(dmax);
            }
            std::vector<float> seq = {(float)mapper2.keyframes().size()};
            seq.insert(seq.end(), ba_flags.begin(), ba_flags.end()); seq.insert(seq.end(), pose_delta.begin(), pose_delta.end());
            wr(d + "out_run_seq.bin", seq.data(), seq.size());
        }
        // error behaviour: exceptions, like the reference's c10::Error
        bool threw = false;
        try { renderer.render_batch_ray(c, decoders, nsb::Tensor({n, 3}, rdv.data()), nsb::Tensor({n, 3}, ro.data()), "bogus", nsb::Tensor(), rgb, depth, var, w); }
        catch (const std::runtime_error&) { threw = true; }
        std::printf("ok n=%lld map_loss0=%.4f trk_loss0=%.4f threw=%d\n", (long long)n, losses[0], tl[0], (int)threw);
        return threw ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
