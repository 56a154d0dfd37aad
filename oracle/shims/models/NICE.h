// TEST INFRASTRUCTURE (oracle build only) -- not part of the product path.
//
// Stand-in for the reference's include/models/NICE.h, used ONLY to build
// oracle/_ref (the reference's own src/Renderer.cpp + include/torchlib/utils.h
// compiled where they lie).  The reference's NICE::forward executes four
// TorchScript files that are not in its tree (src/models/NICE.cpp:8-11), so the
// decoder arithmetic is restated here in libtorch from the architecture the
// reference declares:
//   * layer shapes                       src/models/MLP.cpp:14-46, 104-138
//   * intended forward loop              src/models/MLP.cpp:76-102, 165-181
//     (upstream cvg/nice-slam semantics: h = relu(W_i h) + Fc_i c, skip after
//      block 2; SURVEY.md 8-A.3 lists the transliteration defects not followed)
//   * Fourier embedding sin(p @ B)       src/models/GaussianFFT.cpp:10-15
//   * stage dispatch / occupancy sums    src/models/NICE.cpp:16-51
//   * coordinate normalisation           include/torchlib/utils.h:132-139
//   * trilinear border/align_corners     src/models/MLP.cpp:61
//
// Flat parameter layout of one decoder (shared with include/nsb.h):
//   MLP:        B[3][E] | for i<5: W_i[H][K_i], b_i[H] | for i<5: Fc_i[H][C], bc_i[H] | Wo[O][H], bo[O]
//               K = {E, H, H, E+H, H}; skip input order is cat(e, h)
//   MLP_no_xyz: for i<5: W_i[H][K_i], b_i[H] | Wo[1][H], bo[1];  K = {C, H, H, C+H, H}; skip = cat(c, h)
#pragma once
#include <torch/torch.h>
#include <string>
#include <vector>
#include "torchlib/utils.h"

struct NsbRefDecoder {
    bool no_xyz = false;
    int E = 93, H = 32, C = 32, O = 1;
    torch::Tensor B;
    std::vector<torch::Tensor> W, b, Fc, bc;
    torch::Tensor Wo, bo;

    static int64_t count(bool no_xyz, int E, int H, int C, int O) {
        int64_t n = 0;
        if (!no_xyz) {
            int K[5] = {E, H, H, E + H, H};
            n += 3 * E;
            for (int i = 0; i < 5; ++i) n += H * K[i] + H;
            n += 5 * (H * C + H);
        } else {
            int K[5] = {C, H, H, C + H, H};
            for (int i = 0; i < 5; ++i) n += H * K[i] + H;
        }
        return n + O * H + O;
    }

    // Takes views of one contiguous flat tensor so that flat.grad() is the flat gradient.
    void bind(torch::Tensor flat, bool no_xyz_, int E_, int H_, int C_, int O_) {
        no_xyz = no_xyz_; E = E_; H = H_; C = C_; O = O_;
        W.clear(); b.clear(); Fc.clear(); bc.clear();
        int64_t off = 0;
        auto take = [&](int64_t r, int64_t c) {
            auto t = flat.narrow(0, off, r * c).view({r, c}); off += r * c; return t; };
        auto take1 = [&](int64_t r) { auto t = flat.narrow(0, off, r); off += r; return t; };
        if (!no_xyz) {
            int K[5] = {E, H, H, E + H, H};
            B = take(3, E);
            for (int i = 0; i < 5; ++i) { W.push_back(take(H, K[i])); b.push_back(take1(H)); }
            for (int i = 0; i < 5; ++i) { Fc.push_back(take(H, C)); bc.push_back(take1(H)); }
        } else {
            int K[5] = {C, H, H, C + H, H};
            for (int i = 0; i < 5; ++i) { W.push_back(take(H, K[i])); b.push_back(take1(H)); }
        }
        Wo = take(O, H); bo = take1(O);
        TORCH_CHECK(off == flat.numel(), "decoder flat size mismatch: ", off, " vs ", flat.numel());
    }
};

struct NICE {
    NsbRefDecoder coarse, middle, fine, color;
    torch::Tensor bound;  // (3,2), Renderer.cpp:15 values

    NICE() { bound = torch::tensor({{-4.5, 3.82}, {-1.5, 2.02}, {-3.0, 2.76}}); }

    // utils.h:132-139 (intent) followed by MLP.cpp:58-62 (intent: return the sampled features).
    torch::Tensor sample_grid_feature(torch::Tensor p, torch::Tensor grid) const {
        namespace F = torch::nn::functional;
        auto lo = bound.index({torch::indexing::Slice(), 0});
        auto hi = bound.index({torch::indexing::Slice(), 1});
        auto pn = ((p.reshape({-1, 3}) - lo) / (hi - lo)) * 2 - 1;
        auto vgrid = pn.unsqueeze(0).unsqueeze(2).unsqueeze(2);  // (1,P,1,1,3)
        auto c = F::grid_sample(grid, vgrid,
            F::GridSampleFuncOptions().mode(torch::kBilinear).padding_mode(torch::kBorder).align_corners(true));
        return c.squeeze(-1).squeeze(-1).transpose(1, 2).squeeze(0);  // (P,C)
    }

    torch::Tensor run_mlp(const NsbRefDecoder& d, torch::Tensor p, torch::Tensor c) const {
        auto e = torch::sin(torch::matmul(p.reshape({-1, 3}), d.B));  // GaussianFFT.cpp:12-14
        auto h = e;
        for (int i = 0; i < 5; ++i) {
            h = torch::relu(torch::linear(h, d.W[i], d.b[i]));
            h = h + torch::linear(c, d.Fc[i], d.bc[i]);
            if (i == 2) h = torch::cat({e, h}, -1);
        }
        return torch::linear(h, d.Wo, d.bo);
    }

    torch::Tensor run_mlp_no_xyz(const NsbRefDecoder& d, torch::Tensor c) const {
        auto h = c;
        for (int i = 0; i < 5; ++i) {
            h = torch::relu(torch::linear(h, d.W[i], d.b[i]));
            if (i == 2) h = torch::cat({c, h}, -1);
        }
        return torch::linear(h, d.Wo, d.bo);
    }

    torch::Tensor middle_occ(torch::Tensor p, c10::Dict<std::string, torch::Tensor>& g) const {
        return run_mlp(middle, p, sample_grid_feature(p, g.at("grid_middle"))).squeeze(-1);
    }
    torch::Tensor fine_occ(torch::Tensor p, c10::Dict<std::string, torch::Tensor>& g) const {
        auto c = sample_grid_feature(p, g.at("grid_fine"));
        torch::Tensor cm;
        { torch::NoGradGuard ng; cm = sample_grid_feature(p, g.at("grid_middle")); }  // MLP.cpp:79-84
        return run_mlp(fine, p, torch::cat({c, cm}, 1)).squeeze(-1);
    }

    // NICE.cpp:16-51
    torch::Tensor forward(torch::Tensor p, c10::Dict<std::string, torch::Tensor> g, std::string stage) {
        p = p.squeeze(0);
        if (stage == "coarse") {
            auto occ = run_mlp_no_xyz(coarse, sample_grid_feature(p, g.at("grid_coarse"))).squeeze(-1);
            auto z = torch::zeros({occ.size(0), 3});
            return torch::cat({z, occ.unsqueeze(-1)}, -1);
        } else if (stage == "middle") {
            auto occ = middle_occ(p, g);
            auto z = torch::zeros({occ.size(0), 3});
            return torch::cat({z, occ.unsqueeze(-1)}, -1);
        } else if (stage == "fine") {
            auto occ = fine_occ(p, g) + middle_occ(p, g);
            auto z = torch::zeros({occ.size(0), 3});
            return torch::cat({z, occ.unsqueeze(-1)}, -1);
        } else {
            auto f = fine_occ(p, g);
            auto raw = run_mlp(color, p, sample_grid_feature(p, g.at("grid_color")));
            auto m = middle_occ(p, g);
            // NICE.cpp:49 overwrites channel 3; building it with cat keeps the same values and the same
            // (zero) gradient into the colour decoder's 4th output.
            return torch::cat({raw.index({torch::indexing::Slice(), torch::indexing::Slice(0, 3)}),
                               (f + m).unsqueeze(-1)}, -1);
        }
    }
};
