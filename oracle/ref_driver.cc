// Driver around the REFERENCE's own SPFrontend (test infrastructure only).
// The struct declaration and its constructor / forward() are not in this
// repository: oracle/ref_build.sh extracts them at build time from
// /root/reference/orb_slam2/{include/orb_slam/cv/sp_extractor.h:16-47,
// src/cv/sp_extractor.cpp:16-159} into oracle/_ref/gen/ (git-ignored), dropping
// only the two `.cuda()` calls (:73, :134) because this container has no GPU.
// This file just feeds it weights and an image through a C entry point.
#include <torch/torch.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

using namespace std;

namespace orbslam {
#include "spfrontend_decl.inc"
#include "spfrontend_impl.inc"
}  // namespace orbslam

namespace {
std::shared_ptr<orbslam::SPFrontend> g_model;
int g_h = 0, g_w = 0;
const char *kLayers[12] = {"conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b",
                           "conv4a", "conv4b", "convPa", "convPb", "convDa", "convDb"};
}  // namespace

extern "C" {

// weights: 24 pointers, (weight, bias) per layer in kLayers order, OIHW fp32.
int spref_load(const float *const *weights, int H, int W, int threads) {
  torch::NoGradGuard ng;
  if (threads > 0) at::set_num_threads(threads);
  g_model = std::make_shared<orbslam::SPFrontend>(0.015, H, W, 8);  // sp_extractor.cpp:354
  g_model->eval();
  g_h = H; g_w = W;
  auto params = g_model->named_parameters();
  for (int l = 0; l < 12; l++) {
    for (int k = 0; k < 2; k++) {
      auto &p = params[std::string(kLayers[l]) + (k == 0 ? ".weight" : ".bias")];
      std::memcpy(p.data_ptr<float>(), weights[2 * l + k], p.numel() * sizeof(float));
    }
  }
  return 0;
}

// One frame through SPFrontend::forward, fed exactly like SPExtractor::operator() does
// (sp_extractor.cpp:374-390).  Outputs sized for `max_n` candidates; returns n or -1.
int spref_forward(const uint8_t *img, float *semi_dust, float *dense_dust, float *pixels_in, float *score,
                  float *desc, float *heat, int max_n) {
  torch::NoGradGuard ng;
  if (!g_model) return -1;
  const int H = g_h, W = g_w;
  std::vector<float> im(static_cast<size_t>(H) * W);
  const float a = static_cast<float>(static_cast<double>(1.f / 255.f));  // convertTo(CV_32FC1, 1.f / 255.f)
  for (size_t i = 0; i < im.size(); i++) im[i] = img[i] * a;
  auto x = torch::from_blob(im.data(), {1, H, W, 1}, torch::kFloat32).permute({0, 3, 1, 2});
  auto out = g_model->forward(x);
  const int n = static_cast<int>(out[2].size(1));
  if (n > max_n) return -1;
  auto cp = [](const torch::Tensor &t, float *dst) { auto c = t.contiguous(); std::memcpy(dst, c.data_ptr<float>(), c.numel() * sizeof(float)); };
  cp(out[0], semi_dust); cp(out[1], dense_dust); cp(out[2], pixels_in); cp(out[3], score); cp(out[4], desc); cp(out[5], heat);
  return n;
}
}
