// Implementation of the drop-in classes declared in sp_extractor.h / sp_matcher.h.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "sp_extractor.h"
#include "sp_matcher.h"
#include "optimizer_dust.h"

namespace orbslam {

namespace {
// IEEE binary16 -> binary32 (subnormals and infinities included)
inline float half_to_float(uint16_t h) {
  const uint32_t sign = static_cast<uint32_t>(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {
      int e = -1;
      uint32_t m = man;
      do { e++; m <<= 1; } while (!(m & 0x400u));
      bits = sign | static_cast<uint32_t>(127 - 15 - e) << 23 | (m & 0x3FFu) << 13;
    }
  } else if (exp == 31) bits = sign | 0x7F800000u | man << 13;
  else bits = sign | (exp + 127 - 15) << 23 | man << 13;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
}  // namespace

#ifndef SPFE_WITH_ORBSLAM_CONFIG
namespace camera { int width = 752, height = 480; }
namespace common { std::string model_path; }
#endif

const float SPMatcher::TH_HIGH = 0.7f;  // sp_matcher.cpp:18
const float SPMatcher::TH_LOW = 0.3f;   // sp_matcher.cpp:19
const int SPMatcher::HISTO_LENGTH = 30; // sp_matcher.cpp:20

// Scale-pyramid bookkeeping of the reference base class (base_extractor.h:13-49).
// SuperPoint uses one level with factor 1.0 (sp_extractor.cpp:343).
BaseExtractor::BaseExtractor(int nfeatures_, float scale_factor, int nlevels_, int ini_th_fast, int min_th_fast)
    : nfeatures(nfeatures_), scaleFactor(scale_factor), nlevels(nlevels_), iniThFAST(ini_th_fast), minThFAST(min_th_fast) {
  mvScaleFactor.assign(nlevels, 1.0f);
  mvLevelSigma2.assign(nlevels, 1.0f);
  for (int l = 1; l < nlevels; l++) {
    mvScaleFactor[l] = static_cast<float>(mvScaleFactor[l - 1] * scaleFactor);
    mvLevelSigma2[l] = mvScaleFactor[l] * mvScaleFactor[l];
  }
  mvInvScaleFactor.resize(nlevels);
  mvInvLevelSigma2.resize(nlevels);
  for (int l = 0; l < nlevels; l++) {
    mvInvScaleFactor[l] = 1.0f / mvScaleFactor[l];
    mvInvLevelSigma2[l] = 1.0f / mvLevelSigma2[l];
  }
  mvImagePyramid.resize(nlevels);
  mnFeaturesPerLevel.assign(nlevels, 0);
  // base_extractor.h:35-47, same arithmetic: factor = 1.0f / (double) scaleFactor narrowed to float, cvRound = round to
  // nearest even (lrint).  With one level the quotient is 0 / 0 and never used, as upstream.
  const float factor = static_cast<float>(1.0f / scaleFactor);
  float per_level = nfeatures * (1 - factor) / (1 - static_cast<float>(std::pow(static_cast<double>(factor), static_cast<double>(nlevels))));
  int used = 0;
  for (int l = 0; l < nlevels - 1; l++) {
    mnFeaturesPerLevel[l] = static_cast<int>(std::lrint(per_level));
    used += mnFeaturesPerLevel[l];
    per_level *= factor;
  }
  mnFeaturesPerLevel[nlevels - 1] = nfeatures - used > 0 ? nfeatures - used : 0;
}

SPExtractor::SPExtractor(int nfeatures_) : BaseExtractor(nfeatures_, 1.0f, 1, 1, 1), num_feature_(nfeatures_) {
  spfe_config cfg;
  spfe_default_config(&cfg, camera::height, camera::width, nfeatures_);
  cfg.weights_path = common::model_path.c_str();
  // Opt-in throughput mode: the two largest outputs shrink.  Defaults = the reference's behaviour.
  const char *lz = getenv("SPFE_SHIM_LAZY_HEAT"), *h16 = getenv("SPFE_SHIM_DESC_F16");
  lazy_heat_ = lz && lz[0] == '1';
  desc_f16_ = h16 && h16[0] == '1';
  if (lazy_heat_) cfg.flags = (cfg.flags & ~(SPFE_EMIT_HEAT | SPFE_EMIT_HEAT_INV)) | SPFE_LAZY_HEAT;
  if (desc_f16_) cfg.flags |= SPFE_DESC_F16;
  if (spfe_create(&cfg, &ctx_) != SPFE_OK) throw std::runtime_error(std::string("SPExtractor: ") + spfe_last_error(nullptr));
  // The matcher and the optimiser share the FIRST extractor's device context (the reference has exactly one: the "ini"
  // extractor aliases it, tracker.cpp:131,144); a second extractor (another camera) must not silently redirect them.
  if (!SPMatcher::Backend()) SPMatcher::SetBackend(ctx_);
  if (!Optimizer::DustBackend()) Optimizer::SetBackend(ctx_);
}

SPExtractor::~SPExtractor() {
  // nobody may keep calling into a destroyed context
  if (SPMatcher::Backend() == ctx_) SPMatcher::SetBackend(nullptr);
  if (Optimizer::DustBackend() == ctx_) Optimizer::SetBackend(nullptr);
  spfe_destroy(ctx_);
}

void SPExtractor::operator()(cv::InputArray image_, cv::InputArray /*mask*/, std::vector<cv::KeyPoint> &keypoints,
                             cv::OutputArray descriptors) {
  const cv::Mat &image = image_.getMat();
  if (image.empty()) throw std::runtime_error("input image is empty");  // sp_extractor.cpp:364-365
  if (image.type() != CV_8UC1) throw std::runtime_error("SPExtractor expects CV_8UC1");  // assert at :368
  if (image.rows != camera::height || image.cols != camera::width) throw std::runtime_error("SPExtractor: image size differs from camera::height/width");
  spfe_frame_out o;
  const int rc = spfe_extract(ctx_, image.data, image.step, &o);
  if (rc == SPFE_ERR_EMPTY) throw std::runtime_error("input image is empty");
  if (rc != SPFE_OK) throw std::runtime_error(spfe_last_error(ctx_));
  const int hc = image.rows / 8, wc = image.cols / 8;
  auto wrap = [](int r, int c, int type, const void *src) { return cv::Mat(r, c, type, const_cast<void *>(src)).clone(); };
  semi_dust_ = wrap(hc, wc, CV_32FC1, o.semi_dust);
  dense_dust_ = wrap(hc, wc, CV_32FC1, o.dense_dust);
  occ_grid_ = wrap(hc, wc, CV_16SC1, o.occ_grid);
  if (lazy_heat_) {
    heat_ = cv::Mat();  // fetched by getHeatMap() / getHeatInv() if somebody asks
    heat_inv_ = cv::Mat();
  } else {
    heat_ = wrap(image.rows, image.cols, CV_32FC1, o.heat);
    heat_inv_ = wrap(image.rows, image.cols, CV_32FC1, o.heat_inv);
  }
  keypoints.clear();
  keypoints.reserve(o.n);
  cov2_.clear();
  cov2_inv_.clear();
  for (int i = 0; i < o.n; i++) {
    cv::KeyPoint kp(o.kp_xy[2 * i], o.kp_xy[2 * i + 1], 1.0f);  // size 1, angle -1, octave 0 (sp_extractor.cpp:231-232)
    kp.response = o.kp_response[i];                              // sp_extractor.cpp:271
    keypoints.push_back(kp);
    cov2_.emplace_back(o.cov2[2 * i], o.cov2[2 * i + 1]);
    cov2_inv_.emplace_back(o.cov2_inv[2 * i], o.cov2_inv[2 * i + 1]);
  }
  descriptors.create(o.n, SPFE_DESC_DIM, CV_32FC1);  // sp_extractor.cpp:512-513
  if (o.n > 0 && !desc_f16_) memcpy(descriptors.getMat().data, o.desc, static_cast<size_t>(o.n) * SPFE_DESC_DIM * sizeof(float));
  if (o.n > 0 && desc_f16_) {  // binary16 on the wire, widened here (exact: every fp16 value is an fp32 value)
    float *dst = reinterpret_cast<float *>(descriptors.getMat().data);
    for (size_t i = 0; i < static_cast<size_t>(o.n) * SPFE_DESC_DIM; i++) dst[i] = half_to_float(o.desc_f16[i]);
  }
}

cv::Mat SPExtractor::getHeatMap() {
  if (lazy_heat_ && heat_.empty()) {
    heat_.create(camera::height, camera::width, CV_32FC1);
    if (spfe_fetch_heat(ctx_, 0, 0, reinterpret_cast<float *>(heat_.data), nullptr) != SPFE_OK) throw std::runtime_error(spfe_last_error(ctx_));
  }
  return heat_;
}

cv::Mat SPExtractor::getHeatInv() {
  if (lazy_heat_ && heat_inv_.empty()) {
    heat_inv_.create(camera::height, camera::width, CV_32FC1);
    if (spfe_fetch_heat(ctx_, 0, 0, nullptr, reinterpret_cast<float *>(heat_inv_.data)) != SPFE_OK) throw std::runtime_error(spfe_last_error(ctx_));
  }
  return heat_inv_;
}

}  // namespace orbslam
