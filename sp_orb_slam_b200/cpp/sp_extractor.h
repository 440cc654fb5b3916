// Drop-in replacement for the reference header
// orb_slam2/include/orb_slam/cv/sp_extractor.h (+ base_extractor.h): the same
// class names, constructor, operator(), getters and public side-output members
// (reference sp_extractor.h:49-88, base_extractor.h:54-72), implemented over
// the C ABI in include/spfe.h instead of libtorch.  No torch headers.
#pragma once
#include <string>
#include <vector>

#include "mini_cv.h"
#include "spfe.h"

namespace orbslam {

// In the reference these globals live in orb_slam/config.h (config.h:11-141);
// the extractor constructor reads them (sp_extractor.cpp:354-355).
namespace camera { extern int width, height; }
namespace common { extern std::string model_path; }

class BaseExtractor {
 public:
  BaseExtractor() = delete;
  BaseExtractor(int nfeatures_, float scale_factor, int nlevels_, int ini_th_fast, int min_th_fast);
  virtual ~BaseExtractor() = default;

  // Mask is ignored, as in the reference (base_extractor.h:51-56).
  virtual void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints,
                          cv::OutputArray descriptors) = 0;

  int GetLevels() { return nlevels; }
  float GetScaleFactor() { return static_cast<float>(scaleFactor); }
  std::vector<float> GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  std::vector<cv::Mat> mvImagePyramid;

 protected:
  int nfeatures;
  double scaleFactor;
  int nlevels, iniThFAST, minThFAST;
  std::vector<int> mnFeaturesPerLevel;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

class SPExtractor : public BaseExtractor {
 public:
  explicit SPExtractor(int nfeatures);
  ~SPExtractor() override;
  SPExtractor(const SPExtractor &) = delete;
  SPExtractor &operator=(const SPExtractor &) = delete;

  void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints,
                  cv::OutputArray descriptors) override;

  cv::Mat getMask() { return mask_; }
  // Throughput mode (environment SPFE_SHIM_LAZY_HEAT=1 when the extractor is constructed): heat_ / heat_inv_ are not
  // copied on every call -- they stay on the device, where computeCovariance consumed them -- and these two getters
  // fetch them for the last frame on demand (spfe_fetch_heat, bit-identical to the eager copy).  The reference reads the
  // member directly at frame.cpp:304; in this mode that line becomes `getHeatMap()`.
  cv::Mat getHeatMap();
  cv::Mat getHeatInv();
  const std::vector<Eigen::Vector2f> getCov() { return cov2_; }
  const std::vector<Eigen::Vector2f> getCov2Inv() { return cov2_inv_; }

  // Side outputs, overwritten by every call and cloned by Frame::ExtractORB (frame.cpp:300-309).
  cv::Mat semi_dust_, dense_dust_;
  cv::Mat mask_, heat_, heat_inv_;
  cv::Mat occ_grid_;

  spfe_ctx *handle() { return ctx_; }  // lets SPMatcher share the device context

 protected:
  std::vector<Eigen::Vector2f> cov2_, cov2_inv_;
  int num_feature_;
  spfe_ctx *ctx_ = nullptr;
  bool lazy_heat_ = false, desc_f16_ = false;  // SPFE_SHIM_LAZY_HEAT / SPFE_SHIM_DESC_F16 at construction
};

}  // namespace orbslam
