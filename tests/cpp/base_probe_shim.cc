// Prints the scale-pyramid bookkeeping of the SHIM's BaseExtractor (sp_orb_slam_b200/cpp/sp_extractor.h + sp_shim.cc) in
// the format of oracle/ref_base_driver.cc, for tests/test_capi.py::test_base_extractor_matches_reference.
// usage: base_probe_shim nfeatures scaleFactor nlevels
#include <cstdio>
#include <cstdlib>

#include "sp_extractor.h"

struct Probe : orbslam::BaseExtractor {
  using orbslam::BaseExtractor::BaseExtractor;
  void operator()(cv::InputArray, cv::InputArray, std::vector<cv::KeyPoint> &, cv::OutputArray) override {}
  void dump() {
    printf("%d %.9g\n", GetLevels(), GetScaleFactor());
    for (float v : GetScaleFactors()) printf("%.9g ", v);
    printf("\n");
    for (float v : GetInverseScaleFactors()) printf("%.9g ", v);
    printf("\n");
    for (float v : GetScaleSigmaSquares()) printf("%.9g ", v);
    printf("\n");
    for (float v : GetInverseScaleSigmaSquares()) printf("%.9g ", v);
    printf("\n");
    for (int v : mnFeaturesPerLevel) printf("%d ", v);
    printf("\n%zu\n", mvImagePyramid.size());
  }
};

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  Probe p(atoi(argv[1]), (float)atof(argv[2]), atoi(argv[3]), 1, 1);
  p.dump();
  return 0;
}
