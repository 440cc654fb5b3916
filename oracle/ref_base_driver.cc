// Prints the scale-pyramid bookkeeping of the REFERENCE's own BaseExtractor (test infrastructure only): the class
// (orb_slam2/include/orb_slam/cv/base_extractor.h:7-95, constructor inline) is extracted verbatim by oracle/ref_build.sh
// into oracle/_ref/gen/base_extractor_decl.inc and compiled against oracle/ref_cv_stub.h.
// usage: ref_base_probe nfeatures scaleFactor nlevels
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ref_cv_stub.h"

#include "base_extractor_decl.inc"

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
