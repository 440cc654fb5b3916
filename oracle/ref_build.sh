#!/bin/bash
# Compile the reference's own SPFrontend (network half of the hot path) and its own nms / computeCovariance from the
# sources where they lie under $SPFE_REFERENCE (default /root/reference) against
# this image's libtorch (CPU).  Outputs ONLY into oracle/_ref/ (git-ignored).
# The rest of the reference (OpenCV / Eigen / ROS / g2o / Pangolin) is
# unbuildable here -- see DESIGN.md.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SPFE_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
[ -f "$REF/orb_slam2/src/cv/sp_extractor.cpp" ] || { echo "reference not present at $REF; keeping any prebuilt $OUT"; exit 0; }
mkdir -p "$OUT/gen"
sed -n '16,47p' "$REF/orb_slam2/include/orb_slam/cv/sp_extractor.h" > "$OUT/gen/spfrontend_decl.inc"
sed -n '16,159p' "$REF/orb_slam2/src/cv/sp_extractor.cpp" \
  | sed -e 's/^\( *\)\.cuda();/\1;/' -e 's/\.clone()\.cuda()/.clone()/' > "$OUT/gen/spfrontend_impl.inc"
if grep -q 'cuda()' "$OUT/gen/spfrontend_impl.inc"; then echo "unexpected .cuda() left in extracted source"; exit 1; fi
TORCH="$(python -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
g++ -O2 -std=c++17 -fPIC -shared -D_GLIBCXX_USE_CXX11_ABI=1 -w \
  -I"$OUT/gen" -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" \
  "$HERE/ref_driver.cc" -o "$OUT/libspref.so" \
  -L"$TORCH/lib" -ltorch -ltorch_cpu -lc10 -Wl,-rpath,"$TORCH/lib"
echo "built $OUT/libspref.so"
# the reference's own nms + computeCovariance (sp_extractor.cpp:161-340), verbatim, against oracle/ref_cv_stub.h
sed -n '161,340p' "$REF/orb_slam2/src/cv/sp_extractor.cpp" > "$OUT/gen/sppost_impl.inc"
g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -w -I"$OUT/gen" -I"$HERE" "$HERE/ref_post_driver.cc" -o "$OUT/libsppost_ref.so"
echo "built $OUT/libsppost_ref.so"
# the reference's own EdgeSE3ProjectDustOnlyPose (class + computeError / linearizeOplus / isInImage / getPixelValue),
# verbatim, against oracle/ref_g2o_stub.h
sed -n '22,65p' "$REF/orb_slam2/include/orb_slam/optimization/types_dust_tracking.h" > "$OUT/gen/dust_edge_decl.inc"
sed -n '36,141p' "$REF/orb_slam2/src/optimization/types_dust_tracking.cpp" > "$OUT/gen/dust_edge_impl.inc"
g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -w -I"$OUT/gen" -I"$HERE" "$HERE/ref_dust_driver.cc" -o "$OUT/libspdust_ref.so"
echo "built $OUT/libspdust_ref.so"
# the reference's own guided-search loops: Frame::GetFeaturesInArea, SPMatcher::SearchByProjection(Frame&, MapPoints),
# SearchByProjection(Cur, Last), RadiusByViewingCos, DescriptorDistance and the dust-track association block, verbatim,
# against class skeletons
sed -n '382,474p' "$REF/orb_slam2/src/type/frame.cpp" > "$OUT/gen/frame_area.inc"
sed -n '344,439p' "$REF/orb_slam2/src/cv/sp_matcher.cpp" > "$OUT/gen/matcher_proj.inc"
sed -n '1636,1640p' "$REF/orb_slam2/src/cv/sp_matcher.cpp" > "$OUT/gen/matcher_dist.inc"
sed -n '105,172p' "$REF/orb_slam2/src/tracking/tracker_dust.cpp" > "$OUT/gen/dust_assoc.inc"
sed -n '18,20p' "$REF/orb_slam2/src/cv/sp_matcher.cpp" > "$OUT/gen/matcher_const.inc"
sed -n '1439,1543p' "$REF/orb_slam2/src/cv/sp_matcher.cpp" > "$OUT/gen/matcher_last.inc"
g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -w -I"$OUT/gen" -I"$HERE" "$HERE/ref_guided_driver.cc" -o "$OUT/libspguided_ref.so"
echo "built $OUT/libspguided_ref.so"
# the reference's own SearchByBruteForce overloads, verbatim
sed -n '1642,1674p' "$REF/orb_slam2/src/cv/sp_matcher.cpp" > "$OUT/gen/bf_kf_frame.inc"
sed -n '334,376p' "$REF/orb_slam2/src/cv/sp_matcher_loop.cpp" > "$OUT/gen/bf_kf_kf.inc"
g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -w -I"$OUT/gen" -I"$HERE" "$HERE/ref_bf_driver.cc" -o "$OUT/libspbf_ref.so"
echo "built $OUT/libspbf_ref.so"
# the reference's own BaseExtractor (scale-pyramid bookkeeping the drop-in class must reproduce), verbatim
sed -n '7,95p' "$REF/orb_slam2/include/orb_slam/cv/base_extractor.h" > "$OUT/gen/base_extractor_decl.inc"
g++ -O2 -std=c++17 -ffp-contract=off -w -I"$OUT/gen" -I"$HERE" "$HERE/ref_base_driver.cc" -o "$OUT/ref_base_probe"
echo "built $OUT/ref_base_probe"
# the reference's own SearchForTriByFlann + CheckDistEpipolarLine, verbatim (FLANN itself replaced by an exact k-NN stand-in)
sed -n '183,262p' "$REF/orb_slam2/src/cv/sp_matcher.cpp" > "$OUT/gen/flann_tri.inc"
sed -n '441,469p' "$REF/orb_slam2/src/cv/sp_matcher.cpp" > "$OUT/gen/epi_check.inc"
g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -w -I"$OUT/gen" -I"$HERE" "$HERE/ref_flann_driver.cc" -o "$OUT/libspflann_ref.so"
echo "built $OUT/libspflann_ref.so"
