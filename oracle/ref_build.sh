#!/bin/bash
# Compile the reference's own SPFrontend (network half of the hot path) from the
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
