// SuperPoint weight-file readers (host only, no torch).
//
// The reference loads `common::model_path` with torch::load
// (orb_slam2/src/cv/sp_extractor.cpp:355).  The shipped file
// (orb_ros/data/models/superpoint.pt) is a PyTorch-1.0 "legacy" TorchScript
// archive: a zip with *stored* entries; <root>/model.json lists the tensors
// (dims, key "tensors/<id>") and the parameters of every sub-module
// ("conv1a" .. "convDb": {bias, weight} -> tensorId).  Modern libtorch refuses
// that format, so it is parsed directly.  `.spw` is this repo's flat container
// (see oracle/weights.py for the layout).
#pragma once
#include <map>
#include <string>
#include <vector>

namespace spfe {

struct HostTensor {
  std::vector<int> dims;
  std::vector<float> data;
  size_t numel() const { size_t n = 1; for (int d : dims) n *= static_cast<size_t>(d); return n; }
};

using WeightMap = std::map<std::string, HostTensor>;  // "conv1a.weight" -> OIHW fp32

// Returns true on success; on failure `err` holds the reason.
bool load_weights(const std::string &path, WeightMap &out, std::string &err);

}  // namespace spfe
