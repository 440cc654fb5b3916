// Native throughput harness over the C ABI (no Python in the loop): one camera stream, `slots` batches in flight,
// frames wait in page-locked memory (spfe_host_alloc), results land in the context's pinned buffers.
// usage: stream_bench <weights> <H> <W> <batch> <slots> <steps> [full]
// Prints one JSON line: frames/s end to end (H2D + extract + covariance + match to the previous frame + D2H).
// Default = the throughput output set of bench.py (n fp16 descriptor rows per frame, heat maps stay on the device);
// "full" = every output of SPExtractor::operator() copied eagerly (fp32 descriptors, H x W heat_).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "spfe.h"

// deterministic test pattern: grey background, filled rectangles that drift by one pixel per frame
static void make_frame(uint8_t *img, int H, int W, int t) {
  memset(img, 110, static_cast<size_t>(H) * W);
  uint32_t rng = 12345u;
  auto next = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
  for (int s = 0; s < 300; s++) {
    const int w = 8 + next() % 60, h = 8 + next() % 60;
    const int x0 = static_cast<int>(next() % W) + t, y0 = static_cast<int>(next() % H);
    const uint8_t v = static_cast<uint8_t>(20 + next() % 220);
    for (int y = y0; y < y0 + h && y < H; y++)
      for (int x = x0; x < x0 + w && x < W; x++)
        if (x >= 0) img[static_cast<size_t>(y) * W + x] = v;
  }
}

int main(int argc, char **argv) {
  if (argc < 7) { fprintf(stderr, "usage: stream_bench <weights> <H> <W> <batch> <slots> <steps> [full]\n"); return 2; }
  const bool full = argc > 7 && !strcmp(argv[7], "full");
  const int H = atoi(argv[2]), W = atoi(argv[3]), B = atoi(argv[4]), S = atoi(argv[5]), steps = atoi(argv[6]);
  spfe_config cfg;
  spfe_default_config(&cfg, H, W, 800);
  cfg.weights_path = argv[1];
  cfg.max_batch = B;
  cfg.num_slots = S;
  cfg.flags = full ? (SPFE_EMIT_HEAT | SPFE_EMIT_COV | SPFE_MATCH_PREV)  // everything Frame::ExtractORB reads, eagerly
                   : (SPFE_EMIT_COV | SPFE_MATCH_PREV | SPFE_LAZY_HEAT | SPFE_DESC_F16);
  spfe_ctx *ctx = nullptr;
  if (spfe_create(&cfg, &ctx) != SPFE_OK) { fprintf(stderr, "spfe_create: %s\n", spfe_last_error(nullptr)); return 1; }
  const size_t px = static_cast<size_t>(H) * W;
  const int pool = 4;  // batches of distinct frames, reused round-robin
  uint8_t *frames = static_cast<uint8_t *>(spfe_host_alloc(static_cast<size_t>(pool) * B * px));
  if (!frames) { fprintf(stderr, "spfe_host_alloc failed\n"); return 1; }
  for (int i = 0; i < pool * B; i++) make_frame(frames + i * px, H, W, i);
  std::vector<spfe_frame_out> outs(B);
  long long kps = 0, matches = 0, d2h = 0;
  auto run = [&](int n, bool count) -> int {
    for (int i = 0; i < n + S; i++) {
      const int s = i % S;
      if (i >= S) {
        if (spfe_wait(ctx, s, outs.data()) != SPFE_OK) { fprintf(stderr, "spfe_wait: %s\n", spfe_last_error(ctx)); return 1; }
        if (count) d2h += spfe_last_d2h_bytes(ctx, s);
        if (count)
          for (int b = 0; b < B; b++) {
            kps += outs[b].n;
            for (int k = 0; k < outs[b].n; k++) matches += outs[b].match_prev[k] >= 0;
          }
      }
      if (i < n && spfe_submit_pinned(ctx, s, frames + static_cast<size_t>(i % pool) * B * px, B) != SPFE_OK) {
        fprintf(stderr, "spfe_submit_pinned: %s\n", spfe_last_error(ctx));
        return 1;
      }
    }
    return 0;
  };
  if (run(S + 2, false)) return 1;  // warm-up
  const auto t0 = std::chrono::steady_clock::now();
  if (run(steps, true)) return 1;
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("{\"frames_per_s\": %.1f, \"frames\": %d, \"ms_per_step\": %.3f, \"keypoints_per_frame\": %.1f, \"matches_per_frame\": %.1f, "
         "\"launches\": %lld, \"d2h_bytes_per_frame\": %.0f, \"outputs\": \"%s\"}\n",
         steps * B / sec, steps * B, sec / steps * 1e3, double(kps) / (steps * B), double(matches) / (steps * B),
         static_cast<long long>(spfe_launch_count(ctx)), double(d2h) / (steps * B), full ? "full" : "throughput");
  spfe_host_free(frames);
  spfe_destroy(ctx);
  return 0;
}
