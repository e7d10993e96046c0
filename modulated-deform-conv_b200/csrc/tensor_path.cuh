// tensor_path.cuh -- interface of the tcgen05 / TMEM path (tensor_path.cu).
#pragma once
#include "common.cuh"

namespace mdconv {

// True when the tcgen05 kernels can take this problem (dtype f16/bf16 and the channel / tile
// divisibility the UMMA tiling needs); everything else runs on the generic path.
bool tensor_path_supported(const DevGeom& g, int dtype, int backward);
size_t tensor_workspace_bytes(const DevGeom& g, int dtype, int backward);
// Optional extras of a forward call (mdconv_fwd_extra_t): xt = a channels-last copy of the input some earlier call
// produced (tensor_to_channels_last) -- when it is the head of `ws` itself the layout stays where the backward
// expects it; grid = regular-grid producer convolution (no offsets) with its two outputs.
struct FwdExtra {
  const void* xt = nullptr;
  int grid = 0;
  int co_valid = 0;
  void* y2 = nullptr;
  int split = 0;
};
void tensor_forward(const DevGeom& g, int dtype, const void* x, const void* w, const void* bias, const void* off,
                    const void* mask, void* y, void* ws, const FwdExtra& ex, LaunchCtx& ctx);
// x [B][C][IN] -> xt [B*IN + 1][C] (channels-last + the all-zero position), what FwdExtra::xt points at
void tensor_to_channels_last(const DevGeom& g, int dtype, const void* x, void* xt, LaunchCtx& ctx);
size_t tensor_channels_last_bytes(const DevGeom& g);
size_t tensor_weight_layout_bytes(const DevGeom& g);   // forward workspace when FwdExtra::xt lives elsewhere
// saved: NULL, or the first tensor_saved_bytes() bytes of the workspace the matching tensor_forward call used
// (channels-last input + re-ordered weight): the backward then skips those two layout passes.
void tensor_backward(const DevGeom& g, int dtype, const void* x, const void* w, const void* off, const void* mask,
                     const void* go, void* gx, void* gw, void* gb, void* goff, void* gmask, void* ws, const void* saved,
                     LaunchCtx& ctx);
size_t tensor_saved_bytes(const DevGeom& g, int dtype);

}  // namespace mdconv
