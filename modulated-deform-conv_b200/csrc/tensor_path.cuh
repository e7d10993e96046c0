// tensor_path.cuh -- interface of the tcgen05 / TMEM path (tensor_path.cu).
#pragma once
#include "common.cuh"

namespace mdconv {

// True when the tcgen05 kernels can take this problem (dtype f16/bf16 and the channel / tile
// divisibility the UMMA tiling needs); everything else runs on the generic path.
bool tensor_path_supported(const DevGeom& g, int dtype, int backward);
size_t tensor_workspace_bytes(const DevGeom& g, int dtype, int backward);
void tensor_forward(const DevGeom& g, int dtype, const void* x, const void* w, const void* bias, const void* off,
                    const void* mask, void* y, void* ws, LaunchCtx& ctx);
// saved: NULL, or the first tensor_saved_bytes() bytes of the workspace the matching tensor_forward call used
// (channels-last input + re-ordered weight): the backward then skips those two layout passes.
void tensor_backward(const DevGeom& g, int dtype, const void* x, const void* w, const void* off, const void* mask,
                     const void* go, void* gx, void* gw, void* gb, void* goff, void* gmask, void* ws, const void* saved,
                     LaunchCtx& ctx);
size_t tensor_saved_bytes(const DevGeom& g, int dtype);

}  // namespace mdconv
