// tensor_path.cu -- placeholder until the tcgen05 kernels land (next commit).
#include "tensor_path.cuh"
namespace mdconv {
bool tensor_path_supported(const DevGeom&, int, int) { return false; }
size_t tensor_workspace_bytes(const DevGeom&, int, int) { return 0; }
void tensor_forward(const DevGeom&, int, const void*, const void*, const void*, const void*, const void*, void*,
                    void*, LaunchCtx&) {}
void tensor_backward(const DevGeom&, int, const void*, const void*, const void*, const void*, const void*, void*,
                     void*, void*, void*, void*, void*, LaunchCtx&) {}
}  // namespace mdconv
