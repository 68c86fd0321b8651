// so3d_common.cuh -- host-side helpers shared by the translation units of libso3d (defined in so3d_kernels.cu).
#pragma once
namespace so3d_host {
int fail(int code, const char* what);   // records the thread-local message returned by so3d_last_error(); returns code
int check_launch(const char* name);     // cudaGetLastError() -> 0 or the cudaError_t (message recorded)
int sm_count();                         // SM count of the current device (cached per device)
int current_device();                   // cudaGetDevice(), clamped to [0, 64): index of the per-device caches
}  // namespace so3d_host
