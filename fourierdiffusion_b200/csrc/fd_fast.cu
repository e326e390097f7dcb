// TF32 tensor-core path (placeholder until the kernels land): reports "not supported" so the handle stays generic.
#include "fd_common.cuh"
namespace fd {
int fast_path_supported(const fd_config &) { return 0; }
int fast_finalize(fd_handle *) { return 0; }
int score_fast(fd_handle *, const float *, const float *, float *, int, cudaStream_t) {
    set_error("tensor-core path not built");
    return 1;
}
}  // namespace fd
