// g3d_tdf_exact.cu -- placeholder until the brute-force kernel lands.
#include "g3d_internal.h"
namespace g3d {
size_t tdf_exact_workspace_bytes(int32_t, int64_t, const g3d_tdf_params&) { return 0; }
int tdf_exact_launch(const float*, const int64_t*, int64_t, const uint32_t*, const int64_t*, int64_t, int32_t,
                     const g3d_tdf_params&, const g3d_tdf_outputs&, void*, size_t, cudaStream_t)
{
    set_error("exact mode not built");
    return G3D_ERR_INVALID;
}
}
