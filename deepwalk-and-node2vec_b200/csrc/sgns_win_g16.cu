// Window-resident SGNS kernel, 16 lanes per centre (32 < emb <= 64: two centres per warp), with and without the hot-row cache.
#include "sgns_win.cuh"

namespace se {
int launch_win_g16(const SgnsArgs &a, cudaStream_t stream) {
    int rc = SE_ERR_UNSUPPORTED;
    if (a.hot_rows > 0) rc = launch_win_t<16, false, true>(a, stream);
    if (rc == SE_ERR_UNSUPPORTED) rc = launch_win_t<16, false, false>(a, stream);
    return rc;
}
}  // namespace se
