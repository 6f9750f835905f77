// Window-resident SGNS kernel, 32 lanes per centre (64 < emb <= 128), per-CTA hot-row cache (SE_SGNS_HOT_ROWS).
#include "sgns_win.cuh"

namespace se {
int launch_win_g32_hot(const SgnsArgs &a, cudaStream_t stream) {
    return a.emb == 128 ? launch_win_t<32, true, true>(a, stream) : launch_win_t<32, false, true>(a, stream);
}
}  // namespace se
