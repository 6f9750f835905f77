// Window-resident SGNS kernel, 32 lanes per centre (64 < emb <= 128): the S3 bench kernel.
#include "sgns_win.cuh"

namespace se {
int launch_win_g32(const SgnsArgs &a, cudaStream_t stream) {
    return a.emb == 128 ? launch_win_t<32, true>(a, stream) : launch_win_t<32, false>(a, stream);
}
}  // namespace se
