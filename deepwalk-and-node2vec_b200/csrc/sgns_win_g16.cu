// Window-resident SGNS kernel, 16 lanes per centre (32 < emb <= 64: two centres per warp).
#include "sgns_win.cuh"

namespace se {
int launch_win_g16(const SgnsArgs &a, cudaStream_t stream) { return launch_win_t<16, false>(a, stream); }
}  // namespace se
