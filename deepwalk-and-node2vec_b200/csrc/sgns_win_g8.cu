// Window-resident SGNS kernel, 8 lanes per centre (16 <= emb <= 32: four centres per warp).
#include "sgns_win.cuh"

namespace se {
int launch_win_g8(const SgnsArgs &a, cudaStream_t stream) { return launch_win_t<8, false>(a, stream); }
}  // namespace se
