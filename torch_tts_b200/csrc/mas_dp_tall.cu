// mas_dp_tall.cu -- the standalone MAS kernel for two DP warps with 32-row chunks at 256 < S <= 512 (the layout used
// once decision bits / hops are spilled to the workspace anyway); role code in mas_dp.cuh.
#include "mas_dp_launch.cuh"

namespace mas {

int dp_dispatch_tall(const DpPlan &pl, int C, cudaStream_t stream)
{
    const DpParams &p = pl.p;
#define MAS_DP_CASE(CC, RR, WW) \
    if (C == CC && p.R == RR && p.W == WW) return launch_dp_c<CC, RR, WW>(pl, stream);
    MAS_DP_CASE(5, 32, 2) MAS_DP_CASE(6, 32, 2) MAS_DP_CASE(7, 32, 2) MAS_DP_CASE(8, 32, 2)
#undef MAS_DP_CASE
    return kDpNoCase;
}

}  // namespace mas
