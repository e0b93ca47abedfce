// mas_dp_wide.cu -- the standalone MAS kernel for four DP warps per team (S > 512, and the MAS_DP_WARPS=4 A/B partner
// at S <= 256); role code in mas_dp.cuh.
#include "mas_dp_launch.cuh"

namespace mas {

int dp_dispatch_wide(const DpPlan &pl, int C, cudaStream_t stream)
{
    const DpParams &p = pl.p;
#define MAS_DP_CASE(CC, RR, WW) \
    if (C == CC && p.R == RR && p.W == WW) return launch_dp_c<CC, RR, WW>(pl, stream);
    MAS_DP_CASE(2, 32, 4)   // MAS_DP_WARPS=4 at S <= 256: the A/B partner of the default
    MAS_DP_CASE(5, 8, 4) MAS_DP_CASE(6, 8, 4) MAS_DP_CASE(7, 8, 4) MAS_DP_CASE(8, 8, 4)
    MAS_DP_CASE(5, 16, 4) MAS_DP_CASE(6, 16, 4) MAS_DP_CASE(7, 16, 4) MAS_DP_CASE(8, 16, 4)   // bits / hops spilled
#undef MAS_DP_CASE
    return kDpNoCase;
}

}  // namespace mas
