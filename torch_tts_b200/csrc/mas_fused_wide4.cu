// mas_fused_wide4.cu -- the no-noise fused kernel for 256 < S <= 1024 with 4 single-role DP warps per team
// (role code in mas_cost_tc.cuh / mas_dp.cuh, kernel template in mas_fused_body.cuh).
#include "mas_fused_body.cuh"

namespace mas {

const void *fused_pair_kernel_wide4(int C, int R)
{
#define MAS_WIDE_CASE(CC, RR) \
    if (C == CC && R == RR) return (const void *)mas_fused_pair_kernel<CC, RR, 4, false>;
    MAS_WIDE_CASE(5, 8) MAS_WIDE_CASE(6, 8) MAS_WIDE_CASE(7, 8) MAS_WIDE_CASE(8, 8)
    MAS_WIDE_CASE(5, 16) MAS_WIDE_CASE(6, 16) MAS_WIDE_CASE(7, 16) MAS_WIDE_CASE(8, 16)
#undef MAS_WIDE_CASE
    return nullptr;
}

}  // namespace mas
