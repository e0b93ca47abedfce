// mas_fused_wide2.cu -- the no-noise fused kernel for 256 < S <= 1024 with 2 single-role DP warps per team
// (role code in mas_cost_tc.cuh / mas_dp.cuh, kernel template in mas_fused_body.cuh).
#include "mas_fused_body.cuh"

namespace mas {

const void *fused_pair_kernel_wide2(int C, int R)
{
#define MAS_WIDE_CASE(CC, RR) \
    if (C == CC && R == RR) return (const void *)mas_fused_pair_kernel<CC, RR, 2, false>;
    MAS_WIDE_CASE(5, 16) MAS_WIDE_CASE(6, 16) MAS_WIDE_CASE(7, 16) MAS_WIDE_CASE(8, 16)
    MAS_WIDE_CASE(5, 32) MAS_WIDE_CASE(6, 32) MAS_WIDE_CASE(7, 32) MAS_WIDE_CASE(8, 32)
#undef MAS_WIDE_CASE
    return nullptr;
}

}  // namespace mas
