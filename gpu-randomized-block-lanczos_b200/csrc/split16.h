// Pre-split Krylov slab rows ("split16"): the mixed-precision Krylov buffer can hold every fp32 value x as the two
// f16 terms of x*S = hi + lo (S a power of two fixed per solve, |x| <= 1 for orthonormal columns) instead of the
// fp32 bit pattern - same 4 bytes per element, ~22 significant bits.  The tensor-core reorth kernels
// (reorth_tc16.cu) consume exactly these two terms, so storing them removes the split arithmetic from the two
// kernels that stream the whole slab every second step; the producers (one block per step) pay it once.
// Row layout of a stored block: [B halfs hi | B halfs lo], i.e. B 32-bit words per row; word p holds columns
// 2p, 2p+1 of hi (p < B/2) or of lo (p >= B/2).
#pragma once
#include <cuda_fp16.h>

namespace rbl {

// two fp32 values -> packed (hi, hi) and (lo, lo) f16x2 words of x*scale; element 0 in the low half
__device__ __forceinline__ void split_h2(float x0, float x1, float scale, unsigned& hi, unsigned& lo) {
    const float s0 = x0 * scale, s1 = x1 * scale;
    const __half2 h = __floats2half2_rn(s0, s1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
    hi = *reinterpret_cast<const unsigned*>(&h);
    lo = *reinterpret_cast<const unsigned*>(&l);
}

// inverse of split_h2 (exact: hi + lo has at most 23 significant bits)
__device__ __forceinline__ float2 join_h2(unsigned hi, unsigned lo, float inv_scale) {
    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&lo));
    return make_float2((h.x + l.x) * inv_scale, (h.y + l.y) * inv_scale);
}

}  // namespace rbl
