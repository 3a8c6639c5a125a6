// tsg_f32x2.cuh -- packed fp32 pair arithmetic of sm_100 (SASS FFMA2): one instruction issues two independent IEEE
// fma.rn.f32, each half rounded exactly like a scalar fmaf, so kernels that must reproduce the reference's sequence of
// roundings can use it.  What it buys is issue slots: an FFMA2 keeps the FMA pipe busy for two cycles but takes one slot.
#pragma once

namespace tsg {

__device__ __forceinline__ float2 ffma2(float2 x, float2 w, float2 acc) {
    float2 r;
    asm("{ .reg .b64 a, b, c, d; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; mov.b64 c, {%6, %7}; fma.rn.f32x2 d, a, b, c; mov.b64 {%0, %1}, d; }"
        : "=f"(r.x), "=f"(r.y)
        : "f"(x.x), "f"(x.y), "f"(w.x), "f"(w.y), "f"(acc.x), "f"(acc.y));
    return r;
}

}  // namespace tsg
