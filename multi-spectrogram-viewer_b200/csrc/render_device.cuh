// render_device.cuh -- device-side pieces shared by the render kernels (render_kernel.cu: FP32 paths,
// render_tc_kernel.cu: the tcgen05 path): the colour map of display.rs:24-42 and the per-pass clamp of image 0.23.
#pragma once
#include <cuda_runtime.h>

namespace sgx {
namespace {

// the clamp of image's resize for values that cannot be NaN (sums of finite products): branch-free
__device__ __forceinline__ float clamp_fin(float v) { return fminf(fmaxf(v, 0.0f), 3.4028235e38f); }

// display.rs:24-42 with the colour map stored per channel as (a + 0.5, b - a) for the stops a = stop i,
// b = stop i+1:  round(ratio*b + (1-ratio)*a) = floor(a + 0.5 + ratio*(b - a)), one FMA and one
// conversion per channel.  The single rounding of the FMA can differ from the reference's three only when
// the exact value lies within ~3e-5 of a rounding boundary (about 1 byte in 10^4, by 1 LSB).
// The table lives in constant memory, as floats: {r.a, r.d, g.a, g.d} {b.a, b.d, -, -} per segment.  An indexed constant
// load replays once per DISTINCT index in the warp (neighbouring pixels of a row mostly share one or two segments) and
// does not touch the shared-memory / L1 data pipe that bounds the render kernels (C5 K3 3.92 -> 3.69 ms against a
// 16-byte shared-memory entry per pixel, which was a sixth of the kernel's wavefronts).
struct CmSeg { float4 rg, b; };
__constant__ CmSeg kCmConst[9] = {
#define SGX_SEG(r0, g0, b0, r1, g1, b1) {{r0 + 0.5f, (float)(r1 - r0), g0 + 0.5f, (float)(g1 - g0)}, {b0 + 0.5f, (float)(b1 - b0), 0.0f, 0.0f}}
    SGX_SEG(0, 0, 4, 27, 12, 65), SGX_SEG(27, 12, 65, 74, 12, 107), SGX_SEG(74, 12, 107, 120, 28, 109),
    SGX_SEG(120, 28, 109, 165, 44, 96), SGX_SEG(165, 44, 96, 207, 68, 70), SGX_SEG(207, 68, 70, 237, 105, 37),
    SGX_SEG(237, 105, 37, 251, 155, 6), SGX_SEG(251, 155, 6, 247, 209, 61), SGX_SEG(247, 209, 61, 252, 255, 164)
#undef SGX_SEG
};
__device__ __forceinline__ unsigned grey_to_rgba_const(float x)
{
    const float position = __fmul_rn(10.0f, x);
    const float fl = floorf(position);
    const int idx = min(__float2int_rz(fl), 8);
    const float ratio = __fsub_rn(position, fl);
    const float4 rg = kCmConst[idx].rg;
    const float4 b = kCmConst[idx].b;
    const unsigned cr = __float2uint_rd(fmaf(ratio, rg.y, rg.x));
    const unsigned cg = __float2uint_rd(fmaf(ratio, rg.w, rg.z));
    const unsigned cb = __float2uint_rd(fmaf(ratio, b.y, b.x));
    const unsigned px = cr | (cg << 8) | (cb << 16) | 0xff000000u;
    return fl < 9.0f ? px : 0xffa4fffcu; // index >= len-1 -> (252, 255, 164)
}

} // namespace
} // namespace sgx
