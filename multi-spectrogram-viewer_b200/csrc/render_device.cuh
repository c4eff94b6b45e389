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

// The sliding-window kernel computes the same function without conversion instructions (F2I / FRND run on the quarter-rate
// XU pipe: five per pixel were 1.0 ms of XU time per C5 step).  For 0 <= v < 2^23, v + 2^23 rounded towards minus infinity
// (FADD.RM) is 2^23 + floor(v) exactly, so the integer sits in the low mantissa bits: floor(position) for the segment
// index and the three channel bytes.  And with position = i + ratio (ratio = position - floor(position) is exact in
// f32), ratio * d + (a + 0.5) and position * d + (a + 0.5 - i d) are the SAME real number, and a + 0.5 - i d
// is a small multiple of 0.5, exact in f32: one FMA per channel straight from `position`, bit-identical to the form
// above, with neither floor nor ratio computed.  A tenth table entry (d = 0) returns the last colour for position >= 9.
struct CmPos { float4 rg, b; };
__constant__ CmPos kCmPos[10] = {
#define SGX_SEGP(i, r0, g0, b0, r1, g1, b1)                                                                              \
    {{r0 + 0.5f - (float)(i) * (float)(r1 - r0), (float)(r1 - r0), g0 + 0.5f - (float)(i) * (float)(g1 - g0), (float)(g1 - g0)}, \
     {b0 + 0.5f - (float)(i) * (float)(b1 - b0), (float)(b1 - b0), 0.0f, 0.0f}}
    SGX_SEGP(0, 0, 0, 4, 27, 12, 65), SGX_SEGP(1, 27, 12, 65, 74, 12, 107), SGX_SEGP(2, 74, 12, 107, 120, 28, 109),
    SGX_SEGP(3, 120, 28, 109, 165, 44, 96), SGX_SEGP(4, 165, 44, 96, 207, 68, 70), SGX_SEGP(5, 207, 68, 70, 237, 105, 37),
    SGX_SEGP(6, 237, 105, 37, 251, 155, 6), SGX_SEGP(7, 251, 155, 6, 247, 209, 61), SGX_SEGP(8, 247, 209, 61, 252, 255, 164),
    SGX_SEGP(0, 252, 255, 164, 252, 255, 164)
#undef SGX_SEGP
};
// t: the un-clamped sum of the last resampling pass.  __saturatef is the clamp at 0 of image's resize; values above 1
// all map to the last colour, as do 1 and everything from 0.9 on (display.rs:31), so clipping them to 1 changes nothing.
// Bit-identical to grey_to_rgba_const(clamp(t)) for every float (tools/check_colormap_fadd.cpp: every value in [0, 2]).
// The lane that calls it walks along an image row: neighbouring pixels mostly fall into the same colour segment,
// so the segment's six constants stay in registers and are re-loaded (predicated, per lane) only when the segment
// changes.  An indexed constant load costs one request per DISTINCT index among the lanes that execute it: with the
// lanes of a warp on 32 different rows that was ~3 requests per load and pixel (ncu: the indexed constant cache busy for
// half of the kernel's time); lanes whose segment did not change make no request.  Measured effect on the kernel time:
// none (3.55 -> 3.54 ms per C5 step) -- it was not the limiter; kept because it costs nothing.
struct CmCache { float4 rg; float2 b; int idx; };
__device__ __forceinline__ void cm_cache_init(CmCache &c) { c.idx = -1; c.rg = make_float4(0.f, 0.f, 0.f, 0.f); c.b = make_float2(0.f, 0.f); }
__device__ __forceinline__ unsigned grey_to_rgba_cached(float t, CmCache &c)
{
    const float kMagic = 8388608.0f; // 2^23
    const float position = __fmul_rn(10.0f, __saturatef(t)); // <= 10
    const unsigned ti = __float_as_uint(__fadd_rd(position, kMagic)) & 0xfu; // floor(position), 0 .. 10
    const int idx = min((int)ti, 9);
    if (idx != c.idx) {
        c.rg = kCmPos[idx].rg;
        const float4 b = kCmPos[idx].b;
        c.b = make_float2(b.x, b.y);
        c.idx = idx;
    }
    const unsigned cr = __float_as_uint(__fadd_rd(fmaf(position, c.rg.y, c.rg.x), kMagic));
    const unsigned cg = __float_as_uint(__fadd_rd(fmaf(position, c.rg.w, c.rg.z), kMagic));
    const unsigned cb = __float_as_uint(__fadd_rd(fmaf(position, c.b.y, c.b.x), kMagic));
    return __byte_perm(__byte_perm(cr, cg, 0x0040), cb, 0x7410) | 0xff000000u;
}

} // namespace
} // namespace sgx
