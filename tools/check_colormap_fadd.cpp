// check_colormap_fadd.cpp -- host check that the conversion-free colour map of csrc/render_device.cuh
// (floor taken from the mantissa of v + 2^23 rounded down -- `fast` below; grey_to_rgba_cached: in addition the FMAs
// straight from `position`) return the bytes of grey_to_rgba_const
// for EVERY float x in [0, 2] (and a few beyond).  Build and run:  g++ -O1 -frounding-math -o /tmp/ccf tools/check_colormap_fadd.cpp && /tmp/ccf
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

struct Seg { float ra, rd, ga, gd, ba, bd; };
static const int stops[10][3] = {{0, 0, 4}, {27, 12, 65}, {74, 12, 107}, {120, 28, 109}, {165, 44, 96}, {207, 68, 70}, {237, 105, 37}, {251, 155, 6}, {247, 209, 61}, {252, 255, 164}};
static Seg seg[9];

static uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static float add_rd(float a, float b)
{
    std::fesetround(FE_DOWNWARD);
    volatile float x = a, y = b;
    volatile float r = x + y;
    std::fesetround(FE_TONEAREST);
    return r;
}
static uint32_t ref(float x)
{
    const float position = 10.0f * x;
    const float fl = std::floor(position);
    int idx = fl > 2e9f ? 2000000000 : (int)fl;
    if (idx > 8) idx = 8;
    const float ratio = position - fl;
    const Seg &s = seg[idx];
    const uint32_t cr = (uint32_t)std::floor(std::fmaf(ratio, s.rd, s.ra));
    const uint32_t cg = (uint32_t)std::floor(std::fmaf(ratio, s.gd, s.ga));
    const uint32_t cb = (uint32_t)std::floor(std::fmaf(ratio, s.bd, s.ba));
    const uint32_t px = cr | (cg << 8) | (cb << 16) | 0xff000000u;
    return fl < 9.0f ? px : 0xffa4fffcu;
}
static uint32_t fast(float x)
{
    const float kMagic = 8388608.0f;
    const float position = 10.0f * std::fmin(x, 1.5f);
    const float tf = add_rd(position, kMagic);
    const uint32_t ti = bits(tf) & 0xfu;
    const float fl = tf - kMagic;
    const int idx = ti < 8 ? (int)ti : 8;
    const float ratio = position - fl;
    const Seg &s = seg[idx];
    const uint32_t cr = bits(add_rd(std::fmaf(ratio, s.rd, s.ra), kMagic));
    const uint32_t cg = bits(add_rd(std::fmaf(ratio, s.gd, s.ga), kMagic));
    const uint32_t cb = bits(add_rd(std::fmaf(ratio, s.bd, s.ba), kMagic));
    const uint32_t px = (cr & 0xff) | ((cg & 0xff) << 8) | ((cb & 0xff) << 16) | 0xff000000u;
    return ti < 9u ? px : 0xffa4fffcu;
}

struct SegP { float rc, rd, gc, gd, bc, bd; };
static SegP segp[10];
static uint32_t pos_form(float x)
{
    const float kMagic = 8388608.0f;
    const float position = 10.0f * std::fmin(x, 1.5f);
    const uint32_t ti = bits(add_rd(position, kMagic)) & 0xfu;
    const int idx = ti < 9 ? (int)ti : 9;
    const SegP &s = segp[idx];
    const uint32_t cr = bits(add_rd(std::fmaf(position, s.rd, s.rc), kMagic));
    const uint32_t cg = bits(add_rd(std::fmaf(position, s.gd, s.gc), kMagic));
    const uint32_t cb = bits(add_rd(std::fmaf(position, s.bd, s.bc), kMagic));
    return (cr & 0xff) | ((cg & 0xff) << 8) | ((cb & 0xff) << 16) | 0xff000000u;
}

int main()
{
    for (int i = 0; i < 10; ++i) {
        const int a = i < 9 ? i : 9, b = i < 9 ? i + 1 : 9;
        const float m = i < 9 ? (float)i : 0.0f;
        segp[i] = SegP{stops[a][0] + 0.5f - m * (float)(stops[b][0] - stops[a][0]), (float)(stops[b][0] - stops[a][0]),
                       stops[a][1] + 0.5f - m * (float)(stops[b][1] - stops[a][1]), (float)(stops[b][1] - stops[a][1]),
                       stops[a][2] + 0.5f - m * (float)(stops[b][2] - stops[a][2]), (float)(stops[b][2] - stops[a][2])};
    }
    for (int i = 0; i < 9; ++i)
        seg[i] = Seg{stops[i][0] + 0.5f, (float)(stops[i + 1][0] - stops[i][0]), stops[i][1] + 0.5f, (float)(stops[i + 1][1] - stops[i][1]),
                     stops[i][2] + 0.5f, (float)(stops[i + 1][2] - stops[i][2])};
    unsigned long long n = 0, bad = 0;
    const uint32_t hi = bits(2.0f);
    for (uint32_t u = 0; u <= hi; ++u) {
        float x; std::memcpy(&x, &u, 4);
        ++n;
        const uint32_t want = ref(x);
        if (want != fast(x) || want != pos_form(x)) { if (bad < 5) std::printf("x=%.9g ref=%08x fast=%08x pos=%08x\n", x, want, fast(x), pos_form(x)); ++bad; }
    }
    const float extra[] = {2.5f, 10.0f, 1e10f, 3.4028235e38f};
    for (float x : extra) { ++n; if (ref(x) != fast(x) || ref(x) != pos_form(x)) ++bad; }
    std::printf("%llu values, %llu mismatches\n", n, bad);
    return bad != 0;
}
