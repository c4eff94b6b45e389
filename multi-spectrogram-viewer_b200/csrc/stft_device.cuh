// stft_device.cuh -- device-side pieces shared by the analysis kernels (stft_kernel.cu: the block kernel K1,
// stft_warp2_kernel.cu: the warp-per-frame-pair kernel for n_fft = 2048): packed FP32 pairs, the in-register DFT,
// mbarrier / TMA bulk copy, sample access with reflection, and the tile bookkeeping of a launch.
#pragma once
#include <cstdint>
#include <type_traits>

#include "device_common.cuh"

namespace sgx {
namespace {

__device__ __forceinline__ int padi(int e) { return e + (e >> 3); }

// ---- packed FP32 pairs ---------------------------------------------------------------------------
// sm_100 executes two IEEE-RN FP32 operations per instruction on a 64-bit register pair (SASS FADD2 / FMUL2 /
// FFMA2; a scalar operand is broadcast by the instruction itself, a negated one costs nothing).  A pair holds the
// same quantity of two consecutive FRAMES, so every butterfly, twiddle product, split and mel tap below is one
// instruction for two frames; V frames per thread = V/2 pairs.
typedef float2 pk;
__device__ __forceinline__ pk pk_neg(pk a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ pk pk_add(pk a, pk b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ pk pk_sub(pk a, pk b) { return __fadd2_rn(a, pk_neg(b)); }
__device__ __forceinline__ pk pk_mul(pk a, pk b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ pk pk_fma(pk a, pk b, pk c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ pk pk_muls(pk a, float s) { return __fmul2_rn(a, make_float2(s, s)); }          // a * s
__device__ __forceinline__ pk pk_fmas(pk a, float s, pk c) { return __ffma2_rn(a, make_float2(s, s), c); } // a * s + c

// frame v of a pair array (v is a compile-time constant wherever this is used: the loops are unrolled)
template <int VP> __device__ __forceinline__ void pk_set(pk (&a)[VP], int v, float x) { if (v & 1) a[v >> 1].y = x; else a[v >> 1].x = x; }
template <int VP> __device__ __forceinline__ float pk_get(const pk (&a)[VP], int v) { return (v & 1) ? a[v >> 1].y : a[v >> 1].x; }

template <int VP> __device__ __forceinline__ void ld_vec(const float *p, pk (&v)[VP])
{
    if constexpr (VP == 2) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = make_float2(t.x, t.y); v[1] = make_float2(t.z, t.w);
    } else {
        static_assert(VP == 1, "V is 2 or 4 frames");
        v[0] = *reinterpret_cast<const float2 *>(p);
    }
}
template <int VP> __device__ __forceinline__ void st_vec(float *p, const pk (&v)[VP])
{
    if constexpr (VP == 2) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    } else {
        *reinterpret_cast<float2 *>(p) = v[0];
    }
}

// sqrt.approx.f32: max relative error 2^-23 (PTX ISA) -- one MUFU instead of the IEEE sequence
__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- mbarrier / TMA bulk copy (PTX ISA 8.x, sm_90+; SASS: UBLKCP / SYNCS) ----------------------
__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes,
                                              unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    // try_wait suspends the thread for a hardware-defined time slice per attempt; a copy that never completes
    // (it cannot, unless the descriptor table is corrupt) traps instead of hanging the GPU
    for (unsigned spins = 0;; ++spins) {
        unsigned done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1u << 26)) __trap();
    }
}

// ---- sample access: channel sum + reflect (lib.rs:42, utils.rs:79-85) ---------------------------
struct PcmView {
    const void *pcm; long long n; int ch; int fmt;
    long long origin, avail; // time slices: pcm[0] is global sample `origin`, `avail` samples are present
};
__device__ __forceinline__ float load_sample(const PcmView &pv, long long i)
{
    if (i < 0) i = -i;                         // left reflect, edge sample not repeated
    if (i >= pv.n) i = 2 * (pv.n - 1) - i;     // right reflect
    i = i < 0 ? 0 : (i >= pv.n ? pv.n - 1 : i); // only reachable under zero window weight
    i -= pv.origin;                            // global -> local index of a time slice
    i = i < 0 ? 0 : (i >= pv.avail ? pv.avail - 1 : i);
    float s = 0.0f;
    if (pv.fmt == PCM_F32) {
        const float *p = reinterpret_cast<const float *>(pv.pcm) + i * pv.ch;
        for (int c = 0; c < pv.ch; ++c) s += __ldg(p + c);
    } else {
        const short *p = reinterpret_cast<const short *>(pv.pcm) + i * pv.ch;
        for (int c = 0; c < pv.ch; ++c) s += (float)__ldg(p + c) * (1.0f / 32768.0f); // audio.rs:16-19
    }
    return s;
}

// ---- in-register DFT of R points at re[BASE + i*STRIDE], natural order in and out ----------------
// cos/sin(2 pi k / 32), k < 16
__device__ constexpr float kC32[16] = {
    1.0f, 0.98078528040323044f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
    0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f, -0.19509032201612825f,
    -0.38268343236508977f, -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
    -0.92387953251128674f, -0.98078528040323044f};
__device__ constexpr float kS32[16] = {
    0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f, 0.70710678118654752f,
    0.83146961230254524f, 0.92387953251128674f, 0.98078528040323044f, 1.0f, 0.98078528040323044f,
    0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
    0.38268343236508977f, 0.19509032201612825f};

template <int R, int BASE, int STRIDE, int PTS, int VP>
__device__ __forceinline__ void dft_inplace(pk (&re)[PTS][VP], pk (&im)[PTS][VP])
{
    if constexpr (R == 2) {
#pragma unroll
        for (int v = 0; v < VP; ++v) {
            const pk ar = re[BASE][v], ai = im[BASE][v];
            const pk br = re[BASE + STRIDE][v], bi = im[BASE + STRIDE][v];
            re[BASE][v] = pk_add(ar, br); im[BASE][v] = pk_add(ai, bi);
            re[BASE + STRIDE][v] = pk_sub(ar, br); im[BASE + STRIDE][v] = pk_sub(ai, bi);
        }
    } else if constexpr (R > 2) {
        dft_inplace<R / 2, BASE, 2 * STRIDE, PTS, VP>(re, im);          // even inputs
        dft_inplace<R / 2, BASE + STRIDE, 2 * STRIDE, PTS, VP>(re, im); // odd inputs
        pk tr[R][VP], ti[R][VP];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            const int e = BASE + 2 * k * STRIDE, o = BASE + (2 * k + 1) * STRIDE;
            const int widx = k * (32 / R); // exp(-2 pi i k / R) = kC32[widx] - i kS32[widx]
#pragma unroll
            for (int v = 0; v < VP; ++v) {
                pk pr, pi;
                if (widx == 0) { pr = re[o][v]; pi = im[o][v]; }
                else if (widx == 8) { pr = im[o][v]; pi = pk_neg(re[o][v]); }
                else {
                    const float c = kC32[widx], s = kS32[widx];
                    pr = pk_fmas(im[o][v], s, pk_muls(re[o][v], c));   // re c + im s
                    pi = pk_fmas(re[o][v], -s, pk_muls(im[o][v], c));  // im c - re s
                }
                tr[k][v] = pk_add(re[e][v], pr); ti[k][v] = pk_add(im[e][v], pi);
                tr[k + R / 2][v] = pk_sub(re[e][v], pr); ti[k + R / 2][v] = pk_sub(im[e][v], pi);
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int v = 0; v < VP; ++v) { re[BASE + k * STRIDE][v] = tr[k][v]; im[BASE + k * STRIDE][v] = ti[k][v]; }
    }
}

// Where one CTA tile of a launch lives: its track, its frames and the PCM span they read.
struct TileLoc {
    int trk, t0, nfr, off0, len, len4;
    long long S0, A0;
    bool tma;
    bool raw2; // the tile holds raw interleaved stereo f32 (2 floats per sample), summed when the first pass loads it
    bool i16;  // the tile holds raw mono int16 samples (audio.rs:16-19 scale applied when the first pass loads it)
};
// `lo` is a lower bound of the track index (a CTA visits tiles, hence tracks, in rising order)
__device__ __forceinline__ int find_track(const StftLaunch &L, int tile_id, int lo)
{
    int hi = L.n_tracks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (L.tracks[mid].tile_begin <= tile_id) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// `td` may be the descriptor in global memory or the CTA's shared-memory copy of it
__device__ __forceinline__ void locate_tile(const StftLaunch &L, int F, int tile_id, int trk, const StftTrack *td, TileLoc &o,
                                            bool allow_raw2, bool allow_i16 = false)
{
    o.trk = trk;
    o.t0 = (tile_id - td->tile_begin) * L.frames_per_tile;
    o.nfr = min(L.frames_per_tile, td->n_frames - o.t0);
    const long long origin = td->origin;
    o.S0 = (long long)(td->frame0 + o.t0) * td->hop - td->win / 2 - td->pad_l; // first (global) sample of the tile's first FFT frame
    // 16-byte granules of the bulk copy: 4 f32 samples, 8 int16 samples
    const bool want16 = allow_i16 && td->fmt == PCM_I16 && td->ch == 1;
    const int gran = want16 ? 8 : 4;
    o.off0 = (int)((o.S0 - origin) & (gran - 1));
    o.A0 = o.S0 - o.off0; // global index whose LOCAL position is 16-byte aligned: start of the staged tile
    o.len = o.off0 + (o.nfr - 1) * td->hop + F;
    o.len4 = (o.len + gran - 1) & ~(gran - 1);
    // a tile that lies inside the track (no reflection), 16-byte aligned: one TMA bulk copy -- of the f32 samples (mono),
    // of the raw interleaved f32 pairs (stereo; the channels are summed when the first pass loads them, which needs every
    // frame of the tile to start on an even sample and twice the room), or of the raw int16 samples (mono; converted
    // and scaled when the first pass loads them)
    const bool inside = L.staged && ((reinterpret_cast<uintptr_t>(td->pcm) & 15) == 0) &&
                        o.A0 >= 0 && o.A0 + o.len4 <= td->n && o.A0 - origin >= 0 && o.A0 - origin + o.len4 <= td->avail;
    const bool plain = inside && td->fmt == PCM_F32;
    o.raw2 = allow_raw2 && plain && td->ch == 2 && ((td->hop | o.off0) & 1) == 0 && 2 * o.len4 <= L.tile_floats;
    o.i16 = want16 && inside;
    o.tma = (plain && td->ch == 1) || o.raw2 || o.i16;
}
__device__ __forceinline__ void issue_tile_copy(const StftTrack *td, const TileLoc &o, float *tile, unsigned long long *mbar)
{
    const unsigned bps = o.i16 ? 2u : (o.raw2 ? 8u : 4u); // bytes per sample in the staged tile
    mbar_expect_tx(mbar, (unsigned)o.len4 * bps);
    bulk_copy_g2s(tile, reinterpret_cast<const char *>(td->pcm) + (o.A0 - td->origin) * (long long)bps, (unsigned)o.len4 * bps, mbar);
}

} // namespace
} // namespace sgx
