// stft_kernel.cu -- K1, the fused analysis kernel (sm_100a).
//
// One launch turns PCM of a batch of tracks into their dB spectrograms:
//   channel sum (lib.rs:42) -> reflect-padded framing (lib.rs:412-433, utils.rs:79-85) -> window,
//   centred zero-pad (lib.rs:377-384) -> real FFT as an F/2-point complex FFT + split
//   (realfft.rs:105-159) -> |X| (lib.rs:124) -> banded mel projection (lib.rs:131) -> dB
//   (decibel.rs:33-88) -> per-track max/min (lib.rs:197-200).
// Linear magnitudes and complex spectra never touch HBM in the dB modes.
//
// Layout of the work
//   * a CTA owns a tile of consecutive frames of one track and stages the PCM samples those
//     frames cover in shared memory ONCE: interior, mono, f32 tiles by a TMA bulk copy
//     (cp.async.bulk + mbarrier), edge / stereo / int16 tiles by a reflecting gather;
//   * a group of h/PTS threads transforms V frames at a time: every thread keeps PTS complex
//     points of each of the V frames in registers, so index math and twiddles are shared by V
//     frames and every shared-memory access is a V*4-byte vector;
//   * Stockham autosort passes of radix PTS (then one pass of the remaining radix) exchange
//     through a padded shared buffer (pad(e) = e + e/8 keeps the strided writes of the first pass
//     conflict free for 128-bit accesses);
//   * split / magnitude / mel / dB run on the same registers and buffer; dB rows go to HBM with
//     coalesced stores, per-thread extrema are reduced to two atomics per CTA.
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "device_common.cuh"
#include "kernels.h"

namespace sgx {

namespace {

__device__ __forceinline__ int padi(int e) { return e + (e >> 3); }

template <int V> __device__ __forceinline__ void ld_vec(const float *p, float (&v)[V])
{
    if constexpr (V == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (V == 2) {
        const float2 t = *reinterpret_cast<const float2 *>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = *p;
    }
}
template <int V> __device__ __forceinline__ void st_vec(float *p, const float (&v)[V])
{
    if constexpr (V == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (V == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    } else {
        *p = v[0];
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

template <int R, int BASE, int STRIDE, int PTS, int V>
__device__ __forceinline__ void dft_inplace(float (&re)[PTS][V], float (&im)[PTS][V])
{
    if constexpr (R == 2) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float ar = re[BASE][v], ai = im[BASE][v];
            const float br = re[BASE + STRIDE][v], bi = im[BASE + STRIDE][v];
            re[BASE][v] = ar + br; im[BASE][v] = ai + bi;
            re[BASE + STRIDE][v] = ar - br; im[BASE + STRIDE][v] = ai - bi;
        }
    } else if constexpr (R > 2) {
        dft_inplace<R / 2, BASE, 2 * STRIDE, PTS, V>(re, im);          // even inputs
        dft_inplace<R / 2, BASE + STRIDE, 2 * STRIDE, PTS, V>(re, im); // odd inputs
        float tr[R][V], ti[R][V];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            const int e = BASE + 2 * k * STRIDE, o = BASE + (2 * k + 1) * STRIDE;
            const int widx = k * (32 / R); // exp(-2 pi i k / R) = kC32[widx] - i kS32[widx]
#pragma unroll
            for (int v = 0; v < V; ++v) {
                float pr, pi;
                if (widx == 0) { pr = re[o][v]; pi = im[o][v]; }
                else if (widx == 8) { pr = im[o][v]; pi = -re[o][v]; }
                else {
                    const float c = kC32[widx], s = kS32[widx];
                    pr = re[o][v] * c + im[o][v] * s;
                    pi = im[o][v] * c - re[o][v] * s;
                }
                tr[k][v] = re[e][v] + pr; ti[k][v] = im[e][v] + pi;
                tr[k + R / 2][v] = re[e][v] - pr; ti[k + R / 2][v] = im[e][v] - pi;
            }
        }
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int v = 0; v < V; ++v) { re[BASE + k * STRIDE][v] = tr[k][v]; im[BASE + k * STRIDE][v] = ti[k][v]; }
    }
}

// ---- one Stockham pass of radix R with NS = product of the previous radices ----------------------
// Thread `gt` of the group owns butterflies j = gt + q*NT, q < PTS/R.  Inputs of butterfly j are
// elements j + r*H/R; outputs go to (j - k)*R + k + r*NS with k = j mod NS.  NS == 1 (first pass):
// the caller has already put the inputs into the registers and no twiddle is needed.
// Barrier over one thread group (the groups of a CTA transform different frames and never share data).
template <int G, int NT> __device__ __forceinline__ void group_sync(int grp)
{
    if constexpr (G == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(NT) : "memory");
}

template <int H, int PTS, int V, int G, int R, int NS>
__device__ __forceinline__ void fft_pass(float (&re)[PTS][V], float (&im)[PTS][V], float *sre,
                                         float *sim, int gt, int grp, const float2 *__restrict__ tw)
{
    constexpr int NT = H / PTS, NB = PTS / R;
    if constexpr (NS > 1) {
        // twiddle bases first: the table loads fly while the shared loads and the barrier complete
        float2 w1[NB];
#pragma unroll
        for (int q = 0; q < NB; ++q) w1[q] = __ldg(tw + ((gt + q * NT) & (NS - 1)) * (H / (NS * R)));
        // element gt + q*NT + r*H/R: the offsets are multiples of 8, so pad() is linear in them
        static_assert(NT % 8 == 0 && (H / R) % 8 == 0, "padded addressing assumes multiples of 8");
        const int sbase = padi(gt) * V;
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int src = sbase + ((q * NT + r * (H / R)) / 8 * 9) * V;
                ld_vec<V>(sre + src, re[q * R + r]);
                ld_vec<V>(sim + src, im[q * R + r]);
            }
        group_sync<G, NT>(grp); // every thread has its inputs: the buffer may be overwritten
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            // w[r] = exp(-2 pi i r k / (NS R)): one table load, the powers by complex products
            // (at most 3 products deep: error a few ulp, far inside the 1e-4 magnitude tolerance)
            float2 w[R];
            w[1] = w1[q];
            if constexpr (R >= 4) {
                w[2] = make_float2(w[1].x * w[1].x - w[1].y * w[1].y, 2.0f * w[1].x * w[1].y);
                w[3] = make_float2(w[2].x * w[1].x - w[2].y * w[1].y, w[2].x * w[1].y + w[2].y * w[1].x);
            }
            if constexpr (R >= 8) {
                w[4] = make_float2(w[2].x * w[2].x - w[2].y * w[2].y, 2.0f * w[2].x * w[2].y);
                w[5] = make_float2(w[4].x * w[1].x - w[4].y * w[1].y, w[4].x * w[1].y + w[4].y * w[1].x);
                w[6] = make_float2(w[3].x * w[3].x - w[3].y * w[3].y, 2.0f * w[3].x * w[3].y);
                w[7] = make_float2(w[4].x * w[3].x - w[4].y * w[3].y, w[4].x * w[3].y + w[4].y * w[3].x);
            }
            if constexpr (R >= 16) {
                auto cm = [](float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); };
                w[8] = make_float2(w[4].x * w[4].x - w[4].y * w[4].y, 2.0f * w[4].x * w[4].y);
                w[9] = cm(w[8], w[1]); w[10] = cm(w[8], w[2]); w[11] = cm(w[8], w[3]); w[12] = cm(w[8], w[4]);
                w[13] = cm(w[8], w[5]); w[14] = cm(w[8], w[6]); w[15] = cm(w[8], w[7]);
            }
            static_assert(R <= 16, "twiddle powers are written out for radix <= 16");
#pragma unroll
            for (int r = 1; r < R; ++r) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float xr = re[q * R + r][v], xi = im[q * R + r][v];
                    re[q * R + r][v] = xr * w[r].x - xi * w[r].y;
                    im[q * R + r][v] = xr * w[r].y + xi * w[r].x;
                }
            }
        }
    }
    if constexpr (NB == 1) dft_inplace<R, 0, 1, PTS, V>(re, im);
    else if constexpr (NB == 2) { dft_inplace<R, 0, 1, PTS, V>(re, im); dft_inplace<R, R, 1, PTS, V>(re, im); }
    else if constexpr (NB == 4) {
        dft_inplace<R, 0, 1, PTS, V>(re, im); dft_inplace<R, R, 1, PTS, V>(re, im);
        dft_inplace<R, 2 * R, 1, PTS, V>(re, im); dft_inplace<R, 3 * R, 1, PTS, V>(re, im);
    } else {
        static_assert(NB == 8, "unsupported butterflies per thread");
        dft_inplace<R, 0, 1, PTS, V>(re, im); dft_inplace<R, R, 1, PTS, V>(re, im);
        dft_inplace<R, 2 * R, 1, PTS, V>(re, im); dft_inplace<R, 3 * R, 1, PTS, V>(re, im);
        dft_inplace<R, 4 * R, 1, PTS, V>(re, im); dft_inplace<R, 5 * R, 1, PTS, V>(re, im);
        dft_inplace<R, 6 * R, 1, PTS, V>(re, im); dft_inplace<R, 7 * R, 1, PTS, V>(re, im);
    }
    if constexpr (NS == 1 && R == 8) {
        // d = 8 j, element 8 j + r -> padded 9 j + r
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int dst = (9 * gt + 9 * q * NT + r) * V;
                st_vec<V>(sre + dst, re[q * R + r]);
                st_vec<V>(sim + dst, im[q * R + r]);
            }
    } else if constexpr (NS % 8 == 0) {
        // k = j mod NS is the same for every q when NS divides NT; when NS > NT, j < NS and d = j.
        // All per-(q, r) displacements are multiples of 8: one pad() per pass.
        constexpr bool kSmall = NS <= NT;
        const int k = kSmall ? (gt & (NS - 1)) : gt;
        const int dbase = padi(kSmall ? (gt - k) * R + k : gt) * V;
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int dst = dbase + (((kSmall ? q * NT * R : q * NT) + r * NS) / 8 * 9) * V;
                st_vec<V>(sre + dst, re[q * R + r]);
                st_vec<V>(sim + dst, im[q * R + r]);
            }
    } else {
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int j = gt + q * NT;
            const int k = j & (NS - 1);
            const int d = (j - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int dst = padi(d + r * NS) * V;
                st_vec<V>(sre + dst, re[q * R + r]);
                st_vec<V>(sim + dst, im[q * R + r]);
            }
        }
    }
    group_sync<G, NT>(grp);
}

// Runs the passes whose input stride product NS is below NS_END (H: all passes).
template <int H, int PTS, int V, int G, int NS, int NS_END>
__device__ __forceinline__ void run_passes(float (&re)[PTS][V], float (&im)[PTS][V], float *sre,
                                           float *sim, int gt, int grp, const float2 *__restrict__ tw)
{
    if constexpr (NS < NS_END) {
        constexpr int R = (H / NS >= PTS) ? PTS : (H / NS);
        fft_pass<H, PTS, V, G, R, NS>(re, im, sre, sim, gt, grp, tw);
        run_passes<H, PTS, V, G, NS * R, NS_END>(re, im, sre, sim, gt, grp, tw);
    }
}

// radix of the last pass of the schedule above
__host__ __device__ constexpr int last_radix(int h, int pts)
{
    int n = h;
    while (n >= pts) n /= pts;
    return n > 1 ? n : pts;
}

// Where one CTA tile of a launch lives: its track, its frames and the PCM span they read.
struct TileLoc {
    int trk, t0, nfr, off0, len, len4;
    long long S0, A0;
    bool tma;
    bool raw2; // the tile holds raw interleaved stereo f32 (2 floats per sample), summed when the first pass loads it
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
                                            bool allow_raw2)
{
    o.trk = trk;
    o.t0 = (tile_id - td->tile_begin) * L.frames_per_tile;
    o.nfr = min(L.frames_per_tile, td->n_frames - o.t0);
    const long long origin = td->origin;
    o.S0 = (long long)(td->frame0 + o.t0) * td->hop - td->win / 2 - td->pad_l; // first (global) sample of the tile's first FFT frame
    o.off0 = (int)((o.S0 - origin) & 3);
    o.A0 = o.S0 - o.off0; // global index whose LOCAL position is 16-byte aligned: start of the staged tile
    o.len = o.off0 + (o.nfr - 1) * td->hop + F;
    o.len4 = (o.len + 3) & ~3;
    // a tile that lies inside the track (no reflection), f32, 16-byte aligned: one TMA bulk copy -- of the samples
    // (mono) or of the raw interleaved pairs (stereo; the channels are summed when the first pass loads them, which
    // needs every frame of the tile to start on an even sample and twice the room)
    const bool plain = L.staged && td->fmt == PCM_F32 && ((reinterpret_cast<uintptr_t>(td->pcm) & 15) == 0) &&
                       o.A0 >= 0 && o.A0 + o.len4 <= td->n && o.A0 - origin >= 0 && o.A0 - origin + o.len4 <= td->avail;
    o.raw2 = allow_raw2 && plain && td->ch == 2 && ((td->hop | o.off0) & 1) == 0 && 2 * o.len4 <= L.tile_floats;
    o.tma = (plain && td->ch == 1) || o.raw2;
}
__device__ __forceinline__ void issue_tile_copy(const StftTrack *td, const TileLoc &o, float *tile, unsigned long long *mbar)
{
    const unsigned spf = o.raw2 ? 2u : 1u; // floats per sample in the staged tile
    mbar_expect_tx(mbar, (unsigned)o.len4 * 4u * spf);
    bulk_copy_g2s(tile, reinterpret_cast<const float *>(td->pcm) + (o.A0 - td->origin) * spf, (unsigned)o.len4 * 4u * spf, mbar);
}

template <int LOG2H, int PTS, int V, int G, int MC> struct K1Traits {
    static constexpr int H = 1 << LOG2H;
    static constexpr int NT = H / PTS;
    static constexpr int THREADS = G * NT;
    static constexpr int PADH = ((H + (H >> 3) + 1) + 3) & ~3; // elements of V floats
    static constexpr size_t FFT_SMEM = (size_t)G * 2 * PADH * V * sizeof(float);
    static constexpr int MIN_CTAS = MC; // resident CTAs per SM the register budget is shaped for
};

// MEL: the launch projects onto a mel filterbank (MODE_MEL_DB).  A compile-time flag rather than a test of
// L.mode: the code of the other output modes (and their branch targets) is then absent from the mel kernel's
// instruction stream, which is long enough for instruction fetch to show up in the stall profile.
// RAW2: f32 stereo tiles are staged as raw interleaved pairs by TMA and summed by the first pass (its own
// instantiation: the extra first-pass variant costs the mono kernels 2 % when it merely sits in their code).
template <int LOG2H, int PTS, int V, int G, int MC, bool MEL, bool RAW2>
__global__ void __launch_bounds__(K1Traits<LOG2H, PTS, V, G, MC>::THREADS, MC)
stft_db_kernel(const StftLaunch L)
{
    using TR = K1Traits<LOG2H, PTS, V, G, MC>;
    constexpr int H = TR::H, NT = TR::NT, THREADS = TR::THREADS, PADH = TR::PADH, F = 2 * H;
    static_assert(NT >= 32 && (NT % 32) == 0, "a group must be whole warps");
    static_assert(G <= 15, "one named barrier (1..15) per group");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    float *tile = reinterpret_cast<float *>(smem_raw + 16);
    float *fftbuf = tile + L.tile_floats;

    const int tid = threadIdx.x;
    const int grp = tid / NT, gt = tid % NT;
    float *sre = fftbuf + (size_t)grp * 2 * PADH * V;
    float *sim = sre + PADH * V;

    // ---- persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ... ------------------------------------
    // The PCM tile is only read by the first FFT pass (into registers); as soon as every warp has done
    // that, the last one to check in starts the bulk copy of the CTA's NEXT tile into the same buffer, so
    // the copy runs under the remaining passes, the split and the mel projection of the current one.
    float *bank = fftbuf + (size_t)G * 2 * PADH * V; // dedicated filterbank region (L.bank_floats floats), if any
    const float *bank_src = nullptr;                  // whose taps it currently holds
    const int mode = MEL ? (int)MODE_MEL_DB : L.mode;
    unsigned *done_cnt = reinterpret_cast<unsigned *>(smem_raw + 8); // warps that have consumed the current tile
    // the exchange planes start out finite: the block-padded mel path multiplies whatever lies just beyond a
    // filter's last bin (padding holes, the slack above bin H) by zero weights
    for (int i = gt; i < 2 * PADH * V; i += NT) sre[i] = 0.0f;
    if (L.staged && tid == 0) { mbar_init(mbar, 1); *done_cnt = 0u; }
    __syncthreads();
    unsigned phase = 0;
    // The descriptor of the track the CTA is working on is kept in shared memory: finding a tile's place
    // then costs a few shared loads instead of a chain of dependent global ones at every tile.
    __shared__ StftTrack s_td;
    __shared__ int s_trk, s_trk_end;
    auto enter_track = [&](int tile_id, int lo) { // all threads
        __syncthreads();
        if (tid == 0) {
            const int t = find_track(L, tile_id, lo);
            s_trk = t;
            s_trk_end = t + 1 < L.n_tracks ? L.tracks[t + 1].tile_begin : L.n_tiles;
        }
        __syncthreads();
        const int *src = reinterpret_cast<const int *>(L.tracks + s_trk);
        int *dst = reinterpret_cast<int *>(&s_td);
        for (int i = tid; i < (int)(sizeof(StftTrack) / sizeof(int)); i += THREADS) dst[i] = src[i];
        __syncthreads();
    };
    enter_track(blockIdx.x, 0);
    int trk_end = s_trk_end;
    const StftTrack *td = &s_td;
    TileLoc cur;
    locate_tile(L, F, blockIdx.x, s_trk, td, cur, RAW2);
    if (cur.tma && tid == 0) issue_tile_copy(td, cur, tile, mbar);

    float vmax = -INFINITY, vmin = INFINITY;
    int range_trk = -1;
    // per-track extrema (lib.rs:197-200): flushed when the CTA moves on to another track
    auto flush_range = [&]() {
        if (range_trk < 0 || !(mode == MODE_LIN_DB || mode == MODE_MEL_DB)) return;
        unsigned *slot = L.tracks[range_trk].range_slot;
        if (slot == nullptr) return;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, s));
        }
        if ((tid & 31) == 0 && vmax >= vmin) { // at least one value was produced
            atomicMax(slot, enc_ordered(vmax));
            atomicMin(slot + 1, enc_ordered(vmin));
        }
    };

    for (int tile_id = blockIdx.x; tile_id < L.n_tiles; tile_id += gridDim.x) {
    if (tile_id != (int)blockIdx.x) {
        if (tile_id >= trk_end) { // uniform: the CTA moves on to another track
            flush_range();
            enter_track(tile_id, cur.trk);
            trk_end = s_trk_end;
        }
        locate_tile(L, F, tile_id, s_trk, td, cur, RAW2);
    }
    if (cur.trk != range_trk) { range_trk = cur.trk; vmax = -INFINITY; vmin = INFINITY; }
    const PcmView pv{td->pcm, td->n, td->ch, td->fmt, td->origin, td->avail};
    const int hop = td->hop, T = td->n_frames;
    const float *__restrict__ win_f = td->win_f;
    float *__restrict__ out = td->out;
    const int n_out = td->n_out;
    const int t0 = cur.t0, nfr = cur.nfr, off0 = cur.off0;
    const long long S0 = cur.S0;
    const bool vec_ok = ((hop | off0) & 1) == 0; // all frames of the tile start on an even float
    (void)T;

    // ---- filterbank of this track into its dedicated region (once per CTA and track) ----------------
    // melp: the block-padded copy is used (fused kernels, magnitudes unpadded); else the banded copy, in the
    // region when it fits, staged per round in the imaginary plane or read through L1 otherwise
    constexpr int RL_K = last_radix(H, PTS);
    constexpr bool FUSED_K = (PTS / RL_K) >= 2;
    const bool melp = MEL && FUSED_K && td->melp != nullptr && td->melp_nwb + 34 * td->melp_nblk <= L.bank_floats;
    const bool bank_fits = MEL && (melp || ((__ldg(td->mel_cnt + 1) + 3) & ~3) + 4 * n_out <= L.bank_floats);
    if (MEL && bank_fits && td->mel_w != bank_src) {
        __syncthreads(); // other groups may still be projecting frames of the previous track
        if (melp) {
            const int words = td->melp_nwb + 34 * td->melp_nblk;
            const int *__restrict__ srcw = td->melp;
            int *dstw = reinterpret_cast<int *>(bank);
            for (int i = tid; i < words; i += THREADS) dstw[i] = __ldg(srcw + i);
        } else {
            const int nnz = __ldg(td->mel_cnt + 1);
            const int4 *__restrict__ meta = reinterpret_cast<const int4 *>(td->mel_lo);
            int4 *msm = reinterpret_cast<int4 *>(bank + ((nnz + 3) & ~3));
            for (int i = tid; i < nnz; i += THREADS) bank[i] = __ldg(td->mel_w + i);
            for (int i = tid; i < n_out; i += THREADS) msm[i] = __ldg(meta + i);
        }
        __syncthreads();
        bank_src = td->mel_w;
    }

    // ---- the PCM tile: landed by TMA (issued one tile ago), or gathered here (edges, int16, stereo outside RAW2) -----
    if (L.staged) {
        if (cur.tma) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
        } else {
            __syncthreads(); // every group is past its first pass of the previous tile: the buffer is free
            const long long b0 = cur.A0 - pv.origin; // local index of the tile's first sample
            const bool inside = cur.A0 >= 0 && cur.A0 + cur.len <= pv.n && b0 >= 0 && b0 + cur.len <= pv.avail;
            if (inside && pv.ch <= 2) {
                // no reflection inside this tile: plain strided copies, eight loads in flight per thread
                constexpr int NB = 8;
                for (int s0 = tid; s0 < cur.len; s0 += THREADS * NB) {
                    float v[NB];
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const int s = s0 + j * THREADS;
                        v[j] = 0.0f;
                        if (s < cur.len) {
                            if (pv.fmt == PCM_F32) {
                                const float *p = reinterpret_cast<const float *>(pv.pcm) + (b0 + s) * pv.ch;
                                v[j] = pv.ch == 2 ? __ldg(p) + __ldg(p + 1) : __ldg(p);
                            } else { // audio.rs:16-19
                                const short *p = reinterpret_cast<const short *>(pv.pcm) + (b0 + s) * pv.ch;
                                v[j] = (float)__ldg(p) * (1.0f / 32768.0f);
                                if (pv.ch == 2) v[j] += (float)__ldg(p + 1) * (1.0f / 32768.0f);
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const int s = s0 + j * THREADS;
                        if (s < cur.len) tile[s] = v[j];
                    }
                }
            } else {
                for (int s = tid; s < cur.len; s += THREADS) tile[s] = load_sample(pv, cur.A0 + s);
            }
            __syncthreads();
        }
    }

    const int iters = L.frames_per_tile / (G * V);
    for (int it = 0; it < iters; ++it) {
        if (it * G * V >= nfr) break; // uniform
        const int fl0 = (it * G + grp) * V; // first local frame of this group
        float re[PTS][V], im[PTS][V];

        // ---- first-pass inputs: z[m] = g[2m] + i g[2m+1], g = sample * window ---------------------
        if (L.staged && vec_ok && !(RAW2 && cur.raw2)) {
            // every frame of the tile starts on an even float: one 64-bit shared load per point
            const float *fb[V];
#pragma unroll
            for (int v = 0; v < V; ++v) fb[v] = tile + off0 + min(fl0 + v, nfr - 1) * hop + 2 * gt;
#pragma unroll
            for (int p = 0; p < PTS; ++p) {
                const int n = 2 * (gt + p * NT);
                const float2 w = __ldg(reinterpret_cast<const float2 *>(win_f + n));
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float2 x = *reinterpret_cast<const float2 *>(fb[v] + 2 * p * NT);
                    re[p][v] = x.x * w.x; im[p][v] = x.y * w.y;
                }
            }
        } else if (RAW2 && L.staged && vec_ok) {
            // raw stereo tile: one 128-bit shared load brings (L, R) of two consecutive samples; lib.rs:42 sums them
            const float *fb[V];
#pragma unroll
            for (int v = 0; v < V; ++v) fb[v] = tile + 2 * (off0 + min(fl0 + v, nfr - 1) * hop + 2 * gt);
#pragma unroll
            for (int p = 0; p < PTS; ++p) {
                const int n = 2 * (gt + p * NT);
                const float2 w = __ldg(reinterpret_cast<const float2 *>(win_f + n));
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float4 x = *reinterpret_cast<const float4 *>(fb[v] + 4 * p * NT);
                    re[p][v] = (x.x + x.y) * w.x; im[p][v] = (x.z + x.w) * w.y;
                }
            }
        } else {
#pragma unroll
            for (int p = 0; p < PTS; ++p) {
                const int n = 2 * (gt + p * NT); // NB == 1 in the first pass: point p is element gt + p*H/R0
                const float2 w = __ldg(reinterpret_cast<const float2 *>(win_f + n));
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int fl = min(fl0 + v, nfr - 1);
                    float x0, x1;
                    if (L.staged) {
                        const int b = off0 + fl * hop + n;
                        x0 = tile[b]; x1 = tile[b + 1];
                    } else {
                        const long long i = S0 + (long long)fl * hop + n;
                        x0 = load_sample(pv, i); x1 = load_sample(pv, i + 1);
                    }
                    re[p][v] = x0 * w.x; im[p][v] = x1 * w.y;
                }
            }
        }
        // ---- the tile is in registers: let the next one stream in ---------------------------------------
        // No CTA-wide barrier: every warp checks in on a shared counter once its loads of the tile have
        // been performed, and the warp that checks in last starts the copy.  The groups keep drifting
        // against each other, which is what hides their barrier and shared-memory latencies.
        if (L.staged && (it == iters - 1 || (it + 1) * G * V >= nfr)) {
            __syncwarp();
            if ((tid & 31) == 0) {
                __threadfence_block();
                const unsigned seen = atomicAdd(done_cnt, 1u);
                if (seen % (THREADS / 32) == THREADS / 32 - 1) {
                    const int nt = tile_id + (int)gridDim.x;
                    if (nt < L.n_tiles) {
                        TileLoc nx;
                        const StftTrack *ntd = td;
                        int ntrk = cur.trk;
                        if (nt >= trk_end) { ntrk = find_track(L, nt, cur.trk); ntd = L.tracks + ntrk; }
                        locate_tile(L, F, nt, ntrk, ntd, nx, RAW2);
                        if (nx.tma) issue_tile_copy(ntd, nx, tile, mbar);
                    }
                }
            }
        }
        // ---- real-FFT split (realfft.rs:140-157) -------------------------------------------------------
        // emit(): what becomes of one output bin -- complex / magnitude / dB to HBM, or (mel) its
        // magnitude into the shared buffer at the bin's own (padded) position.
        auto emit = [&](int idx, int spos, const float (&xr)[V], const float (&xi)[V]) {
            if (mode == MODE_COMPLEX) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (fl0 + v < nfr)
                        reinterpret_cast<float2 *>(out)[(size_t)(t0 + fl0 + v) * (H + 1) + idx] =
                            make_float2(xr[v], xi[v]);
                return;
            }
            float mg[V];
#pragma unroll
            for (int v = 0; v < V; ++v) mg[v] = sqrt_approx(fmaf(xr[v], xr[v], xi[v] * xi[v])); // lib.rs:124
            if (MEL) {
                st_vec<V>(sre + (melp ? idx * V : spos), mg); // block-padded bank: magnitudes at their bin index
            } else {
                // frames beyond the tile's last one are copies of it (first-pass clamp): harmless in the extrema
                float *op = out + (size_t)(t0 + fl0) * (H + 1) + idx;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    float y = mg[v];
                    if (mode == MODE_LIN_DB) {
                        y = amp_to_db_dev(y);
                        vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
                    }
                    if (fl0 + v < nfr) op[(size_t)v * (H + 1)] = y;
                }
            }
        };

        // one conjugate-symmetric pair: a = Z[k], b = Z[h-k], k <= h/2 -> X[k] and (optionally) X[h-k]
        auto split_pair = [&](int k, float2 cs, const float (&ar)[V], const float (&ai)[V], const float (&br)[V],
                              const float (&bi)[V], bool emit_partner) {
            float xr[V], xi[V], yr[V], yi[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float sumr = ar[v] + br[v], difr = ar[v] - br[v];
                const float sumi = ai[v] + bi[v], difi = ai[v] - bi[v];
                const float p1 = fmaf(cs.x, sumi, -cs.y * difr);  // c*sumi - s*difr
                const float p2 = fmaf(cs.y, sumi, cs.x * difr);   // s*sumi + c*difr
                xr[v] = 0.5f * (sumr + p1); xi[v] = 0.5f * (difi - p2);
                yr[v] = 0.5f * (sumr - p1); yi[v] = -0.5f * (difi + p2);
            }
            emit(k, padi(k) * V, xr, xi);
            if (emit_partner) { const int kp = k == 0 ? H : H - k; emit(kp, padi(kp) * V, yr, yi); } // k == 0: Nyquist bin
        };

        constexpr int RL = last_radix(H, PTS);
        constexpr bool FUSED = (PTS / RL) >= 2;
        if constexpr (FUSED) {
            // ---- last pass fused with the split ---------------------------------------------------------
            // Butterfly j of the last pass (radix RL, NSL = H/RL) produces Z[j + r NSL]; the conjugate
            // partner of that bin comes out of butterfly NSL - j.  Each thread therefore runs butterfly
            // PAIRS (j, NSL - j): the spectrum never returns to shared memory and the split costs no loads.
            run_passes<H, PTS, V, G, 1, H / RL>(re, im, sre, sim, gt, grp, L.tw);
            constexpr int NSL = H / RL, NPR = (PTS / RL) / 2;
            static_assert(NSL % 8 == 0, "padded addressing assumes multiples of 8");
            int bA[NPR], bB[NPR];
            float2 wA[NPR], wB[NPR];
#pragma unroll
            for (int p = 0; p < NPR; ++p) {
                bA[p] = gt + p * NT;                                          // in [0, NSL/2)
                bB[p] = (p == 0 && gt == 0) ? NSL / 2 : NSL - bA[p];          // thread 0 also owns NSL/2
                wA[p] = __ldg(L.tw + bA[p]); wB[p] = __ldg(L.tw + bB[p]);     // exp(-2 pi i j / H)
            }
            // split twiddles (cos, sin)(k pi / h) of the bins this thread will finish, fetched up front
            float2 csA[NPR][RL / 2], csB[NPR][RL / 2];
#pragma unroll
            for (int p = 0; p < NPR; ++p)
#pragma unroll
                for (int r = 0; r < RL / 2; ++r) {
                    csA[p][r] = __ldg(L.split + bA[p] + r * NSL);
                    csB[p][r] = __ldg(L.split + bB[p] + r * NSL);
                }
#pragma unroll
            for (int p = 0; p < NPR; ++p) {
                const int sa = padi(bA[p]) * V, sb = padi(bB[p]) * V;
#pragma unroll
                for (int r = 0; r < RL; ++r) {
                    ld_vec<V>(sre + sa + (r * NSL / 8 * 9) * V, re[(2 * p) * RL + r]);
                    ld_vec<V>(sim + sa + (r * NSL / 8 * 9) * V, im[(2 * p) * RL + r]);
                    ld_vec<V>(sre + sb + (r * NSL / 8 * 9) * V, re[(2 * p + 1) * RL + r]);
                    ld_vec<V>(sim + sb + (r * NSL / 8 * 9) * V, im[(2 * p + 1) * RL + r]);
                }
            }
            if (MEL) group_sync<G, NT>(grp); // the buffer now becomes the magnitude array
#pragma unroll
            for (int b = 0; b < 2 * NPR; ++b) {
                float2 w[RL];
                w[1] = (b & 1) ? wB[b >> 1] : wA[b >> 1];
                if constexpr (RL >= 4) {
                    w[2] = make_float2(w[1].x * w[1].x - w[1].y * w[1].y, 2.0f * w[1].x * w[1].y);
                    w[3] = make_float2(w[2].x * w[1].x - w[2].y * w[1].y, w[2].x * w[1].y + w[2].y * w[1].x);
                }
                static_assert(RL <= 4, "fused last pass is written for radix 2 and 4");
#pragma unroll
                for (int r = 1; r < RL; ++r)
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const float xr = re[b * RL + r][v], xi = im[b * RL + r][v];
                        re[b * RL + r][v] = xr * w[r].x - xi * w[r].y;
                        im[b * RL + r][v] = xr * w[r].y + xi * w[r].x;
                    }
            }
            if constexpr (NPR >= 1) { dft_inplace<RL, 0, 1, PTS, V>(re, im); dft_inplace<RL, RL, 1, PTS, V>(re, im); }
            if constexpr (NPR >= 2) { dft_inplace<RL, 2 * RL, 1, PTS, V>(re, im); dft_inplace<RL, 3 * RL, 1, PTS, V>(re, im); }
            if constexpr (NPR >= 3) { dft_inplace<RL, 4 * RL, 1, PTS, V>(re, im); dft_inplace<RL, 5 * RL, 1, PTS, V>(re, im); }
            if constexpr (NPR >= 4) { dft_inplace<RL, 6 * RL, 1, PTS, V>(re, im); dft_inplace<RL, 7 * RL, 1, PTS, V>(re, im); }
            static_assert(NPR <= 4, "unsupported butterfly pairs per thread");
#pragma unroll
            for (int p = 0; p < NPR; ++p) {
                const int ba = (2 * p) * RL, bb = (2 * p + 1) * RL; // register blocks of butterflies bA, bB
                // Thread 0 holds the two self-paired butterflies: 0 (Z[r NSL] <-> Z[(RL - r) NSL], r = 0 and RL/2
                // self-conjugate) and NSL/2 (Z[NSL/2 + r NSL] <-> Z[NSL/2 + (RL-1-r) NSL]).  It runs the same
                // calls as everybody else with its partner operands picked from its own registers, so that its
                // warp does not execute a second copy of the split (the other warps of the group would wait
                // for it at the next barrier); only the lone bin H/2 is extra.
                const bool self = p == 0 && gt == 0;
#pragma unroll
                for (int r = 0; r < RL / 2; ++r) {
                    float pr[V], pi[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        pr[v] = re[bb + RL - 1 - r][v]; pi[v] = im[bb + RL - 1 - r][v];
                        if (p == 0 && self) { pr[v] = re[ba + (r == 0 ? 0 : RL - r)][v]; pi[v] = im[ba + (r == 0 ? 0 : RL - r)][v]; }
                    }
                    split_pair(bA[p] + r * NSL, csA[p][r], re[ba + r], im[ba + r], pr, pi, true);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        pr[v] = re[ba + RL - 1 - r][v]; pi[v] = im[ba + RL - 1 - r][v];
                        if (p == 0 && self) { pr[v] = re[bb + RL - 1 - r][v]; pi[v] = im[bb + RL - 1 - r][v]; }
                    }
                    split_pair(bB[p] + r * NSL, csB[p][r], re[bb + r], im[bb + r], pr, pi, true);
                }
                if (p == 0 && self) { // X[H/2] = conj(Z[H/2])
                    float ni[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) ni[v] = -im[ba + RL / 2][v];
                    emit(H / 2, padi(H / 2) * V, re[ba + RL / 2], ni);
                }
            }
        } else {
        run_passes<H, PTS, V, G, 1, H>(re, im, sre, sim, gt, grp, L.tw);
        constexpr int NP = PTS / 2;
        const int pa0 = padi(gt) * V;      // element k = gt + q NT      -> pa0 + 9 q NT / 8
        const int pb0 = padi(H - gt) * V;  // element H - k (k > 0)      -> pb0 - 9 q NT / 8
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int k = gt + q * NT;
            const int pa = pa0 + (q * NT / 8 * 9) * V;
            const int pbm = pb0 - (q * NT / 8 * 9) * V;     // where the partner's magnitude goes (index H when k == 0)
            const int pb = (q == 0 && gt == 0) ? 0 : pbm;   // where the partner's spectrum is read (index 0 when k == 0)
            float ar[V], ai[V], br[V], bi[V];
            ld_vec<V>(sre + pa, ar); ld_vec<V>(sim + pa, ai);
            ld_vec<V>(sre + pb, br); ld_vec<V>(sim + pb, bi);
            const float2 cs = __ldg(L.split + k); // (cos, sin)(k pi / h)
            float xr[V], xi[V], yr[V], yi[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const float sumr = ar[v] + br[v], difr = ar[v] - br[v];
                const float sumi = ai[v] + bi[v], difi = ai[v] - bi[v];
                const float p1 = fmaf(cs.x, sumi, -cs.y * difr);  // c*sumi - s*difr
                const float p2 = fmaf(cs.y, sumi, cs.x * difr);   // s*sumi + c*difr
                xr[v] = 0.5f * (sumr + p1); xi[v] = 0.5f * (difi - p2);
                yr[v] = 0.5f * (sumr - p1); yi[v] = -0.5f * (difi + p2);
            }
            emit(k, pa, xr, xi);
            emit(k == 0 ? H : H - k, pbm, yr, yi); // k == 0: the partner output is the Nyquist bin
        }
        if (gt == 0) {
            float cr[V], ci[V];
            ld_vec<V>(sre + padi(H / 2) * V, cr); ld_vec<V>(sim + padi(H / 2) * V, ci);
#pragma unroll
            for (int v = 0; v < V; ++v) ci[v] = -ci[v];
            emit(H / 2, padi(H / 2) * V, cr, ci);
        }
        } // !FUSED

        // ---- banded mel projection + dB -----------------------------------------------------------
        // Work item = (filter m, lane pl of the 2^lg lanes sharing it); 32 consecutive items form a block
        // and the host hands every warp of the group a balanced list of blocks (longest-first packing).
        // The filterbank taps and descriptors are first staged in the imaginary plane of the exchange
        // buffer -- it is dead once the spectrum has been read -- so the tap loop only touches shared memory.
        if (MEL && melp) {
            // Block-padded bank: work item i of block b reads taps wb[off_b + 32 j + lane] (zeros beyond its own)
            // and the unpadded magnitudes of bins bin0 + j P -- no predicates, no index arithmetic beyond one
            // add per tap; nj4_b taps for the whole block.
            constexpr int NWARPS = NT / 32;
            const int *__restrict__ sched = td->mel_cnt;
            const int nslots = __ldg(sched);
            const int lg = td->mel_log2p, P = 1 << lg;
            const int nwb = td->melp_nwb, nblk = td->melp_nblk;
            const float *wb = bank;
            const int *lo_s = reinterpret_cast<const int *>(bank) + nwb;
            const int2 *desc_s = reinterpret_cast<const int2 *>(lo_s + 32 * nblk);
            group_sync<G, NT>(grp); // magnitudes of all bins are in the buffer
            const int wg = gt >> 5, lane = gt & 31;
            const int stride = V << lg;
            for (int slot = 0; slot < nslots; ++slot) {
                const int blk = __ldg(sched + 4 + slot * NWARPS + wg); // warp-uniform
                if (blk < 0) continue;
                const int2 bd = desc_s[blk];
                const int li = lo_s[blk * 32 + lane];   // first bin | filter << 16
                const int m = (int)((unsigned)li >> 16), pl = lane & (P - 1);
                const float *wp = wb + bd.x + lane;
                const float *mp = sre + (li & 0xffff) * V;
                float acc[V];
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] = 0.0f;
                for (int j4 = 0; j4 < bd.y; j4 += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float wgt = wp[(j4 + u) * 32];
                        float mg[V];
                        ld_vec<V>(mp, mg);
                        mp += stride;
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[v] = fmaf(mg[v], wgt, acc[v]);
                    }
                }
                for (int sh = P >> 1; sh > 0; sh >>= 1)
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], sh);
                if (m < n_out && pl == 0) {
                    // Frames beyond the tile's last one were computed from a copy of that last frame (the clamp in
                    // the first pass), so their dB values may enter the extrema; only the store is predicated.
                    float *op = out + (size_t)(t0 + fl0) * n_out + m;
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const float y = amp_to_db_dev(acc[v]); // decibel.rs:33-88
                        vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
                        if (fl0 + v < nfr) op[(size_t)v * n_out] = y;
                    }
                }
            }
        } else if (MEL) {
            constexpr int NWARPS = NT / 32;
            const int *__restrict__ sched = td->mel_cnt; // {slots, taps, staged, 0, block ids [slots][NWARPS]}
            const int nslots = __ldg(sched), nnz = __ldg(sched + 1);
            const bool staged = bank_fits || __ldg(sched + 2) != 0;
            const int lg = td->mel_log2p, P = 1 << lg;
            const int4 *__restrict__ meta = reinterpret_cast<const int4 *>(td->mel_lo); // {lo, cnt, off, 0}
            const float *__restrict__ mw = td->mel_w;
            const bool dedicated = bank_fits; // taps already sit in the CTA's filterbank region
            float *wsm = dedicated ? bank : sim;
            int4 *msm = reinterpret_cast<int4 *>(wsm + ((nnz + 3) & ~3));
            if (staged && !dedicated) {
                if constexpr (!FUSED) group_sync<G, NT>(grp); // split pairs of other threads still read sim
                for (int i = gt; i < nnz; i += NT) wsm[i] = __ldg(mw + i);
                for (int i = gt; i < n_out; i += NT) msm[i] = __ldg(meta + i);
            }
            group_sync<G, NT>(grp); // magnitudes of all bins (and the staged tables) are in the buffer
            const int wg = gt >> 5, lane = gt & 31;
            auto mel_block = [&](auto staged_tag, int blk) {
                constexpr bool ST = decltype(staged_tag)::value;
                const int wi = blk * 32 + lane;
                const int m = wi >> lg, pl = wi & (P - 1);
                const bool valid = m < n_out;
                int4 mt = make_int4(0, 0, 0, 0);
                if (valid) mt = ST ? msm[m] : __ldg(meta + m);
                const int nj = mt.y > pl ? (mt.y - pl + P - 1) >> lg : 0; // taps of this lane: pl, pl+P, ...
                const int njmax = __reduce_max_sync(0xffffffffu, nj);
                const float *wp = (ST ? wsm : mw) + mt.z + pl;
                const int bin0 = mt.x + pl;
                float acc[V];
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] = 0.0f;
#pragma unroll 8
                for (int j = 0; j < njmax; ++j) {
                    const bool on = j < nj;
                    float wgt = 0.0f;
                    if (on) wgt = ST ? wp[j << lg] : __ldg(wp + (j << lg));
                    float mg[V];
                    ld_vec<V>(sre + padi(on ? bin0 + (j << lg) : 0) * V, mg);
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[v] = fmaf(mg[v], wgt, acc[v]);
                }
                for (int s = P >> 1; s > 0; s >>= 1)
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], s);
                if (valid && pl == 0) {
                    float *op = out + (size_t)(t0 + fl0) * n_out + m;
#pragma unroll
                    for (int v = 0; v < V; ++v) { // frames beyond the tile's last one are copies of it
                        const float y = amp_to_db_dev(acc[v]);
                        vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
                        if (fl0 + v < nfr) op[(size_t)v * n_out] = y;
                    }
                }
            };
            for (int slot = 0; slot < nslots; ++slot) {
                const int blk = __ldg(sched + 4 + slot * NWARPS + wg); // warp-uniform
                if (blk < 0) continue;
                if (staged) mel_block(std::true_type{}, blk); else mel_block(std::false_type{}, blk);
            }
        }
        group_sync<G, NT>(grp); // spectrum / magnitudes consumed before the next iteration overwrites the buffer
    }

    } // tiles of this CTA
    flush_range();
}

// =====================================================================================================
// K1W -- warp-per-frame variant of the fused analysis kernel for n_fft = 2048 (h = 1024 = 32 x 32).
//
// One warp owns one frame: lane m2 first holds the 32 points z[32 m1 + m2] and runs a 32-point DFT in
// registers, the twiddle W_1024^(m2 k1) is applied, one transpose through a private 8.25 KB shared
// buffer re-distributes the data so that lane k1 holds A[k1][m2] for all m2, and a second in-register
// 32-point DFT yields Z[k1 + 32 k2].  The conjugate partner of bin k1 + 32 k2 lives in lane 32 - k1,
// register 31 - k2, so the real-FFT split is one pair of warp shuffles per bin.  No block barrier
// exists on the frame path (16 independent warps per SM hide each other's latencies) and the spectrum
// crosses shared memory once instead of three times.  Frames are read straight from global memory
// with coalesced 64-bit loads; the 4x overlap between neighbouring frames -- which neighbouring warps
// of the same CTA process at the same time -- is served by L1.  Window, twiddle, split and mel tables
// are staged in shared memory once per (persistent) CTA.
// =====================================================================================================
constexpr int kWH = 1024;             // complex points
constexpr int kWWarps = 16;
constexpr int kWThreads = kWWarps * 32;
constexpr int kWXchg = 32 * 33;       // float2 elements of one warp's transpose buffer (pitch 33)

__host__ __device__ inline size_t k1w_smem_bytes(int nnz, int n_mel)
{
    return (size_t)kWWarps * kWXchg * sizeof(float2) + 2048 * sizeof(float) + 2 * kWH * sizeof(float2) +
           (size_t)(((nnz + 3) & ~3) + 64) * sizeof(float) + (size_t)n_mel * sizeof(int4) + 16;
}

__global__ void __launch_bounds__(kWThreads, 1) stft_warp_kernel(const StftLaunch L, int nnz, int n_mel_tab)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *xall = reinterpret_cast<float2 *>(smem_raw);
    float *win_s = reinterpret_cast<float *>(xall + kWWarps * kWXchg);
    float2 *tw2_s = reinterpret_cast<float2 *>(win_s + 2048);
    float2 *spl_s = tw2_s + kWH;
    float *wts_s = reinterpret_cast<float *>(spl_s + kWH);
    int4 *meta_s = reinterpret_cast<int4 *>(wts_s + ((nnz + 3) & ~3) + 64);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mode = L.mode;
    // ---- tables (all tracks of a launch share them; the host groups launches accordingly) --------------
    {
        const StftTrack *__restrict__ t0 = L.tracks;
        for (int i = tid; i < 2048; i += kWThreads) win_s[i] = __ldg(t0->win_f + i);
        for (int i = tid; i < kWH; i += kWThreads) { tw2_s[i] = __ldg(L.tw2 + i); spl_s[i] = __ldg(L.split_full + i); }
        if (mode == MODE_MEL_DB) {
            for (int i = tid; i < ((nnz + 3) & ~3) + 64; i += kWThreads) wts_s[i] = i < nnz ? __ldg(t0->mel_w + i) : 0.0f;
            const int4 *__restrict__ meta = reinterpret_cast<const int4 *>(t0->mel_lo);
            for (int i = tid; i < n_mel_tab; i += kWThreads) meta_s[i] = __ldg(meta + i);
        }
    }
    __syncthreads();

    float2 *xb = xall + warp * kWXchg;
    float *magbuf = reinterpret_cast<float *>(xb); // [1025] magnitudes of the current frame (mel mode)
    const float2 *win2 = reinterpret_cast<const float2 *>(win_s);

    int cur = 0; // current track (frames are enumerated track after track)
    float vmax = -INFINITY, vmin = INFINITY;
    auto flush_range = [&](const StftTrack *td) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, s));
        }
        if (lane == 0 && td->range_slot != nullptr && vmax >= vmin) {
            atomicMax(td->range_slot, enc_ordered(vmax));
            atomicMin(td->range_slot + 1, enc_ordered(vmin));
        }
        vmax = -INFINITY; vmin = INFINITY;
    };

    for (int g = blockIdx.x * kWWarps + warp; g < L.n_tiles; g += gridDim.x * kWWarps) {
        // ---- which track / frame (tile_begin holds the frame prefix: one "tile" per frame) ----------------
        int nxt = cur;
        while (nxt + 1 < L.n_tracks && L.tracks[nxt + 1].tile_begin <= g) ++nxt;
        if (nxt != cur) { flush_range(L.tracks + cur); cur = nxt; }
        const StftTrack *__restrict__ td = L.tracks + cur;
        const PcmView pv{td->pcm, td->n, td->ch, td->fmt, td->origin, td->avail};
        const int t = g - td->tile_begin;
        const long long S0 = (long long)(td->frame0 + t) * td->hop - td->win / 2 - td->pad_l; // global
        const long long Sl = S0 - pv.origin;                                                  // local to the slice
        float *__restrict__ out = td->out;
        const int n_out = td->n_out;

        float re[32][1], im[32][1];
        // ---- A: windowed samples, z[32 m1 + lane] = (g[2m], g[2m+1]) ------------------------------------------
        const bool interior = pv.ch == 1 && pv.fmt == PCM_F32 && S0 >= 0 && S0 + 2 * kWH <= pv.n && Sl >= 0 &&
                              Sl + 2 * kWH <= pv.avail;
        if (interior && ((Sl & 1) == 0) && ((reinterpret_cast<uintptr_t>(pv.pcm) & 7) == 0)) {
            const float2 *__restrict__ p = reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(pv.pcm) + Sl) + lane;
#pragma unroll
            for (int m1 = 0; m1 < 32; ++m1) {
                const float2 x = __ldg(p + 32 * m1);
                const float2 w = win2[32 * m1 + lane];
                re[m1][0] = x.x * w.x; im[m1][0] = x.y * w.y;
            }
        } else if (interior) {
            const float *__restrict__ p = reinterpret_cast<const float *>(pv.pcm) + Sl + 2 * lane;
#pragma unroll
            for (int m1 = 0; m1 < 32; ++m1) {
                const float2 w = win2[32 * m1 + lane];
                re[m1][0] = __ldg(p + 64 * m1) * w.x; im[m1][0] = __ldg(p + 64 * m1 + 1) * w.y;
            }
        } else {
#pragma unroll
            for (int m1 = 0; m1 < 32; ++m1) {
                const float2 w = win2[32 * m1 + lane];
                const long long i = S0 + 64 * m1 + 2 * lane;
                re[m1][0] = load_sample(pv, i) * w.x; im[m1][0] = load_sample(pv, i + 1) * w.y;
            }
        }
        // ---- B: 32-point DFT over m1, twiddle W_1024^(lane k1) ----------------------------------------------------
        dft_inplace<32, 0, 1, 32, 1>(re, im);
#pragma unroll
        for (int k1 = 1; k1 < 32; ++k1) {
            const float2 w = tw2_s[k1 * 32 + lane];
            const float xr = re[k1][0], xi = im[k1][0];
            re[k1][0] = xr * w.x - xi * w.y; im[k1][0] = xr * w.y + xi * w.x;
        }
        // ---- C: transpose through the warp's buffer (pitch 33 float2: both sides conflict free) -----------------
        __syncwarp(); // previous frame's magnitudes are consumed
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) xb[k1 * 33 + lane] = make_float2(re[k1][0], im[k1][0]);
        __syncwarp();
#pragma unroll
        for (int m2 = 0; m2 < 32; ++m2) { const float2 v = xb[lane * 33 + m2]; re[m2][0] = v.x; im[m2][0] = v.y; }
        __syncwarp();
        // ---- D: 32-point DFT over m2 -> Z[lane + 32 k2] in register k2 -----------------------------------------------
        dft_inplace<32, 0, 1, 32, 1>(re, im);
        // ---- E: real-FFT split (realfft.rs:140-157); partner bin lives in lane 32-lane, register 31-k2 ------------
        const int pl = (32 - lane) & 31;
        const bool l0 = lane == 0;
        const size_t row = (size_t)t * (kWH + 1);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            // lane 0 pairs k = 32 j with 32 (32 - j): it offers register (32 - j) & 31 instead of 31 - j
            const float sr = l0 ? re[(32 - j) & 31][0] : re[31 - j][0];
            const float si = l0 ? im[(32 - j) & 31][0] : im[31 - j][0];
            const float br = __shfl_sync(0xffffffffu, sr, pl), bi = __shfl_sync(0xffffffffu, si, pl);
            const float ar = re[j][0], ai = im[j][0];
            const int k = lane + 32 * j;
            const float2 cs = spl_s[k]; // (cos, sin)(k pi / h)
            const float sumr = ar + br, difr = ar - br, sumi = ai + bi, difi = ai - bi;
            const float xr = 0.5f * (sumr + fmaf(cs.x, sumi, -cs.y * difr));
            const float xi = 0.5f * (difi - fmaf(cs.y, sumi, cs.x * difr));
            if (mode == MODE_COMPLEX) {
                reinterpret_cast<float2 *>(out)[row + k] = make_float2(xr, xi);
            } else {
                const float mg = sqrt_approx(fmaf(xr, xr, xi * xi)); // lib.rs:124
                if (mode == MODE_MEL_DB) magbuf[k] = mg;
                else {
                    float y = mg;
                    if (mode == MODE_LIN_DB) { y = amp_to_db_dev(y); vmax = fmaxf(vmax, y); vmin = fminf(vmin, y); }
                    out[row + k] = y;
                }
            }
        }
        if (l0) { // Nyquist bin: Z[0].re - Z[0].im (realfft.rs:157)
            const float xr = re[0][0] - im[0][0];
            if (mode == MODE_COMPLEX) reinterpret_cast<float2 *>(out)[row + kWH] = make_float2(xr, 0.0f);
            else {
                const float mg = fabsf(xr);
                if (mode == MODE_MEL_DB) magbuf[kWH] = mg;
                else {
                    float y = mg;
                    if (mode == MODE_LIN_DB) { y = amp_to_db_dev(y); vmax = fmaxf(vmax, y); vmin = fminf(vmin, y); }
                    out[row + kWH] = y;
                }
            }
        }
        // ---- F: banded mel projection + dB (lanes <-> filters) --------------------------------------------------------
        if (mode == MODE_MEL_DB) {
            __syncwarp();
            const int lg = td->mel_log2p, P = 1 << lg;
            const int items = n_out << lg;
            for (int w0 = 0; w0 < items; w0 += 32) {
                const int wi = w0 + lane;
                const int m = wi >> lg, plm = wi & (P - 1);
                const bool valid = m < n_out;
                int4 mt = make_int4(0, 0, 0, 0);
                if (valid) mt = meta_s[m];
                const int nj = mt.y > plm ? (mt.y - plm + P - 1) >> lg : 0;
                const int njmax = __reduce_max_sync(0xffffffffu, nj);
                const float *wp = wts_s + mt.z + plm;
                const float *mp = magbuf + mt.x + plm;
                float acc = 0.0f;
                // loads are unconditional (the tables are zero-padded and the buffer is larger than the
                // spectrum); taps past this lane's band are zeroed by the select
                if (lg == 0) {
#pragma unroll 4
                    for (int jj = 0; jj < njmax; ++jj) acc = fmaf(mp[jj], jj < nj ? wp[jj] : 0.0f, acc);
                } else {
#pragma unroll 2
                    for (int jj = 0; jj < njmax; ++jj) {
                        const bool on = jj < nj;
                        acc = fmaf(on ? mp[jj << lg] : 0.0f, on ? wp[jj << lg] : 0.0f, acc);
                    }
                }
                for (int s = P >> 1; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
                if (valid && plm == 0) {
                    const float y = amp_to_db_dev(acc);
                    vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
                    out[(size_t)t * n_out + m] = y;
                }
            }
        }
    }
    if (mode == MODE_LIN_DB || mode == MODE_MEL_DB) flush_range(L.tracks + cur);
}

// ---- small-F fallback: one CTA per frame, radix-2 Stockham, any power-of-two F >= 2 --------------------
// Covers the reference's known-answer shapes (n_fft = 4, 256) and anything below the tuned sizes.
__global__ void __launch_bounds__(128) stft_generic_kernel(const StftLaunch L, int h)
{
    extern __shared__ __align__(16) float sm[];
    float2 *a = reinterpret_cast<float2 *>(sm);
    float2 *b = a + h;
    float *mag = reinterpret_cast<float *>(b + h); // [h+1]
    __shared__ float red_max[4], red_min[4];

    const int tid = threadIdx.x;
    const int tile_id = blockIdx.x;
    int lo = 0, hi = L.n_tracks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (L.tracks[mid].tile_begin <= tile_id) lo = mid; else hi = mid - 1;
    }
    const StftTrack *__restrict__ td = L.tracks + lo;
    const PcmView pv{td->pcm, td->n, td->ch, td->fmt, td->origin, td->avail};
    const int t = tile_id - td->tile_begin; // one frame per CTA
    const int F = 2 * h, n_out = td->n_out, mode = L.mode;
    const long long S0 = (long long)(td->frame0 + t) * td->hop - td->win / 2 - td->pad_l;
    float *__restrict__ out = td->out;

    for (int m = tid; m < h; m += blockDim.x) {
        const float w0 = td->win_f[2 * m], w1 = td->win_f[2 * m + 1];
        a[m] = make_float2(load_sample(pv, S0 + 2 * m) * w0, load_sample(pv, S0 + 2 * m + 1) * w1);
    }
    __syncthreads();
    for (int ns = 1; ns < h; ns <<= 1) {
        for (int j = tid; j < h / 2; j += blockDim.x) {
            const int k = j & (ns - 1);
            float sn, cs;
            sincospif(-(float)k / (float)ns, &sn, &cs); // exp(-2 pi i k / (2 ns))
            const float2 u = a[j], x = a[j + h / 2];
            const float2 v = make_float2(x.x * cs - x.y * sn, x.x * sn + x.y * cs);
            const int d = (j - k) * 2 + k;
            b[d] = make_float2(u.x + v.x, u.y + v.y);
            b[d + ns] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
        float2 *tmp = a; a = b; b = tmp;
    }
    float vmax = -INFINITY, vmin = INFINITY;
    for (int k = tid; k <= h; k += blockDim.x) {
        float xr, xi;
        if (k == h) { xr = a[0].x - a[0].y; xi = 0.0f; }  // realfft.rs:157
        else {
            const float2 p = a[k], q = a[(h - k) & (h - 1)];
            float sn, cs;
            sincospif((float)k / (float)h, &sn, &cs);
            xr = 0.5f * (((p.x + q.x) + cs * (p.y + q.y)) - sn * (p.x - q.x));
            xi = 0.5f * (((p.y - q.y) - sn * (p.y + q.y)) - cs * (p.x - q.x));
        }
        if (mode == MODE_COMPLEX) {
            reinterpret_cast<float2 *>(out)[(size_t)t * (h + 1) + k] = make_float2(xr, xi);
        } else {
            const float mg = sqrtf(fmaf(xr, xr, xi * xi));
            if (mode == MODE_MEL_DB) mag[k] = mg;
            else {
                float y = mg;
                if (mode == MODE_LIN_DB) { y = amp_to_db_dev(y); vmax = fmaxf(vmax, y); vmin = fminf(vmin, y); }
                out[(size_t)t * (h + 1) + k] = y;
            }
        }
    }
    if (mode == MODE_MEL_DB) {
        __syncthreads();
        for (int m = tid; m < n_out; m += blockDim.x) {
            const int4 mt = reinterpret_cast<const int4 *>(td->mel_lo)[m];
            const int blo = mt.x, bcnt = mt.y, boff = mt.z;
            float acc = 0.0f;
            for (int i = 0; i < bcnt; ++i) acc = fmaf(mag[blo + i], td->mel_w[boff + i], acc);
            const float y = amp_to_db_dev(acc);
            vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
            out[(size_t)t * n_out + m] = y;
        }
    }
    (void)F;
    if (td->range_slot != nullptr && (mode == MODE_LIN_DB || mode == MODE_MEL_DB)) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, s));
        }
        if ((tid & 31) == 0) { red_max[tid >> 5] = vmax; red_min[tid >> 5] = vmin; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (int)blockDim.x / 32; ++w) { vmax = fmaxf(vmax, red_max[w]); vmin = fminf(vmin, red_min[w]); }
            if (vmax >= vmin) {
                atomicMax(td->range_slot, enc_ordered(vmax));
                atomicMin(td->range_slot + 1, enc_ordered(vmin));
            }
        }
    }
}

int resident_sms()
{
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

template <int LOG2H, int PTS, int V, int G, int MC, bool MEL, bool RAW2>
cudaError_t launch_one_mode(const StftLaunch &L, size_t smem, cudaStream_t stream)
{
    using TR = K1Traits<LOG2H, PTS, V, G, MC>;
    auto kern = stft_db_kernel<LOG2H, PTS, V, G, MC, MEL, RAW2>;
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
    if (e != cudaSuccess) return e;
    // persistent CTAs: as many as are resident at once, each walking tiles blockIdx.x + k gridDim.x
    const int grid = std::min(L.n_tiles, resident_sms() * MC);
    kern<<<grid, TR::THREADS, smem, stream>>>(L);
    count_launch();
    return cudaGetLastError();
}
template <int LOG2H, int PTS, int V, int G, int MC>
cudaError_t launch_one(const StftLaunch &L, size_t smem, cudaStream_t stream)
{
    // raw stereo staging exists for the mel kernels (the viewer's default scale); linear launches gather stereo tiles
    if (L.mode == MODE_MEL_DB)
        return L.stereo_raw ? launch_one_mode<LOG2H, PTS, V, G, MC, true, true>(L, smem, stream)
                            : launch_one_mode<LOG2H, PTS, V, G, MC, true, false>(L, smem, stream);
    return launch_one_mode<LOG2H, PTS, V, G, MC, false, false>(L, smem, stream);
}

} // namespace

// FFT size -> kernel instantiations (LOG2H, PTS, V, G, resident CTAs).  The first entry of a size is the
// default; the others are selectable with SGX_K1_VARIANT="pts,v,g" (tuning experiments).  Measured on B200 for
// h = 1024 (8 tracks x 10 min, K1 ms): (8,4,2) 2.35 | (8,2,2) 3.09 | (8,2,4) 3.39 | (4,4,1) 2.80 | (4,4,2) 2.85 |
// (16,2,4) 3.16 | (16,2,2) 3.38 -- sharing index math, twiddles and mel taps across V = 4 frames outweighs the
// higher occupancy of the V = 2 variants and the fewer exchanges of radix 16.
#ifdef SGX_K1_ALTERNATES // the other CTA shapes of the design-space measurements (make TUNE=-DSGX_K1_ALTERNATES)
#define SGX_K1_ALT(X)   \
    X(9, 8, 4, 8, 1)    \
    X(10, 8, 4, 2, 2)   \
    X(10, 8, 2, 2, 3)   \
    X(10, 4, 4, 1, 4)   \
    X(11, 8, 4, 1, 2)
#else
#define SGX_K1_ALT(X)
#endif
#define SGX_K1_TABLE(X) \
    X(8, 8, 4, 8, 2)    \
    X(9, 8, 4, 4, 2)    \
    X(10, 8, 4, 4, 1)   \
    X(11, 8, 4, 2, 1)   \
    X(12, 8, 4, 1, 1)   \
    X(13, 8, 2, 1, 1)   \
    SGX_K1_ALT(X)

bool stft_config_for(size_t n_fft, StftConfig *cfg)
{
    if (n_fft < 2 || (n_fft & (n_fft - 1)) != 0 || n_fft > 16384) return false;
    const int h = (int)(n_fft / 2);
    cfg->n_fft = (int)n_fft; cfg->h = h; cfg->generic = true; cfg->fused = false;
    cfg->pts = 2; cfg->vec = 1; cfg->groups = 1; cfg->threads = 128; cfg->min_ctas = 1;
    cfg->fft_smem = (size_t)(2 * h) * sizeof(float2) + (size_t)(h + 1) * sizeof(float);
    int want_pts = 0, want_v = 0, want_g = 0;
    if (const char *e = getenv("SGX_K1_VARIANT")) sscanf(e, "%d,%d,%d", &want_pts, &want_v, &want_g);
    bool chosen = false;
#define X(LG, PTS, V, G, MC)                                                                     \
    if (h == (1 << LG)) {                                                                        \
        using TR = K1Traits<LG, PTS, V, G, MC>;                                                      \
        const bool match = want_pts == PTS && want_v == V && want_g == G;                        \
        if (!chosen || match) {                                                                  \
            if (cfg->generic || match) {                                                         \
                cfg->generic = false; cfg->pts = PTS; cfg->vec = V; cfg->groups = G;             \
                cfg->threads = TR::THREADS; cfg->fft_smem = TR::FFT_SMEM; cfg->min_ctas = TR::MIN_CTAS; \
                cfg->fused = (PTS / last_radix(1 << LG, PTS)) >= 2;                                  \
            }                                                                                    \
            chosen = chosen || match;                                                            \
        }                                                                                        \
    }
    SGX_K1_TABLE(X)
#undef X
    // n_fft = 2048: SGX_K1W=1 selects the warp-per-frame kernel instead of the block kernel above.  Measured on
    // B200 (C5): 65 % issue-slot use vs 51 %, but 1.25x the instructions (no sharing of index math, tables and
    // mel taps across 4 frames) -> 9.9 ms vs 9.4 ms; kept as an evaluated alternative, off by default.
    static const bool k1w_on = getenv("SGX_K1W") && atoi(getenv("SGX_K1W")) == 1;
    cfg->warp_per_frame = false;
    if (h == kWH && k1w_on && want_pts == 0) {
        cfg->warp_per_frame = true; cfg->generic = false; cfg->fused = false;
        cfg->pts = 32; cfg->vec = 1; cfg->groups = kWWarps; cfg->threads = kWThreads; cfg->min_ctas = 1;
        cfg->fft_smem = 0;
    }
    return true;
}

size_t stft_warp_smem_bytes(int nnz, int n_mel) { return k1w_smem_bytes(nnz, n_mel); }

size_t stft_max_dynamic_smem() { return 227 * 1024; }

StftTiling plan_stft_tiles(const StftConfig &cfg, int max_hop, int bank_floats, int sample_floats)
{
    StftTiling t{};
    if (cfg.warp_per_frame) { // one "tile" per frame; nothing is staged per tile
        t.frames_per_tile = 1; t.staged = 0; t.tile_floats = 0; t.smem_bytes = 0;
        return t;
    }
    if (cfg.generic) {
        t.frames_per_tile = 1; t.staged = 0; t.tile_floats = 0;
        t.smem_bytes = cfg.fft_smem;
        return t;
    }
    const int unit = cfg.groups * cfg.vec;
    // per-CTA shared-memory budget for the register-limited number of resident CTAs (228 KB per SM,
    // 1 KB reserved per CTA); the staged tile gets what the FFT buffers leave
    const int ctas = cfg.min_ctas > 0 ? cfg.min_ctas : 1;
    size_t budget = (size_t)(228 * 1024) / ctas - 1024 - 512;
    if (budget > stft_max_dynamic_smem()) budget = stft_max_dynamic_smem();
    int want_nfr = 0;
    if (const char *e = getenv("SGX_K1_NFR")) want_nfr = atoi(e);
    static const bool no_bank = getenv("SGX_K1_NOBANK") && atoi(getenv("SGX_K1_NOBANK")) == 1;
    if (no_bank) bank_floats = 0;
    // First choice: the filterbank of a track gets its own region (loaded once per persistent CTA and
    // track) next to a staged tile of at least one round of frames; if that does not fit, the taps are
    // staged per round in the dead imaginary plane (or read through L1), as the mel schedule says.
    for (int with_bank = bank_floats > 0 ? 1 : 0; with_bank >= 0; --with_bank) {
        const size_t fixed = cfg.fft_smem + 16 + (with_bank ? (size_t)bank_floats * sizeof(float) : 0);
        const long cap_floats = budget > fixed ? (long)((budget - fixed) / sizeof(float)) : 0;
        // floats(nfr) = 3 + (nfr-1)*hop + F, rounded up to 4
        int best = 0;
        for (int mult = 1; mult <= 8; ++mult) {
            const int nfr = unit * mult;
            const long need = (3 + (long)(nfr - 1) * max_hop + cfg.n_fft + 4) * sample_floats;
            if (need <= cap_floats && (want_nfr == 0 || nfr <= want_nfr || best == 0)) best = nfr;
        }
        if (best == 0 && with_bank) continue; // rather stage the tile than the taps
        t.bank_floats = with_bank ? bank_floats : 0;
        if (best == 0) {
            t.frames_per_tile = unit; t.staged = 0; t.tile_floats = 0;
        } else {
            t.frames_per_tile = best; t.staged = 1;
            long need = 3 + (long)(best - 1) * max_hop + cfg.n_fft;
            t.tile_floats = ((int)((need + 3) & ~3L) + 4) * sample_floats;
        }
        break;
    }
    t.sample_floats = sample_floats;
    t.smem_bytes = 16 + (size_t)(t.tile_floats + t.bank_floats) * sizeof(float) + cfg.fft_smem;
    return t;
}

void make_warp_fft_tables(float2 *tw2 /*[32*32]*/, float2 *split_full /*[1024]*/)
{
    const double pi = 3.14159265358979323846264338327950288;
    for (int k1 = 0; k1 < 32; ++k1)
        for (int lane = 0; lane < 32; ++lane) {
            const double a = -2.0 * pi * (double)(k1 * lane) / (double)kWH; // W_1024^(lane k1)
            tw2[k1 * 32 + lane] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int k = 0; k < kWH; ++k) {
        const double a = pi * (double)k / (double)kWH;
        split_full[k] = make_float2((float)cos(a), (float)sin(a));
    }
}

void make_fft_tables(int h, float2 *tw, float2 *split)
{
    const double pi = 3.14159265358979323846264338327950288;
    for (int j = 0; j < h; ++j) {
        const double a = -2.0 * pi * (double)j / (double)h;
        tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int k = 0; k <= h / 2; ++k) {
        const double a = pi * (double)k / (double)h;
        split[k] = make_float2((float)cos(a), (float)sin(a));
    }
}

cudaError_t launch_stft(const StftConfig &cfg, const StftLaunch &L, cudaStream_t stream)
{
    if (L.n_tiles <= 0) return cudaSuccess;
    if (cfg.warp_per_frame) {
        const size_t smem = k1w_smem_bytes(L.mel_nnz, L.mel_rows);
        cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(stft_warp_kernel), smem);
        if (e != cudaSuccess) return e;
        int sms = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        const int grid = std::min(sms, (L.n_tiles + kWWarps - 1) / kWWarps);
        stft_warp_kernel<<<grid, kWThreads, smem, stream>>>(L, L.mel_nnz, L.mel_rows);
        count_launch();
        return cudaGetLastError();
    }
    const size_t smem = cfg.generic ? cfg.fft_smem
                                    : 16 + (size_t)(L.tile_floats + L.bank_floats) * sizeof(float) + cfg.fft_smem;
    if (cfg.generic) {
        cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(stft_generic_kernel), smem);
        if (e != cudaSuccess) return e;
        stft_generic_kernel<<<L.n_tiles, 128, smem, stream>>>(L, cfg.h);
        count_launch();
        return cudaGetLastError();
    }
#define X(LG, PTS, V, G, MC) \
    if (cfg.h == (1 << LG) && cfg.pts == PTS && cfg.vec == V && cfg.groups == G) return launch_one<LG, PTS, V, G, MC>(L, smem, stream);
    SGX_K1_TABLE(X)
#undef X
    return cudaErrorInvalidValue;
}

} // namespace sgx
