// stft_warp1_kernel.cu -- K1 for n_fft = 2048 (h = 1024 = 32 x 32): one WARP transforms one frame, complex values as
// packed FP32 pairs.
//
// Same path as stft_kernel.cu (channel sum lib.rs:42 -> reflect-padded framing lib.rs:412-433 -> window and centred
// zero-pad lib.rs:377-384 -> real FFT as an h-point complex FFT + split realfft.rs:105-159 -> |X| lib.rs:124 -> banded
// mel lib.rs:131 -> dB decibel.rs:33-88 -> per-track extrema lib.rs:197-200) and the decomposition of
// stft_warp1_kernel.cu (lane m2 holds the 32 points z[32 m1 + m2], 32-point DFT in registers, twiddle, ONE transpose
// through a private plane, second DFT, conjugate partners met by warp shuffles, no block barrier on the frame path).
// The difference is WHAT a register pair holds.  The warp-pair kernel packs the same point of TWO frames: 128 data
// registers, 8 warps per SM, and it ends up bound by per-warp latency.  Here a pair is ONE complex value (re, im) of one
// frame -- sm_100's packed instructions take a half-swap and a per-half negation as free operand modifiers
// (`R.F32x2.LO_HI`, `.NP` in SASS), so
//     complex add / sub ....................... one FADD2
//     times a twiddle (c, s) ................... FMUL2 by (c, c) + FFMA2 of the swapped, half-negated value by (s, s)
//     times -i (the trivial twiddle) .......... nothing: folded into the operand of the add that consumes it
//     split of a conjugate pair (realfft.rs:140-157) ... 8 packed instructions for both bins
// i.e. the same instruction count per frame as packing two frames, with HALF the registers: 64 data registers, 16 warps
// per SM.  The mel projection feeds U (rising side) and D (falling side) of a segment with ONE FFMA2 per tap: the
// magnitude is the broadcast scalar operand, the weight pair the packed one.
//
// MEASURED (B200, C5, K1 ms per step; profiles/r02_k1_w1_packed_complex_experiment.txt): 16 warps per SM at 128 registers,
// no spills, parity-green -- and 6.48 ms against 6.33 (block kernel) and 6.42 (warp-pair kernel).  The occupancy doubled as
// planned (issue slots 60 % busy instead of 44 %), but a warp that holds ONE frame re-reads every table for every frame:
// window 64 + inter-DFT twiddles 62 + split constants 32 + mel weights 120 shared-memory wavefronts per frame, where the
// warp-pair kernel shares them between two frames and the block kernel between four.  Together with the transpose (128),
// the tile (64) and the magnitudes that is 750 wavefronts per frame -- more than the block kernel's 720 -- and 2,350
// instructions per frame instead of 1,670 / 1,970: the shared-memory pipe is the limit again (74 % + the L1 side, 78 % in
// all).  Registers buy either table reuse (frames per thread) or occupancy, not both: three decompositions, one wall
// (6.33 / 6.42 / 6.48 ms).  Kept as a parity-tested alternative (SGX_K1W1=1; SGX_W1_WARPS = 8 | 12 | 16).
#include <cstdint>
#include <cstdlib>
#include <algorithm>

#include "device_common.cuh"
#include "kernels.h"
#include "stft_device.cuh"

namespace sgx {

namespace {

constexpr int kW1H = 1024;                  // complex points
constexpr int kW1Pitch = 33;                // float2 elements per row of a warp's exchange plane (conflict-free both ways)
constexpr int kW1Plane = 32 * kW1Pitch;     // 8448 bytes; afterwards holds the 1025 magnitudes of the frame

typedef float2 cx; // (re, im) in one register pair
__device__ __forceinline__ cx cadd(cx a, cx b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ cx csub(cx a, cx b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ cx mul_mi(cx z) { return make_float2(z.y, -z.x); } // z * (-i): operand modifiers of the consumer
// z * (c - i s) = (re c + im s, im c - re s): the same two roundings per component as the scalar form of dft_inplace
__device__ __forceinline__ cx cmul_conj(cx z, float c, float s)
{
    return __ffma2_rn(make_float2(z.y, -z.x), make_float2(s, s), __fmul2_rn(z, make_float2(c, c)));
}
// z * (wx + i wy) = (re wx - im wy, im wx + re wy)
__device__ __forceinline__ cx cmul(cx z, float wx, float wy)
{
    return __ffma2_rn(make_float2(-z.y, z.x), make_float2(wy, wy), __fmul2_rn(z, make_float2(wx, wx)));
}
// in-register DFT of R points at z[BASE + i STRIDE], natural order in and out (radix-2 decimation in time, as dft_inplace)
template <int R, int BASE, int STRIDE>
__device__ __forceinline__ void cdft(cx (&z)[32])
{
    if constexpr (R == 2) {
        const cx a = z[BASE], b = z[BASE + STRIDE];
        z[BASE] = cadd(a, b); z[BASE + STRIDE] = csub(a, b);
    } else if constexpr (R > 2) {
        cdft<R / 2, BASE, 2 * STRIDE>(z);          // even inputs
        cdft<R / 2, BASE + STRIDE, 2 * STRIDE>(z); // odd inputs
        cx t[R];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            const int e = BASE + 2 * k * STRIDE, o = BASE + (2 * k + 1) * STRIDE;
            const int widx = k * (32 / R); // exp(-2 pi i k / R) = kC32[widx] - i kS32[widx]
            cx p;
            if (widx == 0) p = z[o];
            else if (widx == 8) p = mul_mi(z[o]);
            else p = cmul_conj(z[o], kC32[widx], kS32[widx]);
            t[k] = cadd(z[e], p); t[k + R / 2] = csub(z[e], p);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) z[BASE + k * STRIDE] = t[k];
    }
}

template <bool MEL, int kW1Warps>
__global__ void __launch_bounds__(kW1Warps * 32, 1) stft_warp1_kernel(const StftLaunch L)
{
    constexpr int H = kW1H, F = 2 * kW1H, kW1Threads = kW1Warps * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned *done_cnt = reinterpret_cast<unsigned *>(smem_raw + 8); // warps that have consumed the current tile
    float *tile = reinterpret_cast<float *>(smem_raw + 16);
    float2 *xall = reinterpret_cast<float2 *>(tile + L.tile_floats);
    float2 *win_s = xall + kW1Warps * kW1Plane;  // [h]     (w[2m], w[2m+1]) of the current track
    float2 *tw_s = win_s + H;                    // [32][32] W_1024^(k1 lane)
    float2 *spl_s = tw_s + H;                    // [h/2]   (cos, sin)(k pi / h)                   realfft.rs:88-93
    float *bank = reinterpret_cast<float *>(spl_s + H / 2); // block-padded mel bank of the current track

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2 *xb = xall + warp * kW1Plane;
    const int mode = MEL ? (int)MODE_MEL_DB : L.mode;

    for (int i = tid; i < H; i += kW1Threads) tw_s[i] = __ldg(L.tw + (((i >> 5) * (i & 31)) & (H - 1)));
    for (int i = tid; i < H / 2; i += kW1Threads) spl_s[i] = __ldg(L.split + i);
    if (tid == 0) { mbar_init(mbar, 1); *done_cnt = 0u; }
    __syncthreads();
    unsigned phase = 0;
    const float *bank_src = nullptr, *win_src = nullptr; // whose tables the shared copies hold

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
        for (int i = tid; i < (int)(sizeof(StftTrack) / sizeof(int)); i += kW1Threads) dst[i] = src[i];
        __syncthreads();
    };
    enter_track(blockIdx.x, 0);
    int trk_end = s_trk_end;
    const StftTrack *td = &s_td;
    TileLoc cur;
    locate_tile(L, F, blockIdx.x, s_trk, td, cur, false);
    if (cur.tma && tid == 0) issue_tile_copy(td, cur, tile, mbar);

    float vmax = -INFINITY, vmin = INFINITY;
    int range_trk = -1;
    auto flush_range = [&]() { // per-track extrema (lib.rs:197-200): flushed when the CTA moves on to another track
        if (range_trk < 0 || !(mode == MODE_LIN_DB || mode == MODE_MEL_DB)) return;
        unsigned *slot = L.tracks[range_trk].range_slot;
        if (slot == nullptr) return;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, s));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, s));
        }
        if (lane == 0 && vmax >= vmin) {
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
            locate_tile(L, F, tile_id, s_trk, td, cur, false);
        }
        if (cur.trk != range_trk) { range_trk = cur.trk; vmax = -INFINITY; vmin = INFINITY; }
        const PcmView pv{td->pcm, td->n, td->ch, td->fmt, td->origin, td->avail};
        const int hop = td->hop;
        float *__restrict__ out = td->out;
        const int n_out = td->n_out;
        const int t0 = cur.t0, nfr = cur.nfr, off0 = cur.off0;

        // ---- window (and filterbank) of this track into shared memory, once per CTA and track ------------------
        if (td->win_f != win_src || (MEL && td->mel_w != bank_src)) {
            __syncthreads(); // other warps may still be working on frames of the previous track
            const float2 *__restrict__ wsrc = reinterpret_cast<const float2 *>(td->win_f);
            for (int i = tid; i < H; i += kW1Threads) win_s[i] = __ldg(wsrc + i);
            if (MEL) {
                const int words = td->seg_words;
                const int *__restrict__ srcw = td->segp;
                int *dstw = reinterpret_cast<int *>(bank);
                for (int i = tid; i < words; i += kW1Threads) dstw[i] = __ldg(srcw + i);
            }
            __syncthreads();
            win_src = td->win_f; bank_src = td->mel_w;
        }

        // ---- the PCM tile: landed by TMA (issued one tile ago), or gathered here (edges, int16, stereo) ---------
        if (cur.tma) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
        } else {
            __syncthreads(); // every warp is past its loads of the previous tile: the buffer is free
            for (int s = tid; s < cur.len; s += kW1Threads) tile[s] = load_sample(pv, cur.A0 + s);
            __syncthreads();
        }

        const int rounds = L.frames_per_tile / kW1Warps;
        for (int r = 0; r < rounds; ++r) {
            if (r * kW1Warps >= nfr) break; // uniform
            const int fl0 = r * kW1Warps + warp; // local frame of this warp
            const bool active = fl0 < nfr;       // warp-uniform
            cx z[32];

            // ---- A: windowed samples, z[32 m1 + lane] = (g[2m] w[2m], g[2m+1] w[2m+1]): one FMUL2 per point ------------
            if (active) {
                const int b0 = off0 + fl0 * hop;
                const float *f0 = tile + b0 + 2 * lane;
                if ((b0 & 1) == 0) { // the frame starts on an even float: one 64-bit shared load per point
#pragma unroll
                    for (int m1 = 0; m1 < 32; ++m1)
                        z[m1] = __fmul2_rn(*reinterpret_cast<const float2 *>(f0 + 64 * m1), win_s[32 * m1 + lane]);
                } else {
#pragma unroll
                    for (int m1 = 0; m1 < 32; ++m1)
                        z[m1] = __fmul2_rn(make_float2(f0[64 * m1], f0[64 * m1 + 1]), win_s[32 * m1 + lane]);
                }
            }
            // ---- the tile is in registers: let the next one stream in (last warp to check in issues the copy) ------
            if (r == rounds - 1 || (r + 1) * kW1Warps >= nfr) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    const unsigned seen = atomicAdd(done_cnt, 1u);
                    if (seen % kW1Warps == kW1Warps - 1) {
                        const int nt = tile_id + (int)gridDim.x;
                        if (nt < L.n_tiles) {
                            TileLoc nx;
                            const StftTrack *ntd = td;
                            int ntrk = cur.trk;
                            if (nt >= trk_end) { ntrk = find_track(L, nt, cur.trk); ntd = L.tracks + ntrk; }
                            locate_tile(L, F, nt, ntrk, ntd, nx, false);
                            if (nx.tma) issue_tile_copy(ntd, nx, tile, mbar);
                        }
                    }
                }
            }
            if (!active) continue;

            // ---- B: 32-point DFT over m1, then the twiddle W_1024^(lane k1) ----------------------------------------
            cdft<32, 0, 1>(z);
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) {
                const float2 w = tw_s[k1 * 32 + lane];
                z[k1] = cmul(z[k1], w.x, w.y);
            }
            // ---- C: transpose (lane m2, register k1) -> (lane k1, register m2) through the warp's plane --------------
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) xb[k1 * kW1Pitch + lane] = z[k1];
            __syncwarp();
#pragma unroll
            for (int m2 = 0; m2 < 32; ++m2) z[m2] = xb[lane * kW1Pitch + m2];
            __syncwarp(); // the plane now becomes the magnitude array (mel)
            // ---- D: 32-point DFT over m2 -> Z[lane + 32 k2] in register k2 --------------------------------------------
            cdft<32, 0, 1>(z);

            // ---- E: real-FFT split (realfft.rs:140-157) + what becomes of a bin ------------------------------------
            float *magbuf = reinterpret_cast<float *>(xb); // [h + 1] magnitudes at their bin index (mel)
            // X: the bin's value; `conj`: its imaginary part is still to be negated (only the complex output cares)
            auto emit = [&](int idx, cx X, bool conj) {
                if (mode == MODE_COMPLEX) {
                    reinterpret_cast<float2 *>(out)[(size_t)(t0 + fl0) * (H + 1) + idx] = make_float2(X.x, conj ? -X.y : X.y);
                    return;
                }
                const float mg = sqrt_approx(fmaf(X.x, X.x, X.y * X.y)); // lib.rs:124
                if (MEL) {
                    magbuf[idx] = mg;
                } else {
                    float y = mg;
                    if (mode == MODE_LIN_DB) { y = amp_to_db_dev(y); vmax = fmaxf(vmax, y); vmin = fminf(vmin, y); }
                    out[(size_t)(t0 + fl0) * (H + 1) + idx] = y;
                }
            };
            // Bin k = lane + 32 j sits in register j; its conjugate partner h - k in lane (32 - lane) & 31, register
            // 31 - j (lane 0: its own register (32 - j) & 31).  Every lane finishes the pairs of its 16 lowest bins:
            // together that is every bin but h/2, which lane 0 adds.
            const int pl = (32 - lane) & 31;
            const bool l0 = lane == 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const cx sv = l0 ? z[(32 - j) & 31] : z[31 - j];
                cx b;
                b.x = __shfl_sync(0xffffffffu, sv.x, pl); b.y = __shfl_sync(0xffffffffu, sv.y, pl);
                const cx a = z[j];
                const int k = lane + 32 * j;
                const float2 cs = spl_s[k]; // (cos, sin)(k pi / h)
                const cx cb = make_float2(b.x, -b.y);
                const cx S = cadd(a, cb);   // (ar + br, ai - bi)
                const cx D = csub(a, cb);   // (ar - br, ai + bi)
                // (p1, p2) = (c (ai + bi) - s (ar - br), s (ai + bi) + c (ar - br))
                const cx P = __ffma2_rn(make_float2(-D.x, D.y), make_float2(cs.y, cs.y), __fmul2_rn(make_float2(D.y, D.x), make_float2(cs.x, cs.x)));
                const cx Xk = __fmul2_rn(cadd(S, make_float2(P.x, -P.y)), make_float2(0.5f, 0.5f));  // (sumr + p1, difi - p2) / 2
                const cx Xh = __fmul2_rn(cadd(S, make_float2(-P.x, P.y)), make_float2(0.5f, 0.5f));  // (sumr - p1, difi + p2) / 2, im to be negated
                emit(k, Xk, false);
                emit(k == 0 ? H : H - k, Xh, true); // k == 0: the Nyquist bin
            }
            if (l0) emit(H / 2, z[16], true); // X[h/2] = conj(Z[h/2])

            // ---- F: mel projection + dB, segment form of the bank (host_tables.h MelBands::seg) ----------------------
            // A lane walks the bins of ONE segment: every magnitude is read once and feeds U (rising side, filter s) and D
            // (falling side, filter s - 1) with one FFMA2; filter m = U_m + D_(m+1), the neighbour's D arriving by a shuffle.
            if (MEL) {
                __syncwarp(); // magnitudes of all bins are in the plane
                const int lg = td->seg_log2p, P = 1 << lg;
                const int nwq = td->seg_nwq, nblk = td->seg_nblk;
                const float2 *wq = reinterpret_cast<const float2 *>(bank);
                const int *lo_s = reinterpret_cast<const int *>(bank) + 2 * nwq;
                const int2 *desc_s = reinterpret_cast<const int2 *>(lo_s + 32 * nblk);
                const int plm = lane & (P - 1);
                float *orow = out + (size_t)(t0 + fl0) * n_out;
                for (int blk = 0; blk < nblk; ++blk) {
                    const int2 bd = desc_s[blk];
                    const int li = lo_s[blk * 32 + lane];   // first bin | filter << 16
                    const int m = (int)((unsigned)li >> 16);
                    const float2 *wp = wq + bd.x + lane;
                    const float *mp = magbuf + (li & 0xffff);
                    float2 ud = make_float2(0.0f, 0.0f); // (U, D)
                    for (int j2 = 0; j2 < bd.y; j2 += 2) {
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const float2 wgt = wp[(j2 + u) * 32];
                            const float mg = *mp;
                            mp += P;
                            ud = __ffma2_rn(make_float2(mg, mg), wgt, ud);
                        }
                    }
                    if (P > 1) {
                        for (int sh = P >> 1; sh > 0; sh >>= 1) {
                            ud.x += __shfl_xor_sync(0xffffffffu, ud.x, sh); ud.y += __shfl_xor_sync(0xffffffffu, ud.y, sh);
                        }
                    }
                    const float a0 = ud.x + __shfl_down_sync(0xffffffffu, ud.y, P); // D of the next segment
                    if (m != 0xffff && plm == 0) {
                        const float y = amp_to_db_dev(a0); // decibel.rs:33-88
                        vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
                        orow[m] = y;
                    }
                }
            }
            __syncwarp(); // magnitudes consumed before the next round's transpose overwrites the plane
        }
    } // tiles of this CTA
    flush_range();
}

} // namespace

size_t stft_warp1_fixed_smem(int bank_floats, int warps)
{
    return 16 + (size_t)warps * kW1Plane * sizeof(float2) + (size_t)(kW1H + kW1H + kW1H / 2) * sizeof(float2) +
           (size_t)bank_floats * sizeof(float);
}

int stft_warp1_warps()
{
    static const int w = [] {
        const char *e = getenv("SGX_W1_WARPS");
        const int v = e ? atoi(e) : 16;
        return (v == 8 || v == 12) ? v : 16;
    }();
    return w;
}

namespace {
template <bool MEL, int NW> cudaError_t launch_w1(const StftLaunch &L, size_t smem, int grid, cudaStream_t stream)
{
    auto kern = stft_warp1_kernel<MEL, NW>;
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, stream>>>(L);
    count_launch();
    return cudaGetLastError();
}
} // namespace

cudaError_t launch_stft_warp1(const StftLaunch &L, cudaStream_t stream)
{
    const int nw = L.warp1;
    const size_t smem = stft_warp1_fixed_smem(L.bank_floats, nw) + (size_t)L.tile_floats * sizeof(float);
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    const int grid = std::min(L.n_tiles, sms);
    const bool mel = L.mode == MODE_MEL_DB;
    switch (nw) {
    case 8: return mel ? launch_w1<true, 8>(L, smem, grid, stream) : launch_w1<false, 8>(L, smem, grid, stream);
    case 12: return mel ? launch_w1<true, 12>(L, smem, grid, stream) : launch_w1<false, 12>(L, smem, grid, stream);
    case 16: return mel ? launch_w1<true, 16>(L, smem, grid, stream) : launch_w1<false, 16>(L, smem, grid, stream);
    }
    return cudaErrorInvalidValue;
}

} // namespace sgx
