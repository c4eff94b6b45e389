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
#include <vector>

#include "device_common.cuh"
#include "kernels.h"
#include "stft_device.cuh"

// The instantiations of K1 are spread over several translation units so that they compile in parallel (the Makefile
// builds this file once per part with -DSGX_K1_PART=k; part 0 holds the host side and the generic kernel, part k > 0
// the table rows tagged k).  Without the macro the file is one self-contained translation unit.
#ifndef SGX_K1_PART
#define SGX_K1_PART -1
#endif

namespace sgx {

namespace {

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

// where the twiddles of the pass with input stride product NS start in the per-pass table (make_fft_pass_tables)
__host__ __device__ constexpr int pass_table_offset(int h, int pts, int ns_of_pass)
{
    int off = 0;
    for (int ns = 1; ns < ns_of_pass;) {
        const int r = (h / ns >= pts) ? pts : h / ns;
        if (ns > 1) off += (r - 1) * ns;
        ns *= r;
    }
    return off;
}

template <int H, int PTS, int V, int G, int R, int NS>
__device__ __forceinline__ void fft_pass(pk (&re)[PTS][V / 2], pk (&im)[PTS][V / 2], float *sre,
                                         float *sim, int gt, int grp, const float2 *__restrict__ twr)
{
    constexpr int VP = V / 2;
    constexpr int NT = H / PTS, NB = PTS / R;
    if constexpr (NS > 1) {
        // Twiddles first: the table loads fly while the shared loads and the barrier complete.  Every w[r] =
        // exp(-2 pi i r k / (NS R)) comes from a table, correctly rounded from f64 -- like the precomputed twiddles of
        // rustfft's Radix4 (realfft.rs:94).  Forming the powers by complex products from one table entry left the
        // twiddles 3-5 ulp off, which put the noise floor of the spectrum (the bins 60+ dB below a frame's peak) 5-6x
        // above the reference's: see tests/parity.py, the f64-truth gate.  The table of a pass is r-major,
        // twr[(r - 1) NS + k], so that the lanes of a warp (consecutive k) read consecutive entries: one or two
        // 128-byte lines per load (the natural table tw[r k H / (NS R)] is strided: 8 to 28 lines per load, and cost
        // 11 % of the kernel).
        constexpr int OFF = pass_table_offset(H, PTS, NS);
        float2 wt[NB][R];
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int k = (gt + q * NT) & (NS - 1);
#pragma unroll
            for (int r = 1; r < R; ++r) wt[q][r] = __ldg(twr + OFF + (r - 1) * NS + k);
        }
        // element gt + q*NT + r*H/R: the offsets are multiples of 8, so pad() is linear in them
        static_assert(NT % 8 == 0 && (H / R) % 8 == 0, "padded addressing assumes multiples of 8");
        const int sbase = padi(gt) * V;
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int src = sbase + ((q * NT + r * (H / R)) / 8 * 9) * V;
                ld_vec<VP>(sre + src, re[q * R + r]);
                ld_vec<VP>(sim + src, im[q * R + r]);
            }
        group_sync<G, NT>(grp); // every thread has its inputs: the buffer may be overwritten
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const float2 (&w)[R] = wt[q];
#pragma unroll
            for (int r = 1; r < R; ++r) {
#pragma unroll
                for (int v = 0; v < VP; ++v) {
                    const pk xr = re[q * R + r][v], xi = im[q * R + r][v];
                    re[q * R + r][v] = pk_fmas(xi, -w[r].y, pk_muls(xr, w[r].x));
                    im[q * R + r][v] = pk_fmas(xi, w[r].x, pk_muls(xr, w[r].y));
                }
            }
        }
    }
    if constexpr (NB == 1) dft_inplace<R, 0, 1, PTS, VP>(re, im);
    else if constexpr (NB == 2) { dft_inplace<R, 0, 1, PTS, VP>(re, im); dft_inplace<R, R, 1, PTS, VP>(re, im); }
    else if constexpr (NB == 4) {
        dft_inplace<R, 0, 1, PTS, VP>(re, im); dft_inplace<R, R, 1, PTS, VP>(re, im);
        dft_inplace<R, 2 * R, 1, PTS, VP>(re, im); dft_inplace<R, 3 * R, 1, PTS, VP>(re, im);
    } else {
        static_assert(NB == 8, "unsupported butterflies per thread");
        dft_inplace<R, 0, 1, PTS, VP>(re, im); dft_inplace<R, R, 1, PTS, VP>(re, im);
        dft_inplace<R, 2 * R, 1, PTS, VP>(re, im); dft_inplace<R, 3 * R, 1, PTS, VP>(re, im);
        dft_inplace<R, 4 * R, 1, PTS, VP>(re, im); dft_inplace<R, 5 * R, 1, PTS, VP>(re, im);
        dft_inplace<R, 6 * R, 1, PTS, VP>(re, im); dft_inplace<R, 7 * R, 1, PTS, VP>(re, im);
    }
    if constexpr (NS == 1 && R == 8) {
        // d = 8 j, element 8 j + r -> padded 9 j + r
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int dst = (9 * gt + 9 * q * NT + r) * V;
                st_vec<VP>(sre + dst, re[q * R + r]);
                st_vec<VP>(sim + dst, im[q * R + r]);
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
                st_vec<VP>(sre + dst, re[q * R + r]);
                st_vec<VP>(sim + dst, im[q * R + r]);
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
                st_vec<VP>(sre + dst, re[q * R + r]);
                st_vec<VP>(sim + dst, im[q * R + r]);
            }
        }
    }
    group_sync<G, NT>(grp);
}

// Runs the passes whose input stride product NS is below NS_END (H: all passes).
template <int H, int PTS, int V, int G, int NS, int NS_END>
__device__ __forceinline__ void run_passes(pk (&re)[PTS][V / 2], pk (&im)[PTS][V / 2], float *sre,
                                           float *sim, int gt, int grp, const float2 *__restrict__ twr)
{
    if constexpr (NS < NS_END) {
        constexpr int R = (H / NS >= PTS) ? PTS : (H / NS);
        fft_pass<H, PTS, V, G, R, NS>(re, im, sre, sim, gt, grp, twr);
        run_passes<H, PTS, V, G, NS * R, NS_END>(re, im, sre, sim, gt, grp, twr);
    }
}

// radix of the last pass of the schedule above
__host__ __device__ constexpr int last_radix(int h, int pts)
{
    int n = h;
    while (n >= pts) n /= pts;
    return n > 1 ? n : pts;
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
// LOADER: which raw tiles the first pass can read besides mono f32 -- 1: f32 stereo tiles, staged as raw interleaved
// pairs by TMA and summed by the first pass; 2: int16 mono tiles, staged as 16-bit samples and converted by it.
// Instantiations of their own: an extra first-pass variant costs 2-8 % when it merely sits in a kernel's code.
template <int LOG2H, int PTS, int V, int G, int MC, bool MEL, int LOADER>
__global__ void __launch_bounds__(K1Traits<LOG2H, PTS, V, G, MC>::THREADS, MC)
stft_db_kernel(const StftLaunch L)
{
    using TR = K1Traits<LOG2H, PTS, V, G, MC>;
    constexpr int H = TR::H, NT = TR::NT, THREADS = TR::THREADS, PADH = TR::PADH, F = 2 * H;
    constexpr bool RAW2 = LOADER == 1, RAW16 = LOADER == 2; // which raw tiles the first pass of this instantiation reads
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
    locate_tile(L, F, blockIdx.x, s_trk, td, cur, RAW2, RAW16);
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
        locate_tile(L, F, tile_id, s_trk, td, cur, RAW2, RAW16);
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
    // melp: the segment form of the bank is used (fused kernels, magnitudes unpadded); else the banded copy, in the
    // region when it fits, staged per round in the imaginary plane or read through L1 otherwise
    constexpr int RL_K = last_radix(H, PTS);
    constexpr bool FUSED_K = (PTS / RL_K) >= 2;
    const bool melp = MEL && FUSED_K && td->segp != nullptr && td->seg_words <= L.bank_floats;
    const bool bank_fits = MEL && (melp || ((__ldg(td->mel_cnt + 1) + 3) & ~3) + 4 * n_out <= L.bank_floats);
    if (MEL && bank_fits && td->mel_w != bank_src) {
        __syncthreads(); // other groups may still be projecting frames of the previous track
        if (melp) {
            const int words = td->seg_words;
            const int *__restrict__ srcw = td->segp;
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
        constexpr int VP = V / 2;
        pk re[PTS][VP], im[PTS][VP]; // pair i = frames 2i, 2i+1 of the group

        // ---- first-pass inputs: z[m] = g[2m] + i g[2m+1], g = sample * window ---------------------
        if (L.staged && vec_ok && !(RAW2 && cur.raw2) && !(RAW16 && cur.i16)) {
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
                    pk_set<VP>(re[p], v, x.x * w.x); pk_set<VP>(im[p], v, x.y * w.y);
                }
            }
        } else if (RAW16 && L.staged && vec_ok && cur.i16) {
            // raw int16 tile: one 32-bit shared load brings the two samples of a point; audio.rs:16-19 scales them
            const short *tb = reinterpret_cast<const short *>(tile);
            const short *fb[V];
#pragma unroll
            for (int v = 0; v < V; ++v) fb[v] = tb + off0 + min(fl0 + v, nfr - 1) * hop + 2 * gt;
#pragma unroll
            for (int p = 0; p < PTS; ++p) {
                const int n = 2 * (gt + p * NT);
                const float2 w = __ldg(reinterpret_cast<const float2 *>(win_f + n));
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const unsigned u = *reinterpret_cast<const unsigned *>(fb[v] + 2 * p * NT);
                    const float x0 = (float)(short)(u & 0xffffu) * (1.0f / 32768.0f), x1 = (float)(short)(u >> 16) * (1.0f / 32768.0f);
                    pk_set<VP>(re[p], v, x0 * w.x); pk_set<VP>(im[p], v, x1 * w.y);
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
                    pk_set<VP>(re[p], v, (x.x + x.y) * w.x); pk_set<VP>(im[p], v, (x.z + x.w) * w.y);
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
                    if (RAW16 && L.staged && cur.i16) { // odd hop: frames of an int16 tile that start on an odd sample
                        const short *tb = reinterpret_cast<const short *>(tile);
                        const int b = off0 + fl * hop + n;
                        x0 = (float)tb[b] * (1.0f / 32768.0f); x1 = (float)tb[b + 1] * (1.0f / 32768.0f);
                    } else if (L.staged) {
                        const int b = off0 + fl * hop + n;
                        x0 = tile[b]; x1 = tile[b + 1];
                    } else {
                        const long long i = S0 + (long long)fl * hop + n;
                        x0 = load_sample(pv, i); x1 = load_sample(pv, i + 1);
                    }
                    pk_set<VP>(re[p], v, x0 * w.x); pk_set<VP>(im[p], v, x1 * w.y);
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
                        locate_tile(L, F, nt, ntrk, ntd, nx, RAW2, RAW16);
                        if (nx.tma) issue_tile_copy(ntd, nx, tile, mbar);
                    }
                }
            }
        }
        // ---- real-FFT split (realfft.rs:140-157) -------------------------------------------------------
        // emit(): what becomes of one output bin -- complex / magnitude / dB to HBM, or (mel) its
        // magnitude into the shared buffer at the bin's own (padded) position.
        auto emit = [&](int idx, int spos, const pk (&xr)[VP], const pk (&xi)[VP]) {
            if (mode == MODE_COMPLEX) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (fl0 + v < nfr)
                        reinterpret_cast<float2 *>(out)[(size_t)(t0 + fl0 + v) * (H + 1) + idx] =
                            make_float2(pk_get<VP>(xr, v), pk_get<VP>(xi, v));
                return;
            }
            pk mg[VP];
#pragma unroll
            for (int v = 0; v < VP; ++v) { // lib.rs:124
                const pk q = pk_fma(xr[v], xr[v], pk_mul(xi[v], xi[v]));
                mg[v] = make_float2(sqrt_approx(q.x), sqrt_approx(q.y));
            }
            if (MEL) {
                st_vec<VP>(sre + (melp ? idx * V : spos), mg); // block-padded bank: magnitudes at their bin index
            } else {
                // frames beyond the tile's last one are copies of it (first-pass clamp): harmless in the extrema
                float *op = out + (size_t)(t0 + fl0) * (H + 1) + idx;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    float y = pk_get<VP>(mg, v);
                    if (mode == MODE_LIN_DB) {
                        y = amp_to_db_dev(y);
                        vmax = fmaxf(vmax, y); vmin = fminf(vmin, y);
                    }
                    if (fl0 + v < nfr) op[(size_t)v * (H + 1)] = y;
                }
            }
        };

        // one conjugate-symmetric pair: a = Z[k], b = Z[h-k], k <= h/2 -> X[k] and (optionally) X[h-k]
        auto split_pair = [&](int k, float2 cs, const pk (&ar)[VP], const pk (&ai)[VP], const pk (&br)[VP],
                              const pk (&bi)[VP], bool emit_partner) {
            pk xr[VP], xi[VP], yr[VP], yi[VP];
#pragma unroll
            for (int v = 0; v < VP; ++v) {
                const pk sumr = pk_add(ar[v], br[v]), difr = pk_sub(ar[v], br[v]);
                const pk sumi = pk_add(ai[v], bi[v]), difi = pk_sub(ai[v], bi[v]);
                const pk p1 = pk_fmas(sumi, cs.x, pk_muls(difr, -cs.y));  // c*sumi - s*difr
                const pk p2 = pk_fmas(sumi, cs.y, pk_muls(difr, cs.x));   // s*sumi + c*difr
                xr[v] = pk_muls(pk_add(sumr, p1), 0.5f); xi[v] = pk_muls(pk_sub(difi, p2), 0.5f);
                yr[v] = pk_muls(pk_sub(sumr, p1), 0.5f); yi[v] = pk_muls(pk_add(difi, p2), -0.5f);
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
            run_passes<H, PTS, V, G, 1, H / RL>(re, im, sre, sim, gt, grp, L.twr);
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
                    ld_vec<VP>(sre + sa + (r * NSL / 8 * 9) * V, re[(2 * p) * RL + r]);
                    ld_vec<VP>(sim + sa + (r * NSL / 8 * 9) * V, im[(2 * p) * RL + r]);
                    ld_vec<VP>(sre + sb + (r * NSL / 8 * 9) * V, re[(2 * p + 1) * RL + r]);
                    ld_vec<VP>(sim + sb + (r * NSL / 8 * 9) * V, im[(2 * p + 1) * RL + r]);
                }
            }
            if (MEL) group_sync<G, NT>(grp); // the buffer now becomes the magnitude array
#pragma unroll
            for (int b = 0; b < 2 * NPR; ++b) {
                float2 w[RL];
                w[1] = (b & 1) ? wB[b >> 1] : wA[b >> 1];
                if constexpr (RL >= 4) { // table twiddles here too (j < H / RL, so r j < H)
                    const int jb = (b & 1) ? bB[b >> 1] : bA[b >> 1];
                    w[2] = __ldg(L.tw + 2 * jb); w[3] = __ldg(L.tw + 3 * jb);
                }
                static_assert(RL <= 4, "fused last pass is written for radix 2 and 4");
#pragma unroll
                for (int r = 1; r < RL; ++r)
#pragma unroll
                    for (int v = 0; v < VP; ++v) {
                        const pk xr = re[b * RL + r][v], xi = im[b * RL + r][v];
                        re[b * RL + r][v] = pk_fmas(xi, -w[r].y, pk_muls(xr, w[r].x));
                        im[b * RL + r][v] = pk_fmas(xi, w[r].x, pk_muls(xr, w[r].y));
                    }
            }
            if constexpr (NPR >= 1) { dft_inplace<RL, 0, 1, PTS, VP>(re, im); dft_inplace<RL, RL, 1, PTS, VP>(re, im); }
            if constexpr (NPR >= 2) { dft_inplace<RL, 2 * RL, 1, PTS, VP>(re, im); dft_inplace<RL, 3 * RL, 1, PTS, VP>(re, im); }
            if constexpr (NPR >= 3) { dft_inplace<RL, 4 * RL, 1, PTS, VP>(re, im); dft_inplace<RL, 5 * RL, 1, PTS, VP>(re, im); }
            if constexpr (NPR >= 4) { dft_inplace<RL, 6 * RL, 1, PTS, VP>(re, im); dft_inplace<RL, 7 * RL, 1, PTS, VP>(re, im); }
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
                    pk pr[VP], pi[VP];
#pragma unroll
                    for (int v = 0; v < VP; ++v) {
                        pr[v] = re[bb + RL - 1 - r][v]; pi[v] = im[bb + RL - 1 - r][v];
                        if (p == 0 && self) { pr[v] = re[ba + (r == 0 ? 0 : RL - r)][v]; pi[v] = im[ba + (r == 0 ? 0 : RL - r)][v]; }
                    }
                    split_pair(bA[p] + r * NSL, csA[p][r], re[ba + r], im[ba + r], pr, pi, true);
#pragma unroll
                    for (int v = 0; v < VP; ++v) {
                        pr[v] = re[ba + RL - 1 - r][v]; pi[v] = im[ba + RL - 1 - r][v];
                        if (p == 0 && self) { pr[v] = re[bb + RL - 1 - r][v]; pi[v] = im[bb + RL - 1 - r][v]; }
                    }
                    split_pair(bB[p] + r * NSL, csB[p][r], re[bb + r], im[bb + r], pr, pi, true);
                }
                if (p == 0 && self) { // X[H/2] = conj(Z[H/2])
                    pk ni[VP];
#pragma unroll
                    for (int v = 0; v < VP; ++v) ni[v] = pk_neg(im[ba + RL / 2][v]);
                    emit(H / 2, padi(H / 2) * V, re[ba + RL / 2], ni);
                }
            }
        } else {
        run_passes<H, PTS, V, G, 1, H>(re, im, sre, sim, gt, grp, L.twr);
        constexpr int NP = PTS / 2;
        const int pa0 = padi(gt) * V;      // element k = gt + q NT      -> pa0 + 9 q NT / 8
        const int pb0 = padi(H - gt) * V;  // element H - k (k > 0)      -> pb0 - 9 q NT / 8
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int k = gt + q * NT;
            const int pa = pa0 + (q * NT / 8 * 9) * V;
            const int pbm = pb0 - (q * NT / 8 * 9) * V;     // where the partner's magnitude goes (index H when k == 0)
            const int pb = (q == 0 && gt == 0) ? 0 : pbm;   // where the partner's spectrum is read (index 0 when k == 0)
            pk ar[VP], ai[VP], br[VP], bi[VP];
            ld_vec<VP>(sre + pa, ar); ld_vec<VP>(sim + pa, ai);
            ld_vec<VP>(sre + pb, br); ld_vec<VP>(sim + pb, bi);
            const float2 cs = __ldg(L.split + k); // (cos, sin)(k pi / h)
            pk xr[VP], xi[VP], yr[VP], yi[VP];
#pragma unroll
            for (int v = 0; v < VP; ++v) {
                const pk sumr = pk_add(ar[v], br[v]), difr = pk_sub(ar[v], br[v]);
                const pk sumi = pk_add(ai[v], bi[v]), difi = pk_sub(ai[v], bi[v]);
                const pk p1 = pk_fmas(sumi, cs.x, pk_muls(difr, -cs.y));  // c*sumi - s*difr
                const pk p2 = pk_fmas(sumi, cs.y, pk_muls(difr, cs.x));   // s*sumi + c*difr
                xr[v] = pk_muls(pk_add(sumr, p1), 0.5f); xi[v] = pk_muls(pk_sub(difi, p2), 0.5f);
                yr[v] = pk_muls(pk_sub(sumr, p1), 0.5f); yi[v] = pk_muls(pk_add(difi, p2), -0.5f);
            }
            emit(k, pa, xr, xi);
            emit(k == 0 ? H : H - k, pbm, yr, yi); // k == 0: the partner output is the Nyquist bin
        }
        if (gt == 0) {
            pk cr[VP], ci[VP];
            ld_vec<VP>(sre + padi(H / 2) * V, cr); ld_vec<VP>(sim + padi(H / 2) * V, ci);
#pragma unroll
            for (int v = 0; v < VP; ++v) ci[v] = pk_neg(ci[v]);
            emit(H / 2, padi(H / 2) * V, cr, ci);
        }
        } // !FUSED

        // ---- banded mel projection + dB -----------------------------------------------------------
        // Work item = (filter m, lane pl of the 2^lg lanes sharing it); 32 consecutive items form a block
        // and the host hands every warp of the group a balanced list of blocks (longest-first packing).
        // The filterbank taps and descriptors are first staged in the imaginary plane of the exchange
        // buffer -- it is dead once the spectrum has been read -- so the tap loop only touches shared memory.
        if (MEL && melp) {
            // Segment form of the bank (host_tables.h MelBands::seg): a lane walks the bins of ONE segment, reads each
            // magnitude vector once and feeds two accumulators -- U (rising side, filter s) and D (falling side, filter
            // s - 1); filter m = U_m + D_(m+1), the neighbour's D arriving by one shuffle.  Taps are block-padded and
            // tap-major: no predicates, one 64-bit and one vector shared load per tap.
            constexpr int NWARPS = NT / 32;
            const int lg = td->seg_log2p, P = 1 << lg;
            const int nwq = td->seg_nwq, nblk = td->seg_nblk;
            const float2 *wq = reinterpret_cast<const float2 *>(bank);
            const int *lo_s = reinterpret_cast<const int *>(bank) + 2 * nwq;
            const int2 *desc_s = reinterpret_cast<const int2 *>(lo_s + 32 * nblk);
            const int *sched = lo_s + 34 * nblk; // {slots, block ids [slots][NWARPS]}
            group_sync<G, NT>(grp); // magnitudes of all bins are in the buffer
            const int wg = gt >> 5, lane = gt & 31;
            const int stride = V << lg;
            const int nslots = sched[0];
            for (int slot = 0; slot < nslots; ++slot) {
                const int blk = sched[1 + slot * NWARPS + wg]; // warp-uniform
                if (blk < 0) continue;
                const int2 bd = desc_s[blk];
                const int li = lo_s[blk * 32 + lane];   // first bin | filter << 16
                const int m = (int)((unsigned)li >> 16), pl = lane & (P - 1);
                const float2 *wp = wq + bd.x + lane;
                const float *mp = sre + (li & 0xffff) * V;
                pk up[VP], dn[VP];
#pragma unroll
                for (int v = 0; v < VP; ++v) { up[v] = make_float2(0.0f, 0.0f); dn[v] = make_float2(0.0f, 0.0f); }
                for (int j2 = 0; j2 < bd.y; j2 += 2) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float2 wgt = wp[(j2 + u) * 32];
                        pk mg[VP];
                        ld_vec<VP>(mp, mg);
                        mp += stride;
#pragma unroll
                        for (int v = 0; v < VP; ++v) { up[v] = pk_fmas(mg[v], wgt.x, up[v]); dn[v] = pk_fmas(mg[v], wgt.y, dn[v]); }
                    }
                }
                float acc[V];
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    float a = pk_get<VP>(up, v), d = pk_get<VP>(dn, v);
                    for (int sh = P >> 1; sh > 0; sh >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, sh);
                        d += __shfl_xor_sync(0xffffffffu, d, sh);
                    }
                    acc[v] = a + __shfl_down_sync(0xffffffffu, d, P); // D of the next segment
                }
                if (m != 0xffff && pl == 0) {
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
                pk accp[VP];
#pragma unroll
                for (int v = 0; v < VP; ++v) accp[v] = make_float2(0.0f, 0.0f);
#pragma unroll 8
                for (int j = 0; j < njmax; ++j) {
                    const bool on = j < nj;
                    float wgt = 0.0f;
                    if (on) wgt = ST ? wp[j << lg] : __ldg(wp + (j << lg));
                    pk mg[VP];
                    ld_vec<VP>(sre + padi(on ? bin0 + (j << lg) : 0) * V, mg);
#pragma unroll
                    for (int v = 0; v < VP; ++v) accp[v] = pk_fmas(mg[v], wgt, accp[v]);
                }
                float acc[V];
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] = pk_get<VP>(accp, v);
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

#if SGX_K1_PART <= 0
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

#endif // SGX_K1_PART <= 0

int resident_sms()
{
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

template <int LOG2H, int PTS, int V, int G, int MC, bool MEL, int LOADER>
cudaError_t launch_one_mode(const StftLaunch &L, size_t smem, cudaStream_t stream)
{
    using TR = K1Traits<LOG2H, PTS, V, G, MC>;
    auto kern = stft_db_kernel<LOG2H, PTS, V, G, MC, MEL, LOADER>;
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
    // launches that hold f32 stereo (1) or int16 mono (2) tracks use an instantiation whose first pass also reads
    // that kind of raw staged tile; mono f32 launches keep the lean one (0)
    const int ld = L.stereo_raw;
    if (L.mode == MODE_MEL_DB)
        return ld == 1 ? launch_one_mode<LOG2H, PTS, V, G, MC, true, 1>(L, smem, stream)
             : ld == 2 ? launch_one_mode<LOG2H, PTS, V, G, MC, true, 2>(L, smem, stream)
                       : launch_one_mode<LOG2H, PTS, V, G, MC, true, 0>(L, smem, stream);
    return ld == 1 ? launch_one_mode<LOG2H, PTS, V, G, MC, false, 1>(L, smem, stream)
         : ld == 2 ? launch_one_mode<LOG2H, PTS, V, G, MC, false, 2>(L, smem, stream)
                   : launch_one_mode<LOG2H, PTS, V, G, MC, false, 0>(L, smem, stream);
}

} // namespace

// FFT size -> kernel instantiations (LOG2H, PTS, V, G, resident CTAs).  The first entry of a size is the
// default; the others are selectable with SGX_K1_VARIANT="pts,v,g" (tuning experiments).  Measured on B200 for
// h = 1024 (8 tracks x 10 min, K1 ms): (8,4,2) 2.35 | (8,2,2) 3.09 | (8,2,4) 3.39 | (4,4,1) 2.80 | (4,4,2) 2.85 |
// (16,2,4) 3.16 | (16,2,2) 3.38 -- sharing index math, twiddles and mel taps across V = 4 frames outweighs the
// higher occupancy of the V = 2 variants and the fewer exchanges of radix 16.
#ifdef SGX_K1_ALTERNATES // the other CTA shapes of the design-space measurements (make TUNE=-DSGX_K1_ALTERNATES)
#define SGX_K1_ALT(X)      \
    X(5, 9, 8, 4, 8, 1)    \
    X(5, 10, 8, 4, 2, 2)   \
    X(5, 10, 8, 2, 2, 3)   \
    X(5, 10, 4, 4, 1, 4)   \
    X(6, 10, 16, 2, 8, 1)  \
    X(6, 11, 8, 4, 1, 2)
#else
#define SGX_K1_ALT(X)
#endif
// (part, LOG2H, PTS, V, G, resident CTAs)
#define SGX_K1_TABLE(X)    \
    X(1, 8, 8, 4, 8, 2)    \
    X(1, 9, 8, 4, 4, 2)    \
    X(2, 10, 8, 4, 4, 1)   \
    X(3, 11, 8, 4, 2, 1)   \
    X(4, 12, 8, 4, 1, 1)   \
    X(4, 13, 8, 2, 1, 1)   \
    SGX_K1_ALT(X)
#define SGX_K1_NPARTS 6

namespace {
template <int PART>
cudaError_t launch_rows(const StftConfig &cfg, const StftLaunch &L, size_t smem, cudaStream_t stream)
{
#define X(P, LG, PTS, V, G, MC)                                                                   \
    if constexpr (PART < 0 || PART == P) {                                                        \
        if (cfg.h == (1 << LG) && cfg.pts == PTS && cfg.vec == V && cfg.groups == G)              \
            return launch_one<LG, PTS, V, G, MC>(L, smem, stream);                                \
    }
    SGX_K1_TABLE(X)
#undef X
    return cudaErrorInvalidValue; // not a row of this part
}
} // namespace

#define SGX_K1_CAT2(a, b) a##b
#define SGX_K1_CAT(a, b) SGX_K1_CAT2(a, b)
#if SGX_K1_PART > 0
cudaError_t SGX_K1_CAT(launch_stft_rows_, SGX_K1_PART)(const StftConfig &cfg, const StftLaunch &L, size_t smem, cudaStream_t stream)
{
    return launch_rows<SGX_K1_PART>(cfg, L, smem, stream);
}
#endif

#if SGX_K1_PART <= 0
#if SGX_K1_PART == 0
cudaError_t launch_stft_rows_1(const StftConfig &, const StftLaunch &, size_t, cudaStream_t);
cudaError_t launch_stft_rows_2(const StftConfig &, const StftLaunch &, size_t, cudaStream_t);
cudaError_t launch_stft_rows_3(const StftConfig &, const StftLaunch &, size_t, cudaStream_t);
cudaError_t launch_stft_rows_4(const StftConfig &, const StftLaunch &, size_t, cudaStream_t);
#ifdef SGX_K1_ALTERNATES
cudaError_t launch_stft_rows_5(const StftConfig &, const StftLaunch &, size_t, cudaStream_t);
cudaError_t launch_stft_rows_6(const StftConfig &, const StftLaunch &, size_t, cudaStream_t);
#endif
#endif

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
#define X(P, LG, PTS, V, G, MC)                                                                  \
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
    // n_fft = 2048: SGX_K1W2=1 selects the warp-per-frame-pair kernel (stft_warp2_kernel.cu; measured 6.42 vs 6.32 ms
    // per C5 step, see its header) -- an evaluated alternative, off by default
    static const bool w2_on = getenv("SGX_K1W2") && atoi(getenv("SGX_K1W2")) == 1;
    cfg->warp2 = h == 1024 && !cfg->generic && cfg->fused && w2_on && want_pts == 0;
    // SGX_K1W1=1: one warp per frame on packed complex values (stft_warp1_kernel.cu)
    constexpr bool kW1Default = false;
    static const bool block_only = getenv("SGX_K1BLOCK") && atoi(getenv("SGX_K1BLOCK")) == 1;
    static const bool w1_on = !block_only && (getenv("SGX_K1W1") ? atoi(getenv("SGX_K1W1")) == 1 : kW1Default);
    cfg->warp1 = h == 1024 && !cfg->generic && cfg->fused && w1_on && want_pts == 0;
    if (cfg->warp1 || block_only) cfg->warp2 = false;
    return true;
}

size_t stft_max_dynamic_smem() { return 227 * 1024; }

StftTiling plan_stft_tiles(const StftConfig &cfg, int max_hop, int bank_floats, int sample_floats, bool warp2_ok)
{
    StftTiling t{};
    if (cfg.generic) {
        t.frames_per_tile = 1; t.staged = 0; t.tile_floats = 0;
        t.smem_bytes = cfg.fft_smem;
        return t;
    }
    int want_nfr = 0;
    if (const char *e = getenv("SGX_K1_NFR")) want_nfr = atoi(e);
    if (cfg.warp1 && warp2_ok) {
        // one frame per warp and round; the tile gets what the exchange planes, tables and filterbank leave
        const int nw = stft_warp1_warps();
        const size_t fixed = stft_warp1_fixed_smem(bank_floats, nw);
        const long cap_floats = stft_max_dynamic_smem() > fixed ? (long)((stft_max_dynamic_smem() - fixed) / sizeof(float)) : 0;
        int best = 0;
        for (int mult = 1; mult <= 4; ++mult) {
            const int nfr = nw * mult;
            const long need = 3 + (long)(nfr - 1) * max_hop + cfg.n_fft + 4;
            if (need <= cap_floats && (nfr <= (want_nfr ? want_nfr : 2 * nw) || best == 0)) best = nfr;
        }
        if (best > 0) {
            t.frames_per_tile = best; t.staged = 1; t.warp1 = nw; t.bank_floats = bank_floats; t.sample_floats = 1;
            t.tile_floats = (int)((3 + (long)(best - 1) * max_hop + cfg.n_fft + 3) & ~3L) + 4;
            t.smem_bytes = fixed + (size_t)t.tile_floats * sizeof(float);
            return t;
        }
    }
    if (cfg.warp2 && warp2_ok) {
        // eight warps x two frames per round; the tile gets what the exchange planes, tables and filterbank leave
        const int nw = stft_warp2_warps();
        const size_t fixed = stft_warp2_fixed_smem(bank_floats, nw);
        const long cap_floats = stft_max_dynamic_smem() > fixed ? (long)((stft_max_dynamic_smem() - fixed) / sizeof(float)) : 0;
        int best = 0;
        for (int mult = 1; mult <= 4; ++mult) {
            const int nfr = 2 * nw * mult;
            const long need = 3 + (long)(nfr - 1) * max_hop + cfg.n_fft + 4;
            if (need <= cap_floats && (nfr <= (want_nfr ? want_nfr : 6 * nw) || best == 0)) best = nfr;
        }
        if (best > 0) {
            t.frames_per_tile = best; t.staged = 1; t.warp2 = nw; t.bank_floats = bank_floats; t.sample_floats = 1;
            t.tile_floats = (int)((3 + (long)(best - 1) * max_hop + cfg.n_fft + 3) & ~3L) + 4;
            t.smem_bytes = fixed + (size_t)t.tile_floats * sizeof(float);
            return t;
        }
    }
    const int unit = cfg.groups * cfg.vec;
    // per-CTA shared-memory budget for the register-limited number of resident CTAs (228 KB per SM,
    // 1 KB reserved per CTA); the staged tile gets what the FFT buffers leave
    const int ctas = cfg.min_ctas > 0 ? cfg.min_ctas : 1;
    size_t budget = (size_t)(228 * 1024) / ctas - 1024 - 512;
    if (budget > stft_max_dynamic_smem()) budget = stft_max_dynamic_smem();
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

// Twiddles of the Stockham passes of stft_db_kernel<.., PTS, ..>, pass after pass, r-major inside a pass:
// exp(-2 pi i r k / (NS R)) at pass_table_offset(h, pts, NS) + (r - 1) NS + k  (see fft_pass)
std::vector<float2> make_fft_pass_tables(int h, int pts)
{
    const double pi = 3.14159265358979323846264338327950288;
    std::vector<float2> t;
    for (int ns = 1; ns < h;) {
        const int r = (h / ns >= pts) ? pts : h / ns;
        if (ns > 1)
            for (int rr = 1; rr < r; ++rr)
                for (int k = 0; k < ns; ++k) {
                    const double a = -2.0 * pi * (double)rr * (double)k / ((double)ns * (double)r);
                    t.push_back(make_float2((float)cos(a), (float)sin(a)));
                }
        ns *= r;
    }
    if (t.empty()) t.push_back(make_float2(1.0f, 0.0f));
    return t;
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
    if (L.warp1) return launch_stft_warp1(L, stream);
    if (L.warp2) return launch_stft_warp2(L, stream);
    const size_t smem = cfg.generic ? cfg.fft_smem
                                    : 16 + (size_t)(L.tile_floats + L.bank_floats) * sizeof(float) + cfg.fft_smem;
    if (cfg.generic) {
        cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(stft_generic_kernel), smem);
        if (e != cudaSuccess) return e;
        stft_generic_kernel<<<L.n_tiles, 128, smem, stream>>>(L, cfg.h);
        count_launch();
        return cudaGetLastError();
    }
#if SGX_K1_PART < 0
    return launch_rows<-1>(cfg, L, smem, stream);
#else
    cudaError_t e = launch_stft_rows_1(cfg, L, smem, stream);
    if (e == cudaErrorInvalidValue) e = launch_stft_rows_2(cfg, L, smem, stream);
    if (e == cudaErrorInvalidValue) e = launch_stft_rows_3(cfg, L, smem, stream);
    if (e == cudaErrorInvalidValue) e = launch_stft_rows_4(cfg, L, smem, stream);
#ifdef SGX_K1_ALTERNATES
    if (e == cudaErrorInvalidValue) e = launch_stft_rows_5(cfg, L, smem, stream);
    if (e == cudaErrorInvalidValue) e = launch_stft_rows_6(cfg, L, smem, stream);
#endif
    return e;
#endif
}
#endif // SGX_K1_PART <= 0

} // namespace sgx
