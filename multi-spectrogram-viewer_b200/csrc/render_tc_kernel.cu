// render_tc_kernel.cu -- K3 on the 5th-generation tensor cores (tcgen05 + TMEM), for the viewer's default geometry.
//
// display.rs:44-61 + image 0.23 `resize` are two banded contractions with a clamp after each:
//     T[oy][f]    = clamp( sum_r  Av[oy][r] * grey[r][f] )          vertical_sample   (Lanczos3 rows)
//     out[oy][ox] = clamp( sum_f  T[oy][f]  * Ah[ox][f]  )          horizontal_sample (Lanczos3 columns)
// followed by the colour map.  The FP32 kernels (render_kernel.cu) spend their time moving taps through the
// shared-memory pipe: 17 FMAs per pixel, but ~2,700 shared-memory wavefronts and ~100 instructions per pixel-tile row
// (ncu: L1 / shared data pipe 92 % busy).  Here both contractions run as dense 128 x N x K MMAs -- the operands are
// read from shared memory by the tensor core itself, not by LDS instructions:
//   pass B   D1[128 oy][NF f]  = Av[128][KV] * grey[KV][NF]       A, B in shared memory (K-major, no swizzle), D1 in TMEM
//   epi 1    T = max(D1, 0), split into TF32 hi + lo, written back to TMEM as the A operand of pass C
//   pass C   D2[128 oy][NX ox] = T[128][NF]  * Ah[NF][NX]          A in TMEM, B in shared memory, D2 in TMEM
//   epi 2    max(D2, 0) -> colour map -> RGB(A) stores
// The band structure is given up (a 128-row tile multiplies a [128 x ~100] weight matrix that is 92 % zeros), which the
// tensor core absorbs: 69 MMAs per 128 x 64 pixel tile.  Precision: one TF32 pass (10-bit mantissa) moves 28 % of the
// bytes by one LSB (tools/tf32_proto.py); every product is therefore formed as hi*hi + hi*lo + lo*hi with both operands
// split into two TF32 terms ("3xTF32", ~21 mantissa bits), which holds the +-1 LSB / < 1 % tolerance like the FP32 path.
// Applicable when both axes are in the 8-tap class (n_in / n_out < 1.16, i.e. the viewer's default mel geometry at
// 100 px/s x 500) and the operands fit in shared memory.
//
// MEASURED (B200, C5: 32 tracks, 120,064 tiles of 128 x 64 pixels; pixels identical in tolerance to the FP32 path, all
// parity tests green with SGX_K3_TC=1):
//   v1  256 threads, one load in flight per warp while filling ........ 21.3 ms per step
//   v2  512 threads, loads batched ..................................... 9.5 ms   cycles per tile (clock64, CTA 0): grey tile 6,900 |
//       column weights 3,250 | pass B (39 MMAs) 3,060 | epilogue 1 1,330 | pass C (30 MMAs) 2,370 | epilogue 2 4,850
//   v3  this file: MMA-issuer warp, weights under pass B, next grey tile under pass C, staged epilogue ... 8.3 ms
//   FP32 path (render_fast_kernel, 4 CTAs of 8 warps per SM) ........... 3.7 ms
// The tensor core is not the problem (69 MMAs = ~5,400 of 20,000 cycles per tile, and they overlap); the work AROUND it
// is: per tile the CUDA cores still convert 8,320 dB values to TF32 pairs, build 5,120 column weights, split 10,240
// intermediates and colour 8,192 pixels -- ~11,600 warp instructions -- and with 210 KB of operands only ONE CTA fits
// per SM, so every one of those phases runs at the latency of a single 16-warp CTA between block barriers, while the FP32
// kernel hides the same latencies with 32 warps from four independent CTAs.  Dense tiles also cost 3 x 12 x the MACs of the
// band.  Projected with two-deep register prefetch of the grey tile and the weights: ~9,000 cycles per tile = parity with
// the FP32 path, not better.  Kept as a parity-tested alternative (SGX_K3_TC=1); the FP32 kernels stay the default.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#include "device_common.cuh"
#include "kernels.h"
#include "render_device.cuh"

namespace sgx {

namespace {

constexpr int kTcM = 128;        // output rows per tile = UMMA M
constexpr int kTcWorkers = 16;       // worker warps: four per TMEM lane quadrant
constexpr int kTcThreads = (kTcWorkers + 1) * 32; // + the warp that issues the MMAs
constexpr unsigned kTcColsD1 = 0, kTcColsThi = 128, kTcColsTlo = 256, kTcColsD2 = 384, kTcTmemCols = 512;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"), K-major, no swizzle: the operand is stored as
// [K/4][rows][4 floats], i.e. 16-byte rows of 4 K-values; 8 consecutive rows form a core matrix (128 B);
// LBO = bytes between the two 16-byte K-chunks of one MMA (= rows_pitch * 16), SBO = bytes between 8-row groups (= 128).
__device__ __forceinline__ uint64_t tc_desc(const void *base, unsigned lbo_bytes, unsigned sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(base) >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46; // descriptor version of sm_100
    return d;               // base offset 0, layout type 0 (no swizzle)
}
// Instruction descriptor: F32 accumulate, TF32 x TF32, both operands K-major, M x N
__device__ __forceinline__ uint32_t tc_idesc(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_ss(unsigned d_tmem, uint64_t a, uint64_t b, uint32_t idesc, unsigned acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_ts(unsigned d_tmem, unsigned a_tmem, uint64_t b, uint32_t idesc, unsigned acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// bounded wait: an MMA batch that never completes traps instead of hanging the GPU
__device__ __forceinline__ void tc_wait(unsigned long long *bar, unsigned parity)
{
    for (unsigned spins = 0;; ++spins) {
        unsigned done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, float (&v)[8])
{
    unsigned r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const float (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
// x = hi + lo with both terms representable in TF32 (10-bit mantissa, round to nearest)
__device__ __forceinline__ float tf32_rna(float x)
{
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo)
{
    hi = tf32_rna(x);
    lo = tf32_rna(x - hi);
}

// Thread roles: warps 0..15 are workers (fill, weights, the two epilogues; worker w touches the TMEM lanes of quadrant
// w & 3 and every fourth group of 8 columns), warp 16 issues the MMAs (one lane) and otherwise only joins the barriers.
// Per tile, with the grey tile of tile i already in shared memory:
//     issuer : pass B(i) ............................ | pass C(i) ........................ |
//     workers: weights(i) | wait B | epilogue 1(i)   | grey tile (i+1) | wait C | epilogue 2(i)
// so the weights are built under pass B and the next grey tile is fetched under pass C.
template <int CH>
__global__ void __launch_bounds__(kTcThreads, 1) render_tc_kernel(const RenderLaunch L)
{
    extern __shared__ __align__(128) float tsm[];
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ unsigned tmem_base_s;
    const int KV = L.rv_max, NF = L.fc, NX = L.px, NFp = NF + 1, NXp = NX + 1, NXs = NX + 4;
    float *A_hi = tsm, *A_lo = A_hi + (size_t)KV * kTcM;                 // [KV/4][128][4]   vertical weights of this CTA's row tile
    float *G_hi = A_lo + (size_t)KV * kTcM, *G_lo = G_hi + (size_t)(KV / 4) * NFp * 4; // [KV/4][NF+1][4] grey tile
    float *W_hi = G_lo + (size_t)(KV / 4) * NFp * 4, *W_lo = W_hi + (size_t)(NF / 4) * NXp * 4; // [NF/4][NX+1][4] horizontal weights
    float *stage = W_hi;                                                 // [128][NX+4] resampled tile, aliases the weights (dead after pass C)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool worker = warp < kTcWorkers, issuer = tid == kTcWorkers * 32;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcWorkers) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- geometry shared by every track of the launch (same T, height, nwidth, nheight) ------------------------
    const RenderTrack *__restrict__ tr0 = L.tracks;
    const int nheight = tr0->nheight, nwidth = tr0->nwidth, height = tr0->height, n_out = tr0->n_out, width = tr0->width;
    const int n_mt = (nheight + kTcM - 1) / kTcM;
    const int mt = (int)blockIdx.x % n_mt, oy0 = mt * kTcM;
    const int k0 = __ldg(tr0->v_left + oy0); // grey row of the tile's first tap
    const int pad_rows = height - n_out;      // display.rs:47-52: the top rows of the grey image are zero

    // ---- A: the vertical weights of the row tile, dense [128 x KV], split into TF32 hi / lo; once per CTA ------
    for (int idx = tid; idx < (KV / 4) * kTcM; idx += kTcThreads) {
        const int i = idx % kTcM, c = idx / kTcM, oy = oy0 + i;
        float h[4] = {0.0f, 0.0f, 0.0f, 0.0f}, l[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (oy < nheight) {
            const int left = __ldg(tr0->v_left + oy), cnt = __ldg(tr0->v_cnt + oy);
            const float rs = __frcp_rn(__ldg(tr0->v_sum + oy));
            const float *__restrict__ wrow = tr0->v_w + (size_t)oy * tr0->v_taps;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = k0 + 4 * c + j - left;
                if (t >= 0 && t < cnt) split_tf32(__ldg(wrow + t) * rs, h[j], l[j]);
            }
        }
        reinterpret_cast<float4 *>(A_hi)[idx] = make_float4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<float4 *>(A_lo)[idx] = make_float4(l[0], l[1], l[2], l[3]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tmem_base_s;
    const unsigned lane_base = (unsigned)((warp & 3) * 32) << 16; // the TMEM lanes this worker may touch
    const int hsel = (warp >> 2) & 3;                              // which groups of 8 columns it handles

    const float min_db = L.range[1], inv_span = __frcp_rn(L.range[0] - L.range[1]);
    const int n_xt = (L.rv_cols + NX - 1) / NX;                    // x tiles per track (widest track of the launch)
    const int n_items = L.n_tracks * n_xt, stride = (int)gridDim.x / n_mt;
    unsigned ph0 = 0, ph1 = 0;
    // descriptors of the operands in shared memory; a K step of 8 advances the start address by two 16-byte-row chunks
    const uint64_t dAh = tc_desc(A_hi, kTcM * 16, 128), dAl = tc_desc(A_lo, kTcM * 16, 128);
    const uint64_t dGh = tc_desc(G_hi, NFp * 16, 128), dGl = tc_desc(G_lo, NFp * 16, 128);
    const uint64_t dWh = tc_desc(W_hi, NXp * 16, 128), dWl = tc_desc(W_lo, NXp * 16, 128);
    const uint64_t sA = 2 * kTcM, sG = 2 * NFp, sW = 2 * NXp;     // in 16-byte units (the start-address field)
    const uint32_t idB = tc_idesc(kTcM, NF), idC = tc_idesc(kTcM, NX);

    // where a tile lives; `ok` false: past the end, or a tile beyond this track's columns
    struct Tile { const RenderTrack *tr; int ox0, pxc, fl0; bool ok; };
    auto tile_of = [&](int item) {
        Tile t{nullptr, 0, 0, 0, false};
        if (item >= n_items) return t;
        t.tr = L.tracks + item / n_xt;
        const int ox_end = t.tr->ox_begin + t.tr->ox_count;
        t.ox0 = t.tr->ox_begin + (item % n_xt) * NX;
        if (t.ox0 >= ox_end) return t;
        t.pxc = min(NX, ox_end - t.ox0);
        t.fl0 = __ldg(t.tr->h_left + t.ox0); // first source frame of the tile's window
        t.ok = true;
        return t;
    };
    // ---- grey tile: dB -> [0,1] (display.rs:44-54: normalise, clip, flip, top padding), TF32 hi / lo -------------
    // A warp takes one frame and 32 consecutive grey rows per load: the dB bins are contiguous (descending) in memory, and
    // with the row pitch NF + 1 the 32 stores of a request fall into 32 different banks.  All loads of a 32-row block (up
    // to six frames per warp) are issued before the first is used.
    auto fill_grey = [&](const Tile &t) {
        const float *__restrict__ src = t.tr->src;
        const int frame0 = t.tr->frame0, src_frames = t.tr->src_frames;
        for (int kb = 0; kb < (KV + 31) / 32; ++kb) {
            const int k = kb * 32 + lane, r = k0 + k;
            const bool kok = k < KV && r >= pad_rows && r < height;
            const float *__restrict__ col = src + (height - 1 - r);
            float v[6];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int f = warp + u * kTcWorkers, frame = t.fl0 + f, lf = frame - frame0;
                v[u] = -INFINITY; // -> grey 0
                if (f < NF && kok && frame < width && lf >= 0 && lf < src_frames) v[u] = __ldg(col + (size_t)lf * n_out);
            }
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int f = warp + u * kTcWorkers;
                if (f < NF && k < KV) {
                    float hi, lo;
                    split_tf32(__saturatef((v[u] - min_db) * inv_span), hi, lo);
                    const int at = ((k >> 2) * NFp + f) * 4 + (k & 3);
                    G_hi[at] = hi; G_lo[at] = lo;
                }
            }
        }
    };

    int item = (int)blockIdx.x / n_mt;
    Tile cur = tile_of(item);
    while (item < n_items && !cur.ok) { item += stride; cur = tile_of(item); } // (tiles beyond a narrow track's columns)
    if (cur.ok && worker) fill_grey(cur);

    while (cur.ok) {
        int nitem = item + stride;
        Tile nxt = tile_of(nitem);
        while (nitem < n_items && !nxt.ok) { nitem += stride; nxt = tile_of(nitem); }

        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the grey tile -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                             // grey tile complete; weights / staging / TMEM of the last tile free
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- pass B: D1 = Av * grey, 3xTF32 (small terms first) -- under it, the workers build the column weights --------
        if (issuer) {
            uint64_t ah = dAh, al = dAl, gh = dGh, gl = dGl;
            for (int ks = 0; ks < KV / 8; ++ks, ah += sA, al += sA, gh += sG, gl += sG) {
                tc_mma_ss(tb + kTcColsD1, al, gh, idB, ks > 0);
                tc_mma_ss(tb + kTcColsD1, ah, gl, idB, 1u);
                tc_mma_ss(tb + kTcColsD1, ah, gh, idB, 1u);
            }
            tc_commit(&bars[0]);
        }
        if (worker) {
            // horizontal weights of the tile's columns, dense [NX x NF]: 4 consecutive frames of one column per store
            const RenderTrack *__restrict__ tr = cur.tr;
            const int i = tid % NX, ox = cur.ox0 + i;
            int left = 0, cnt = 0;
            float rs = 0.0f;
            if (i < cur.pxc) { left = __ldg(tr->h_left + ox); cnt = __ldg(tr->h_cnt + ox); rs = __frcp_rn(__ldg(tr->h_sum + ox)); }
            const float *__restrict__ hw = tr->h_w + ox;
            const int cl = tid / NX, cstep = (kTcWorkers * 32) / NX; // >= 8 chunk lanes: at most three chunks per thread
            float wv[3][4];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int c = cl + m * cstep;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = cur.fl0 + 4 * c + j - left;
                    wv[m][j] = 0.0f;
                    if (c < NF / 4 && t >= 0 && t < cnt) wv[m][j] = __ldg(hw + (size_t)t * nwidth);
                }
            }
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int c = cl + m * cstep;
                if (c < NF / 4) {
                    float h[4], l[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split_tf32(wv[m][j] * rs, h[j], l[j]);
                    reinterpret_cast<float4 *>(W_hi)[c * NXp + i] = make_float4(h[0], h[1], h[2], h[3]);
                    reinterpret_cast<float4 *>(W_lo)[c * NXp + i] = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

            // ---- epilogue 1: the clamp of vertical_sample, then T as the A operand of pass C (hi / lo) ---------------
            tc_wait(&bars[0], ph0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int c = hsel * 8; c < NF; c += 32) {
                float v[8], hi[8], lo[8];
                tmem_ld8(tb + kTcColsD1 + lane_base + c, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) split_tf32(clamp_fin(v[j]), hi[j], lo[j]);
                tmem_st8(tb + kTcColsThi + lane_base + c, hi);
                tmem_st8(tb + kTcColsTlo + lane_base + c, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        ph0 ^= 1u;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                             // T and the weights are ready
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- pass C: D2 = T * Ah^T, A from TMEM -- under it, the workers fetch the grey tile of the next tile -------------
        if (issuer) {
            uint64_t wh = dWh, wl = dWl;
            for (int kf = 0; kf < NF / 8; ++kf, wh += sW, wl += sW) {
                tc_mma_ts(tb + kTcColsD2, tb + kTcColsTlo + kf * 8, wh, idC, kf > 0);
                tc_mma_ts(tb + kTcColsD2, tb + kTcColsThi + kf * 8, wl, idC, 1u);
                tc_mma_ts(tb + kTcColsD2, tb + kTcColsThi + kf * 8, wh, idC, 1u);
            }
            tc_commit(&bars[1]);
        }
        if (worker) {
            if (nxt.ok) fill_grey(nxt); // pass B of this tile has completed (bars[0]): the grey buffer is free

            // ---- epilogue 2: the clamp of horizontal_sample; through shared memory so that lanes run along a row -------
            tc_wait(&bars[1], ph1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                float *srow = stage + (size_t)((warp & 3) * 32 + lane) * NXs;
                for (int c = hsel * 8; c < NX; c += 32) {
                    float v[8];
                    tmem_ld8(tb + kTcColsD2 + lane_base + c, v);
                    reinterpret_cast<float4 *>(srow + c)[0] = make_float4(clamp_fin(v[0]), clamp_fin(v[1]), clamp_fin(v[2]), clamp_fin(v[3]));
                    reinterpret_cast<float4 *>(srow + c)[1] = make_float4(clamp_fin(v[4]), clamp_fin(v[5]), clamp_fin(v[6]), clamp_fin(v[7]));
                }
            }
            asm volatile("bar.sync 1, %0;" :: "r"(kTcWorkers * 32) : "memory"); // workers only: the staged tile is complete
            // colour map + stores: a warp walks rows, its lanes run along the columns (neighbouring pixels of a row mostly
            // share a colour segment, and a store covers 128 contiguous bytes)
            {
                const RenderTrack *__restrict__ tr = cur.tr;
                unsigned char *__restrict__ outp = tr->out;
                const int opitch = tr->ox_count, oxrel = cur.ox0 - tr->ox_begin;
                for (int rrow = warp; rrow < kTcM; rrow += kTcWorkers) {
                    const int oy = oy0 + rrow;
                    if (oy >= nheight) break;
                    const size_t pix0 = (size_t)oy * opitch + (size_t)oxrel;
                    for (int c = lane; c < cur.pxc; c += 32) {
                        const unsigned px = grey_to_rgba_const(stage[(size_t)rrow * NXs + c]);
                        if (CH == 4) reinterpret_cast<unsigned *>(outp)[pix0 + c] = px;
                        else {
                            unsigned char *o = outp + (pix0 + c) * 3;
                            o[0] = (unsigned char)px; o[1] = (unsigned char)(px >> 8); o[2] = (unsigned char)(px >> 16);
                        }
                    }
                }
            }
        }
        ph1 ^= 1u;
        item = nitem; cur = nxt;
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kTcWorkers) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(kTcTmemCols) : "memory");
}

} // namespace

// Shared memory of one CTA for a row-window of kv grey rows, nf source frames and nx output columns.
size_t render_tc_smem(int kv, int nf, int nx)
{
    const size_t weights = (size_t)2 * (nf / 4) * (nx + 1) * 4, stage = (size_t)kTcM * (nx + 4); // the staged tile aliases the weights
    return ((size_t)2 * kv * kTcM + (size_t)2 * (kv / 4) * (nf + 1) * 4 + std::max(weights, stage)) * sizeof(float);
}

cudaError_t launch_render_tc(const RenderLaunch &L, size_t smem_bytes, cudaStream_t s)
{
    if (L.n_tracks <= 0) return cudaSuccess;
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    const int n_mt = L.rv_rows; // row tiles of 128
    const RenderLaunch &P = L;
    const int n_xt = (L.rv_cols + L.px - 1) / L.px;
    const long items = (long)L.n_tracks * n_xt;
    int per_mt = (int)std::min<long>(items, std::max(1, sms / n_mt));
    const int grid = per_mt * n_mt;
    cudaError_t e;
    if (L.channels == 4) {
        e = ensure_dynamic_smem(reinterpret_cast<const void *>(render_tc_kernel<4>), smem_bytes);
        if (e != cudaSuccess) return e;
        render_tc_kernel<4><<<grid, kTcThreads, smem_bytes, s>>>(P);
    } else {
        e = ensure_dynamic_smem(reinterpret_cast<const void *>(render_tc_kernel<3>), smem_bytes);
        if (e != cudaSuccess) return e;
        render_tc_kernel<3><<<grid, kTcThreads, smem_bytes, s>>>(P);
    }
    count_launch();
    return cudaGetLastError();
}

} // namespace sgx
