// render_slide_kernel.cu -- K3 for the 8-tap class on both axes (n_in / n_out < 7/6: the bench geometry of
// benches/bench.rs:57 and every zoom level that magnifies), display.rs:44-61 in one kernel:
//     dB -> grey (normalise, clip, flip, top-pad; display.rs:44-54) -> Lanczos3 rows, clamp -> Lanczos3 columns, clamp
//     (image 0.23 resize = vertical_sample then horizontal_sample) -> colour map (display.rs:24-42) -> RGBA.
//
// Why another kernel: render_fast_kernel gives every output its own eight 128-bit shared-memory loads per pass (one
// lane <-> one output index, 4 FMAs per load): 96 instructions and 0.8 shared-memory wavefronts per pixel, the
// shared-memory / L1 data pipe 91 % busy (ncu, profiles/r01_v8_k3_render_fast.txt).  Here the lanes of a warp run ACROSS
// the resampled axis and every lane SLIDES an 8-tap window along it in registers, two or four samples per lane as packed
// FP32 pairs (FFMA2: the tap weight is broadcast, one instruction serves both samples):
//   A  raw source tile G[row][frame] by 4-byte asynchronous copies (LDGSTS), all of a lane's copies in flight at once,
//   B  vertical pass: a lane owns frames l, l+32, l+64, l+96; a warp walks down 8 output rows; the window is a ring of
//      eight register pairs indexed by (steps taken) mod 8 at compile time, so advancing it by one source row is ONE
//      shared load per frame and no register moves; dB -> grey (display.rs:49-51) is applied as a value enters the
//      window; the eight weights of an output row are warp-uniform (two broadcast 128-bit loads); clamp; Tm[frame][row],
//   C  horizontal pass: a lane owns output rows l, l+32; a warp walks along 16 output columns; where eight columns start
//      exactly one frame apart (ratio ~ 1: the bench geometry) their windows are w[j .. j+7] of fifteen consecutive
//      frames (seven loads per eight columns, no moves), otherwise the window shifts by h_left[c+1] - h_left[c]; clamp,
//      colour map without conversion instructions (render_device.cuh) with the segment constants cached per row;
//      pixels leave through a per-warp staging tile as 128-bit stores (16 rows x 32 bytes per instruction).
// Shared memory per CTA at the bench geometry (tile 120 x 64 pixels, 128 frames x 56 grey rows): G 29.6 KB (reused as the
// pixel staging) + Tm 33.3 KB + tables 6.8 KB = 70 KB -> 3 CTAs per SM, 80 registers.
// MEASURED (B200, C5, ms per step; profiles/r02_k3_slide_*): 3.54 against render_fast_kernel's 3.69 with 81 instead of 96
// instructions per pixel and 0.6 instead of 0.8 shared-memory wavefronts per pixel.  The gain is small because the
// kernel is no longer bound by one pipe: issue slots 59 % busy, L1/shared data pipe 73 % (a third of it the 32-byte
// sectors of the global stores and the LDGSTS fill), and with 24 warps per SM the dependent FFMA2 chains and
// shared-memory round trips are not fully hidden (stall reasons: short scoreboard 2.3, wait 2.2 per issue).
// Arithmetic (order of the taps, normalised weights, clamps, colour map) is that of render_fast_kernel operation for
// operation, so the two kernels produce identical pixels; SGX_K3_SLIDE=0 selects render_fast_kernel.
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "device_common.cuh"
#include "host_tables.h"
#include "kernels.h"
#include "render_device.cuh"

namespace sgx {

namespace {

constexpr int kSlThreads = 256, kSlWarps = kSlThreads / 32;
constexpr int kSlPX = 120, kSlPY = 64; // output tile; 120 columns need 127-128 frames at ratio 1: four full lane groups
constexpr int kSlTP = kSlPY + 1;       // pitch of Tm [frame][out row]: odd, so both the frame-wise stores and the row-wise loads are conflict-free
constexpr int kSlFR = 128;             // source frames a tile can hold: a lane of phase B owns frames l, l + 32, l + 64, l + 96
constexpr int kSlGP = kSlFR + 4;       // pitch of G [row][frame]: 4 mod 32, the transposing stores of phase A are conflict-free
constexpr int kSlRowsPerItem = 8;      // phase B: output rows a warp walks per item (for all 128 frames)
constexpr int kSlColsPerItem = 16;     // phase C: output columns a warp walks per item (for 64 rows; multiple of 8)

__host__ __device__ constexpr size_t slide_table_floats() { return (size_t)8 * kSlPY + 8 * kSlPX + kSlPY + 128 + 16; }

__device__ __forceinline__ void st_global_v4(void *p, unsigned a, unsigned b, unsigned c, unsigned d)
{
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 4-byte asynchronous copy global -> shared (LDGSTS): no register in between, so a lane keeps all its copies in flight
__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(__cvta_generic_to_global(gsrc)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void st_global_u32(void *p, unsigned a)
{
    asm volatile("st.global.u32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(a) : "memory");
}
// One output of a pass for the TWO samples a lane owns: eight taps in ascending order into one accumulator each (the
// order of image 0.23's sample loops), as packed FP32 pairs -- one FFMA2 per tap serves both (the weight is broadcast).
typedef float2 pk;
__device__ __forceinline__ pk taps8(pk w0, pk w1, pk w2, pk w3, pk w4, pk w5, pk w6, pk w7, const float4 a, const float4 b)
{
    pk t = __fmul2_rn(w0, make_float2(a.x, a.x));
    t = __ffma2_rn(w1, make_float2(a.y, a.y), t); t = __ffma2_rn(w2, make_float2(a.z, a.z), t);
    t = __ffma2_rn(w3, make_float2(a.w, a.w), t); t = __ffma2_rn(w4, make_float2(b.x, b.x), t);
    t = __ffma2_rn(w5, make_float2(b.y, b.y), t); t = __ffma2_rn(w6, make_float2(b.z, b.z), t);
    t = __ffma2_rn(w7, make_float2(b.w, b.w), t);
    return t;
}
// the per-pass clamp of image 0.23's resize, clamp(t, 0, f32::MAX), for sums of finite products (never NaN or inf)
__device__ __forceinline__ float clamp0(float v) { return fmaxf(v, 0.0f); }
constexpr int kSlSP = 68; // pitch of the pixel staging tile [8 columns][64 rows + 4]: 4 * 68 = 16 (mod 32)

template <bool FROM_DB, int CH>
__global__ void __launch_bounds__(kSlThreads, 3) render_slide_kernel(const RenderLaunch L)
{
    constexpr int FR = kSlFR, GP = kSlGP;
    const int RCAP = L.rv_max;
    extern __shared__ __align__(16) float rsm[];
    float *Tm = rsm;                                               // [FR][kSlTP]
    float4 *vW = reinterpret_cast<float4 *>(Tm + (size_t)FR * kSlTP); // [kSlPY][2] normalised weights of an output row
    float4 *hW = vW + 2 * kSlPY;                                   // [kSlPX][2]
    int *vL = reinterpret_cast<int *>(hW + 2 * kSlPX);            // [kSlPY]  first source row, tile coordinates
    int *hL = vL + kSlPY;                                          // [128]    first source frame, tile coordinates
    int *hU = hL + 128;                                            // [16]     columns 8k .. 8k+7 start one frame apart
    float *G = reinterpret_cast<float *>(hU + 16);                // [RCAP][GP]; phase C reuses it as pixel staging

    const RenderTrack *__restrict__ tr = L.tracks + blockIdx.z;
    const int nwidth = tr->nwidth, nheight = tr->nheight;
    const int ox_begin = tr->ox_begin, ox_count = tr->ox_count, frame0 = tr->frame0, src_frames = tr->src_frames;
    const int ox_end = ox_begin + ox_count; // this launch renders columns [ox_begin, ox_end)
    const int ox0 = ox_begin + blockIdx.x * kSlPX, oy0 = blockIdx.y * kSlPY;
    if (ox0 >= ox_end || oy0 >= nheight) return;
    const int pxc = min(kSlPX, ox_end - ox0), pyc = min(kSlPY, nheight - oy0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int *__restrict__ h_left = tr->h_left;
    const int *__restrict__ v_left = tr->v_left;
    const float *__restrict__ src = tr->src;
    const int width = tr->width, height = tr->height, n_out = tr->n_out;

    const int fl = __ldg(h_left + ox0);
    const int nfr = min(__ldg(h_left + ox0 + pxc - 1) + 8 - fl, FR);
    const int nfq = (nfr + 3) >> 2;                 // frame quads
    const int yl = __ldg(v_left + oy0);
    const int nrow = min(__ldg(v_left + oy0 + pyc - 1) + 8 - yl, RCAP);

    // ---- tables of this tile: weights divided by their sum once, window starts in tile coordinates ---------------
    if (tid < kSlPY) {
        const int oy = oy0 + min(tid, pyc - 1);
        const float4 *__restrict__ wrow = reinterpret_cast<const float4 *>(tr->v_w + (size_t)oy * tr->v_taps); // 16-byte aligned rows
        float4 a = __ldg(wrow), b = __ldg(wrow + 1);
        const float rs = __frcp_rn(__ldg(tr->v_sum + oy));
        a.x *= rs; a.y *= rs; a.z *= rs; a.w *= rs; b.x *= rs; b.y *= rs; b.z *= rs; b.w *= rs;
        vW[2 * tid] = a; vW[2 * tid + 1] = b;
        vL[tid] = min(__ldg(v_left + oy) - yl, RCAP - 8);
    } else if (tid - kSlPY < kSlPX) {
        const int c = tid - kSlPY;
        const int ox = ox0 + min(c, pxc - 1);
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __ldg(tr->h_w + (size_t)i * nwidth + ox);
        const float rs = __frcp_rn(__ldg(tr->h_sum + ox));
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] *= rs;
        hW[2 * c] = make_float4(w[0], w[1], w[2], w[3]); hW[2 * c + 1] = make_float4(w[4], w[5], w[6], w[7]);
        hL[c] = min(__ldg(h_left + ox) - fl, FR - 8);
    }

    // ---- A: source tile G[row][frame], RAW (dB or grey): a warp copies 8 rows x 4 frames per request (one 32-byte sector
    //         per frame) with 4-byte asynchronous copies, all of a lane's copies in flight at once; the dB -> grey
    //         normalisation (display.rs:49-51) is applied when phase B loads a value into its window ------------------
    float min_db = 0.0f, inv_span = 0.0f;
    if (FROM_DB) { min_db = L.range[1]; inv_span = __frcp_rn(L.range[0] - L.range[1]); }
    {
        // rows of the tile that hold data: grey rows [pad_rows, height) (display.rs:47-52), in tile coordinates
        const int pad_rows = FROM_DB ? height - n_out : 0;
        const int yy_lo = max(pad_rows - yl, 0), yy_hi = min(height - yl, nrow);
        const int fsub = lane & 3, rsub = lane >> 2;
        const int nsteps = (nrow + 7) >> 3;          // 8-row steps; rows up to 8 * nsteps <= RCAP are written (zeros past nrow)
        const unsigned yy_span = (unsigned)max(yy_hi - yy_lo, 0);
        const int gstep = 8 * GP;
        for (int fq = warp; fq < nfq; fq += kSlWarps) {
            const int fx = fq * 4 + fsub;
            const int f = fl + fx;
            const int lf = FROM_DB ? f - frame0 : f; // row of the (possibly time-sliced) dB array
            const bool fok = f < width && lf >= 0 && (!FROM_DB || lf < src_frames);
            // FROM_DB: element (frame f, grey row y) is dB[lf][height-1-y]; else grey[y][f]
            const float *__restrict__ p0 = FROM_DB ? src + (size_t)(fok ? lf : 0) * n_out + (height - 1 - yl - rsub)
                                                   : src + (size_t)(yl + rsub) * width + (fok ? f : 0);
            float *gq = G + rsub * GP + fx;
            // a lane's rows are rsub + 8 k: valid while (unsigned)(row - yy_lo) < yy_span; no frame -> no valid row
            const unsigned span = fok ? yy_span : 0u;
            int rel = rsub - yy_lo;
            for (int s0 = 0; s0 < nsteps; s0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (s0 + u < nsteps) { // warp-uniform
                        if ((unsigned)(rel + 8 * u) < span) cp_async4(gq + u * gstep, FROM_DB ? p0 - 8 * u : p0 + (size_t)(8 * u) * width);
                        else gq[u * gstep] = FROM_DB ? -INFINITY : 0.0f; // -inf -> grey 0 after the saturate
                    }
                }
                rel += 64; gq += 8 * gstep;
                p0 = FROM_DB ? p0 - 64 : p0 + (size_t)64 * width;
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (tid < kSlPX / 8) { // blocks of eight columns whose windows start exactly one frame apart (ratio ~ 1)
        const int b0 = hL[8 * tid];
        bool unit = true;
#pragma unroll
        for (int j = 1; j < 8; ++j) unit = unit && hL[8 * tid + j] == b0 + j;
        hU[tid] = unit ? 1 : 0;
    }

    // ---- B: vertical pass: a lane owns frames l, l + 32, l + 64, l + 96; the window slides down the output rows ---------
    {
        const int nrc = (pyc + kSlRowsPerItem - 1) / kSlRowsPerItem;
        // lanes past nfr work on a copy of the last frame; nobody reads their Tm rows
        const float *__restrict__ g0 = G + min(lane, nfr - 1), *__restrict__ g1 = G + min(lane + 32, nfr - 1);
        const float *__restrict__ g2 = G + min(lane + 64, nfr - 1), *__restrict__ g3 = G + min(lane + 96, nfr - 1);
        float *t0 = Tm + lane * kSlTP;
        auto grey = [&](float v) { return FROM_DB ? __saturatef((v - min_db) * inv_span) : v; }; // display.rs:49-51
        for (int it = warp; it < nrc; it += kSlWarps) {
            int r = it * kSlRowsPerItem;
            const int r1 = min(r + kSlRowsPerItem, pyc);
            int cur = vL[r];
            pk wa[8], wb[8]; // (frame l, l + 32) and (l + 64, l + 96) of source rows cur .. cur + 7
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int o = (cur + i) * GP;
                wa[i] = make_float2(grey(g0[o]), grey(g1[o])); wb[i] = make_float2(grey(g2[o]), grey(g3[o]));
            }
            int nxt = (cur + 8) * GP; // the next source row to enter the window
            // The window holds source rows cur .. cur + 7 in w[(k + i) & 7] (k = steps taken, mod 8): every output row whose
            // first tap is `cur` is finished, then the oldest row is replaced by row cur + 8 -- no register moves.
            bool more = true;
            while (more) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (more) { // warp-uniform
                        while (r < r1 && vL[r] <= cur) { // v_left never decreases
                            const float4 a = vW[2 * r], b = vW[2 * r + 1];
                            const pk ta = taps8(wa[k & 7], wa[(k + 1) & 7], wa[(k + 2) & 7], wa[(k + 3) & 7], wa[(k + 4) & 7],
                                                wa[(k + 5) & 7], wa[(k + 6) & 7], wa[(k + 7) & 7], a, b);
                            const pk tb = taps8(wb[k & 7], wb[(k + 1) & 7], wb[(k + 2) & 7], wb[(k + 3) & 7], wb[(k + 4) & 7],
                                                wb[(k + 5) & 7], wb[(k + 6) & 7], wb[(k + 7) & 7], a, b);
                            float *to = t0 + r;
                            to[0] = clamp0(ta.x); to[32 * kSlTP] = clamp0(ta.y);
                            to[64 * kSlTP] = clamp0(tb.x); to[96 * kSlTP] = clamp0(tb.y);
                            ++r;
                        }
                        if (r >= r1) more = false;
                        else {
                            wa[k & 7] = make_float2(grey(g0[nxt]), grey(g1[nxt])); wb[k & 7] = make_float2(grey(g2[nxt]), grey(g3[nxt]));
                            nxt += GP; ++cur;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- C: horizontal pass: a lane owns output rows (l, l + 32), the window slides along the output columns; colour; store --
    {
        const int ncc = (pxc + kSlColsPerItem - 1) / kSlColsPerItem;
        // Pixels leave through a per-warp staging tile [8 columns][kSlSP]: a lane writes its two rows of a column (conflict-free),
        // then reads FOUR columns of one row and stores them as 16 bytes -- per store instruction 16 rows x 8 columns, row
        // segments of 32 bytes.  (Two 128-bit stores per lane straight from registers, 32 rows apart, measured slower.)
        unsigned *stage = reinterpret_cast<unsigned *>(G) + warp * (8 * kSlSP);
        unsigned char *__restrict__ outp = tr->out;
        // x0 is a multiple of 8 pixels by construction: four RGBA pixels are 16 aligned bytes, four RGB pixels 12 bytes on a 4-byte boundary
        const bool wide = (ox_count & 3) == 0 && (reinterpret_cast<size_t>(outp) & (CH == 4 ? 15 : 3)) == 0;
        unsigned *sa = stage + lane, *sb = sa + 32;
        const float *__restrict__ ta = Tm + min(lane, pyc - 1);
        const float *__restrict__ tb = Tm + min(lane + 32, pyc - 1);
        for (int it = warp; it < ncc; it += kSlWarps) {
            const int c0 = it * kSlColsPerItem, c1 = min(c0 + kSlColsPerItem, pxc);
            int cur = hL[c0];
            pk w0, w1, w2, w3, w4, w5, w6, w7;
            {
                const int o = cur * kSlTP;
                w0 = make_float2(ta[o], tb[o]); w1 = make_float2(ta[o + kSlTP], tb[o + kSlTP]);
                w2 = make_float2(ta[o + 2 * kSlTP], tb[o + 2 * kSlTP]); w3 = make_float2(ta[o + 3 * kSlTP], tb[o + 3 * kSlTP]);
                w4 = make_float2(ta[o + 4 * kSlTP], tb[o + 4 * kSlTP]); w5 = make_float2(ta[o + 5 * kSlTP], tb[o + 5 * kSlTP]);
                w6 = make_float2(ta[o + 6 * kSlTP], tb[o + 6 * kSlTP]); w7 = make_float2(ta[o + 7 * kSlTP], tb[o + 7 * kSlTP]);
            }
            auto slide_to = [&](int lft) { // warp-uniform: h_left never decreases
#pragma unroll 1
                while (cur < lft) {
                    w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6; w6 = w7;
                    const int o = (cur + 8) * kSlTP;
                    w7 = make_float2(ta[o], tb[o]);
                    ++cur;
                }
            };
            CmCache cma, cmb; // colour segment constants of the lane's two rows
            cm_cache_init(cma); cm_cache_init(cmb);
            auto emit = [&](int j, pk t) { // the saturate is the clamp at 0; above 1 every value is the last colour (display.rs:31)
                sa[j * kSlSP] = grey_to_rgba_cached(t.x, cma);
                sb[j * kSlSP] = grey_to_rgba_cached(t.y, cmb);
            };
            for (int cb = c0; cb < c1; cb += 8) {
                const int nb = min(8, c1 - cb);
                if (nb == 8 && hU[cb >> 3]) {
                    // the eight windows of this block are w[j .. j+7] of fifteen consecutive frames: seven loads, no moves
                    slide_to(hL[cb]);
                    const int o = (cur + 8) * kSlTP;
                    const float4 *__restrict__ hw = hW + 2 * cb;
                    const pk w8 = make_float2(ta[o], tb[o]), w9 = make_float2(ta[o + kSlTP], tb[o + kSlTP]);
                    const pk w10 = make_float2(ta[o + 2 * kSlTP], tb[o + 2 * kSlTP]), w11 = make_float2(ta[o + 3 * kSlTP], tb[o + 3 * kSlTP]);
                    const pk w12 = make_float2(ta[o + 4 * kSlTP], tb[o + 4 * kSlTP]), w13 = make_float2(ta[o + 5 * kSlTP], tb[o + 5 * kSlTP]);
                    const pk w14 = make_float2(ta[o + 6 * kSlTP], tb[o + 6 * kSlTP]);
                    emit(0, taps8(w0, w1, w2, w3, w4, w5, w6, w7, hw[0], hw[1]));
                    emit(1, taps8(w1, w2, w3, w4, w5, w6, w7, w8, hw[2], hw[3]));
                    emit(2, taps8(w2, w3, w4, w5, w6, w7, w8, w9, hw[4], hw[5]));
                    emit(3, taps8(w3, w4, w5, w6, w7, w8, w9, w10, hw[6], hw[7]));
                    emit(4, taps8(w4, w5, w6, w7, w8, w9, w10, w11, hw[8], hw[9]));
                    emit(5, taps8(w5, w6, w7, w8, w9, w10, w11, w12, hw[10], hw[11]));
                    emit(6, taps8(w6, w7, w8, w9, w10, w11, w12, w13, hw[12], hw[13]));
                    emit(7, taps8(w7, w8, w9, w10, w11, w12, w13, w14, hw[14], hw[15]));
                    w0 = w7; w1 = w8; w2 = w9; w3 = w10; w4 = w11; w5 = w12; w6 = w13; w7 = w14;
                    cur += 7;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < nb) { // warp-uniform
                            const int c = cb + j;
                            slide_to(hL[c]);
                            emit(j, taps8(w0, w1, w2, w3, w4, w5, w6, w7, hW[2 * c], hW[2 * c + 1]));
                        }
                    }
                }
                __syncwarp();
                const size_t x0 = (size_t)(ox0 - ox_begin + cb);
                if (wide && nb == 8) {
                    const unsigned *sp = stage + (lane & 1) * (4 * kSlSP) + (lane >> 1);
                    unsigned char *dst = outp + ((size_t)(oy0 + (lane >> 1)) * ox_count + x0 + 4 * (lane & 1)) * 4;
                    const size_t dstep = (size_t)16 * ox_count * 4;
                    unsigned char *dst3 = outp + ((size_t)(oy0 + (lane >> 1)) * ox_count + x0 + 4 * (lane & 1)) * 3;
                    const size_t dstep3 = (size_t)16 * ox_count * 3;
#pragma unroll
                    for (int g = 0; g < 4; ++g) { // rows 16 g + lane / 2, columns 4 (lane & 1) .. + 3
                        const unsigned p0 = sp[16 * g], p1 = sp[16 * g + kSlSP], p2 = sp[16 * g + 2 * kSlSP], p3 = sp[16 * g + 3 * kSlSP];
                        if (16 * g + (lane >> 1) < pyc) {
                            if (CH == 4) st_global_v4(dst + g * dstep, p0, p1, p2, p3);
                            else { // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3  (the reference's RGB layout, lib.rs:294-298)
                                unsigned char *d3 = dst3 + g * dstep3;
                                st_global_u32(d3, __byte_perm(p0, p1, 0x4210));
                                st_global_u32(d3 + 4, __byte_perm(p1, p2, 0x5421));
                                st_global_u32(d3 + 8, __byte_perm(p2, p3, 0x6542));
                            }
                        }
                    }
                } else {
                    const int sub_r = lane >> 3, sub_c = lane & 7;
                    if (sub_c < nb) {
                        for (int q = 0; q < 16; ++q) { // per store instruction 4 rows x 8 columns
                            const int rl = 4 * q + sub_r;
                            const unsigned c = stage[sub_c * kSlSP + rl];
                            if (rl < pyc) {
                                const size_t pix = (size_t)(oy0 + rl) * ox_count + x0 + sub_c;
                                if (CH == 4) st_global_u32(outp + pix * 4, c);
                                else { outp[pix * 3] = (unsigned char)c; outp[pix * 3 + 1] = (unsigned char)(c >> 8); outp[pix * 3 + 2] = (unsigned char)(c >> 16); }
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
}

} // namespace

// Largest number of source samples any window of `tile` consecutive outputs needs (first tap of its first output to
// the eighth tap of its last), from the very f32 expressions the tap tables are built with -- exact, so that a tile's
// capacity can be the next multiple of 32 above it and not one group more.  Cached per geometry.
static int slide_max_span(int n_in, int n_out, int tile)
{
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, int> cache;
    const auto key = std::make_tuple(n_in, n_out, tile);
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
    }
    std::vector<int> left((size_t)n_out);
    for (int o = 0; o < n_out; ++o) {
        uint32_t l, r;
        lanczos3_span((uint32_t)n_in, (uint32_t)n_out, (uint32_t)o, &l, &r);
        left[(size_t)o] = (int)l;
    }
    int span = 8;
    for (int o = 0; o < n_out; ++o) {
        const int last = std::min(o + tile, n_out) - 1;
        span = std::max(span, left[(size_t)last] + 8 - left[(size_t)o]);
    }
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = span;
    return span;
}

// Capacities of a tile for the geometry (width -> nwidth, height -> nheight); false when the geometry is outside the
// 8-tap class or the tile does not fit.
bool render_slide_plan(int width, int height, int nwidth, int nheight, RenderTiling *out)
{
    if (width <= 0 || height <= 0 || nwidth <= 0 || nheight <= 0) return false;
    const float rhf = (float)width / (float)nwidth, rvf = (float)height / (float)nheight;
    if (!(rhf < 1.16f && rvf < 1.16f)) return false; // same bound as the 8-tap class of render_fast_kernel
    const int frames = slide_max_span(width, nwidth, kSlPX);
    const int rows = slide_max_span(height, nheight, kSlPY);
    if (frames > kSlFR) return false; // the horizontal ratio is above ~1: render_fast_kernel's geometry
    const int fr = kSlFR;
    const int rcap = (rows + 7) & ~7;
    const size_t g_floats = std::max((size_t)rcap * kSlGP, (size_t)kSlWarps * 8 * kSlSP);
    const size_t smem = ((size_t)fr * kSlTP + slide_table_floats() + g_floats) * sizeof(float);
    if (smem + 1024 > (size_t)227 * 1024) return false;
    *out = RenderTiling{kSlPX, kSlPY, fr, rcap, smem, 3};
    return true;
}

cudaError_t launch_render_slide(const RenderLaunch &L_in, int max_nwidth, int max_nheight, size_t smem_bytes, cudaStream_t s)
{
    dim3 grid((max_nwidth + kSlPX - 1) / kSlPX, (max_nheight + kSlPY - 1) / kSlPY, L_in.n_tracks);
    cudaError_t err = cudaErrorInvalidValue;
    const RenderLaunch &L = L_in;
#define SGX_SLIDE(DB, CHN)                                                                                    \
    if ((L.from_db != 0) == DB && L.channels == CHN) {                                                        \
        auto kern = render_slide_kernel<DB, CHN>;                                                             \
        cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem_bytes);               \
        if (e != cudaSuccess) return e;                                                                       \
        kern<<<grid, kSlThreads, smem_bytes, s>>>(L);                                                         \
        err = cudaSuccess;                                                                                    \
    }
    SGX_SLIDE(true, 4) SGX_SLIDE(true, 3) SGX_SLIDE(false, 4) SGX_SLIDE(false, 3)
#undef SGX_SLIDE
    if (err != cudaSuccess) return err;
    count_launch();
    return cudaGetLastError();
}

} // namespace sgx
