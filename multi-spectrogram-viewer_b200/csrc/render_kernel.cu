// render_kernel.cu -- K2 (global dB range) and K3 (fused rasteriser), sm_100a.
//
// K3 turns the cached dB spectrogram of a batch of tracks into pixels in one pass:
//   normalise / clip / vertical flip / top padding    display.rs:44-54  (spec_to_grey)
//   separable Lanczos3 resample, vertical then horizontal, each pass followed by the clamp to
//   [0, f32::MAX] of image 0.23's imageops::resize                     display.rs:57
//   10-stop colour map                                                  display.rs:24-42
//   RGB / RGBA bytes, row 0 = highest frequency                         display.rs:56-61
// The grey image and the vertically resampled intermediate only ever exist as shared-memory
// tiles; HBM sees one read of dB and one write of pixels.
#include <cstdint>
#include <cmath>
#include <cuda_fp16.h>
#include <cstdlib>
#include <type_traits>

#include "device_common.cuh"
#include "kernels.h"
#include "render_device.cuh"

namespace sgx {

namespace {

// ---------------------------------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------------------------------
__global__ void range_init_kernel(unsigned *slots, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        slots[2 * i] = enc_ordered(-INFINITY);    // lib.rs:203 identity of the reduce
        slots[2 * i + 1] = enc_ordered(INFINITY);
    }
}

__global__ void range_reset_kernel(const StftTrack *__restrict__ descs, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && descs[i].range_slot != nullptr) {
        descs[i].range_slot[0] = enc_ordered(-INFINITY);
        descs[i].range_slot[1] = enc_ordered(INFINITY);
    }
}

__device__ __forceinline__ void range_commit_dev(const float *max_negmin, float db_range, float *state);

// state != nullptr: the commit of lib.rs:208-218 follows in the same launch (single handle, no exchange in between)
__global__ void range_reduce_kernel(const unsigned *__restrict__ slots, int n, float *out, float max_sr, float max_sec,
                                    float db_range, float *state)
{
    __shared__ float smax[32], smin[32];
    float mx = -INFINITY, mn = INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        mx = fmaxf(mx, dec_ordered(slots[2 * i]));
        mn = fminf(mn, dec_ordered(slots[2 * i + 1]));
    }
    for (int s = 16; s > 0; s >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
    }
    if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = mx; smin[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)blockDim.x / 32; ++w) { mx = fmaxf(mx, smax[w]); mn = fminf(mn, smin[w]); }
        out[0] = mx;  // all-reduce(MAX) friendly: {max, -min, max sample rate, longest track in seconds}
        out[1] = -mn;
        out[2] = max_sr;  // lib.rs:220-224 and lib.rs:178-182: metadata of this handle's tracks, so that one
        out[3] = max_sec; // exchange serves everything update_spec_greys needs from the other shards
        if (state != nullptr) range_commit_dev(out, db_range, state);
    }
}

// approx::abs_diff_ne!(a, b, epsilon = e): !( |a-b| <= e ), with the subtraction ordered as approx does
__device__ __forceinline__ bool abs_diff_ne(float a, float b, float eps)
{
    const float d = a > b ? a - b : b - a;
    return !(d <= eps);
}

__device__ __forceinline__ void range_commit_dev(const float *max_negmin, float db_range, float *state)
{
    // lib.rs:208-209: max = max.min(0.); min = min.max(max - db_range)
    const float mx = fminf(max_negmin[0], 0.0f);
    const float mn = fmaxf(-max_negmin[1], mx - db_range);
    // lib.rs:210-218: the stored range only moves when it differs by more than 1e-3
    if (abs_diff_ne(state[0], mx, 1e-3f)) { state[0] = mx; state[2] = 1.0f; }
    if (abs_diff_ne(state[1], mn, 1e-3f)) { state[1] = mn; state[2] = 1.0f; }
    state[3] = max_negmin[2]; state[4] = max_negmin[3]; // max_sr, max_sec over all shards
}

__global__ void range_commit_kernel(const float *__restrict__ max_negmin, float db_range, float *state)
{
    range_commit_dev(max_negmin, db_range, state);
}

// ---------------------------------------------------------------------------------------------------
// Lanczos3 tap tables (image 0.23 horizontal_sample / vertical_sample geometry).  One thread per
// output index; every f32 operation is a separately rounded IEEE op (no FMA contraction) so the
// tap windows are exactly the ones the CPU code derives.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sinc_dev(float t)
{
    const float a = __fmul_rn(t, 3.14159265358979323846f);
    return t == 0.0f ? 1.0f : __fdiv_rn(sinf(a), a);
}
__device__ __forceinline__ float lanczos3_dev(float x)
{
    return fabsf(x) < 3.0f ? __fmul_rn(sinc_dev(x), sinc_dev(__fdiv_rn(x, 3.0f))) : 0.0f;
}

// upper bound of the taps one output index has: (r - l) of the table builder below
__host__ __device__ inline int lanczos3_taps_bound(int n_in, int n_out)
{
    const float ratio = (float)n_in / (float)n_out;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    return (int)(2.0f * 3.0f * sratio) + 3;
}

__global__ void build_axis_table_kernel(int n_in, int n_out, int taps, int tap_major, int *left_o,
                                        int *cnt_o, float *sum_o, float *w_o)
{
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    const float ratio = __fdiv_rn((float)n_in, (float)n_out);
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = __fmul_rn(3.0f, sratio);
    float inputx = __fmul_rn(__fadd_rn((float)o, 0.5f), ratio);
    long long l = (long long)floorf(__fsub_rn(inputx, support));
    l = l < 0 ? 0 : (l > (long long)n_in - 1 ? (long long)n_in - 1 : l);
    long long r = (long long)ceilf(__fadd_rn(inputx, support));
    r = r < l + 1 ? l + 1 : (r > (long long)n_in ? (long long)n_in : r);
    inputx = __fsub_rn(inputx, 0.5f);
    int cnt = (int)(r - l);
    if (cnt > taps) cnt = taps; // cannot happen when taps == lanczos3_max_taps(n_in, n_out)
    float sum = 0.0f;
    for (int i = 0; i < taps; ++i) {
        float w = 0.0f;
        if (i < cnt) {
            w = lanczos3_dev(__fdiv_rn(__fsub_rn((float)(l + i), inputx), sratio));
            sum = __fadd_rn(sum, w);
        }
        if (tap_major) w_o[(size_t)i * n_out + o] = w; else w_o[(size_t)o * taps + i] = w;
    }
    left_o[o] = (int)l; cnt_o[o] = cnt; sum_o[o] = sum;
}

// ---------------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------------
__constant__ unsigned char kColormap[10][3] = { // display.rs:10-21
    {0, 0, 4},     {27, 12, 65},  {74, 12, 107}, {120, 28, 109}, {165, 44, 96},
    {207, 68, 70}, {237, 105, 37}, {251, 155, 6}, {247, 209, 61}, {252, 255, 164}};

__device__ __forceinline__ float clamp_pos(float v)
{
    // image's clamp(v, 0, f32::MAX): `if a < min {min} else if a > max {max} else {a}`
    return v < 0.0f ? 0.0f : (v > 3.4028235e38f ? 3.4028235e38f : v);
}


// display.rs:24-42 convert_grey_to_color; cm = colour map as floats in shared memory
__device__ __forceinline__ uchar4 grey_to_color(float x, const float *cm)
{
    const float position = __fmul_rn(10.0f, x);
    const float fl = floorf(position);
    if (!(fl < 9.0f)) return make_uchar4(252, 255, 164, 255); // index >= len-1 (also +inf)
    const int idx = (int)fl;
    const float ratio = __fsub_rn(position, fl);
    const float om = __fsub_rn(1.0f, ratio);
    unsigned char c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float a = cm[idx * 3 + k], b = cm[idx * 3 + 3 + k];
        const float v = roundf(__fadd_rn(__fmul_rn(ratio, b), __fmul_rn(om, a))); // half away from 0
        c[k] = (unsigned char)fminf(fmaxf(v, 0.0f), 255.0f);
    }
    return make_uchar4(c[0], c[1], c[2], 255);
}

constexpr int kRenderThreads = 256;
constexpr int kMaxPpt = 16; // output pixels per thread

__global__ void __launch_bounds__(kRenderThreads) render_kernel(const RenderLaunch L)
{
    extern __shared__ __align__(16) float rsm[];
    __shared__ float cm[30];
    const RenderTrack *__restrict__ tr = L.tracks + blockIdx.z;
    const int nwidth = tr->nwidth, nheight = tr->nheight;
    const int ox_end = tr->ox_begin + tr->ox_count; // this launch renders columns [ox_begin, ox_end)
    const int ox0 = tr->ox_begin + blockIdx.x * L.px, oy0 = blockIdx.y * L.py;
    if (ox0 >= ox_end || oy0 >= nheight) return;
    const int pxc = min(L.px, ox_end - ox0), pyc = min(L.py, nheight - oy0);
    const int tid = threadIdx.x;
    if (tid < 30) cm[tid] = (float)kColormap[tid / 3][tid % 3];

    const int *__restrict__ h_left = tr->h_left; const int *__restrict__ h_cnt = tr->h_cnt;
    const int *__restrict__ v_left = tr->v_left; const int *__restrict__ v_cnt = tr->v_cnt;
    const float *__restrict__ v_w = tr->v_w; const float *__restrict__ h_w = tr->h_w;
    const int v_taps = tr->v_taps;
    const float *__restrict__ src = tr->src;
    const int width = tr->width, height = tr->height, n_out = tr->n_out;

    // source window of this tile
    const int fl = h_left[ox0];
    const int fr = h_left[ox0 + pxc - 1] + h_cnt[ox0 + pxc - 1];
    const int yl = v_left[oy0];
    const int yr = v_left[oy0 + pyc - 1] + v_cnt[oy0 + pyc - 1];
    const int rv = yr - yl;                  // grey rows needed (<= L.rv_max)
    const int gp = L.rv_max | 1;             // odd pitch of the grey tile   [frame][row]
    const int tp = L.fc + 1;                 // pitch of the vertical result [row][frame]
    float *gs = rsm;
    float *ts = rsm + (size_t)L.fc * gp;

    float max_db = 0.0f, inv_span = 0.0f, min_db = 0.0f;
    if (L.from_db) { max_db = L.range[0]; min_db = L.range[1]; inv_span = max_db - min_db; }
    const int pad_rows = height - n_out; // grey rows above the spectrogram are 0 (display.rs:47-52)

    float acc[kMaxPpt];
#pragma unroll
    for (int i = 0; i < kMaxPpt; ++i) acc[i] = 0.0f;

    for (int c0 = fl; c0 < fr; c0 += L.fc) {
        const int nf = min(L.fc, fr - c0);
        __syncthreads(); // previous chunk fully consumed
        // ---- grey tile: gs[fx][y - yl] ---------------------------------------------------------------
        for (int e = tid; e < nf * rv; e += kRenderThreads) {
            const int fx = e / rv, yy = e - fx * rv;
            const int y = yl + yy;
            float g;
            if (L.from_db) {
                const int lf = c0 + fx - tr->frame0; // row of the (possibly time-sliced) dB array
                if (y >= pad_rows && lf >= 0 && lf < tr->src_frames) {
                    const float db = __ldg(src + (size_t)lf * n_out + (height - 1 - y));
                    g = fminf(fmaxf(__fdiv_rn(__fsub_rn(db, min_db), inv_span), 0.0f), 1.0f);
                } else g = 0.0f;
            } else {
                g = __ldg(src + (size_t)y * width + (c0 + fx));
            }
            gs[fx * gp + yy] = g;
        }
        __syncthreads();
        // ---- vertical pass: ts[oy][fx] ------------------------------------------------------------------
        for (int e = tid; e < pyc * nf; e += kRenderThreads) {
            const int oyl = e / nf, fx = e - oyl * nf;
            const int oy = oy0 + oyl;
            const int l = v_left[oy] - yl, cnt = v_cnt[oy];
            const float *wrow = v_w + (size_t)oy * v_taps;
            const float *g = gs + fx * gp + l;
            float t = 0.0f;
            for (int i = 0; i < cnt; ++i) t = fmaf(g[i], __ldg(wrow + i), t);
            ts[oyl * tp + fx] = clamp_pos(__fdiv_rn(t, tr->v_sum[oy]));
        }
        __syncthreads();
        // ---- horizontal pass, accumulated per pixel across chunks ----------------------------------------
#pragma unroll
        for (int q = 0; q < kMaxPpt; ++q) {
            const int p = tid + q * kRenderThreads;
            if (p < pxc * pyc) {
                const int oyl = p / pxc, oxl = p - oyl * pxc;
                const int ox = ox0 + oxl;
                const int l = h_left[ox], cnt = h_cnt[ox];
                const int i0 = max(l, c0), i1 = min(l + cnt, c0 + nf);
                float t = acc[q];
                for (int i = i0; i < i1; ++i)
                    t = fmaf(ts[oyl * tp + (i - c0)], __ldg(h_w + (size_t)(i - l) * nwidth + ox), t);
                acc[q] = t;
            }
        }
    }
    // ---- clamp, colour, store ------------------------------------------------------------------------------
    unsigned char *__restrict__ outp = tr->out;
#pragma unroll
    for (int q = 0; q < kMaxPpt; ++q) {
        const int p = tid + q * kRenderThreads;
        if (p < pxc * pyc) {
            const int oyl = p / pxc, oxl = p - oyl * pxc;
            const int ox = ox0 + oxl, oy = oy0 + oyl;
            const float g = clamp_pos(__fdiv_rn(acc[q], tr->h_sum[ox]));
            const uchar4 c = grey_to_color(g, cm);
            const size_t pix = (size_t)oy * tr->ox_count + (ox - tr->ox_begin);
            if (L.channels == 4) reinterpret_cast<uchar4 *>(outp)[pix] = c;
            else { outp[pix * 3] = c.x; outp[pix * 3 + 1] = c.y; outp[pix * 3 + 2] = c.z; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K3 fast path: both axes magnify or keep size (n_in / n_out < 7/6), so every output index has at most
// 8 Lanczos3 taps.  A CTA renders a 64 x 64 pixel tile:
//   A  dB -> grey tile G[row][frame] in shared memory; a warp loads 8 rows x 4 frames per request
//      (one 32-byte sector per frame) so that the transposing store is bank-conflict free,
//   B  vertical pass: lane <-> output row with its 8 normalised weights in registers; every 128-bit
//      shared load brings one source row of FOUR frames, i.e. 4 FMAs per load; clamp at 0 (the per-pass
//      clamp of image 0.23's resize) into Tm[frame][out row],
//   C  horizontal pass: lane <-> output column, 8 weights in registers, 128-bit loads bring one source
//      frame of FOUR output rows; clamp, colour map, and per output row one coalesced 128-byte store
//      of 32 RGBA pixels per warp.
// Pitches of 84 and 68 floats (odd multiples of 4) keep the 8 lanes of every 128-bit phase on
// distinct 16-byte bank groups while neighbouring lanes address neighbouring (or equal) rows.
// ---------------------------------------------------------------------------------------------------
#ifndef SGX_K3_ABATCH
#define SGX_K3_ABATCH 8
#endif
#ifndef SGX_K3_CTAS
#define SGX_K3_CTAS 4
#endif
constexpr int kABatch = SGX_K3_ABATCH; // 8-row load batches in flight per thread while the grey tile is filled
constexpr int kFpCtas = SGX_K3_CTAS;   // resident CTAs the 8-tap kernel is compiled for
constexpr int kFpTile = 64;     // output pixels per tile edge
constexpr int kFpTP = 68;       // pitch of Tm [frame][out row]  (odd multiple of 4)
// source frames / rows a tile can need with T taps per output index (n_in/n_out < (T-1)/6):
// 63 * ratio + T, rounded up to a multiple of 4 whose quarter is odd (conflict-free 128-bit phases)
__host__ __device__ constexpr int fp_cap(int taps) { return taps <= 8 ? 84 : 164; }
__host__ __device__ constexpr float fp_max_ratio(int taps) { return taps <= 8 ? 1.16f : 2.33f; }
__host__ __device__ constexpr size_t fp_smem(int tv, int th)
{
    return (size_t)(fp_cap(tv) * fp_cap(th) + fp_cap(th) * kFpTP) * sizeof(float);
}

// (packed FP32 pairs -- FFMA2 -- in the two tap loops were measured and change nothing, 3.905 vs 3.923 ms: the kernel is
// bound by the shared-memory / L1 data pipe, not by issue slots)
template <int TV, int TH, bool FROM_DB, int CH>
__global__ void __launch_bounds__(kRenderThreads, (TV <= 8 && TH <= 8) ? kFpCtas : 1) render_fast_kernel(const RenderLaunch L)
{
    constexpr int RCAP = fp_cap(TV), FCAP = fp_cap(TH), GP = FCAP; // G [row][frame], pitch = frame capacity
    extern __shared__ __align__(16) float rsm[];
     float *G = rsm;
    float *Tm = rsm + RCAP * GP;         // [frame][out row]
    const RenderTrack *__restrict__ tr = L.tracks + blockIdx.z;
    const int nwidth = tr->nwidth, nheight = tr->nheight;
    const int ox_begin = tr->ox_begin, ox_count = tr->ox_count, frame0 = tr->frame0, src_frames = tr->src_frames;
    const int ox_end = ox_begin + ox_count; // this launch renders columns [ox_begin, ox_end)
    const int ox0 = ox_begin + blockIdx.x * kFpTile, oy0 = blockIdx.y * kFpTile;
    if (ox0 >= ox_end || oy0 >= nheight) return;
    const int pxc = min(kFpTile, ox_end - ox0), pyc = min(kFpTile, nheight - oy0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int *__restrict__ h_left = tr->h_left;
    const int *__restrict__ v_left = tr->v_left;
    const float *__restrict__ src = tr->src;
    const int width = tr->width, height = tr->height, n_out = tr->n_out;

    const int fl = __ldg(h_left + ox0);
    const int nfr = min(__ldg(h_left + ox0 + pxc - 1) + TH - fl, FCAP);
    const int nfq = (nfr + 3) >> 2;                 // frame quads
    const int yl = __ldg(v_left + oy0);
    const int nrow = min(__ldg(v_left + oy0 + pyc - 1) + TV - yl, RCAP);

    // weights of phases B and C first: their global loads complete while phase A runs
    const int oyl_b = (warp & 1) * 32 + lane;
    const int oy_b = oy0 + min(oyl_b, pyc - 1);
    const int oxl_c = (warp & 1) * 32 + lane;
    const int ox_c = ox0 + min(oxl_c, pxc - 1);
    float wv[TV], wh[TH];
    {
        // v_w rows are 16-byte aligned (row width is a multiple of 4 floats, >= 16)
        const float4 *__restrict__ wrow = reinterpret_cast<const float4 *>(tr->v_w + (size_t)oy_b * tr->v_taps);
#pragma unroll
        for (int i = 0; i < TV / 4; ++i) {
            const float4 q = __ldg(wrow + i);
            wv[4 * i] = q.x; wv[4 * i + 1] = q.y; wv[4 * i + 2] = q.z; wv[4 * i + 3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < TH; ++i) wh[i] = __ldg(tr->h_w + (size_t)i * nwidth + ox_c);
    }
    const float vsum = __ldg(tr->v_sum + oy_b), hsum = __ldg(tr->h_sum + ox_c);
    const int voff = __ldg(v_left + oy_b) - yl, hoff = __ldg(h_left + ox_c) - fl;

    // ---- A: grey tile ------------------------------------------------------------------------------
    float min_db = 0.0f, inv_span = 0.0f;
    if (FROM_DB) { min_db = L.range[1]; inv_span = __frcp_rn(L.range[0] - L.range[1]); }
    {
        // rows of the tile that hold data: grey rows [pad_rows, height) (display.rs:47-52), in tile coordinates
        const int pad_rows = FROM_DB ? height - n_out : 0;
        const int yy_lo = max(pad_rows - yl, 0), yy_hi = min(height - yl, nrow);
        const int fsub = lane & 3, rsub = lane >> 2;
        for (int fq = warp; fq < nfq; fq += kRenderThreads / 32) {
            const int fx = fq * 4 + fsub;
            const int f = fl + fx;
            const int lf = FROM_DB ? f - frame0 : f; // row of the (possibly time-sliced) dB array
            const bool fok = f < width && lf >= 0 && (!FROM_DB || lf < src_frames);
            // FROM_DB: element (frame f, grey row y) is dB[lf][height-1-y]; else grey[y][f]
            const float *__restrict__ p0 = FROM_DB ? src + (size_t)(fok ? lf : 0) * n_out + (height - 1 - yl - rsub)
                                                   : src + (size_t)(yl + rsub) * width + (fok ? f : 0);
            float *g0 = G + rsub * GP + fx;
            for (int j0 = 0; j0 < nrow; j0 += 8 * kABatch) {
                // kABatch 8-row batches per trip, all loads issued before the first use
                float v[kABatch];
#pragma unroll
                for (int j = 0; j < kABatch; ++j) {
                    const int yy = j0 + j * 8 + rsub;
                    const bool ok = fok && yy >= yy_lo && yy < yy_hi;
                    v[j] = FROM_DB ? -INFINITY : 0.0f; // -inf -> grey 0 after the saturate
                    if (ok) v[j] = FROM_DB ? __ldg(p0 - (j0 + j * 8)) : __ldg(p0 + (size_t)(j0 + j * 8) * width);
                }
#pragma unroll
                for (int j = 0; j < kABatch; ++j) {
                    const int yy = j0 + j * 8 + rsub;
                    const float g = FROM_DB ? __saturatef((v[j] - min_db) * inv_span) : v[j];
                    if (yy < RCAP) g0[(j0 + j * 8) * GP] = g;
                }
            }
        }
    }
    __syncthreads();

    // ---- B: vertical pass ----------------------------------------------------------------------------
    {
        const int oyl = oyl_b;
        const float rs = __frcp_rn(vsum);
#pragma unroll
        for (int i = 0; i < TV; ++i) wv[i] *= rs;
        const float *g = G + voff * GP;
        if (oyl < pyc) {
            for (int fq = warp >> 1; fq < nfq; fq += kRenderThreads / 64) {
                float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
#pragma unroll
                for (int i = 0; i < TV; ++i) {
                    const float4 v = *reinterpret_cast<const float4 *>(g + i * GP + fq * 4);
                    t0 = fmaf(v.x, wv[i], t0); t1 = fmaf(v.y, wv[i], t1);
                    t2 = fmaf(v.z, wv[i], t2); t3 = fmaf(v.w, wv[i], t3);
                }
                float *t_out = Tm + (fq * 4) * kFpTP + oyl;
                t_out[0] = clamp_fin(t0); t_out[kFpTP] = clamp_fin(t1);
                t_out[2 * kFpTP] = clamp_fin(t2); t_out[3 * kFpTP] = clamp_fin(t3);
            }
        }
    }
    __syncthreads();

    // ---- C: horizontal pass, colour, store -----------------------------------------------------------------
    {
        const int oxl = oxl_c, ox = ox_c;
        const float rs = __frcp_rn(hsum);
#pragma unroll
        for (int i = 0; i < TH; ++i) wh[i] *= rs;
        const float *t_in = Tm + hoff * kFpTP;
        unsigned char *__restrict__ outp = tr->out;
        if (oxl < pxc) {
            const int nrq = (pyc + 3) >> 2;
            for (int rq = warp >> 1; rq < nrq; rq += kRenderThreads / 64) {
                float t[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int i = 0; i < TH; ++i) {
                    const float4 v = *reinterpret_cast<const float4 *>(t_in + i * kFpTP + rq * 4);
                    t[0] = fmaf(v.x, wh[i], t[0]); t[1] = fmaf(v.y, wh[i], t[1]);
                    t[2] = fmaf(v.z, wh[i], t[2]); t[3] = fmaf(v.w, wh[i], t[3]);
                }
                const int opitch = ox_count;
                size_t pix = (size_t)(oy0 + rq * 4) * opitch + (ox - ox_begin);
#pragma unroll
                for (int j = 0; j < 4; ++j, pix += opitch) {
                    if (rq * 4 + j < pyc) {
                        const unsigned c = grey_to_rgba_const(clamp_fin(t[j]));
                        if (CH == 4) reinterpret_cast<unsigned *>(outp)[pix] = c;
                        else { outp[pix * 3] = (unsigned char)c; outp[pix * 3 + 1] = (unsigned char)(c >> 8); outp[pix * 3 + 2] = (unsigned char)(c >> 16); }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K3 wide path: any number of taps per axis, as long as the source window of one tile fits in shared
// memory (strong minification on either axis: short FFTs at 100 px/s, long FFTs at 500 rows).  Same
// three phases and layouts as the fast path, with the tap loops running over the tables:
//   B  a warp owns output rows and its lanes run along the frames (up to 4 frames per lane), so the
//      weight of a tap is one uniform 128-bit load per 4 taps and the grey loads are conflict-free
//      whatever the vertical ratio is,
//   C  lane <-> output column, weights read tap-major (coalesced), every 128-bit shared load brings one
//      source frame of four output rows, up to 4 row quads per lane share a weight load.
// Taps are summed in ascending order into one accumulator per output and divided by the weight sum
// afterwards, as the reference's resize does; the rows beyond an index's own tap count hold zero weights.
// ---------------------------------------------------------------------------------------------------
constexpr int kWideMaxSmem = 200 * 1024;
__host__ __device__ constexpr int wide_pitch(int n) // multiple of 4 whose quarter is odd, >= n
{
    int q = (n + 3) / 4;
    if ((q & 1) == 0) ++q;
    return q * 4;
}

template <bool FROM_DB, int CH>
__global__ void __launch_bounds__(kRenderThreads, 4) render_wide_kernel(const RenderLaunch L)
{
    const int RCAP = L.rv_max, GP = L.fc; // G [row][frame], pitch = frame capacity
    extern __shared__ __align__(16) float rsm[];
     float *G = rsm;
    float *Tm = rsm + (size_t)RCAP * GP;  // [frame][out row], pitch TP
    const int TP = L.py + 4;              // 68 / 36 / 20: multiples of 4 with an odd quarter
    const RenderTrack *__restrict__ tr = L.tracks + blockIdx.z;
    const int nwidth = tr->nwidth, nheight = tr->nheight;
    const int ox_begin = tr->ox_begin, ox_count = tr->ox_count, frame0 = tr->frame0, src_frames = tr->src_frames;
    const int ox_end = ox_begin + ox_count; // this launch renders columns [ox_begin, ox_end)
    const int ox0 = ox_begin + blockIdx.x * L.px, oy0 = blockIdx.y * L.py;
    if (ox0 >= ox_end || oy0 >= nheight) return;
    const int pxc = min(L.px, ox_end - ox0), pyc = min(L.py, nheight - oy0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int *__restrict__ h_left = tr->h_left;
    const int *__restrict__ v_left = tr->v_left;
    const float *__restrict__ src = tr->src;
    const int width = tr->width, height = tr->height, n_out = tr->n_out;
    const int v_taps = tr->v_taps;
    // taps an index can really have (the tables are at least 16 wide for the fast path; the rest is zeros)
    const int h_need = min(tr->h_taps, lanczos3_taps_bound(width, nwidth));
    const int v_need = min(v_taps, (lanczos3_taps_bound(height, nheight) + 3) & ~3);

    const int fl = __ldg(h_left + ox0);
    const int nfr = min(__ldg(h_left + ox0 + pxc - 1) + h_need - fl, GP);
    const int nfq = (nfr + 3) >> 2;                 // frame quads
    const int yl = __ldg(v_left + oy0);
    const int nrow = min(__ldg(v_left + oy0 + pyc - 1) + v_need - yl, RCAP);

    // ---- A: grey tile ---------------------------------------------------------------------------------
    float min_db = 0.0f, inv_span = 0.0f;
    if (FROM_DB) { min_db = L.range[1]; inv_span = __frcp_rn(L.range[0] - L.range[1]); }
    const int pad_rows = FROM_DB ? height - n_out : 0;
    const int yy_lo = max(pad_rows - yl, 0), yy_hi = min(height - yl, nrow);
    if (FROM_DB && nfq < kRenderThreads / 32) {
        // few frames, many rows (strong vertical minification): a warp reads 32 consecutive rows of ONE frame
        // -- they are contiguous in the dB array -- so that every warp has loads in flight
        const int nblk = (nrow + 32 * kABatch - 1) / (32 * kABatch);
        for (int item = warp; item < nfq * 4 * nblk; item += kRenderThreads / 32) {
            const int fx = item / nblk, j0 = (item - fx * nblk) * 32 * kABatch;
            const int f = fl + fx, lf = f - frame0;
            const bool fok = f < width && lf >= 0 && lf < src_frames;
            const float *__restrict__ p0 = src + (size_t)(fok ? lf : 0) * n_out + (height - 1 - yl - lane);
            float v[kABatch];
#pragma unroll
            for (int j = 0; j < kABatch; ++j) {
                const int yy = j0 + j * 32 + lane;
                v[j] = -INFINITY; // -inf -> grey 0 after the saturate
                if (fok && yy >= yy_lo && yy < yy_hi) v[j] = __ldg(p0 - (j0 + j * 32));
            }
#pragma unroll
            for (int j = 0; j < kABatch; ++j) {
                const int yy = j0 + j * 32 + lane;
                if (yy < nrow) G[yy * GP + fx] = __saturatef((v[j] - min_db) * inv_span);
            }
        }
    } else {
        // as in the fast path: a warp loads 8 rows x 4 frames per request
        const int fsub = lane & 3, rsub = lane >> 2;
        for (int fq = warp; fq < nfq; fq += kRenderThreads / 32) {
            const int fx = fq * 4 + fsub;
            const int f = fl + fx;
            const int lf = FROM_DB ? f - frame0 : f; // row of the (possibly time-sliced) dB array
            const bool fok = f < width && lf >= 0 && (!FROM_DB || lf < src_frames);
            const float *__restrict__ p0 = FROM_DB ? src + (size_t)(fok ? lf : 0) * n_out + (height - 1 - yl - rsub)
                                                   : src + (size_t)(yl + rsub) * width + (fok ? f : 0);
            float *g0 = G + rsub * GP + fx;
            for (int j0 = 0; j0 < nrow; j0 += 8 * kABatch) {
                float v[kABatch];
#pragma unroll
                for (int j = 0; j < kABatch; ++j) {
                    const int yy = j0 + j * 8 + rsub;
                    const bool ok = fok && yy >= yy_lo && yy < yy_hi;
                    v[j] = FROM_DB ? -INFINITY : 0.0f; // -inf -> grey 0 after the saturate
                    if (ok) v[j] = FROM_DB ? __ldg(p0 - (j0 + j * 8)) : __ldg(p0 + (size_t)(j0 + j * 8) * width);
                }
#pragma unroll
                for (int j = 0; j < kABatch; ++j) {
                    const int yy = j0 + j * 8 + rsub;
                    const float g = FROM_DB ? __saturatef((v[j] - min_db) * inv_span) : v[j];
                    if (yy < RCAP) g0[(j0 + j * 8) * GP] = g;
                }
            }
        }
    }
    __syncthreads();

    // ---- B: vertical pass: warp <-> output row, lanes <-> frames ----------------------------------------
    {
        const int *__restrict__ v_cnt = tr->v_cnt;
        const float *__restrict__ v_sum = tr->v_sum;
        const int nfx = nfq * 4;
        auto rows = [&](auto na_tag) {
            constexpr int NA = decltype(na_tag)::value; // frames per lane
            for (int oyl = warp; oyl < pyc; oyl += kRenderThreads / 32) {
                const int oy = oy0 + oyl;
                const int voff = __ldg(v_left + oy) - yl;
                const int cnt4 = (__ldg(v_cnt + oy) + 3) >> 2;
                const float rvs = __frcp_rn(__ldg(v_sum + oy));
                const float4 *__restrict__ wrow = reinterpret_cast<const float4 *>(tr->v_w + (size_t)oy * v_taps);
                for (int fb = 0; fb < nfx; fb += 32 * NA) {
                    const float *g[NA];
                    float t[NA];
#pragma unroll
                    for (int a = 0; a < NA; ++a) { g[a] = G + voff * GP + min(fb + a * 32 + lane, nfx - 1); t[a] = 0.0f; }
                    for (int i4 = 0; i4 < cnt4; ++i4) {
                        const float4 w = __ldg(wrow + i4);
#pragma unroll
                        for (int a = 0; a < NA; ++a) {
                            const float *gp = g[a] + (i4 * 4) * GP;
                            const float g0 = gp[0], g1 = gp[GP], g2 = gp[2 * GP], g3 = gp[3 * GP];
                            t[a] = fmaf(g0, w.x, t[a]); t[a] = fmaf(g1, w.y, t[a]);
                            t[a] = fmaf(g2, w.z, t[a]); t[a] = fmaf(g3, w.w, t[a]);
                        }
                    }
#pragma unroll
                    for (int a = 0; a < NA; ++a) {
                        const int fx = fb + a * 32 + lane;
                        if (fx < nfx) Tm[fx * TP + oyl] = clamp_fin(t[a] * rvs);
                    }
                }
            }
        };
        if (nfx <= 32) rows(std::integral_constant<int, 1>{});
        else if (nfx <= 64) rows(std::integral_constant<int, 2>{});
        else rows(std::integral_constant<int, 4>{});
    }
    __syncthreads();

    // ---- C: horizontal pass, colour, store -----------------------------------------------------------------
    for (int cb = 0; cb < pxc; cb += kFpTile) { // 64-column blocks of the tile
        const int oxl = cb + (warp & 1) * 32 + lane;
        const int ox = ox0 + min(oxl, pxc - 1);
        const int hoff = __ldg(h_left + ox) - fl;
        const int cnt = __reduce_max_sync(0xffffffffu, __ldg(tr->h_cnt + ox)); // zero weights beyond a lane's own count
        const float rhs = __frcp_rn(__ldg(tr->h_sum + ox));
        const float *__restrict__ wcol = tr->h_w + ox; // tap-major: tap i of column ox at wcol[i * nwidth]
        const float *t_in = Tm + hoff * TP;
        unsigned char *__restrict__ outp = tr->out;
        const int nrq = (pyc + 3) >> 2;
        // this lane's row quads: (warp >> 1) + 4 k, k < 4 (a tile is at most 64 rows = 16 quads)
        float t[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) t[k][j] = 0.0f;
        const int rq0 = warp >> 1;
        auto tap = [&](int i, float w) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int rq = rq0 + 4 * k;
                if (rq < nrq) {
                    const float4 v = *reinterpret_cast<const float4 *>(t_in + i * TP + rq * 4);
                    t[k][0] = fmaf(v.x, w, t[k][0]); t[k][1] = fmaf(v.y, w, t[k][1]);
                    t[k][2] = fmaf(v.z, w, t[k][2]); t[k][3] = fmaf(v.w, w, t[k][3]);
                }
            }
        };
        int i = 0;
        for (; i + 4 <= cnt; i += 4) { // four weight loads in flight before the first use
            const float w0 = __ldg(wcol + (size_t)i * nwidth), w1 = __ldg(wcol + (size_t)(i + 1) * nwidth);
            const float w2 = __ldg(wcol + (size_t)(i + 2) * nwidth), w3 = __ldg(wcol + (size_t)(i + 3) * nwidth);
            tap(i, w0); tap(i + 1, w1); tap(i + 2, w2); tap(i + 3, w3);
        }
        for (; i < cnt; ++i) tap(i, __ldg(wcol + (size_t)i * nwidth));
        if (oxl < pxc) {
            const int opitch = ox_count;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int rq = rq0 + 4 * k;
                if (rq >= nrq) continue;
                size_t pix = (size_t)(oy0 + rq * 4) * opitch + (ox - ox_begin);
#pragma unroll
                for (int j = 0; j < 4; ++j, pix += opitch) {
                    if (rq * 4 + j < pyc) {
                        const unsigned c = grey_to_rgba_const(clamp_fin(t[k][j] * rhs));
                        if (CH == 4) reinterpret_cast<unsigned *>(outp)[pix] = c;
                        else { outp[pix * 3] = (unsigned char)c; outp[pix * 3 + 1] = (unsigned char)(c >> 8); outp[pix * 3 + 2] = (unsigned char)(c >> 16); }
                    }
                }
            }
        }
    }
}

// display.rs:44-54 as a stand-alone stage (surface 2)
__global__ void spec_to_grey_kernel(const float *__restrict__ spec, int T, int n_out, int height,
                                    float max_db, float min_db, float *__restrict__ grey)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)T * height) return;
    const int y = (int)(i / T), x = (int)(i - (size_t)y * T);
    float g = 0.0f;
    if (y >= height - n_out) {
        const float db = spec[(size_t)x * n_out + (height - 1 - y)];
        g = fminf(fmaxf(__fdiv_rn(__fsub_rn(db, min_db), __fsub_rn(max_db, min_db)), 0.0f), 1.0f);
    }
    grey[i] = g;
}

// decibel.rs:33-88 as a stand-alone stage
__global__ void amp_to_db_kernel(float *x, size_t n, int *bad)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    if (!(v >= 0.0f)) { atomicExch(bad, 1); return; } // decibel.rs:34 assert
    x[i] = v > 1e-18f ? __fmul_rn(20.0f, log10f(v)) : -360.0f;
}

// display.rs:63-115 wav_to_image: one thread per pixel column
__global__ void wav_image_kernel(const void *pcm, int fmt, int ch, long long n_in, int nwidth,
                                 int nheight, float amp_min, float amp_max,
                                 unsigned char *__restrict__ out, int *err)
{
    const int i_px = blockIdx.x * blockDim.x + threadIdx.x;
    if (i_px >= nwidth) return;
    const PcmFormat f = (PcmFormat)fmt;
    auto sample = [&](long long i) -> float {
        float s = 0.0f;
        if (f == PCM_F32) { const float *p = (const float *)pcm + i * ch; for (int c = 0; c < ch; ++c) s += p[c]; }
        else { const short *p = (const short *)pcm + i * ch; for (int c = 0; c < ch; ++c) s += (float)p[c] * (1.0f / 32768.0f); }
        return s;
    };
    const float spp = __fdiv_rn((float)n_in, (float)nwidth); // samples_per_px, display.rs:75
    long long n = n_in;
    long long factor = 1;
    const bool up = spp < 1.0f;
    if (up) { factor = (long long)ceilf(__fdiv_rn(1.0f, spp)); n = factor * n_in; }
    auto wav_at = [&](long long i) -> float {
        if (!up) return sample(i);
        // display.rs:78-87 linear interpolation by `factor`
        const long long i0 = i / factor;
        const float b = (i0 + 1 < n_in) ? sample(i0 + 1) : 0.0f;
        const float fr = __fdiv_rn((float)(i % factor), (float)factor);
        return __fadd_rn(__fmul_rn(b, fr), __fmul_rn(sample(i0), __fsub_rn(1.0f, fr)));
    };
    const float fs = fmaxf(roundf(__fmul_rn(__fsub_rn((float)i_px, 1.5f), spp)), 0.0f);
    const long long i_start = (long long)fs;
    const float fe = roundf(__fmul_rn(__fadd_rn((float)i_px, 1.5f), spp));
    long long i_end = fe <= 0.0f ? 0 : (long long)fe;
    if (i_end > n) i_end = n;
    if (i_start >= i_end) { atomicExch(err, 1); return; } // reference: max() of empty slice panics
    float mx = wav_at(i_start), mn = mx;
    for (long long i = i_start + 1; i < i_end; ++i) { const float v = wav_at(i); mx = fmaxf(mx, v); mn = fminf(mn, v); }
    const float span = __fsub_rn(amp_max, amp_min);
    long long top = (long long)roundf(__fdiv_rn(__fmul_rn(__fsub_rn(amp_max, mx), (float)nheight), span));
    long long bottom = (long long)roundf(__fdiv_rn(__fmul_rn(__fsub_rn(amp_max, mn), (float)nheight), span));
    if (bottom - top < 3) { // display.rs:100-105
        const float d = (float)(3 - bottom + top);
        const long long pad_bottom = (long long)ceilf(d / 2.0f), pad_top = (long long)floorf(d / 2.0f);
        top -= pad_top; bottom += pad_bottom;
    }
    if (top < 0) top = 0;
    if (bottom > nheight) bottom = nheight;
    if (bottom + 1 > nheight) bottom = nheight - 1; // ndarray slice top..bottom+1 must stay inside
    for (long long y = top; y <= bottom; ++y)
        reinterpret_cast<uchar4 *>(out)[(size_t)y * nwidth + i_px] = make_uchar4(200, 21, 103, 255);
}

} // namespace

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
cudaError_t launch_range_init(unsigned *slots, int n_slots, cudaStream_t s)
{
    if (n_slots <= 0) return cudaSuccess;
    range_init_kernel<<<(n_slots + 255) / 256, 256, 0, s>>>(slots, n_slots);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_range_reset(const StftTrack *descs, int n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    range_reset_kernel<<<(n + 127) / 128, 128, 0, s>>>(descs, n);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_range_reduce(const unsigned *slots, int n_slots, float *out, float max_sr, float max_sec, float db_range,
                                float *commit_state, cudaStream_t s)
{
    range_reduce_kernel<<<1, 256, 0, s>>>(slots, n_slots, out, max_sr, max_sec, db_range, commit_state);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_range_commit(const float *max_negmin, float db_range, float *state,
                                cudaStream_t s)
{
    range_commit_kernel<<<1, 1, 0, s>>>(max_negmin, db_range, state);
    count_launch();
    return cudaGetLastError();
}

uint32_t lanczos3_max_taps(uint32_t n_in, uint32_t n_out) { return (uint32_t)lanczos3_taps_bound((int)n_in, (int)n_out); }

cudaError_t launch_build_axis_table(int n_in, int n_out, int taps, bool tap_major, int *left,
                                    int *cnt, float *sum, float *w, cudaStream_t s)
{
    build_axis_table_kernel<<<(n_out + 127) / 128, 128, 0, s>>>(n_in, n_out, taps, tap_major ? 1 : 0,
                                                               left, cnt, sum, w);
    count_launch();
    return cudaGetLastError();
}

RenderTiling plan_render_tiles(int width, int height, int nwidth, int nheight, bool from_db)
{
    // fast path: at most 8 / 16 taps per output index on each axis (same f32 ratio the tap tables use)
    const float rhf = (float)width / (float)nwidth, rvf = (float)height / (float)nheight;
    // tensor-core path (render_tc_kernel.cu): both axes in the 8-tap class, dB input, and a tile whose operands fit in
    // shared memory -- 128 output rows need kv grey rows, nx output columns need nf source frames
    // Measured on B200 (C5): 8.3 ms per step against 3.7 ms of the FP32 path -- see the header of render_tc_kernel.cu --
    // so it is an evaluated alternative, selected with SGX_K3_TC=1.
    static const bool tc_on = getenv("SGX_K3_TC") && atoi(getenv("SGX_K3_TC")) == 1;
    if (from_db && tc_on && rhf < fp_max_ratio(8) && rvf < fp_max_ratio(8) && nheight >= 64 && nwidth >= 64) {
        const float sv = rvf < 1.0f ? 1.0f : rvf, sh = rhf < 1.0f ? 1.0f : rhf;
        const int kv = ((int)std::ceil(127.0 * rvf + 6.0 * sv) + 3 + 7) & ~7;
        for (int nx : {64, 48, 32}) {
            const int nf = ((int)std::ceil((nx - 1) * (double)rhf + 6.0 * sh) + 3 + 15) & ~15;
            const size_t smem = render_tc_smem(kv, nf, nx);
            if (nf <= 96 && smem + 1024 <= (size_t)227 * 1024) return RenderTiling{nx, 128, nf, kv, smem, 2};
        }
    }
    // sliding-window path (render_slide_kernel.cu): both axes in the 8-tap class; SGX_K3_SLIDE=0 keeps render_fast_kernel
    static const bool slide_on = !(getenv("SGX_K3_SLIDE") && atoi(getenv("SGX_K3_SLIDE")) == 0);
    if (slide_on) {
        RenderTiling t{};
        if (render_slide_plan(width, height, nwidth, nheight, &t)) return t;
    }
    if (rhf < fp_max_ratio(16) && rvf < fp_max_ratio(16)) {
        const int th = rhf < fp_max_ratio(8) ? 8 : 16, tv = rvf < fp_max_ratio(8) ? 8 : 16;
        RenderTiling t{};
        t.px = kFpTile; t.py = kFpTile; t.fc = fp_cap(th); t.rv_max = fp_cap(tv); t.smem_bytes = fp_smem(tv, th);
        t.fast = tv * 100 + th;
        return t;
    }
    // wide path: runtime tap counts, the whole source window of a px x py tile in shared memory.  Under
    // horizontal magnification a tile is made wider, so that the frames it owns outnumber the Lanczos halo
    // (otherwise most of the vertical pass would be repeated by the neighbouring tiles).
    {
        const int vt = (lanczos3_taps_bound(height, nheight) + 3) & ~3;
        const int ht = lanczos3_taps_bound(width, nwidth);
        int px = rhf <= 0.25f ? 4 * kFpTile : (rhf <= 0.5f ? 2 * kFpTile : kFpTile);
        int py_max = kFpTile;
        if (const char *e = getenv("SGX_K3_WIDE_PX")) px = atoi(e);   // tuning: 64 / 128 / 256
        if (const char *e = getenv("SGX_K3_WIDE_PY")) py_max = atoi(e); // tuning: 64 / 32 / 16
        const int gp = wide_pitch((int)std::ceil((px - 1) * (double)rhf) + ht + 2);
        // tile height: fewer rows per tile mean more redundant source rows (the Lanczos halo) but more resident
        // CTAs to hide the load latency of phase A; weights fitted to the C4 sweep on B200
        RenderTiling pick{};
        double best = 1e300;
        for (int py = py_max; py >= 16; py >>= 1) {
            const int rcap = ((int)std::ceil((py - 1) * (double)rvf) + vt + 2 + 7) & ~7;
            const size_t smem = ((size_t)rcap * gp + (size_t)gp * (py + 4)) * sizeof(float);
            if (smem > (size_t)kWideMaxSmem) continue;
            const int ctas = std::min(4, (int)((size_t)(227 * 1024) / (smem + 1024)));
            const double halo = (double)rcap / std::max(1.0, py * (double)rvf);
            const double occ = ctas >= 4 ? 1.0 : (ctas == 3 ? 1.05 : (ctas == 2 ? 1.25 : 1.6));
            const double cost = (0.5 + 0.5 * halo) * occ;
            if (cost < best) { best = cost; pick = RenderTiling{px, py, gp, rcap, smem, 1}; }
        }
        static const bool no_wide = getenv("SGX_K3_NOWIDE") && atoi(getenv("SGX_K3_NOWIDE")) == 1;
        if (pick.fast == 1 && !no_wide) return pick;
    }
    // frames a tile of px output columns needs, rows a tile of py output rows needs
    const double rh = (double)width / nwidth, rv = (double)height / nheight;
    const double sh = 3.0 * (rh < 1 ? 1 : rh), sv = 3.0 * (rv < 1 ? 1 : rv);
    auto frames_for = [&](int px) { return (int)(px * rh + 2 * sh) + 4; };
    auto rows_for = [&](int py) { return (int)(py * rv + 2 * sv) + 4; };
    const size_t budget = 72 * 1024;
    RenderTiling best{};
    double best_cost = 1e300;
    for (int px = 128; px >= 1; px >>= 1) {
        if (px > 1 && px / 2 >= nwidth) continue;
        for (int py = 256; py >= 1; py >>= 1) {
            if (py > 1 && py / 2 >= nheight) continue;
            if ((long)px * py > (long)kRenderThreads * kMaxPpt) continue;
            int fc = frames_for(px);
            const int rvm = rows_for(py);
            if (fc > 96) fc = 64; // chunked accumulation for strong horizontal minification
            const size_t smem = ((size_t)fc * (rvm | 1) + (size_t)py * (fc + 1)) * sizeof(float);
            if (smem > budget) continue;
            // cost per output pixel: redundant vertical work + redundant loads + underfilled threads
            const double halo_h = (double)frames_for(px) / (px * rh), halo_v = (double)rows_for(py) / (py * rv);
            double cost = halo_h * (1.0 + 0.5 * halo_v);
            if (px * py < kRenderThreads * 4) cost *= 1.0 + (double)(kRenderThreads * 4) / (px * py) * 0.25;
            if (px < 32) cost *= 1.0 + (32.0 / px - 1.0) * 0.5; // sub-line pixel stores
            if (cost < best_cost) { best_cost = cost; best = RenderTiling{px, py, fc, rvm, smem, 0}; }
        }
    }
    if (best.px == 0) { // nothing fits: one pixel per tile, minimum chunk
        const int rvm = rows_for(1);
        int fc = 16;
        while (fc > 1 && ((size_t)fc * (rvm | 1) + (fc + 1)) * sizeof(float) > 200 * 1024) fc >>= 1;
        best = RenderTiling{1, 1, fc, rvm, ((size_t)fc * (rvm | 1) + (fc + 1)) * sizeof(float), 0};
    }
    return best;
}

cudaError_t launch_render(const RenderLaunch &L, int max_nwidth, int max_nheight, size_t smem_bytes,
                          int fast, cudaStream_t s)
{
    if (L.n_tracks <= 0 || max_nwidth <= 0 || max_nheight <= 0) return cudaSuccess;
    if (fast == 2) { // tensor-core path
        RenderLaunch T = L;
        T.rv_cols = max_nwidth; T.rv_rows = (max_nheight + 127) / 128;
        return launch_render_tc(T, smem_bytes, s);
    }
    if (fast == 3) return launch_render_slide(L, max_nwidth, max_nheight, smem_bytes, s);
    if (fast == 1) { // wide path: tile L.px x L.py, capacities in L.fc / L.rv_max
        dim3 grid((max_nwidth + L.px - 1) / L.px, (max_nheight + L.py - 1) / L.py, L.n_tracks);
        cudaError_t err = cudaErrorInvalidValue;
#define SGX_WIDE(DB, CHN)                                                                                     \
        if ((L.from_db != 0) == DB && L.channels == CHN) {                                                    \
            auto kern = render_wide_kernel<DB, CHN>;                                                          \
            cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem_bytes);           \
            if (e != cudaSuccess) return e;                                                                   \
            kern<<<grid, kRenderThreads, smem_bytes, s>>>(L);                                                 \
            err = cudaSuccess;                                                                                \
        }
        SGX_WIDE(true, 4) SGX_WIDE(true, 3) SGX_WIDE(false, 4) SGX_WIDE(false, 3)
#undef SGX_WIDE
        if (err != cudaSuccess) return err;
        count_launch();
        return cudaGetLastError();
    }
    if (fast) {
        dim3 grid((max_nwidth + kFpTile - 1) / kFpTile, (max_nheight + kFpTile - 1) / kFpTile, L.n_tracks);
        const int tv = fast / 100, th = fast % 100;
        cudaError_t err = cudaErrorInvalidValue;
#define SGX_FAST(TV, TH, DB, CHN)                                                                             \
        if (tv == TV && th == TH && (L.from_db != 0) == DB && L.channels == CHN) {                           \
            auto kern = render_fast_kernel<TV, TH, DB, CHN>;                                                  \
            cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), fp_smem(TV, TH));              \
            if (e != cudaSuccess) return e;                                                                   \
            kern<<<grid, kRenderThreads, fp_smem(TV, TH), s>>>(L);                                            \
            err = cudaSuccess;                                                                                \
        }
#define SGX_FAST4(TV, TH) SGX_FAST(TV, TH, true, 4) SGX_FAST(TV, TH, true, 3) SGX_FAST(TV, TH, false, 4) SGX_FAST(TV, TH, false, 3)
        SGX_FAST4(8, 8) SGX_FAST4(8, 16) SGX_FAST4(16, 8) SGX_FAST4(16, 16)
#undef SGX_FAST4
#undef SGX_FAST
        if (err != cudaSuccess) return err;
        count_launch();
        return cudaGetLastError();
    }
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(render_kernel), smem_bytes);
    if (e != cudaSuccess) return e;
    dim3 grid((max_nwidth + L.px - 1) / L.px, (max_nheight + L.py - 1) / L.py, L.n_tracks);
    render_kernel<<<grid, kRenderThreads, smem_bytes, s>>>(L);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_spec_to_grey(const float *spec, int n_frames, int n_out, int height, float max_db,
                                float min_db, float *grey, cudaStream_t s)
{
    const size_t n = (size_t)n_frames * height;
    if (n == 0) return cudaSuccess;
    spec_to_grey_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(spec, n_frames, n_out, height,
                                                                     max_db, min_db, grey);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_amp_to_db(float *x, size_t n, int *bad_flag, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    amp_to_db_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n, bad_flag);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_wav_image(const void *pcm, int fmt, int ch, long long n, int nwidth, int nheight,
                             float amp_min, float amp_max, unsigned char *out, int *err_flag,
                             cudaStream_t s)
{
    if (nwidth <= 0 || nheight <= 0) return cudaSuccess;
    wav_image_kernel<<<(nwidth + 127) / 128, 128, 0, s>>>(pcm, fmt, ch, n, nwidth, nheight, amp_min,
                                                         amp_max, out, err_flag);
    count_launch();
    return cudaGetLastError();
}

} // namespace sgx
