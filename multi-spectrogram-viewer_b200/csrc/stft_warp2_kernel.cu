// stft_warp2_kernel.cu -- K1 for n_fft = 2048 (h = 1024 = 32 x 32): one WARP transforms two frames at a time.
//
// Same path as stft_kernel.cu (channel sum lib.rs:42 -> reflect-padded framing lib.rs:412-433 -> window and centred
// zero-pad lib.rs:377-384 -> real FFT as an h-point complex FFT + split realfft.rs:105-159 -> |X| lib.rs:124 -> banded
// mel lib.rs:131 -> dB decibel.rs:33-88 -> per-track extrema lib.rs:197-200), other decomposition:
//   * the block kernel spreads a frame over 128 threads (8 points each) and needs three shared-memory exchanges plus
//     the partner loads of the split -- 738 shared-memory wavefronts per frame, the pipe that bounds it (ncu);
//   * here lane m2 of a warp holds the 32 points z[32 m1 + m2] of TWO frames as packed FP32 pairs (64 register
//     pairs), runs a 32-point DFT in registers (FADD2 / FMUL2 / FFMA2: one instruction serves both frames), applies
//     W_1024^(m2 k1), transposes ONCE through a private 8.25 KB plane, runs the second 32-point DFT and meets the
//     conjugate partner of bin k1 + 32 k2 (lane 32 - k1, register 31 - k2) by warp shuffles: every lane finishes
//     the bin pairs (k, h - k) of its 16 lowest bins, so the split costs 32 shuffles per frame and no loads.
// No block barrier sits on the frame path; the eight warps of a CTA only meet at the PCM tile, which is staged by
// TMA exactly as in the block kernel (bulk copy + mbarrier, next tile prefetched by the last warp to consume this one).
//
// MEASURED (B200, C5, K1 ms per step; profiles/r02_k1_w2_*): 535 shared-memory wavefronts per frame instead of 715 and
// 1,670 instead of 2,000 instructions, yet 6.42 ms against the block kernel's 6.32: 128 data registers per thread leave
// 8 warps per SM (2 per scheduler), and a packed FADD2 / FFMA2 holds its warp for two issue cycles, so the kernel is
// bound by per-warp latency (issue slots 44 % busy, shared-memory pipe 51 %) where the block kernel, with 16 warps, is
// bound by the shared-memory pipe (72 %).  10 and 12 warps (168 registers, a few spills) are no faster (8.0 / 7.1 ms).
// Kept as a parity-tested alternative (SGX_K1W2=1); the block kernel stays the default at every FFT size.
#include <cstdint>
#include <cstdlib>
#include <algorithm>

#include "device_common.cuh"
#include "kernels.h"
#include "stft_device.cuh"

namespace sgx {

namespace {

constexpr int kW2H = 1024;                  // complex points
constexpr int kW2MaxWarps = 12;             // 8 by default; 10 / 12 trade registers per thread for resident warps (SGX_W2_WARPS)
constexpr int kW2Pitch = 33;                // float2 elements per row of a warp's exchange plane (conflict-free both ways)
constexpr int kW2Plane = 32 * kW2Pitch;     // 8448 bytes; also holds the 1025 magnitude pairs of the two frames

template <bool MEL, int kW2Warps>
__global__ void __launch_bounds__(kW2Warps * 32, 1) stft_warp2_kernel(const StftLaunch L)
{
    constexpr int H = kW2H, F = 2 * kW2H, kW2Threads = kW2Warps * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned *done_cnt = reinterpret_cast<unsigned *>(smem_raw + 8); // warps that have consumed the current tile
    float *tile = reinterpret_cast<float *>(smem_raw + 16);
    float2 *xall = reinterpret_cast<float2 *>(tile + L.tile_floats);
    float2 *win_s = xall + kW2Warps * kW2Plane;  // [h]     (w[2m], w[2m+1]) of the current track
    float2 *tw_s = win_s + H;                    // [32][32] W_1024^(k1 lane)
    float2 *spl_s = tw_s + H;                    // [h/2]   (cos, sin)(k pi / h)                   realfft.rs:88-93
    float *bank = reinterpret_cast<float *>(spl_s + H / 2); // block-padded mel bank of the current track

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2 *xb = xall + warp * kW2Plane;
    const int mode = MEL ? (int)MODE_MEL_DB : L.mode;

    for (int i = tid; i < H; i += kW2Threads) tw_s[i] = __ldg(L.tw + (((i >> 5) * (i & 31)) & (H - 1)));
    for (int i = tid; i < H / 2; i += kW2Threads) spl_s[i] = __ldg(L.split + i);
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
        for (int i = tid; i < (int)(sizeof(StftTrack) / sizeof(int)); i += kW2Threads) dst[i] = src[i];
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
            for (int i = tid; i < H; i += kW2Threads) win_s[i] = __ldg(wsrc + i);
            if (MEL) {
                const int words = td->seg_words;
                const int *__restrict__ srcw = td->segp;
                int *dstw = reinterpret_cast<int *>(bank);
                for (int i = tid; i < words; i += kW2Threads) dstw[i] = __ldg(srcw + i);
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
            for (int s = tid; s < cur.len; s += kW2Threads) tile[s] = load_sample(pv, cur.A0 + s);
            __syncthreads();
        }

        const int rounds = L.frames_per_tile / (2 * kW2Warps);
        for (int r = 0; r < rounds; ++r) {
            if (r * 2 * kW2Warps >= nfr) break; // uniform
            const int fl0 = (r * kW2Warps + warp) * 2; // first local frame of this warp
            const bool active = fl0 < nfr;             // warp-uniform
            pk re[32][1], im[32][1];                   // pair = (frame fl0, frame fl0 + 1)

            // ---- A: windowed samples, z[32 m1 + lane] = (g[2m], g[2m+1]) ------------------------------------------
            if (active) {
                const int b0 = off0 + fl0 * hop, b1 = off0 + min(fl0 + 1, nfr - 1) * hop; // a frame past the end repeats the last
                const float *f0 = tile + b0 + 2 * lane, *f1 = tile + b1 + 2 * lane;
                const int al = (b0 & 1) | ((b1 & 1) << 1);
                if (al == 0) { // both frames start on an even float: one 64-bit shared load per point and frame
#pragma unroll
                    for (int m1 = 0; m1 < 32; ++m1) {
                        const float2 w = win_s[32 * m1 + lane];
                        const float2 x0 = *reinterpret_cast<const float2 *>(f0 + 64 * m1);
                        const float2 x1 = *reinterpret_cast<const float2 *>(f1 + 64 * m1);
                        re[m1][0] = make_float2(x0.x * w.x, x1.x * w.x); im[m1][0] = make_float2(x0.y * w.y, x1.y * w.y);
                    }
                } else if (al == 1) { // odd hop: one frame of the pair is odd
#pragma unroll
                    for (int m1 = 0; m1 < 32; ++m1) {
                        const float2 w = win_s[32 * m1 + lane];
                        const float xa = f0[64 * m1], xb0 = f0[64 * m1 + 1];
                        const float2 x1 = *reinterpret_cast<const float2 *>(f1 + 64 * m1);
                        re[m1][0] = make_float2(xa * w.x, x1.x * w.x); im[m1][0] = make_float2(xb0 * w.y, x1.y * w.y);
                    }
                } else if (al == 2) {
#pragma unroll
                    for (int m1 = 0; m1 < 32; ++m1) {
                        const float2 w = win_s[32 * m1 + lane];
                        const float2 x0 = *reinterpret_cast<const float2 *>(f0 + 64 * m1);
                        const float xa = f1[64 * m1], xb1 = f1[64 * m1 + 1];
                        re[m1][0] = make_float2(x0.x * w.x, xa * w.x); im[m1][0] = make_float2(x0.y * w.y, xb1 * w.y);
                    }
                } else {
#pragma unroll
                    for (int m1 = 0; m1 < 32; ++m1) {
                        const float2 w = win_s[32 * m1 + lane];
                        const float xa = f0[64 * m1], xb0 = f0[64 * m1 + 1];
                        const float xc = f1[64 * m1], xd = f1[64 * m1 + 1];
                        re[m1][0] = make_float2(xa * w.x, xc * w.x); im[m1][0] = make_float2(xb0 * w.y, xd * w.y);
                    }
                }
            }
            // ---- the tile is in registers: let the next one stream in (last warp to check in issues the copy) ------
            if (r == rounds - 1 || (r + 1) * 2 * kW2Warps >= nfr) {
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    const unsigned seen = atomicAdd(done_cnt, 1u);
                    if (seen % kW2Warps == kW2Warps - 1) {
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
            dft_inplace<32, 0, 1, 32, 1>(re, im);
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) {
                const float2 w = tw_s[k1 * 32 + lane];
                const pk xr = re[k1][0], xi = im[k1][0];
                re[k1][0] = pk_fmas(xi, -w.y, pk_muls(xr, w.x));
                im[k1][0] = pk_fmas(xi, w.x, pk_muls(xr, w.y));
            }
            // ---- C: transpose (lane m2, register k1) -> (lane k1, register m2), one plane at a time ------------------
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) xb[k1 * kW2Pitch + lane] = re[k1][0];
            __syncwarp();
#pragma unroll
            for (int m2 = 0; m2 < 32; ++m2) re[m2][0] = xb[lane * kW2Pitch + m2];
            __syncwarp();
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) xb[k1 * kW2Pitch + lane] = im[k1][0];
            __syncwarp();
#pragma unroll
            for (int m2 = 0; m2 < 32; ++m2) im[m2][0] = xb[lane * kW2Pitch + m2];
            __syncwarp(); // the plane now becomes the magnitude array (mel)
            // ---- D: 32-point DFT over m2 -> Z[lane + 32 k2] in register k2 --------------------------------------------
            dft_inplace<32, 0, 1, 32, 1>(re, im);

            // ---- E: real-FFT split (realfft.rs:140-157) + what becomes of a bin ------------------------------------
            float2 *magbuf = xb; // [h + 1] magnitude pairs at their bin index (mel)
            auto emit = [&](int idx, pk xr, pk xi) {
                if (mode == MODE_COMPLEX) {
                    float2 *op = reinterpret_cast<float2 *>(out) + (size_t)(t0 + fl0) * (H + 1) + idx;
                    op[0] = make_float2(xr.x, xi.x);
                    if (fl0 + 1 < nfr) op[H + 1] = make_float2(xr.y, xi.y);
                    return;
                }
                const pk q = pk_fma(xr, xr, pk_mul(xi, xi));
                const pk mg = make_float2(sqrt_approx(q.x), sqrt_approx(q.y)); // lib.rs:124
                if (MEL) {
                    magbuf[idx] = mg;
                } else {
                    float y0 = mg.x, y1 = mg.y;
                    if (mode == MODE_LIN_DB) { // the second frame may be a copy of the first: harmless in the extrema
                        y0 = amp_to_db_dev(y0); y1 = amp_to_db_dev(y1);
                        vmax = fmaxf(vmax, fmaxf(y0, y1)); vmin = fminf(vmin, fminf(y0, y1));
                    }
                    float *op = out + (size_t)(t0 + fl0) * (H + 1) + idx;
                    op[0] = y0;
                    if (fl0 + 1 < nfr) op[H + 1] = y1;
                }
            };
            // Bin k = lane + 32 j sits in register j; its conjugate partner h - k in lane (32 - lane) & 31, register
            // 31 - j (lane 0: its own register (32 - j) & 31).  Every lane finishes the pairs of its 16 lowest bins:
            // together that is every bin but h/2, which lane 0 adds.
            const int pl = (32 - lane) & 31;
            const bool l0 = lane == 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const pk sr = l0 ? re[(32 - j) & 31][0] : re[31 - j][0];
                const pk si = l0 ? im[(32 - j) & 31][0] : im[31 - j][0];
                pk br, bi;
                br.x = __shfl_sync(0xffffffffu, sr.x, pl); br.y = __shfl_sync(0xffffffffu, sr.y, pl);
                bi.x = __shfl_sync(0xffffffffu, si.x, pl); bi.y = __shfl_sync(0xffffffffu, si.y, pl);
                const pk ar = re[j][0], ai = im[j][0];
                const int k = lane + 32 * j;
                const float2 cs = spl_s[k]; // (cos, sin)(k pi / h)
                const pk sumr = pk_add(ar, br), difr = pk_sub(ar, br);
                const pk sumi = pk_add(ai, bi), difi = pk_sub(ai, bi);
                const pk p1 = pk_fmas(sumi, cs.x, pk_muls(difr, -cs.y));  // c*sumi - s*difr
                const pk p2 = pk_fmas(sumi, cs.y, pk_muls(difr, cs.x));   // s*sumi + c*difr
                emit(k, pk_muls(pk_add(sumr, p1), 0.5f), pk_muls(pk_sub(difi, p2), 0.5f));
                emit(k == 0 ? H : H - k, pk_muls(pk_sub(sumr, p1), 0.5f), pk_muls(pk_add(difi, p2), -0.5f)); // k == 0: Nyquist bin
            }
            if (l0) emit(H / 2, re[16][0], pk_neg(im[16][0])); // X[h/2] = conj(Z[h/2])

            // ---- F: mel projection + dB, segment form of the bank (host_tables.h MelBands::seg) ----------------------
            // A lane walks the bins of ONE segment: every magnitude pair is read once and feeds U (rising side, filter s)
            // and D (falling side, filter s - 1); filter m = U_m + D_(m+1), the neighbour's D arriving by a shuffle.
            if (MEL) {
                __syncwarp(); // magnitudes of all bins are in the plane
                const int lg = td->seg_log2p, P = 1 << lg;
                const int nwq = td->seg_nwq, nblk = td->seg_nblk;
                const float2 *wq = reinterpret_cast<const float2 *>(bank);
                const int *lo_s = reinterpret_cast<const int *>(bank) + 2 * nwq;
                const int2 *desc_s = reinterpret_cast<const int2 *>(lo_s + 32 * nblk);
                const int plm = lane & (P - 1);
                float *orow = out + (size_t)(t0 + fl0) * n_out;
                const bool two = fl0 + 1 < nfr;
                for (int blk = 0; blk < nblk; ++blk) {
                    const int2 bd = desc_s[blk];
                    const int li = lo_s[blk * 32 + lane];   // first bin | filter << 16
                    const int m = (int)((unsigned)li >> 16);
                    const float2 *wp = wq + bd.x + lane;
                    const float2 *mp = magbuf + (li & 0xffff);
                    pk up = make_float2(0.0f, 0.0f), dn = up;
                    for (int j2 = 0; j2 < bd.y; j2 += 2) { // (fully unrolled tap bodies behind a switch were slower: 7.1 vs 6.4 ms)
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const float2 wgt = wp[(j2 + u) * 32];
                            const pk mg = *mp;
                            mp += P;
                            up = pk_fmas(mg, wgt.x, up); dn = pk_fmas(mg, wgt.y, dn);
                        }
                    }
                    if (P > 1) {
                        for (int sh = P >> 1; sh > 0; sh >>= 1) {
                            up.x += __shfl_xor_sync(0xffffffffu, up.x, sh); up.y += __shfl_xor_sync(0xffffffffu, up.y, sh);
                            dn.x += __shfl_xor_sync(0xffffffffu, dn.x, sh); dn.y += __shfl_xor_sync(0xffffffffu, dn.y, sh);
                        }
                    }
                    const float a0 = up.x + __shfl_down_sync(0xffffffffu, dn.x, P); // D of the next segment
                    const float a1 = up.y + __shfl_down_sync(0xffffffffu, dn.y, P);
                    if (m != 0xffff && plm == 0) {
                        const float y0 = amp_to_db_dev(a0), y1 = amp_to_db_dev(a1); // decibel.rs:33-88
                        vmax = fmaxf(vmax, fmaxf(y0, y1)); vmin = fminf(vmin, fminf(y0, y1));
                        float *op = orow + m;
                        op[0] = y0;
                        if (two) op[n_out] = y1;
                    }
                }
            }
            __syncwarp(); // magnitudes consumed before the next round's transpose overwrites the plane
        }
    } // tiles of this CTA
    flush_range();
}

} // namespace

size_t stft_warp2_fixed_smem(int bank_floats, int warps)
{
    return 16 + (size_t)warps * kW2Plane * sizeof(float2) + (size_t)(kW2H + kW2H + kW2H / 2) * sizeof(float2) +
           (size_t)bank_floats * sizeof(float);
}

int stft_warp2_warps()
{
    static const int w = [] {
        const char *e = getenv("SGX_W2_WARPS");
        const int v = e ? atoi(e) : 8;
        return (v == 10 || v == 12) ? v : 8;
    }();
    return w;
}

namespace {
template <bool MEL, int NW> cudaError_t launch_w2(const StftLaunch &L, size_t smem, int grid, cudaStream_t stream)
{
    auto kern = stft_warp2_kernel<MEL, NW>;
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(kern), smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, stream>>>(L);
    count_launch();
    return cudaGetLastError();
}
} // namespace

cudaError_t launch_stft_warp2(const StftLaunch &L, cudaStream_t stream)
{
    const int nw = L.warp2;
    const size_t smem = stft_warp2_fixed_smem(L.bank_floats, nw) + (size_t)L.tile_floats * sizeof(float);
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    const int grid = std::min(L.n_tiles, sms);
    const bool mel = L.mode == MODE_MEL_DB;
    switch (nw) {
    case 8: return mel ? launch_w2<true, 8>(L, smem, grid, stream) : launch_w2<false, 8>(L, smem, grid, stream);
    case 10: return mel ? launch_w2<true, 10>(L, smem, grid, stream) : launch_w2<false, 10>(L, smem, grid, stream);
    case 12: return mel ? launch_w2<true, 12>(L, smem, grid, stream) : launch_w2<false, 12>(L, smem, grid, stream);
    }
    return cudaErrorInvalidValue;
}

} // namespace sgx
