// device_common.cuh -- descriptors shared by the host launchers and the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sgx {

// What K1 writes per frame.
enum StftMode : int {
    MODE_COMPLEX = 0, // perform_stft            lib.rs:388-471   out [T][h+1] (re,im)
    MODE_MAG = 1,     // .mapv(norm)             lib.rs:124       out [T][h+1]
    MODE_LIN_DB = 2,  // Linear + amp_to_db      lib.rs:126-129   out [T][h+1]
    MODE_MEL_DB = 3   // .dot(mel_fb) + dB       lib.rs:130-134   out [T][n_mel]
};

enum PcmFormat : int { PCM_F32 = 0, PCM_I16 = 1 };

// One track of a K1 launch (device-resident array of these).
struct StftTrack {
    const void *pcm;       // interleaved [n][ch], f32 or i16 (audio.rs:16-19 scale applied on load)
    long long n;           // samples per channel of the WHOLE track (reflection happens at 0 and n-1)
    long long origin;      // global sample index of pcm[0]  (0 unless the handle holds a time slice)
    long long avail;       // samples per channel present at `pcm` (n unless a time slice)
    int frame0;            // global index of the first frame this descriptor produces (row 0 of `out`)
    int ch;
    int fmt;               // PcmFormat
    int win, hop, pad_l;   // W, H, (F-W)/2                              lib.rs:400
    int n_frames;          // frames to produce (T of lib.rs:435 unless a time slice)
    const float *win_f;    // [F]: window centred in the FFT frame, zeros outside (lib.rs:377-384)
    float *out;            // see StftMode
    int n_out;             // h+1 or n_mel
    // banded mel filterbank (MODE_MEL_DB)
    const int *mel_lo, *mel_cnt, *mel_off; // mel_lo points at packed int4 {lo, cnt, off, 0} per filter
    const float *mel_w;
    int mel_log2p;
    const int *segp;       // segment form of the bank (host_tables.h MelBands::seg), or null when the bank is not mel-like
    int seg_nwq, seg_nblk;   // its weight pairs and blocks
    int seg_log2p;           // lanes cooperating on one segment
    int seg_words;           // 32-bit words in total (pairs, lane and block descriptors, schedule)
    unsigned *range_slot;  // [2] order-preserving encodings of (max, min) dB; may be null
    int tile_begin;        // first CTA tile of this track inside the launch
};

struct StftLaunch {
    const StftTrack *tracks;
    int n_tracks;
    int n_tiles;
    int mode;
    int frames_per_tile;   // multiple of (groups * V)
    int staged;            // 1: PCM tile staged in shared memory, 0: direct global loads
    int tile_floats;       // capacity of the staged tile
    int bank_floats;       // > 0: the CTA keeps the track's mel taps + descriptors in a region of its own
    int stereo_raw;        // raw-tile loader of the launch -- 1: f32 stereo tracks (tiles have room for the pairs), 2: int16 mono tracks, 0: none
    int warp2;             // > 0: tiles were planned for the warp-per-frame-pair kernel (n_fft = 2048) with that many warps
    int warp1;             // > 0: tiles were planned for the warp-per-frame kernel (n_fft = 2048) with that many warps
    const float2 *tw;      // [h]      exp(-2 pi i j / h)
    const float2 *twr;     // twiddles of the Stockham passes, pass after pass, r-major (kernels.h make_fft_pass_tables)
    const float2 *split;   // [h/2+1]  (cos, sin)(k pi / h)                realfft.rs:88-93
};

// order-preserving float <-> unsigned mapping for atomicMax / atomicMin on dB values
__host__ __device__ inline unsigned enc_ordered(float f)
{
#ifdef __CUDA_ARCH__
    unsigned u = __float_as_uint(f);
#else
    union { float f; unsigned u; } c; c.f = f; unsigned u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float dec_ordered(unsigned e)
{
    unsigned u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; unsigned u; } c; c.u = u; return c.f;
#endif
}

// decibel.rs:33-88 with ref = 1, amin = 1e-18:  20*log10(x) above amin, else 20*(-18).
// lg2.approx has an absolute error <= 2^-22 so the dB value is within ~2e-6 dB of libm.
__device__ __forceinline__ float amp_to_db_dev(float x)
{
    // x > 1e-18 is a normal number: the bare MUFU, without __log2f's subnormal pre-scaling (same result)
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    return x > 1e-18f ? 6.02059991327962390f * l : -360.0f;
}

// One track of a K3 launch.
struct RenderTrack {
    const float *src;      // dB [rows][n_out]  (from_db)  or grey [height][width]
    int width;             // T of the whole track
    int frame0, src_frames; // from_db: src row r holds global frame frame0 + r, src_frames rows exist
    int ox_begin, ox_count; // output columns rendered by this launch; `out` is [nheight][ox_count]
    int n_out;             // rows of the dB array (from_db)
    int height;            // grey height = round(n_out * up_ratio)     display.rs:45
    int nwidth, nheight;
    unsigned char *out;    // [nheight][nwidth][channels]
    // separable Lanczos3 tables built by build_axis_table (image 0.23 resize semantics)
    const int *v_left, *v_cnt; const float *v_sum, *v_w; int v_taps;   // v_w [nheight][v_taps]
    const int *h_left, *h_cnt; const float *h_sum, *h_w; int h_taps;   // h_w [h_taps][nwidth]
};

struct RenderLaunch {
    const RenderTrack *tracks;
    int n_tracks;
    int from_db;           // 1: normalise/clip/flip/top-pad on load (display.rs:44-54)
    const float *range;    // device {max_db, min_db} (from_db)
    int channels;          // 3 or 4
    int px, py;            // output tile
    int fc;                // frames per chunk
    int rv_max;            // grey rows a tile may need
    int rv_cols, rv_rows;  // tensor-core path: widest column window of the launch, row tiles of 128
    int debug;             // tensor-core path: print per-phase cycle counts of CTA 0 (SGX_K3_TC_PROF)
};

} // namespace sgx
