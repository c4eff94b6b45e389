// host_tables.h -- host-side parameter derivation and table generation of the path.
// These are the parts of the reference that stay on the CPU because they run once per sample
// rate (lib.rs:143-158) and their outputs are INPUTS of the device kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace sgx {

// utils.rs:17-19
size_t calc_proper_n_fft(size_t win_length);
// windows.rs:7-30
void hann(size_t size, bool symmetric, float *out);
// lib.rs:138-140
void calc_window(size_t win_length, size_t n_fft, float *out);
// mel.rs:14-31
float mel_to_hz(float mel);
float hz_to_mel(float hz);
// mel.rs:33-85 ; fmax < 0 == None ; out [n_fft/2+1][n_mel]
void calc_mel_fb(uint32_t sr, size_t n_fft, size_t n_mel, float fmin, float fmax, bool do_norm,
                 float *out);
// mel.rs:87-99
size_t calc_mel_fb_default(uint32_t sr, size_t n_fft, std::vector<float> &fb);
// lib.rs:412-435 frame count of perform_stft; <0 where the reference panics
long stft_num_frames(size_t n, size_t win, size_t hop);
// lib.rs:296
uint32_t calc_nwidth(float px_per_sec, size_t n, uint32_t sr);
// lib.rs:231-248
float calc_up_ratio(uint32_t max_sr, uint32_t sr, bool mel);
// display.rs:45
uint32_t grey_height(size_t n_out, float up_ratio);
// tap window [left, right) of output index o when resampling n_in -> n_out with image 0.23's Lanczos3
// (same f32 operations as the device table builder)
void lanczos3_span(uint32_t n_in, uint32_t n_out, uint32_t o, uint32_t *left, uint32_t *right);

// Banded (CSR-by-filter) form of a [n_freq][n_mel] filterbank for the device projection.
struct MelBands {
    std::vector<int> lo, cnt, off; // per filter: first bin, tap count, offset into w
    std::vector<float> w;          // taps, filter-major
    int max_cnt = 0;
    int log2_split = 0;            // lanes cooperating on one filter (power of two <= 32)
    std::vector<int> sched;        // {slots, taps, staged, 0, block ids [slots][warps]}: balanced block lists per warp
    // Segment form of the bank for the kernels that keep the magnitudes unpadded (fused last pass, warp kernel); see
    // build_mel_segments in host_tables.cpp.  32-bit words: weight pairs {rising, falling} [2 * 32 * sum_b nj_b]
    // (block b, tap j of lane i at 2 * (off_b + 32 j + i), zero outside the lane's own bins), {first bin | filter << 16}
    // per lane [32 * n_blocks] (0xffff: the lane produces no output), {off_b, nj_b} per block, then the schedule
    // {slots, block ids [slots][warps]}.  Block b walks segments b (U - 1) ... b (U - 1) + U - 1 with U = 32 / P units of
    // P adjacent lanes; unit u < U - 1 outputs filter b (U - 1) + u = U_u + D_(u+1).  Empty when the bank is not mel-like.
    std::vector<int> seg;
    int seg_nwq = 0, seg_nblk = 0, seg_log2p = 0, seg_slots = 0;
    int seg_words(int warps) const { return 2 * seg_nwq + 34 * seg_nblk + 1 + seg_slots * warps; }
};
// `vec`: frames a thread group transforms together (magnitudes sit in shared memory as vectors of `vec` floats).
MelBands make_mel_bands(const float *fb, size_t n_freq, size_t n_mel, int threads_per_group,
                        size_t stage_capacity_floats, int vec);

} // namespace sgx
