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
    // Block-padded copy of the bank for the kernels whose last FFT pass is fused with the split (they keep the
    // magnitudes unpadded).  32-bit words: taps [sum_b 32 * nj4_b] (block b, tap j of work item i at
    // off_b + 32 j + i, zero beyond an item's own taps; nj4_b = longest item of the block rounded up to 4),
    // {first bin | filter << 16} per lane [32 * n_blocks] (0xffff: no filter), then {off_b, nj4_b} per block.  The P lanes
    // of a filter are adjacent; the filters of a block are ordered so that the lanes of a shared-memory phase hit
    // different bank groups.
    std::vector<int> packed;
    int packed_nwb = 0, packed_nblk = 0;
};
MelBands make_mel_bands(const float *fb, size_t n_freq, size_t n_mel, int threads_per_group,
                        size_t stage_capacity_floats);

} // namespace sgx
