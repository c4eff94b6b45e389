// kernels.h -- host-callable launchers of the sm_100a kernels (implemented in *.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "device_common.cuh"

namespace sgx {

// Static configuration of the fused analysis kernel K1 for one FFT size.
struct StftConfig {
    int n_fft;             // F
    int h;                 // F/2 complex points
    int pts;               // points per thread (radix of the main passes)
    int vec;               // frames processed together by one thread group (V)
    int groups;            // thread groups per CTA (G)
    int threads;           // G * h / pts
    int min_ctas;          // resident CTAs per SM the kernel is compiled for
    size_t fft_smem;       // bytes of FFT exchange buffers
    bool generic;          // small-F fallback kernel (one CTA per frame)
    bool fused;            // last FFT pass fused with the split: magnitudes unpadded, block-padded mel bank
    bool warp2;            // n_fft = 2048: the warp-per-frame-pair kernel (stft_warp2_kernel.cu) is preferred when its tile fits
    bool warp1;            // n_fft = 2048: the warp-per-frame kernel on packed complex values (stft_warp1_kernel.cu), likewise
};
bool stft_config_for(size_t n_fft, StftConfig *cfg);
size_t stft_max_dynamic_smem();

// Enqueues K1.  `launch` carries device pointers; tile geometry must come from plan_stft_tiles.
cudaError_t launch_stft(const StftConfig &cfg, const StftLaunch &launch, cudaStream_t stream);

// Chooses frames per tile / staging for a set of (hop) values sharing one FFT size.
// `bank_floats`: shared-memory floats the largest mel filterbank of the launch needs (taps rounded up to 4, plus
// 4 per filter for its descriptor), 0 when not a mel launch; the planner reports whether it got its own region.
struct StftTiling { int frames_per_tile; int staged; int tile_floats; size_t smem_bytes; int bank_floats; int sample_floats; int warp2; int warp1; };
// `sample_floats`: 2 when the launch holds f32 stereo tracks (their tiles are staged as raw interleaved pairs), else 1.
// `warp2_ok`: every mel track of the launch has a segment-form bank (the warp kernel has no other mel path).
StftTiling plan_stft_tiles(const StftConfig &cfg, int max_hop, int bank_floats = 0, int sample_floats = 1, bool warp2_ok = true);

// the warp-per-frame-pair kernel (n_fft = 2048); `launch.warp2` routes launch_stft here
cudaError_t launch_stft_warp2(const StftLaunch &launch, cudaStream_t stream);
size_t stft_warp2_fixed_smem(int bank_floats, int warps); // shared memory besides the PCM tile
int stft_warp2_warps();                                    // warps per CTA (8 unless SGX_W2_WARPS = 10 | 12)
// the warp-per-frame kernel on packed complex values (n_fft = 2048); `launch.warp1` routes launch_stft here
cudaError_t launch_stft_warp1(const StftLaunch &launch, cudaStream_t stream);
size_t stft_warp1_fixed_smem(int bank_floats, int warps);
int stft_warp1_warps();                                    // warps per CTA (16 unless SGX_W1_WARPS = 8 | 12)

// FFT twiddle tables for one size (host vectors -> caller uploads).
void make_fft_tables(int h, float2 *tw /*[h]*/, float2 *split /*[h/2+1]*/);
// twiddles of the Stockham passes, per pass and r-major (what the block kernel of this size loads)
std::vector<float2> make_fft_pass_tables(int h, int pts);

// K2: global dB range (lib.rs:193-209)
cudaError_t launch_range_init(unsigned *slots, int n_slots, cudaStream_t s);
// resets the range slot of every track of a K1 descriptor array to the identity of the reduce
cudaError_t launch_range_reset(const StftTrack *descs, int n, cudaStream_t s);
// reduces slots [n][2] -> local {max, -min, max_sr, max_sec} (the last two are host metadata passed through)
// commit_state != nullptr: launch_range_commit's work is done by the same launch (no exchange in between)
cudaError_t launch_range_reduce(const unsigned *slots, int n_slots, float *local_max_negmin, float max_sr,
                                float max_sec, float db_range, float *commit_state, cudaStream_t s);
// {max, -min, max_sr, max_sec} -> state {max_db, min_db, changed, max_sr, max_sec}: the clamps of lib.rs:208-209
// and the sticky 1e-3 change detection of lib.rs:210-218, all on the device
cudaError_t launch_range_commit(const float *max_negmin, float db_range, float *state,
                                cudaStream_t s);

// K3: render
struct AxisTable { int *left, *cnt; float *sum, *w; int taps; int n_in, n_out; bool tap_major; };
uint32_t lanczos3_max_taps(uint32_t n_in, uint32_t n_out);
cudaError_t launch_build_axis_table(int n_in, int n_out, int taps, bool tap_major, int *left,
                                    int *cnt, float *sum, float *w, cudaStream_t s);
// fast: 0 general kernel, 1 wide path, 2 tensor-core path (render_tc_kernel.cu), 3 sliding-window path
// (render_slide_kernel.cu), TV * 100 + TH fast FP32 path
struct RenderTiling { int px, py, fc, rv_max; size_t smem_bytes; int fast; };
RenderTiling plan_render_tiles(int width, int height, int nwidth, int nheight, bool from_db = false);
// the tcgen05 path: tiles of 128 rows x L.px columns, L.fc source frames, L.rv_max grey rows per tile
size_t render_tc_smem(int kv, int nf, int nx);
cudaError_t launch_render_tc(const RenderLaunch &launch, size_t smem_bytes, cudaStream_t s);
// the sliding-window path: tiles of 120 x 64 pixels, L.fc source frames (multiple of 32), L.rv_max grey rows per tile
bool render_slide_plan(int width, int height, int nwidth, int nheight, RenderTiling *out);
cudaError_t launch_render_slide(const RenderLaunch &launch, int max_nwidth, int max_nheight, size_t smem_bytes, cudaStream_t s);
cudaError_t launch_render(const RenderLaunch &launch, int max_nwidth, int max_nheight,
                          size_t smem_bytes, int fast, cudaStream_t s);

// small elementwise kernels of the stage API
cudaError_t launch_spec_to_grey(const float *spec, int n_frames, int n_out, int height, float max_db,
                                float min_db, float *grey, cudaStream_t s);
cudaError_t launch_amp_to_db(float *x, size_t n, int *bad_flag, cudaStream_t s);
cudaError_t launch_wav_image(const void *pcm, int fmt, int ch, long long n, int nwidth, int nheight,
                             float amp_min, float amp_max, unsigned char *out, int *err_flag,
                             cudaStream_t s);

// Opt a kernel into `bytes` of dynamic shared memory on the CURRENT device (the attribute is per device and
// per function; thread-safe, remembers what was already granted).
cudaError_t ensure_dynamic_smem(const void *func, size_t bytes);

void count_launch(int n = 1);
uint64_t launch_count();

} // namespace sgx
