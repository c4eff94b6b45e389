// engine.h -- host side of the engine: device buffers, plan caches and the MultiTrack container
// that mirrors the reference's `MultiTrack` (src_rust/lib.rs:72-365) on top of the kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sgx.h"
#include "host_tables.h"
#include "kernels.h"
#include "nccl_dyn.h"

namespace sgx {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void cuda_check(cudaError_t e, const char *what, const char *file, int line);
#define SGX_CUDA(x) ::sgx::cuda_check((x), #x, __FILE__, __LINE__)

template <class T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count)
    {
        release();
        if (count == 0) return;
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, count * sizeof(T));
        if (e != cudaSuccess) { cudaGetLastError(); throw Error(SGX_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e)); }
        p = static_cast<T *>(q); n = count;
    }
    void ensure(size_t count) { if (count > n) alloc(count); }
    void upload(const T *h, size_t count, cudaStream_t s)
    {
        if (count > n) alloc(count);
        if (count) SGX_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
};

// FFT tables of one size on one device
struct FftPlan {
    StftConfig cfg;
    DevBuf<float2> tw, split, twr;
};

// everything that depends on (sr, win, n_fft, n_mel): the `windows` / `mel_fbs` caches, lib.rs:76-77
struct TrackTables {
    size_t win = 0, n_fft = 0, n_mel = 0; // n_mel == 0: linear scale
    DevBuf<float> win_f;                   // [n_fft]
    DevBuf<int> mel_lo, mel_cnt, mel_off;
    DevBuf<float> mel_w;
    int mel_log2p = 0;
    int mel_nnz = 0;
    DevBuf<int> segp;                      // segment form of the bank (MelBands::seg); empty when the bank is not mel-like
    int seg_nwq = 0, seg_nblk = 0, seg_log2p = 0, seg_words = 0;
    // shared-memory floats a K1 launch reserves for the bank of this track
    int bank_floats(bool fused) const { return fused && seg_nblk > 0 ? seg_words : ((mel_nnz + 3) & ~3) + 4 * (int)n_mel; }
};

struct AxisTableDev {
    DevBuf<int> left, cnt;
    DevBuf<float> sum, w;
    int taps = 0;
};

struct Track {
    std::string path;
    uint32_t sr = 0, ch = 1;
    int fmt = PCM_F32;
    size_t n = 0;             // samples per channel of the whole track
    size_t avail = 0, origin = 0; // samples present at d_pcm and the global index of the first one
    bool is_slice = false;    // the handle holds a time slice of the track (n3: one track over several GPUs)
    size_t t_total = 0, frame0 = 0; // frames of the whole track; global index of spec row 0
    size_t win = 0, hop = 0, n_fft = 0;
    const void *d_pcm = nullptr;
    DevBuf<unsigned char> owned_pcm;
    TrackTables *tables = nullptr;
    DevBuf<float> spec;       // dB [T][n_out]   == MultiTrack.specs[id]  (lib.rs:78)
    size_t n_frames = 0, n_out = 0;
    int slot = -1;
};

struct PcmSource {
    const void *data; int fmt; size_t n; uint32_t sr; uint32_t ch; bool on_device; std::string path;
    // time slice of a longer track (n_total != 0): `data` holds samples [origin, origin + n) of a track of
    // n_total samples and frames [frame_begin, frame_begin + frame_count) are to be analysed
    size_t n_total = 0, origin = 0, frame_begin = 0, frame_count = 0;
};

class DeviceCtx {
public:
    static DeviceCtx &get(int device);
    FftPlan &plan(size_t n_fft);
    int device() const { return device_; }
    int sm_count = 0;
private:
    explicit DeviceCtx(int d);
    int device_;
    std::map<size_t, std::unique_ptr<FftPlan>> plans_;
};

class MultiTrack {
public:
    MultiTrack(const sgx_settings &s, int device, cudaStream_t stream);
    ~MultiTrack();

    // lib.rs:171-191.  want_changed == false: no host synchronisation.
    // `sliced`: the sources are time slices of tracks every shard holds a piece of (n3): no ownership filter
    bool add_tracks(const std::vector<size_t> &ids, std::vector<PcmSource> &srcs, bool want_changed, bool sliced = false);
    bool remove_track(size_t id, bool want_changed);    // lib.rs:265-292
    // ---- sharding: the tracks of a batch spread over several handles -- GPUs of one process or ranks ----
    // With a communicator attached, track `id` lives on the handle with id % world == rank; add_tracks takes the
    // WHOLE id list on every handle and keeps its own share, remove_track of a foreign id only joins the exchange.
    // Every add / remove then all-reduces {max, -min, max_sr, max_sec} in-stream before the range is committed.
    void attach_comm(NcclComm comm, int rank, int world, bool owned);
    bool owns(size_t id) const { return world_ <= 1 || (int)(id % (size_t)world_) == rank_; }
    int rank() const { return rank_; }
    int world() const { return world_; }
    // the phases of add_tracks / remove_track, exposed so that a multi-device handle can issue the exchanges of its
    // sub-engines as one NCCL group:  analyse | drop  ->  exchange  ->  commit  [-> synchronize]
    void analyse(const std::vector<size_t> &ids, std::vector<PcmSource> &srcs);
    void analyse_owned(const std::vector<size_t> &ids, std::vector<PcmSource> &srcs); // this handle's share of a batch
    void drop(size_t id);
    void exchange();
    void commit();
    // batched host-buffer images (lib.rs:294-298 for a list of ids): renders go to per-image device staging, the
    // device->host copies run on a second stream under the remaining renders -- and under the uploads of the NEXT
    // add_tracks; wait_images blocks until the host buffers of the last request are complete
    void images_async(const std::vector<size_t> &ids, float px_per_sec, uint32_t nheight, int channels,
                      uint8_t *const *out, const size_t *cap, size_t *written);
    void wait_images();
    // lib.rs:294-298 (channels 3) / RGBA; device output, asynchronous
    void render(const std::vector<size_t> &ids, float px_per_sec, uint32_t nheight, int channels,
                uint8_t *const *d_out, const size_t *cap, size_t *written, const uint32_t *ox_begin = nullptr,
                const uint32_t *ox_count = nullptr);
    void render_host(size_t id, float px_per_sec, uint32_t nheight, int channels, uint8_t *out, size_t need);
    void set_profiling(bool on);
    void stage_times(float *analysis_ms, float *render_ms);
    std::vector<uint8_t> wav_image(size_t id, float px_per_sec, uint32_t nheight, float amp_min, float amp_max);

    const Track &track(size_t id) const;
    bool synchronize();                                 // returns `changed` accumulated since last call
    void synchronize_keep_changed() { const bool c = synchronize(); changed_acc_ = changed_acc_ || c; }
    float max_db() { synchronize_keep_changed(); return max_db_; }
    float min_db() { synchronize_keep_changed(); return min_db_; }
    float max_sec() { if (comm_) synchronize_keep_changed(); return world_ > 1 ? std::max(max_sec_, global_max_sec_) : max_sec_; }
    float frequency_hz(size_t id, float rel) const;     // lib.rs:315-322
    uint32_t image_width(size_t id, float px_per_sec) const;
    float *range_device_ptr() { return d_local_.p; }
    void commit_range_device();
    void set_global_max_sr(uint32_t sr) { global_max_sr_ = sr; global_sr_fixed_ = sr != 0; }
    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }
    const sgx_settings &settings() const { return set_; }
    void derive_params(uint32_t sr, size_t *win, size_t *hop, size_t *n_fft) const;

private:
    TrackTables *tables_for(uint32_t sr, size_t win, size_t n_fft);
    AxisTableDev *axis_table(int n_in, int n_out, bool tap_major);
    void drop_track(size_t id);
    void reduce_local();
    int alloc_slot();
    uint32_t effective_max_sr() const;

    sgx_settings set_;
    int device_;
    cudaStream_t stream_;
    bool own_stream_ = false;
    DeviceCtx *ctx_;
    std::map<size_t, Track> tracks_;
    std::map<std::tuple<uint32_t, size_t, size_t, size_t>, std::unique_ptr<TrackTables>> tables_;
    std::map<std::tuple<int, int, bool>, std::unique_ptr<AxisTableDev>> axis_;
    DevBuf<unsigned> slots_;
    std::vector<int> free_slots_;
    int n_slots_ = 0;
    DevBuf<float> d_local_;  // {max, -min, max_sr, max_sec} over local tracks (un-clamped); all-reduced in place
    DevBuf<float> d_state_;  // {max_db, min_db, changed flag, max_sr, max_sec}  sticky like lib.rs:210-218
    DevBuf<StftTrack> d_stft_;
    DevBuf<RenderTrack> d_render_;
    DevBuf<uint8_t> d_img_;      // staging for host-buffer image requests
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
    bool profiling_ = false, ev_valid_[2] = {false, false};
    cudaStream_t copy_stream_ = nullptr;          // H2D uploads of host-resident PCM
    std::vector<cudaEvent_t> copy_events_;
    cudaStream_t out_stream_ = nullptr;           // D2H copies of rendered images
    cudaEvent_t out_ev_ = nullptr, out_done_ = nullptr;
    bool out_pending_ = false;
    std::vector<DevBuf<uint8_t>> d_imgs_;         // per-image staging of images_async
    NcclComm comm_ = nullptr;
    bool comm_owned_ = false;
    int rank_ = 0, world_ = 1;
    float global_max_sec_ = 0.0f;
    bool global_sr_fixed_ = false;
    bool fuse_commit_ = false, committed_in_reduce_ = false; // the range commit rides on the reduce launch (lone handle)
    float max_db_, min_db_;
    float max_sec_ = 0.0f;
    size_t id_max_sec_ = 0;
    uint32_t max_sr_ = 0, global_max_sr_ = 0;
    bool pending_ = false;      // device state newer than the host copy
    bool changed_acc_ = false;  // host-side changes (max_sr) since the last synchronize
};

// ---- stage functions (surface 2) --------------------------------------------------------------------
struct StageOut { size_t n_frames; size_t n_out; };
StageOut stage_stft(int mode, const float *input, size_t n, size_t win, size_t hop, size_t n_fft,
                    const float *window, const float *mel_fb, size_t n_mel, float *out,
                    size_t cap_elems);
void stage_amp_to_db(float *x, size_t n);
uint32_t stage_spec_to_grey(const float *spec, size_t T, size_t n_out, float up_ratio, float max_db,
                            float min_db, float *grey, size_t cap);
void stage_grey_to_rgb(const float *grey, uint32_t width, uint32_t height, uint32_t nwidth,
                       uint32_t nheight, int channels, uint8_t *out, size_t cap);
void stage_wav_to_image(const float *wav, size_t n, uint32_t nwidth, uint32_t nheight, float amp_min,
                        float amp_max, uint8_t *out, size_t cap);

// audio.rs:9-37 (WAV branch).  Decodes to interleaved samples; 16-bit PCM stays int16.
struct WavData {
    uint32_t sr = 0, ch = 0; size_t n = 0; bool is_i16 = false;
    std::vector<int16_t> i16; std::vector<float> f32;
};
WavData read_wav(const std::string &path);

} // namespace sgx
