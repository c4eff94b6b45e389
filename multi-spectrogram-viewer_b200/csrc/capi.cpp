// capi.cpp -- the extern "C" boundary declared in include/sgx.h.  Every entry point catches all
// exceptions and maps them to a status + thread-local message; nothing unwinds across the ABI.
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "engine.h"

using namespace sgx;

static thread_local std::string g_last_error;

// One handle = one engine per device.  sgx_mt_new / _ex: a single engine.  sgx_mt_new_sharded: an engine per listed
// device inside this process, track id t on engine t mod G (lib.rs:161-166 parallelises over tracks; here the tracks
// spread over GPUs), the engines joined by an NCCL communicator for the one exchange of the path (lib.rs:194-209).
struct sgx_multitrack {
    std::vector<std::unique_ptr<MultiTrack>> subs;
    bool sharded() const { return subs.size() > 1; }
    MultiTrack &one() // entry points that only make sense for a single engine
    {
        if (sharded()) throw Error(SGX_ERR_STATE, "not available on a multi-device handle (use one handle per device)");
        return *subs[0];
    }
    MultiTrack &of(size_t id) { return *subs[id % subs.size()]; }
    // the exchanges of all engines as one NCCL group (one host thread drives every device), then the commits
    void exchange_and_commit(bool force_commit)
    {
        if (sharded()) {
            nccl_check(nccl().GroupStart(), "ncclGroupStart");
            for (auto &e : subs) e->exchange();
            nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
            for (auto &e : subs) e->commit();
        } else {
            subs[0]->exchange();
            if (force_commit || subs[0]->world() > 1) subs[0]->commit();
        }
    }
    bool synchronize()
    {
        bool c = false;
        for (auto &e : subs) c = e->synchronize() || c;
        return c;
    }
};

// Entry points may switch the CUDA device of the calling thread (a handle lives on its own device, a multi-device
// handle walks several); the caller's current device is put back on the way out -- a host that also uses CUDA
// (PyTorch does) must not find its "current device" changed by a call into this library.
struct DeviceRestore {
    int prev = -1;
    DeviceRestore() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
    ~DeviceRestore() { if (prev >= 0) { int now = -1; if (cudaGetDevice(&now) == cudaSuccess && now != prev) cudaSetDevice(prev); } }
};

template <class F> static int guarded(F &&f)
{
    DeviceRestore restore;
    try {
        f();
        return SGX_OK;
    } catch (const Error &e) {
        g_last_error = e.what();
        return e.code;
    } catch (const std::bad_alloc &) {
        g_last_error = "out of host memory";
        return SGX_ERR_NOMEM;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return SGX_ERR_STATE;
    } catch (...) {
        g_last_error = "unknown error";
        return SGX_ERR_STATE;
    }
}

#define REQUIRE(cond, msg) do { if (!(cond)) throw Error(SGX_ERR_BAD_ARG, msg); } while (0)

extern "C" {

const char *sgx_last_error(void) { return g_last_error.c_str(); }

int sgx_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem)
{
    return guarded([&] {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0) { cudaGetLastError(); throw Error(SGX_ERR_CUDA, "no usable CUDA device"); }
        REQUIRE(device >= 0 && device < count, "device ordinal out of range");
        cudaDeviceProp p{};
        SGX_CUDA(cudaGetDeviceProperties(&p, device));
        if (sm_count) *sm_count = p.multiProcessorCount;
        if (cc_major) *cc_major = p.major;
        if (cc_minor) *cc_minor = p.minor;
        if (total_mem) *total_mem = p.totalGlobalMem;
    });
}

uint64_t sgx_kernel_launch_count(void) { return launch_count(); }

int sgx_host_pin(void *ptr, size_t bytes)
{
    return guarded([&] {
        REQUIRE(ptr && bytes, "NULL argument");
        SGX_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    });
}
int sgx_host_unpin(void *ptr)
{
    return guarded([&] { REQUIRE(ptr, "NULL argument"); SGX_CUDA(cudaHostUnregister(ptr)); });
}

void sgx_settings_default(sgx_settings *s)
{
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    s->win_ms = 40.0f; s->t_overlap = 4; s->f_overlap = 1; s->freq_scale = SGX_FREQ_MEL; s->db_range = 120.0f;
}

int sgx_mt_new(sgx_multitrack **out) { return sgx_mt_new_ex(nullptr, 0, nullptr, out); }

int sgx_mt_new_ex(const sgx_settings *settings, int device, void *cuda_stream, sgx_multitrack **out)
{
    return guarded([&] {
        REQUIRE(out, "out is NULL");
        *out = nullptr;
        sgx_settings s;
        if (settings) s = *settings; else sgx_settings_default(&s);
        REQUIRE(s.freq_scale == SGX_FREQ_LINEAR || s.freq_scale == SGX_FREQ_MEL, "freq_scale");
        std::unique_ptr<sgx_multitrack> h(new sgx_multitrack());
        h->subs.emplace_back(new MultiTrack(s, device, static_cast<cudaStream_t>(cuda_stream)));
        *out = h.release();
    });
}

int sgx_mt_new_sharded(const sgx_settings *settings, const int *devices, size_t n_devices, sgx_multitrack **out)
{
    return guarded([&] {
        REQUIRE(out, "out is NULL");
        *out = nullptr;
        sgx_settings s;
        if (settings) s = *settings; else sgx_settings_default(&s);
        REQUIRE(s.freq_scale == SGX_FREQ_LINEAR || s.freq_scale == SGX_FREQ_MEL, "freq_scale");
        std::vector<int> devs;
        if (devices && n_devices) devs.assign(devices, devices + n_devices);
        else { // every visible device
            int count = 0;
            if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); throw Error(SGX_ERR_CUDA, "no usable CUDA device (this engine has no CPU path)"); }
            for (int d = 0; d < count; ++d) devs.push_back(d);
        }
        for (size_t i = 0; i < devs.size(); ++i)
            for (size_t j = 0; j < i; ++j) REQUIRE(devs[i] != devs[j], "a device is listed twice");
        std::unique_ptr<sgx_multitrack> h(new sgx_multitrack());
        for (int d : devs) h->subs.emplace_back(new MultiTrack(s, d, nullptr));
        if (devs.size() > 1) {
            std::vector<NcclComm> comms(devs.size(), nullptr);
            nccl_check(nccl().CommInitAll(comms.data(), (int)devs.size(), devs.data()), "ncclCommInitAll");
            for (size_t g = 0; g < devs.size(); ++g) h->subs[g]->attach_comm(comms[g], (int)g, (int)devs.size(), true);
        }
        *out = h.release();
    });
}

int sgx_nccl_unique_id(uint8_t out[128])
{
    return guarded([&] {
        REQUIRE(out, "out is NULL");
        NcclUniqueId id;
        nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(out, id.internal, 128);
    });
}

int sgx_mt_attach_nccl(sgx_multitrack *mt, const uint8_t unique_id[128], int rank, int world)
{
    return guarded([&] {
        REQUIRE(mt && unique_id, "NULL argument");
        REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
        MultiTrack &e = mt->one();
        REQUIRE(e.world() == 1, "a communicator is already attached");
        SGX_CUDA(cudaSetDevice(e.device()));
        NcclUniqueId id;
        std::memcpy(id.internal, unique_id, 128);
        NcclComm comm = nullptr;
        nccl_check(nccl().CommInitRank(&comm, world, id, rank), "ncclCommInitRank"); // collective over all ranks
        e.attach_comm(comm, rank, world, true);
    });
}

int sgx_mt_get_device_count(sgx_multitrack *mt, int *n_devices, int *rank, int *world)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        if (n_devices) *n_devices = (int)mt->subs.size();
        if (rank) *rank = mt->sharded() ? 0 : mt->subs[0]->rank();
        if (world) *world = mt->sharded() ? 1 : mt->subs[0]->world();
    });
}

void sgx_mt_free(sgx_multitrack *mt)
{
    guarded([&] { delete mt; });
}

static int add_generic(sgx_multitrack *mt, const size_t *id_list, size_t n_ids, const void *const *pcm, int fmt,
                       const size_t *n_samples, const uint32_t *sr, const uint32_t *channels, bool on_device,
                       int *changed)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        REQUIRE(n_ids == 0 || (id_list && pcm && n_samples && sr && channels), "NULL argument");
        std::vector<size_t> ids(id_list, id_list + n_ids);
        std::vector<PcmSource> srcs(n_ids);
        for (size_t i = 0; i < n_ids; ++i)
            srcs[i] = PcmSource{pcm[i], fmt, n_samples[i], sr[i], channels[i], on_device, std::string()};
        const bool want = changed != nullptr || !on_device;
        bool c = false;
        if (mt->sharded()) {
            for (auto &e : mt->subs) e->analyse_owned(ids, srcs);
            mt->exchange_and_commit(true);
            if (want) c = mt->synchronize();
        } else {
            c = mt->subs[0]->add_tracks(ids, srcs, want);
        }
        if (changed) *changed = c ? 1 : 0;
    });
}

int sgx_mt_add_tracks_pcm(sgx_multitrack *mt, const size_t *id_list, size_t n_ids, const float *const *pcm,
                          const size_t *n_samples, const uint32_t *sr, const uint32_t *channels, int *changed)
{
    return add_generic(mt, id_list, n_ids, reinterpret_cast<const void *const *>(pcm), PCM_F32, n_samples, sr,
                       channels, false, changed);
}

int sgx_mt_add_tracks_pcm_i16(sgx_multitrack *mt, const size_t *id_list, size_t n_ids, const int16_t *const *pcm,
                              const size_t *n_samples, const uint32_t *sr, const uint32_t *channels, int *changed)
{
    return add_generic(mt, id_list, n_ids, reinterpret_cast<const void *const *>(pcm), PCM_I16, n_samples, sr,
                       channels, false, changed);
}

int sgx_mt_add_tracks_pcm_device(sgx_multitrack *mt, const size_t *id_list, size_t n_ids,
                                 const float *const *d_pcm, const size_t *n_samples, const uint32_t *sr,
                                 const uint32_t *channels, int *changed)
{
    return add_generic(mt, id_list, n_ids, reinterpret_cast<const void *const *>(d_pcm), PCM_F32, n_samples, sr,
                       channels, true, changed);
}

int sgx_mt_add_tracks(sgx_multitrack *mt, const size_t *id_list, size_t n_ids, const char *path_list, int *changed)
{
    return guarded([&] {
        REQUIRE(mt && path_list && (id_list || n_ids == 0), "NULL argument");
        // lib.rs:173: ids zipped with path_list.split("\n")
        std::vector<std::string> paths;
        const char *p = path_list;
        for (;;) {
            const char *nl = std::strchr(p, '\n');
            paths.emplace_back(nl ? std::string(p, nl) : std::string(p));
            if (!nl) break;
            p = nl + 1;
        }
        const size_t n = std::min(n_ids, paths.size()); // zip stops at the shorter
        std::vector<size_t> ids(id_list, id_list + n);
        // a rank of a sharded job only opens the files of the tracks it owns (t mod world == rank)
        std::vector<WavData> wavs(n);
        std::vector<PcmSource> srcs(n);
        for (size_t i = 0; i < n; ++i) {
            bool mine = false;
            for (auto &e : mt->subs) mine = mine || (mt->sharded() ? true : e->owns(ids[i]));
            if (!mine) { srcs[i] = PcmSource{nullptr, PCM_F32, 0, 0, 0, false, paths[i]}; continue; }
            wavs[i] = read_wav(paths[i]);
            const WavData &w = wavs[i];
            srcs[i] = PcmSource{w.is_i16 ? (const void *)w.i16.data() : (const void *)w.f32.data(),
                                w.is_i16 ? PCM_I16 : PCM_F32, w.n, w.sr, w.ch, false, paths[i]};
        }
        bool c = false;
        if (mt->sharded()) {
            for (auto &e : mt->subs) e->analyse_owned(ids, srcs);
            mt->exchange_and_commit(true);
            c = mt->synchronize();
        } else {
            c = mt->subs[0]->add_tracks(ids, srcs, true);
        }
        if (changed) *changed = c ? 1 : 0;
    });
}

int sgx_mt_add_track_slice_device(sgx_multitrack *mt, size_t id, const float *d_pcm, size_t chunk_offset,
                                  size_t chunk_len, size_t n_total, uint32_t sr, uint32_t channels,
                                  size_t frame_begin, size_t frame_count)
{
    return guarded([&] {
        REQUIRE(mt && d_pcm, "NULL argument");
        REQUIRE(n_total > 0 && chunk_len > 0, "empty slice");
        std::vector<size_t> ids{id};
        std::vector<PcmSource> srcs(1);
        srcs[0] = PcmSource{d_pcm, PCM_F32, chunk_len, sr, channels, true, std::string()};
        srcs[0].n_total = n_total; srcs[0].origin = chunk_offset;
        srcs[0].frame_begin = frame_begin; srcs[0].frame_count = frame_count;
        mt->one().add_tracks(ids, srcs, false, true);
    });
}

int sgx_mt_get_spec_image_slice_device(sgx_multitrack *mt, size_t id, float px_per_sec, uint32_t nheight,
                                       int channels, uint32_t ox_begin, uint32_t ox_count, uint8_t *d_out,
                                       size_t cap, size_t *written)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        uint8_t *outs[1] = {d_out};
        size_t caps[1] = {cap};
        std::vector<size_t> ids{id};
        mt->one().render(ids, px_per_sec, nheight, channels, d_out ? outs : nullptr, caps, written, &ox_begin, &ox_count);
    });
}

int sgx_slice_plan(size_t n_total, uint32_t sr, const sgx_settings *settings, float px_per_sec, uint32_t ox_begin,
                   uint32_t ox_count, size_t *frame_begin, size_t *frame_count, size_t *sample_begin,
                   size_t *sample_count)
{
    return guarded([&] {
        REQUIRE(frame_begin && frame_count && sample_begin && sample_count, "NULL argument");
        size_t win = 0, hop = 0, n_fft = 0;
        if (sgx_track_params(sr, settings, &win, &hop, &n_fft) != SGX_OK) throw Error(SGX_ERR_BAD_ARG, sgx_last_error());
        REQUIRE(n_fft >= win, "win_length > n_fft");
        const long T = stft_num_frames(n_total, win, hop);
        if (T <= 0) throw Error(SGX_ERR_BAD_ARG, "track shorter than one window");
        const uint32_t nwidth = calc_nwidth(px_per_sec, n_total, sr);
        REQUIRE(ox_count > 0 && ox_begin < nwidth && ox_count <= nwidth - ox_begin, "column window outside the image");
        uint32_t l0, r0, l1, r1;
        lanczos3_span((uint32_t)T, nwidth, ox_begin, &l0, &r0);
        lanczos3_span((uint32_t)T, nwidth, ox_begin + ox_count - 1, &l1, &r1);
        *frame_begin = l0; *frame_count = r1 - l0;
        // samples those frames read (zero-padded FFT frame, a few samples of slack so that the aligned bulk
        // copies of K1 stay inside the chunk), extended by what the reflection at either end of the track needs
        const long long n = (long long)n_total, pad_l = (long long)(n_fft - win) / 2;
        long long lo = (long long)l0 * (long long)hop - (long long)(win / 2) - pad_l - 8;
        long long hi = (long long)(r1 - 1) * (long long)hop - (long long)(win / 2) - pad_l + (long long)n_fft + 8;
        if (lo < 0) hi = std::max(hi, -lo + 1);
        if (hi > n) lo = std::min(lo, 2 * (n - 1) - (hi - 1));
        lo = std::max(0LL, std::min(lo, n)); hi = std::max(0LL, std::min(hi, n));
        *sample_begin = (size_t)lo; *sample_count = (size_t)(hi - lo);
    });
}

int sgx_mt_remove_track(sgx_multitrack *mt, size_t id, int *changed)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        bool c = false;
        if (mt->sharded()) {
            for (auto &e : mt->subs) e->drop(id); // the owner drops it, the others only refresh their local extrema
            mt->exchange_and_commit(true);
            if (changed) c = mt->synchronize();
        } else {
            c = mt->subs[0]->remove_track(id, changed != nullptr);
        }
        if (changed) *changed = c ? 1 : 0;
    });
}

static int spec_image_host(sgx_multitrack *mt, size_t id, float px_per_sec, uint32_t nheight, int channels,
                           uint8_t *out, size_t cap, size_t *written)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        const size_t need = (size_t)mt->of(id).image_width(id, px_per_sec) * nheight * channels;
        if (written) *written = need;
        if (!out) return;
        if (cap < need) throw Error(SGX_ERR_BUFFER, "output buffer too small");
        mt->of(id).render_host(id, px_per_sec, nheight, channels, out, need);
    });
}

int sgx_mt_get_spec_image(sgx_multitrack *mt, size_t id, float px_per_sec, uint32_t nheight, uint8_t *out,
                          size_t cap, size_t *written)
{
    return spec_image_host(mt, id, px_per_sec, nheight, 3, out, cap, written);
}

int sgx_mt_get_spec_image_rgba(sgx_multitrack *mt, size_t id, float px_per_sec, uint32_t nheight, uint8_t *out,
                               size_t cap, size_t *written)
{
    return spec_image_host(mt, id, px_per_sec, nheight, 4, out, cap, written);
}

int sgx_mt_get_spec_image_device(sgx_multitrack *mt, size_t id, float px_per_sec, uint32_t nheight, int channels,
                                 uint8_t *d_out, size_t cap, size_t *written)
{
    uint8_t *outs[1] = {d_out};
    size_t caps[1] = {cap};
    return sgx_mt_get_spec_images_device(mt, &id, 1, px_per_sec, nheight, channels, outs, caps, written);
}

int sgx_mt_get_spec_images_device(sgx_multitrack *mt, const size_t *id_list, size_t n_ids, float px_per_sec,
                                  uint32_t nheight, int channels, uint8_t *const *d_out, const size_t *cap,
                                  size_t *written)
{
    return guarded([&] {
        REQUIRE(mt && (id_list || n_ids == 0), "NULL argument");
        REQUIRE(!d_out || cap, "cap is NULL");
        std::vector<size_t> ids(id_list, id_list + n_ids);
        if (!mt->sharded()) { mt->subs[0]->render(ids, px_per_sec, nheight, channels, d_out, cap, written); return; }
        for (size_t g = 0; g < mt->subs.size(); ++g) { // every engine renders its own tracks into buffers of its own device
            std::vector<size_t> sub, at;
            for (size_t i = 0; i < n_ids; ++i) if (ids[i] % mt->subs.size() == g) { sub.push_back(ids[i]); at.push_back(i); }
            if (sub.empty()) continue;
            std::vector<uint8_t *> o(sub.size(), nullptr);
            std::vector<size_t> c(sub.size(), 0), w(sub.size(), 0);
            for (size_t k = 0; k < sub.size(); ++k) { if (d_out) o[k] = d_out[at[k]]; if (cap) c[k] = cap[at[k]]; }
            mt->subs[g]->render(sub, px_per_sec, nheight, channels, d_out ? o.data() : nullptr, c.data(), w.data());
            if (written) for (size_t k = 0; k < sub.size(); ++k) written[at[k]] = w[k];
        }
    });
}

static void images_async_all(sgx_multitrack *mt, const size_t *id_list, size_t n_ids, float px_per_sec, uint32_t nheight,
                             int channels, uint8_t *const *out, const size_t *cap, size_t *written)
{
    REQUIRE(mt && (id_list || n_ids == 0), "NULL argument");
    REQUIRE(!out || cap, "cap is NULL");
    REQUIRE(channels == 3 || channels == 4, "channels must be 3 or 4");
    std::vector<size_t> ids(id_list, id_list + n_ids);
    for (size_t g = 0; g < mt->subs.size(); ++g) {
        std::vector<size_t> sub, at;
        for (size_t i = 0; i < n_ids; ++i) if (ids[i] % mt->subs.size() == g) { sub.push_back(ids[i]); at.push_back(i); }
        if (sub.empty()) continue;
        std::vector<uint8_t *> o(sub.size(), nullptr);
        std::vector<size_t> c(sub.size(), 0), w(sub.size(), 0);
        for (size_t k = 0; k < sub.size(); ++k) { if (out) o[k] = out[at[k]]; if (cap) c[k] = cap[at[k]]; }
        mt->subs[g]->images_async(sub, px_per_sec, nheight, channels, out ? o.data() : nullptr, c.data(), w.data());
        if (written) for (size_t k = 0; k < sub.size(); ++k) written[at[k]] = w[k];
    }
}

int sgx_mt_get_spec_images_async(sgx_multitrack *mt, const size_t *id_list, size_t n_ids, float px_per_sec,
                                 uint32_t nheight, int channels, uint8_t *const *out, const size_t *cap,
                                 size_t *written)
{
    return guarded([&] { images_async_all(mt, id_list, n_ids, px_per_sec, nheight, channels, out, cap, written); });
}

int sgx_mt_wait_images(sgx_multitrack *mt)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        for (auto &e : mt->subs) e->wait_images();
    });
}

int sgx_mt_get_spec_images(sgx_multitrack *mt, const size_t *id_list, size_t n_ids, float px_per_sec,
                           uint32_t nheight, int channels, uint8_t *const *out, const size_t *cap, size_t *written)
{
    return guarded([&] {
        images_async_all(mt, id_list, n_ids, px_per_sec, nheight, channels, out, cap, written);
        for (auto &e : mt->subs) e->wait_images();
    });
}

int sgx_mt_get_wav_image(sgx_multitrack *mt, size_t id, float px_per_sec, uint32_t nheight, float amp_min,
                         float amp_max, uint8_t *out, size_t cap, size_t *written)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        const size_t need = (size_t)mt->of(id).image_width(id, px_per_sec) * nheight * 4;
        if (written) *written = need;
        if (!out) return;
        if (cap < need) throw Error(SGX_ERR_BUFFER, "output buffer too small");
        std::vector<uint8_t> img = mt->of(id).wav_image(id, px_per_sec, nheight, amp_min, amp_max);
        if (!img.empty()) std::memcpy(out, img.data(), img.size());
    });
}

int sgx_mt_get_frequency_hz(sgx_multitrack *mt, size_t id, float relative_freq, float *out)
{
    return guarded([&] { REQUIRE(mt && out, "NULL argument"); *out = mt->of(id).frequency_hz(id, relative_freq); });
}
int sgx_mt_get_max_db(sgx_multitrack *mt, float *out)
{
    return guarded([&] { REQUIRE(mt && out, "NULL argument"); *out = mt->subs[0]->max_db(); });
}
int sgx_mt_get_min_db(sgx_multitrack *mt, float *out)
{
    return guarded([&] { REQUIRE(mt && out, "NULL argument"); *out = mt->subs[0]->min_db(); });
}
int sgx_mt_get_max_sec(sgx_multitrack *mt, float *out)
{
    return guarded([&] { REQUIRE(mt && out, "NULL argument"); float m = 0.0f; for (auto &e : mt->subs) m = std::max(m, e->max_sec()); *out = m; });
}
int sgx_mt_get_sec(sgx_multitrack *mt, size_t id, float *out)
{
    return guarded([&] {
        REQUIRE(mt && out, "NULL argument");
        const Track &t = mt->of(id).track(id);
        *out = (float)t.n / (float)t.sr; // lib.rs:338
    });
}
int sgx_mt_get_sr(sgx_multitrack *mt, size_t id, uint32_t *out)
{
    return guarded([&] { REQUIRE(mt && out, "NULL argument"); *out = mt->of(id).track(id).sr; });
}

static void copy_string(const std::string &s, char *out, size_t cap, size_t *written)
{
    if (written) *written = s.size() + 1;
    if (!out) return;
    if (cap < s.size() + 1) throw Error(SGX_ERR_BUFFER, "string buffer too small");
    std::memcpy(out, s.c_str(), s.size() + 1);
}
int sgx_mt_get_path(sgx_multitrack *mt, size_t id, char *out, size_t cap, size_t *written)
{
    return guarded([&] { REQUIRE(mt, "handle is NULL"); copy_string(mt->of(id).track(id).path, out, cap, written); });
}
int sgx_mt_get_filename(sgx_multitrack *mt, size_t id, char *out, size_t cap, size_t *written)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        const std::string &p = mt->of(id).track(id).path;
        const size_t slash = p.find_last_of('/');
        copy_string(slash == std::string::npos ? p : p.substr(slash + 1), out, cap, written);
    });
}

int sgx_get_colormap(uint8_t out[30])
{
    static const uint8_t cm[30] = {0, 0, 4, 27, 12, 65, 74, 12, 107, 120, 28, 109, 165, 44, 96,
                                   207, 68, 70, 237, 105, 37, 251, 155, 6, 247, 209, 61, 252, 255, 164};
    if (!out) { g_last_error = "out is NULL"; return SGX_ERR_BAD_ARG; }
    std::memcpy(out, cm, 30);
    return SGX_OK;
}

int sgx_mt_get_spec_shape(sgx_multitrack *mt, size_t id, size_t *n_frames, size_t *n_out)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        const Track &t = mt->of(id).track(id);
        if (n_frames) *n_frames = t.n_frames;
        if (n_out) *n_out = t.n_out;
    });
}

int sgx_mt_get_spec_db(sgx_multitrack *mt, size_t id, float *out, size_t cap_elems, size_t *written_elems)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        MultiTrack &e = mt->of(id);
        const Track &t = e.track(id);
        SGX_CUDA(cudaSetDevice(e.device()));
        const size_t need = t.n_frames * t.n_out;
        if (written_elems) *written_elems = need;
        if (!out) return;
        if (cap_elems < need) throw Error(SGX_ERR_BUFFER, "output buffer too small");
        SGX_CUDA(cudaMemcpyAsync(out, t.spec.p, need * sizeof(float), cudaMemcpyDeviceToHost, e.stream()));
        SGX_CUDA(cudaStreamSynchronize(e.stream()));
    });
}

int sgx_mt_get_image_width(sgx_multitrack *mt, size_t id, float px_per_sec, uint32_t *nwidth)
{
    return guarded([&] { REQUIRE(mt && nwidth, "NULL argument"); *nwidth = mt->of(id).image_width(id, px_per_sec); });
}

int sgx_mt_range_device_ptr(sgx_multitrack *mt, float **d_max_negmin)
{
    return guarded([&] { REQUIRE(mt && d_max_negmin, "NULL argument"); *d_max_negmin = mt->one().range_device_ptr(); });
}
int sgx_mt_commit_range_device(sgx_multitrack *mt)
{
    return guarded([&] { REQUIRE(mt, "handle is NULL"); mt->one().commit_range_device(); });
}
int sgx_mt_set_global_max_sr(sgx_multitrack *mt, uint32_t max_sr)
{
    return guarded([&] { REQUIRE(mt, "handle is NULL"); for (auto &e : mt->subs) e->set_global_max_sr(max_sr); });
}
int sgx_mt_set_profiling(sgx_multitrack *mt, int on)
{
    return guarded([&] { REQUIRE(mt, "handle is NULL"); for (auto &e : mt->subs) e->set_profiling(on != 0); });
}
int sgx_mt_get_stage_times(sgx_multitrack *mt, float *analysis_ms, float *render_ms)
{
    return guarded([&] {
        REQUIRE(mt && analysis_ms && render_ms, "NULL argument");
        *analysis_ms = *render_ms = -1.0f;
        for (auto &e : mt->subs) { // the slowest device
            float a = -1.0f, r = -1.0f;
            e->stage_times(&a, &r);
            *analysis_ms = std::max(*analysis_ms, a); *render_ms = std::max(*render_ms, r);
        }
    });
}
int sgx_mt_synchronize(sgx_multitrack *mt, int *changed)
{
    return guarded([&] {
        REQUIRE(mt, "handle is NULL");
        const bool c = mt->synchronize();
        if (changed) *changed = c ? 1 : 0;
    });
}

// ---- surface 2 ------------------------------------------------------------------------------------------
size_t sgx_calc_proper_n_fft(size_t win_length) { return calc_proper_n_fft(win_length); }

int sgx_track_params(uint32_t sr, const sgx_settings *s, size_t *win_length, size_t *hop_length, size_t *n_fft)
{
    return guarded([&] {
        REQUIRE(win_length && hop_length && n_fft, "NULL argument");
        sgx_settings d;
        if (s) d = *s; else sgx_settings_default(&d);
        REQUIRE(d.t_overlap > 0 && d.f_overlap > 0, "overlap factors must be positive");
        const float w0 = d.win_ms * (float)sr / 1000.0f;
        size_t h = (size_t)std::round(w0 / (float)d.t_overlap);
        if (d.hop_length) h = d.hop_length;
        size_t w = h * d.t_overlap;
        if (d.win_length) w = d.win_length;
        size_t f = calc_proper_n_fft(w) * d.f_overlap;
        if (d.n_fft) f = d.n_fft;
        *win_length = w; *hop_length = h; *n_fft = f;
    });
}

int sgx_hann(size_t size, int symmetric, float *out)
{
    return guarded([&] { REQUIRE(out || size == 0, "out is NULL"); hann(size, symmetric != 0, out); });
}
int sgx_calc_window(size_t win_length, size_t n_fft, float *out)
{
    return guarded([&] { REQUIRE(out || win_length == 0, "out is NULL"); calc_window(win_length, n_fft, out); });
}
float sgx_hz_to_mel(float hz) { return hz_to_mel(hz); }
float sgx_mel_to_hz(float mel) { return mel_to_hz(mel); }

int sgx_calc_mel_fb(uint32_t sr, size_t n_fft, size_t n_mel, float fmin, float fmax, int do_norm, float *out)
{
    return guarded([&] {
        REQUIRE(out, "out is NULL");
        REQUIRE(n_fft % 2 == 0, "n_fft must be even (mel.rs:50)");
        REQUIRE(n_mel != 0, "n_mel must not be 0 (mel.rs:51)");
        calc_mel_fb(sr, n_fft, n_mel, fmin, fmax, do_norm != 0, out);
    });
}

int sgx_calc_mel_fb_default(uint32_t sr, size_t n_fft, float *out, size_t cap_elems, size_t *n_mel)
{
    return guarded([&] {
        REQUIRE(n_mel, "n_mel is NULL");
        REQUIRE(n_fft % 2 == 0 && n_fft >= 2, "n_fft must be even (mel.rs:50)");
        std::vector<float> fb;
        *n_mel = calc_mel_fb_default(sr, n_fft, fb);
        if (out) {
            if (cap_elems < fb.size()) throw Error(SGX_ERR_BUFFER, "filterbank buffer too small");
            std::memcpy(out, fb.data(), fb.size() * sizeof(float));
        }
    });
}

long sgx_stft_num_frames(size_t n, size_t win_length, size_t hop_length)
{
    return stft_num_frames(n, win_length, hop_length);
}

static int stft_stage(int mode, const float *input, size_t n, size_t win, size_t hop, size_t n_fft,
                      const float *window, const float *mel_fb, size_t n_mel, float *out, size_t cap,
                      size_t *n_frames)
{
    return guarded([&] {
        REQUIRE(input, "input is NULL");
        const StageOut so = stage_stft(mode, input, n, win, hop, n_fft, window, mel_fb, n_mel, out, cap);
        if (n_frames) *n_frames = so.n_frames;
    });
}

int sgx_perform_stft(const float *input, size_t n, size_t win_length, size_t hop_length, size_t n_fft,
                     const float *window, float *out, size_t cap_elems, size_t *n_frames)
{
    return stft_stage(MODE_COMPLEX, input, n, win_length, hop_length, n_fft, window, nullptr, 0, out, cap_elems, n_frames);
}
int sgx_stft_magnitude(const float *input, size_t n, size_t win_length, size_t hop_length, size_t n_fft,
                       const float *window, float *out, size_t cap_elems, size_t *n_frames)
{
    return stft_stage(MODE_MAG, input, n, win_length, hop_length, n_fft, window, nullptr, 0, out, cap_elems, n_frames);
}
int sgx_melspectrogram_db(const float *input, size_t n, size_t win_length, size_t hop_length, size_t n_fft,
                          const float *window, const float *mel_fb, size_t n_mel, float *out, size_t cap_elems,
                          size_t *n_frames)
{
    return stft_stage(mel_fb ? MODE_MEL_DB : MODE_LIN_DB, input, n, win_length, hop_length, n_fft, window, mel_fb,
                      n_mel, out, cap_elems, n_frames);
}

int sgx_amp_to_db_default(float *x, size_t n)
{
    return guarded([&] { REQUIRE(x || n == 0, "x is NULL"); stage_amp_to_db(x, n); });
}

int sgx_spec_to_grey(const float *spec, size_t n_frames, size_t n_out, float up_ratio, float max_db, float min_db,
                     float *grey, size_t cap_elems, uint32_t *height)
{
    return guarded([&] {
        REQUIRE(spec || !grey, "spec is NULL");
        const uint32_t h = stage_spec_to_grey(spec, n_frames, n_out, up_ratio, max_db, min_db, grey, cap_elems);
        if (height) *height = h;
    });
}

int sgx_grey_to_rgb(const float *grey, uint32_t width, uint32_t height, uint32_t nwidth, uint32_t nheight,
                    int channels, uint8_t *out, size_t cap)
{
    return guarded([&] {
        REQUIRE(grey && out, "NULL argument");
        stage_grey_to_rgb(grey, width, height, nwidth, nheight, channels, out, cap);
    });
}

int sgx_wav_to_image(const float *wav, size_t n, uint32_t nwidth, uint32_t nheight, float amp_min, float amp_max,
                     uint8_t *out, size_t cap)
{
    return guarded([&] {
        REQUIRE(wav && out, "NULL argument");
        stage_wav_to_image(wav, n, nwidth, nheight, amp_min, amp_max, out, cap);
    });
}

int sgx_open_wav(const char *path, float *out, size_t cap_elems, size_t *n_samples, uint32_t *channels, uint32_t *sr)
{
    return guarded([&] {
        REQUIRE(path, "path is NULL");
        const WavData w = read_wav(path);
        if (n_samples) *n_samples = w.n;
        if (channels) *channels = w.ch;
        if (sr) *sr = w.sr;
        if (!out) return;
        const size_t total = w.n * w.ch;
        if (cap_elems < total) throw Error(SGX_ERR_BUFFER, "sample buffer too small");
        if (w.is_i16) for (size_t i = 0; i < total; ++i) out[i] = (float)w.i16[i] / 32768.0f; // audio.rs:16-19
        else std::memcpy(out, w.f32.data(), total * sizeof(float));
    });
}

} // extern "C"
