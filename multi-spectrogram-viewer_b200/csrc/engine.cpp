// engine.cpp -- host orchestration above the kernels (no arithmetic of the path happens here
// except the once-per-sample-rate tables of host_tables.cpp).
#include "engine.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace sgx {

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

cudaError_t ensure_dynamic_smem(const void *func, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> granted;
    if (bytes <= 48 * 1024) return cudaSuccess; // the default limit needs no opt-in
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    size_t &g = granted[std::make_pair(func, dev)];
    if (bytes <= g) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) g = bytes;
    return e;
}

void cuda_check(cudaError_t e, const char *what, const char *file, int line)
{
    if (e == cudaSuccess) return;
    cudaGetLastError(); // clear the sticky non-fatal error state
    const char *base = std::strrchr(file, '/');
    throw Error(SGX_ERR_CUDA, std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " in " +
                                  what + " (" + (base ? base + 1 : file) + ":" + std::to_string(line) + ")");
}

// ---------------------------------------------------------------------------------------------------
// DeviceCtx
// ---------------------------------------------------------------------------------------------------
static std::mutex g_ctx_mutex;
static std::map<int, std::unique_ptr<DeviceCtx>> g_ctx;

DeviceCtx::DeviceCtx(int d) : device_(d)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        throw Error(SGX_ERR_CUDA, "no usable CUDA device (this engine has no CPU path)");
    }
    if (d < 0 || d >= count) throw Error(SGX_ERR_BAD_ARG, "CUDA device ordinal out of range");
    cudaDeviceProp prop{};
    SGX_CUDA(cudaGetDeviceProperties(&prop, d));
    if (prop.major < 10)
        throw Error(SGX_ERR_CUDA, std::string("kernels are built for sm_100a only; device is ") + prop.name);
    sm_count = prop.multiProcessorCount;
}

DeviceCtx &DeviceCtx::get(int device)
{
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    auto it = g_ctx.find(device);
    if (it == g_ctx.end()) it = g_ctx.emplace(device, std::unique_ptr<DeviceCtx>(new DeviceCtx(device))).first;
    return *it->second;
}

FftPlan &DeviceCtx::plan(size_t n_fft)
{
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    auto it = plans_.find(n_fft);
    if (it != plans_.end()) return *it->second;
    std::unique_ptr<FftPlan> p(new FftPlan());
    if (!stft_config_for(n_fft, &p->cfg))
        throw Error(SGX_ERR_BAD_ARG, "n_fft must be a power of two in [2, 16384] (rustfft Radix4, realfft.rs:94)");
    if (!p->cfg.generic) {
        const int h = p->cfg.h;
        std::vector<float2> tw(h), sp(h / 2 + 1);
        make_fft_tables(h, tw.data(), sp.data());
        p->tw.alloc(h); p->split.alloc(h / 2 + 1);
        SGX_CUDA(cudaMemcpy(p->tw.p, tw.data(), sizeof(float2) * h, cudaMemcpyHostToDevice));
        SGX_CUDA(cudaMemcpy(p->split.p, sp.data(), sizeof(float2) * (h / 2 + 1), cudaMemcpyHostToDevice));
        const std::vector<float2> pt = make_fft_pass_tables(h, p->cfg.pts);
        p->twr.alloc(pt.size());
        SGX_CUDA(cudaMemcpy(p->twr.p, pt.data(), sizeof(float2) * pt.size(), cudaMemcpyHostToDevice));
    }
    it = plans_.emplace(n_fft, std::move(p)).first;
    return *it->second;
}

// ---------------------------------------------------------------------------------------------------
// shared helper: upload tables for one (win, n_fft, filterbank)
// ---------------------------------------------------------------------------------------------------
static void fill_tables(TrackTables &tt, size_t win, size_t n_fft, const float *window,
                        const float *mel_fb, size_t n_mel, const StftConfig &cfg, cudaStream_t s)
{
    tt.win = win; tt.n_fft = n_fft; tt.n_mel = mel_fb ? n_mel : 0;
    std::vector<float> wf(n_fft, 0.0f);
    const size_t pad_l = (n_fft - win) / 2; // lib.rs:400
    for (size_t j = 0; j < win; ++j) wf[pad_l + j] = window[j];
    tt.win_f.upload(wf.data(), n_fft, s);
    if (mel_fb) {
        const int tpg = cfg.generic ? 1 : cfg.h / cfg.pts;
        // capacity of the imaginary plane of one group's exchange buffer (floats), where K1 stages the bank
        const size_t plane = cfg.generic ? 0 : cfg.fft_smem / sizeof(float) / (2 * (size_t)cfg.groups);
        // the warp kernel walks all blocks with one warp and reads magnitude PAIRS; the block kernel spreads them
        // over the warps of a group and reads vectors of cfg.vec frames
        MelBands mb = make_mel_bands(mel_fb, n_fft / 2 + 1, n_mel, tpg, plane, cfg.warp1 ? 1 : (cfg.warp2 ? 2 : cfg.vec));
        if (cfg.generic) mb.log2_split = 0;
        std::vector<int> meta(4 * n_mel, 0); // {lo, cnt, off, 0} per filter: one 16-byte load on the device
        for (size_t m = 0; m < n_mel; ++m) { meta[4 * m] = mb.lo[m]; meta[4 * m + 1] = mb.cnt[m]; meta[4 * m + 2] = mb.off[m]; }
        tt.mel_lo.upload(meta.data(), meta.size(), s);
        tt.mel_cnt.upload(mb.sched.data(), mb.sched.size(), s);
        tt.mel_w.upload(mb.w.data(), mb.w.size(), s);
        tt.mel_log2p = mb.log2_split;
        tt.mel_nnz = (int)mb.w.size();
        tt.segp.upload(mb.seg.data(), mb.seg.size(), s);
        tt.seg_nwq = mb.seg_nwq; tt.seg_nblk = mb.seg_nblk; tt.seg_log2p = mb.seg_log2p; tt.seg_words = (int)mb.seg.size();
    }
    SGX_CUDA(cudaStreamSynchronize(s)); // host vectors die here
}

static StftTrack make_desc(const void *d_pcm, int fmt, size_t n, uint32_t ch, size_t win, size_t hop,
                           size_t n_fft, size_t T, const TrackTables &tt, float *out, size_t n_out,
                           unsigned *slot, size_t origin = 0, size_t avail = (size_t)-1, size_t frame0 = 0)
{
    StftTrack d{};
    d.pcm = d_pcm; d.n = (long long)n; d.ch = (int)ch; d.fmt = fmt;
    d.origin = (long long)origin; d.avail = (long long)(avail == (size_t)-1 ? n : avail); d.frame0 = (int)frame0;
    d.win = (int)win; d.hop = (int)hop; d.pad_l = (int)((n_fft - win) / 2);
    d.n_frames = (int)T; d.win_f = tt.win_f.p; d.out = out; d.n_out = (int)n_out;
    d.mel_lo = tt.mel_lo.p; d.mel_cnt = tt.mel_cnt.p; d.mel_off = tt.mel_off.p; d.mel_w = tt.mel_w.p;
    d.mel_log2p = tt.mel_log2p; d.range_slot = slot; d.tile_begin = 0;
    d.segp = tt.seg_nblk > 0 ? tt.segp.p : nullptr; d.seg_nwq = tt.seg_nwq; d.seg_nblk = tt.seg_nblk;
    d.seg_log2p = tt.seg_log2p; d.seg_words = tt.seg_words;
    return d;
}

static void check_stft_args(size_t n, size_t win, size_t hop, size_t n_fft, long *T)
{
    if (n_fft < 2 || (n_fft & (n_fft - 1)) != 0 || n_fft > 16384)
        throw Error(SGX_ERR_BAD_ARG, "n_fft must be a power of two in [2, 16384]");
    if (win > n_fft) throw Error(SGX_ERR_BAD_ARG, "win_length > n_fft (usize underflow at lib.rs:400)");
    if (hop > (size_t)1 << 24) throw Error(SGX_ERR_BAD_ARG, "hop_length too large");
    *T = stft_num_frames(n, win, hop);
    if (*T < 0)
        throw Error(SGX_ERR_BAD_ARG, "input shorter than the window / reflect padding (the reference panics on its slices, lib.rs:412-433)");
    if (*T > 0x7fffffffL / 2) throw Error(SGX_ERR_BAD_ARG, "too many frames");
}

// ---------------------------------------------------------------------------------------------------
// MultiTrack
// ---------------------------------------------------------------------------------------------------
MultiTrack::MultiTrack(const sgx_settings &s, int device, cudaStream_t stream)
    : set_(s), device_(device), stream_(stream), max_db_(-INFINITY), min_db_(INFINITY)
{
    if (!(s.win_ms > 0.0f) || s.t_overlap == 0 || s.f_overlap == 0)
        throw Error(SGX_ERR_BAD_ARG, "settings: win_ms, t_overlap, f_overlap must be positive");
    ctx_ = &DeviceCtx::get(device);
    SGX_CUDA(cudaSetDevice(device));
    if (!stream_) { SGX_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking)); own_stream_ = true; }
    n_slots_ = 1024;
    slots_.alloc((size_t)n_slots_ * 2);
    SGX_CUDA(launch_range_init(slots_.p, n_slots_, stream_));
    for (int i = n_slots_ - 1; i >= 0; --i) free_slots_.push_back(i);
    d_local_.alloc(4); d_state_.alloc(8);
    const float st[8] = {-INFINITY, INFINITY, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f}; // lib.rs:104-105
    const float lc[4] = {-INFINITY, -INFINITY, 0.0f, 0.0f};
    SGX_CUDA(cudaMemcpyAsync(d_state_.p, st, sizeof(st), cudaMemcpyHostToDevice, stream_));
    SGX_CUDA(cudaMemcpyAsync(d_local_.p, lc, sizeof(lc), cudaMemcpyHostToDevice, stream_));
    SGX_CUDA(cudaStreamSynchronize(stream_));
}

MultiTrack::~MultiTrack()
{
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    tracks_.clear(); tables_.clear(); axis_.clear();
    for (auto &e : ev_) if (e) cudaEventDestroy(e);
    for (auto &e : copy_events_) cudaEventDestroy(e);
    if (copy_stream_) { cudaStreamSynchronize(copy_stream_); cudaStreamDestroy(copy_stream_); }
    if (out_stream_) { cudaStreamSynchronize(out_stream_); cudaStreamDestroy(out_stream_); }
    if (out_ev_) cudaEventDestroy(out_ev_);
    if (out_done_) cudaEventDestroy(out_done_);
    d_imgs_.clear();
    if (comm_ && comm_owned_) { try { nccl().CommDestroy(comm_); } catch (...) {} }
    if (own_stream_) cudaStreamDestroy(stream_);
}

void MultiTrack::derive_params(uint32_t sr, size_t *win, size_t *hop, size_t *n_fft) const
{
    // lib.rs:43-46
    const float w0 = set_.win_ms * (float)sr / 1000.0f;
    size_t h = (size_t)std::round(w0 / (float)set_.t_overlap);
    if (set_.hop_length) h = set_.hop_length;
    size_t w = h * set_.t_overlap;
    if (set_.win_length) w = set_.win_length;
    size_t f = calc_proper_n_fft(w) * set_.f_overlap;
    if (set_.n_fft) f = set_.n_fft;
    *win = w; *hop = h; *n_fft = f;
}

TrackTables *MultiTrack::tables_for(uint32_t sr, size_t win, size_t n_fft)
{
    const auto key = std::make_tuple(sr, win, n_fft, set_.n_mel);
    auto it = tables_.find(key);
    if (it != tables_.end()) return it->second.get();
    std::unique_ptr<TrackTables> tt(new TrackTables());
    std::vector<float> window(win);
    calc_window(win, n_fft, window.data()); // lib.rs:138-140
    std::vector<float> fb;
    size_t n_mel = 0;
    if (set_.freq_scale == SGX_FREQ_MEL) {
        if (set_.n_mel) {
            n_mel = set_.n_mel;
            fb.assign((n_fft / 2 + 1) * n_mel, 0.0f);
            calc_mel_fb(sr, n_fft, n_mel, 0.0f, -1.0f, true, fb.data());
        } else {
            n_mel = calc_mel_fb_default(sr, n_fft, fb); // lib.rs:151-158
            if (n_mel == 0) throw Error(SGX_ERR_BAD_ARG, "default mel filterbank is empty for this sr / n_fft");
        }
    }
    fill_tables(*tt, win, n_fft, window.data(), n_mel ? fb.data() : nullptr, n_mel, ctx_->plan(n_fft).cfg, stream_);
    TrackTables *raw = tt.get();
    tables_.emplace(key, std::move(tt));
    return raw;
}

int MultiTrack::alloc_slot()
{
    if (free_slots_.empty()) {
        // grow: new array, identity-initialised, old contents copied
        const int nn = n_slots_ * 2;
        DevBuf<unsigned> ns;
        ns.alloc((size_t)nn * 2);
        SGX_CUDA(launch_range_init(ns.p, nn, stream_));
        SGX_CUDA(cudaMemcpyAsync(ns.p, slots_.p, sizeof(unsigned) * 2 * n_slots_, cudaMemcpyDeviceToDevice, stream_));
        SGX_CUDA(cudaStreamSynchronize(stream_));
        slots_ = std::move(ns);
        for (int i = nn - 1; i >= n_slots_; --i) free_slots_.push_back(i);
        n_slots_ = nn;
    }
    const int s = free_slots_.back();
    free_slots_.pop_back();
    return s;
}

void MultiTrack::drop_track(size_t id)
{
    auto it = tracks_.find(id);
    if (it == tracks_.end()) return;
    // buffers may still be in use by enqueued work
    SGX_CUDA(cudaStreamSynchronize(stream_));
    if (it->second.slot >= 0) {
        SGX_CUDA(launch_range_init(slots_.p + 2 * it->second.slot, 1, stream_));
        free_slots_.push_back(it->second.slot);
    }
    tracks_.erase(it);
}

void MultiTrack::reduce_local()
{
    // lib.rs:220-224 / 178-182: metadata maxima over ALL tracks this handle holds, recomputed on every add and remove
    uint32_t msr = 0;
    for (auto &kv : tracks_) msr = std::max(msr, kv.second.sr);
    if (msr != max_sr_) { max_sr_ = msr; changed_acc_ = true; }
    // a lone handle whose caller waits for `changed` commits in the same launch (one launch less per add / remove)
    const bool fuse = fuse_commit_ && !comm_;
    SGX_CUDA(launch_range_reduce(slots_.p, n_slots_, d_local_.p, (float)max_sr_, max_sec_, set_.db_range, fuse ? d_state_.p : nullptr, stream_));
    committed_in_reduce_ = fuse;
    if (fuse) pending_ = true;
}

void MultiTrack::attach_comm(NcclComm comm, int rank, int world, bool owned)
{
    if (world < 1 || rank < 0 || rank >= world) throw Error(SGX_ERR_BAD_ARG, "attach: bad rank / world");
    if (!tracks_.empty()) throw Error(SGX_ERR_STATE, "attach a communicator before the first track is added");
    comm_ = comm; comm_owned_ = owned; rank_ = rank; world_ = world;
}

void MultiTrack::exchange()
{
    if (!comm_) return;
    SGX_CUDA(cudaSetDevice(device_));
    // {max, -min, max_sr, max_sec}: 16 bytes, MAX, in place, on the stream that carries K1 and K3 -- the render that
    // follows is ordered behind it on the device; no host round trip (lib.rs:194-209 across shards)
    nccl_check(nccl().AllReduce(d_local_.p, d_local_.p, 4, kNcclFloat32, kNcclMax, comm_, stream_), "ncclAllReduce");
}

void MultiTrack::commit()
{
    SGX_CUDA(cudaSetDevice(device_));
    SGX_CUDA(launch_range_commit(d_local_.p, set_.db_range, d_state_.p, stream_));
    pending_ = true;
}

void MultiTrack::commit_range_device() { commit(); }

bool MultiTrack::synchronize()
{
    SGX_CUDA(cudaSetDevice(device_));
    if (pending_) {
        float st[8];
        SGX_CUDA(cudaMemcpyAsync(st, d_state_.p, sizeof(st), cudaMemcpyDeviceToHost, stream_));
        SGX_CUDA(cudaStreamSynchronize(stream_));
        max_db_ = st[0]; min_db_ = st[1];
        if (comm_) { // what the other shards hold (lib.rs:220-224, 178-182)
            const uint32_t gsr = (uint32_t)st[3];
            if (!global_sr_fixed_ && gsr != global_max_sr_) { global_max_sr_ = gsr; changed_acc_ = true; }
            global_max_sec_ = st[4];
        }
        if (st[2] != 0.0f) {
            changed_acc_ = true;
            const float zero = 0.0f;
            SGX_CUDA(cudaMemcpyAsync(d_state_.p + 2, &zero, sizeof(float), cudaMemcpyHostToDevice, stream_));
            SGX_CUDA(cudaStreamSynchronize(stream_));
        }
        pending_ = false;
    } else {
        SGX_CUDA(cudaStreamSynchronize(stream_));
    }
    const bool c = changed_acc_;
    changed_acc_ = false;
    return c;
}

uint32_t MultiTrack::effective_max_sr() const { return global_max_sr_ ? std::max(global_max_sr_, max_sr_) : max_sr_; }

bool MultiTrack::add_tracks(const std::vector<size_t> &ids, std::vector<PcmSource> &srcs, bool want_changed, bool sliced)
{
    if (ids.size() != srcs.size()) throw Error(SGX_ERR_BAD_ARG, "id_list and track list differ in length");
    fuse_commit_ = want_changed;
    if (sliced) analyse(ids, srcs); else analyse_owned(ids, srcs);
    fuse_commit_ = false;
    exchange();
    // without a communicator and without a waiting caller the range stays uncommitted: a driver that shards by its
    // own means all-reduces range_device_ptr() in-stream and calls commit_range_device()
    if ((want_changed || comm_) && !committed_in_reduce_) commit();
    if (!want_changed) return false;
    return synchronize();
}

void MultiTrack::analyse_owned(const std::vector<size_t> &ids, std::vector<PcmSource> &srcs)
{
    if (ids.size() != srcs.size()) throw Error(SGX_ERR_BAD_ARG, "id_list and track list differ in length");
    if (world_ <= 1) { analyse(ids, srcs); return; }
    std::vector<size_t> mine; // track t lives where t % world == rank
    std::vector<PcmSource> msrc;
    for (size_t i = 0; i < ids.size(); ++i) if (owns(ids[i])) { mine.push_back(ids[i]); msrc.push_back(srcs[i]); }
    analyse(mine, msrc);
}

void MultiTrack::analyse(const std::vector<size_t> &ids, std::vector<PcmSource> &srcs)
{
    SGX_CUDA(cudaSetDevice(device_));
    if (ids.size() != srcs.size()) throw Error(SGX_ERR_BAD_ARG, "id_list and track list differ in length");
    // ---- validate everything before touching state (atomic, unlike lib.rs:174-189) -------------------
    struct Pre { size_t win, hop, n_fft; long T; };
    std::vector<Pre> pre(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {
        const PcmSource &s = srcs[i];
        if (!s.data || s.n == 0 || s.ch == 0 || s.sr == 0) throw Error(SGX_ERR_BAD_ARG, "empty track");
        derive_params(s.sr, &pre[i].win, &pre[i].hop, &pre[i].n_fft);
        check_stft_args(s.n_total ? s.n_total : s.n, pre[i].win, pre[i].hop, pre[i].n_fft, &pre[i].T);
        if (s.n_total) { // time slice of a longer track
            if (s.origin + s.n > s.n_total || s.frame_count == 0 || s.frame_begin + s.frame_count > (size_t)pre[i].T)
                throw Error(SGX_ERR_BAD_ARG, "time slice outside the track");
            // the chunk must hold every sample its frames read, reflections at the ends of the track included
            // (the loader clamps silently otherwise); same arithmetic as sgx_slice_plan without its slack
            const long long n = (long long)s.n_total, pad_l = (long long)(pre[i].n_fft - pre[i].win) / 2;
            long long lo = (long long)s.frame_begin * (long long)pre[i].hop - (long long)(pre[i].win / 2) - pad_l;
            long long hi = (long long)(s.frame_begin + s.frame_count - 1) * (long long)pre[i].hop - (long long)(pre[i].win / 2) - pad_l + (long long)pre[i].n_fft;
            if (lo < 0) hi = std::max(hi, -lo + 1);
            if (hi > n) lo = std::min(lo, 2 * (n - 1) - (hi - 1));
            lo = std::max(0LL, std::min(lo, n)); hi = std::max(0LL, std::min(hi, n));
            if ((long long)s.origin > lo || (long long)(s.origin + s.n) < hi)
                throw Error(SGX_ERR_BAD_ARG, "time slice does not hold the samples its frames read (see sgx_slice_plan)");
        }
    }
    // ---- insert tracks (lib.rs:174-187) --------------------------------------------------------------
    // host-resident PCM is uploaded on a second stream, track by track, and every track's analysis starts as
    // soon as its own copy has landed: H2D of track i+1 overlaps K1 of track i
    bool pipelined = false;
    for (const PcmSource &s : srcs) pipelined = pipelined || !s.on_device;
    std::map<size_t, size_t> copy_event_of; // id -> index into copy_events_
    if (pipelined) {
        if (!copy_stream_) SGX_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
        while (copy_events_.size() < ids.size() + 1) {
            cudaEvent_t e;
            SGX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            copy_events_.push_back(e);
        }
        // the upload may overwrite buffers that kernels already enqueued on stream_ still read
        SGX_CUDA(cudaEventRecord(copy_events_[ids.size()], stream_));
        SGX_CUDA(cudaStreamWaitEvent(copy_stream_, copy_events_[ids.size()], 0));
    }
    for (size_t i = 0; i < ids.size(); ++i) {
        const PcmSource &s = srcs[i];
        // HashMap::insert replaces an existing id; its device buffers and range slot are re-used
        // (all work is ordered on one stream, so no synchronisation is needed for that)
        auto it = tracks_.find(ids[i]);
        if (it == tracks_.end()) it = tracks_.emplace(ids[i], Track()).first;
        Track &t = it->second;
        t.path = s.path; t.sr = s.sr; t.ch = s.ch; t.fmt = s.fmt;
        t.n = s.n_total ? s.n_total : s.n; t.avail = s.n; t.origin = s.n_total ? s.origin : 0;
        t.is_slice = s.n_total != 0;
        t.win = pre[i].win; t.hop = pre[i].hop; t.n_fft = pre[i].n_fft;
        t.tables = tables_for(s.sr, t.win, t.n_fft);
        const size_t esz = s.fmt == PCM_I16 ? 2 : 4;
        if (s.on_device) t.d_pcm = s.data;
        else {
            const size_t bytes = s.n * s.ch * esz;
            t.owned_pcm.ensure(bytes + 64);
            SGX_CUDA(cudaMemcpyAsync(t.owned_pcm.p, s.data, bytes, cudaMemcpyHostToDevice, copy_stream_));
            SGX_CUDA(cudaEventRecord(copy_events_[i], copy_stream_));
            copy_event_of[ids[i]] = i;
            t.d_pcm = t.owned_pcm.p;
        }
        t.t_total = (size_t)pre[i].T;
        t.frame0 = t.is_slice ? s.frame_begin : 0;
        t.n_frames = t.is_slice ? s.frame_count : (size_t)pre[i].T;
        t.n_out = t.tables->n_mel ? t.tables->n_mel : t.n_fft / 2 + 1;
        t.spec.ensure(t.n_frames * t.n_out);
        if (t.slot < 0) t.slot = alloc_slot();
        const float sec = (float)t.n / (float)t.sr; // lib.rs:178-182
        if (sec > max_sec_) { max_sec_ = sec; id_max_sec_ = ids[i]; }
    }
    // ---- update_specs (lib.rs:142-168): one K1 launch per FFT size -------------------------------------
    // tracks that share an FFT size go into one launch
    std::map<size_t, std::vector<size_t>> by_fft;
    for (size_t id : ids) by_fft[tracks_.at(id).n_fft].push_back(id);
    std::vector<StftTrack> descs;
    std::vector<size_t> desc_ids;
    struct Group { size_t n_fft; size_t first, count; StftTiling tiling; int n_tiles; bool raw_loader = false; };
    std::vector<Group> groups;
    for (auto &kv : by_fft) {
        const StftConfig &cfg = ctx_->plan(kv.first).cfg;
        int max_hop = 1;
        for (size_t id : kv.second) max_hop = std::max<int>(max_hop, (int)tracks_.at(id).hop);
        int bank_floats = 0; // room the largest filterbank of the launch needs in shared memory
        if (set_.freq_scale == SGX_FREQ_MEL)
            for (size_t id : kv.second) {
                const TrackTables &tt = *tracks_.at(id).tables;
                bank_floats = std::max(bank_floats, tt.bank_floats(cfg.fused));
            }
        int sample_floats = 1; // f32 stereo tracks stage raw interleaved pairs: twice the room per sample
        bool warp2_ok = true, raw_loader = false; // int16 mono tracks stage their 16-bit samples
        for (size_t id : kv.second) {
            const Track &t = tracks_.at(id);
            if (t.ch == 2 && t.fmt == PCM_F32) sample_floats = 2;
            if (t.ch == 1 && t.fmt == PCM_I16) raw_loader = true;
            if (set_.freq_scale == SGX_FREQ_MEL && t.tables->seg_nblk == 0) warp2_ok = false;
        }
        Group g{kv.first, descs.size(), 0, plan_stft_tiles(cfg, max_hop, bank_floats, sample_floats, warp2_ok), 0};
        g.raw_loader = raw_loader;
        // a track id may appear twice in id_list; the last one wins, launch it once
        std::vector<size_t> uniq;
        for (size_t id : kv.second) if (std::find(uniq.begin(), uniq.end(), id) == uniq.end()) uniq.push_back(id);
        for (size_t id : uniq) {
            Track &t = tracks_.at(id);
            StftTrack d = make_desc(t.d_pcm, t.fmt, t.n, t.ch, t.win, t.hop, t.n_fft, t.n_frames, *t.tables,
                                    t.spec.p, t.n_out, slots_.p + 2 * t.slot, t.origin, t.avail, t.frame0);
            d.tile_begin = pipelined ? 0 : g.n_tiles; // pipelined: one launch per track
            g.n_tiles += (int)((t.n_frames + g.tiling.frames_per_tile - 1) / g.tiling.frames_per_tile);
            descs.push_back(d);
            desc_ids.push_back(id);
            ++g.count;
        }
        groups.push_back(g);
    }
    d_stft_.ensure(descs.size());
    SGX_CUDA(cudaMemcpyAsync(d_stft_.p, descs.data(), sizeof(StftTrack) * descs.size(), cudaMemcpyHostToDevice, stream_));
    SGX_CUDA(launch_range_reset(d_stft_.p, (int)descs.size(), stream_)); // extrema of re-analysed tracks start over
    if (profiling_) SGX_CUDA(cudaEventRecord(ev_[0], stream_));
    for (const Group &g : groups) {
        FftPlan &pl = ctx_->plan(g.n_fft);
        StftLaunch L{};
        L.tracks = d_stft_.p + g.first; L.n_tracks = (int)g.count; L.n_tiles = g.n_tiles;
        L.mode = set_.freq_scale == SGX_FREQ_MEL ? MODE_MEL_DB : MODE_LIN_DB;
        L.frames_per_tile = g.tiling.frames_per_tile; L.staged = g.tiling.staged;
        L.tile_floats = g.tiling.tile_floats; L.bank_floats = g.tiling.bank_floats; L.tw = pl.tw.p; L.split = pl.split.p; L.twr = pl.twr.p;
        L.stereo_raw = g.tiling.sample_floats == 2 ? 1 : (g.raw_loader ? 2 : 0); L.warp2 = g.tiling.warp2; L.warp1 = g.tiling.warp1;
        if (!pipelined) { SGX_CUDA(launch_stft(pl.cfg, L, stream_)); continue; }
        for (size_t k = 0; k < g.count; ++k) {
            const size_t di = g.first + k;
            const Track &t = tracks_.at(desc_ids[di]);
            auto ce = copy_event_of.find(desc_ids[di]);
            if (ce != copy_event_of.end()) SGX_CUDA(cudaStreamWaitEvent(stream_, copy_events_[ce->second], 0));
            L.tracks = d_stft_.p + di; L.n_tracks = 1;
            L.n_tiles = (int)((t.n_frames + g.tiling.frames_per_tile - 1) / g.tiling.frames_per_tile);
            SGX_CUDA(launch_stft(pl.cfg, L, stream_));
        }
    }
    if (profiling_) { SGX_CUDA(cudaEventRecord(ev_[1], stream_)); ev_valid_[0] = true; }
    // ---- update_spec_greys, range part (lib.rs:193-229) ------------------------------------------------
    reduce_local();
}

bool MultiTrack::remove_track(size_t id, bool want_changed)
{
    fuse_commit_ = want_changed;
    drop(id);
    fuse_commit_ = false;
    exchange();
    if ((want_changed || comm_) && !committed_in_reduce_) commit();
    if (!want_changed) return false;
    return synchronize();
}

void MultiTrack::drop(size_t id)
{
    SGX_CUDA(cudaSetDevice(device_));
    auto it = tracks_.find(id);
    if (it == tracks_.end()) {
        if (!owns(id)) { reduce_local(); return; } // another shard's track: only the exchange concerns this handle
        throw Error(SGX_ERR_UNKNOWN_ID, "remove_track: unknown track id " + std::to_string(id));
    }
    drop_track(id);
    if (id_max_sec_ == id) { // lib.rs:269-286
        size_t best_id = 0; float best = 0.0f;
        for (auto &kv : tracks_) {
            const float sec = (float)kv.second.n / (float)kv.second.sr;
            if (sec > best) { best = sec; best_id = kv.first; }
        }
        id_max_sec_ = best_id; max_sec_ = best;
    }
    reduce_local();
}

const Track &MultiTrack::track(size_t id) const
{
    auto it = tracks_.find(id);
    if (it == tracks_.end()) throw Error(SGX_ERR_UNKNOWN_ID, "unknown track id " + std::to_string(id));
    return it->second;
}

float MultiTrack::frequency_hz(size_t id, float rel) const
{
    const float half_sr = (float)track(id).sr / 2.0f;
    if (set_.freq_scale == SGX_FREQ_LINEAR) return half_sr * rel;
    return mel_to_hz(hz_to_mel(half_sr) * rel);
}

uint32_t MultiTrack::image_width(size_t id, float px_per_sec) const
{
    const Track &t = track(id);
    return calc_nwidth(px_per_sec, t.n, t.sr);
}

AxisTableDev *MultiTrack::axis_table(int n_in, int n_out, bool tap_major)
{
    const auto key = std::make_tuple(n_in, n_out, tap_major);
    auto it = axis_.find(key);
    if (it != axis_.end()) return it->second.get();
    std::unique_ptr<AxisTableDev> t(new AxisTableDev());
    // the fast render kernels read 8 or 16 taps per output index unconditionally: rows are at least 16 wide, zero-filled
    t->taps = (std::max(16, (int)lanczos3_max_taps((uint32_t)n_in, (uint32_t)n_out)) + 3) & ~3;
    t->left.alloc(n_out); t->cnt.alloc(n_out); t->sum.alloc(n_out);
    t->w.alloc((size_t)n_out * t->taps);
    SGX_CUDA(launch_build_axis_table(n_in, n_out, t->taps, tap_major, t->left.p, t->cnt.p, t->sum.p, t->w.p, stream_));
    AxisTableDev *raw = t.get();
    axis_.emplace(key, std::move(t));
    return raw;
}

void MultiTrack::render(const std::vector<size_t> &ids, float px_per_sec, uint32_t nheight, int channels,
                        uint8_t *const *d_out, const size_t *cap, size_t *written, const uint32_t *ox_begin,
                        const uint32_t *ox_count)
{
    SGX_CUDA(cudaSetDevice(device_));
    if (channels != 3 && channels != 4) throw Error(SGX_ERR_BAD_ARG, "channels must be 3 or 4");
    if (nheight > 65535u) throw Error(SGX_ERR_BAD_ARG, "nheight too large");
    const bool mel = set_.freq_scale == SGX_FREQ_MEL;
    // the image geometry depends on the highest sample rate of ALL shards (lib.rs:220-248): with a communicator it
    // arrives with the range exchange and is read back here (one stream synchronisation per add / remove) unless
    // the driver supplied it with set_global_max_sr
    if (comm_ && pending_ && !global_sr_fixed_) synchronize_keep_changed();
    const uint32_t msr = effective_max_sr();
    // The axis-table cache is bounded, but it is only ever emptied HERE, before this call resolves its first key:
    // descriptors assembled below hold raw pointers into it (and launches already enqueued read its device tables,
    // hence the synchronisation).  Inside a call it grows by at most two entries per track.
    if (axis_.size() >= 128) {
        SGX_CUDA(cudaStreamSynchronize(stream_));
        axis_.clear();
    }
    struct Item { size_t idx; int T, height, nwidth, cols; };
    std::vector<Item> items;
    std::vector<RenderTrack> descs;
    bool short_buf = false;
    for (size_t i = 0; i < ids.size(); ++i) {
        const Track &t = track(ids[i]);
        const uint32_t nwidth = calc_nwidth(px_per_sec, t.n, t.sr); // lib.rs:296
        // column window of this request (whole image unless a driver renders a strip of a time slice)
        const uint32_t ob = ox_begin ? ox_begin[i] : 0, oc = ox_count ? ox_count[i] : nwidth;
        if (ob > nwidth || oc > nwidth - ob) throw Error(SGX_ERR_BAD_ARG, "column window outside the image");
        const size_t need = (size_t)oc * nheight * channels;
        if (written) written[i] = need;
        if (need == 0) continue;
        if (!d_out || !d_out[i]) continue; // size query
        if (cap[i] < need) { short_buf = true; continue; }
        const float up = calc_up_ratio(msr, t.sr, mel);            // lib.rs:231-248
        const uint32_t height = grey_height(t.n_out, up);          // display.rs:45
        if (height < t.n_out)
            throw Error(SGX_ERR_STATE, "up_ratio < 1: u32 underflow at display.rs:47 (global max_sr not set?)");
        RenderTrack r{};
        r.src = t.spec.p; r.width = (int)t.t_total; r.n_out = (int)t.n_out; r.height = (int)height;
        r.frame0 = (int)t.frame0; r.src_frames = (int)t.n_frames; r.ox_begin = (int)ob; r.ox_count = (int)oc;
        r.nwidth = (int)nwidth; r.nheight = (int)nheight; r.out = d_out[i];
        if (t.is_slice) { // the strip may only need frames this handle holds
            uint32_t l0, r0, l1, r1;
            lanczos3_span((uint32_t)t.t_total, nwidth, ob, &l0, &r0);
            lanczos3_span((uint32_t)t.t_total, nwidth, ob + oc - 1, &l1, &r1);
            if (l0 < t.frame0 || r1 > t.frame0 + t.n_frames)
                throw Error(SGX_ERR_BAD_ARG, "time slice does not hold the frames this column window needs (see sgx_slice_plan)");
        }
        AxisTableDev *v = axis_table((int)height, (int)nheight, false);
        AxisTableDev *h = axis_table((int)t.t_total, (int)nwidth, true);
        r.v_left = v->left.p; r.v_cnt = v->cnt.p; r.v_sum = v->sum.p; r.v_w = v->w.p; r.v_taps = v->taps;
        r.h_left = h->left.p; r.h_cnt = h->cnt.p; r.h_sum = h->sum.p; r.h_w = h->w.p; r.h_taps = h->taps;
        items.push_back(Item{descs.size(), r.width, r.height, r.nwidth, r.ox_count});
        descs.push_back(r);
    }
    if (!descs.empty()) {
        // Tracks that can share a launch sit next to each other: the FP32 fast paths read every geometry value from the
        // track's own descriptor, so all tracks of one tap class go into ONE launch whatever their length, sample rate or
        // mel height (six rates x six geometries: one render launch instead of six); the wide, general and tensor-core
        // paths size their tiles for a geometry and take identical geometries only.
        struct Group { std::tuple<int, int, int, int> key; RenderTiling tl; };
        std::vector<Group> plan(items.size());
        for (size_t i = 0; i < items.size(); ++i) {
            const RenderTiling tl = plan_render_tiles(items[i].T, items[i].height, items[i].nwidth, (int)nheight, true);
            plan[i].tl = tl;
            plan[i].key = tl.fast >= 100 ? std::make_tuple(tl.fast, 0, 0, 0)
                          : tl.fast == 3 ? std::make_tuple(3, 0, 0, 0) // sliding-window path: one launch, sized for the tallest source window
                                         : std::make_tuple(tl.fast, items[i].T, items[i].height, items[i].nwidth);
        }
        std::vector<size_t> order(items.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return plan[x].key < plan[y].key; });
        std::vector<RenderTrack> sorted;
        for (size_t i : order) sorted.push_back(descs[items[i].idx]);
        d_render_.ensure(sorted.size());
        SGX_CUDA(cudaMemcpyAsync(d_render_.p, sorted.data(), sizeof(RenderTrack) * sorted.size(), cudaMemcpyHostToDevice, stream_));
        if (profiling_) SGX_CUDA(cudaEventRecord(ev_[2], stream_));
        size_t a = 0;
        while (a < order.size()) {
            size_t b = a + 1;
            while (b < order.size() && plan[order[b]].key == plan[order[a]].key) ++b;
            RenderTiling tl = plan[order[a]].tl;
            int max_cols = 0;
            for (size_t c = a; c < b; ++c) {
                max_cols = std::max(max_cols, items[order[c]].cols);
                if (tl.fast == 3) { // capacities of the group = those of its most demanding track
                    tl.rv_max = std::max(tl.rv_max, plan[order[c]].tl.rv_max);
                    tl.smem_bytes = std::max(tl.smem_bytes, plan[order[c]].tl.smem_bytes);
                }
            }
            for (size_t c = a; c < b; c += 65535) { // gridDim.z limit
                RenderLaunch L{};
                L.tracks = d_render_.p + c; L.n_tracks = (int)std::min<size_t>(65535, b - c);
                L.from_db = 1; L.range = d_state_.p; L.channels = channels;
                L.px = tl.px; L.py = tl.py; L.fc = tl.fc; L.rv_max = tl.rv_max;
                SGX_CUDA(launch_render(L, max_cols, (int)nheight, tl.smem_bytes, tl.fast, stream_));
            }
            a = b;
        }
        if (profiling_) { SGX_CUDA(cudaEventRecord(ev_[3], stream_)); ev_valid_[1] = true; }
    }
    if (short_buf) throw Error(SGX_ERR_BUFFER, "output buffer too small");
}

void MultiTrack::render_host(size_t id, float px_per_sec, uint32_t nheight, int channels, uint8_t *out, size_t need)
{
    if (need == 0) return;
    // the staging buffer must live on THIS engine's device: a multi-device handle enters with the caller's device current
    // (found on 8 GPUs: allocated on device 0, written by a kernel on device 5 -- it only worked where NCCL had enabled
    // peer access between the two)
    SGX_CUDA(cudaSetDevice(device_));
    d_img_.ensure(need);
    uint8_t *outs[1] = {d_img_.p};
    size_t caps[1] = {need}, wr[1] = {0};
    render({id}, px_per_sec, nheight, channels, outs, caps, wr);
    SGX_CUDA(cudaMemcpyAsync(out, d_img_.p, need, cudaMemcpyDeviceToHost, stream_));
    SGX_CUDA(cudaStreamSynchronize(stream_));
}

void MultiTrack::images_async(const std::vector<size_t> &ids, float px_per_sec, uint32_t nheight, int channels,
                              uint8_t *const *out, const size_t *cap, size_t *written)
{
    SGX_CUDA(cudaSetDevice(device_));
    wait_images(); // the staging buffers of the previous request are being read by its copies
    std::vector<size_t> need(ids.size(), 0);
    bool short_buf = false;
    for (size_t i = 0; i < ids.size(); ++i) {
        need[i] = (size_t)image_width(ids[i], px_per_sec) * nheight * (size_t)channels;
        if (written) written[i] = need[i];
        if (out && out[i] && cap[i] < need[i]) short_buf = true;
    }
    if (!out) return; // size query
    if (short_buf) throw Error(SGX_ERR_BUFFER, "output buffer too small");
    if (!out_stream_) {
        SGX_CUDA(cudaStreamCreateWithFlags(&out_stream_, cudaStreamNonBlocking));
        SGX_CUDA(cudaEventCreateWithFlags(&out_ev_, cudaEventDisableTiming));
        SGX_CUDA(cudaEventCreateWithFlags(&out_done_, cudaEventDisableTiming));
    }
    // one staging buffer per image: every render of the batch is enqueued at once (K3 of image i+1 runs while the copy
    // of image i is on the wire), and the compute stream is free again after the last render -- the uploads and K1 of
    // the next add_tracks then overlap the rest of the downloads (the two directions of the link are independent)
    if (d_imgs_.size() < ids.size()) d_imgs_.resize(ids.size());
    std::vector<uint8_t *> d_out(ids.size(), nullptr);
    std::vector<size_t> caps(ids.size(), 0), wr(ids.size(), 0);
    for (size_t i = 0; i < ids.size(); ++i) {
        if (!out[i] || need[i] == 0) continue;
        d_imgs_[i].ensure(need[i]);
        d_out[i] = d_imgs_[i].p; caps[i] = need[i];
    }
    // renders in chunks of a few images so that the first copy starts early
    const size_t chunk = 4;
    for (size_t a = 0; a < ids.size(); a += chunk) {
        const size_t b = std::min(ids.size(), a + chunk);
        std::vector<size_t> sub(ids.begin() + a, ids.begin() + b);
        render(sub, px_per_sec, nheight, channels, d_out.data() + a, caps.data() + a, wr.data() + a);
        SGX_CUDA(cudaEventRecord(out_ev_, stream_));
        SGX_CUDA(cudaStreamWaitEvent(out_stream_, out_ev_, 0));
        for (size_t i = a; i < b; ++i)
            if (d_out[i]) SGX_CUDA(cudaMemcpyAsync(out[i], d_out[i], need[i], cudaMemcpyDeviceToHost, out_stream_));
    }
    SGX_CUDA(cudaEventRecord(out_done_, out_stream_));
    out_pending_ = true;
}

void MultiTrack::wait_images()
{
    if (!out_pending_) return;
    SGX_CUDA(cudaSetDevice(device_));
    SGX_CUDA(cudaEventSynchronize(out_done_));
    out_pending_ = false;
}

void MultiTrack::set_profiling(bool on)
{
    SGX_CUDA(cudaSetDevice(device_));
    if (on && !ev_[0]) for (auto &e : ev_) SGX_CUDA(cudaEventCreate(&e));
    profiling_ = on;
    ev_valid_[0] = ev_valid_[1] = false;
}

void MultiTrack::stage_times(float *analysis_ms, float *render_ms)
{
    SGX_CUDA(cudaSetDevice(device_));
    SGX_CUDA(cudaStreamSynchronize(stream_));
    *analysis_ms = *render_ms = -1.0f;
    if (ev_valid_[0]) SGX_CUDA(cudaEventElapsedTime(analysis_ms, ev_[0], ev_[1]));
    if (ev_valid_[1]) SGX_CUDA(cudaEventElapsedTime(render_ms, ev_[2], ev_[3]));
}

std::vector<uint8_t> MultiTrack::wav_image(size_t id, float px_per_sec, uint32_t nheight, float amp_min, float amp_max)
{
    SGX_CUDA(cudaSetDevice(device_));
    const Track &t = track(id);
    if (t.is_slice) throw Error(SGX_ERR_STATE, "get_wav_image is not available for a time slice of a track");
    const uint32_t nwidth = calc_nwidth(px_per_sec, t.n, t.sr); // lib.rs:308
    const size_t need = (size_t)nwidth * nheight * 4;
    std::vector<uint8_t> host(need);
    if (need == 0) return host;
    DevBuf<uint8_t> d; d.alloc(need);
    DevBuf<int> flag; flag.alloc(1);
    SGX_CUDA(cudaMemsetAsync(d.p, 0, need, stream_));
    SGX_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), stream_));
    SGX_CUDA(launch_wav_image(t.d_pcm, t.fmt, (int)t.ch, (long long)t.n, (int)nwidth, (int)nheight, amp_min, amp_max, d.p, flag.p, stream_));
    int bad = 0;
    SGX_CUDA(cudaMemcpyAsync(host.data(), d.p, need, cudaMemcpyDeviceToHost, stream_));
    SGX_CUDA(cudaMemcpyAsync(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    SGX_CUDA(cudaStreamSynchronize(stream_));
    if (bad) throw Error(SGX_ERR_BAD_ARG, "wav_to_image: empty sample slice for a pixel column (the reference panics, display.rs:95)");
    return host;
}

// ---------------------------------------------------------------------------------------------------
// stage functions (surface 2): host in, host out, current device, default stream
// ---------------------------------------------------------------------------------------------------
static int current_device()
{
    int d = 0;
    cudaError_t e = cudaGetDevice(&d);
    if (e != cudaSuccess) { cudaGetLastError(); throw Error(SGX_ERR_CUDA, "no usable CUDA device (this engine has no CPU path)"); }
    return d;
}

// 64-bit content hash (four interleaved multiplicative lanes) of a float array: lets the stage functions see that
// a caller passes the same window / filterbank again (bench.rs builds both once, outside the timed closure)
static uint64_t hash_floats(const float *p, size_t n, uint64_t seed)
{
    uint64_t h[4] = {seed ^ 0x9E3779B97F4A7C15ull, seed ^ 0xC2B2AE3D27D4EB4Full, seed ^ 0x165667B19E3779F9ull, seed ^ 0x27D4EB2F165667C5ull};
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
    size_t i = 0;
    for (; i + 4 <= n; i += 4)
        for (int k = 0; k < 4; ++k) h[k] = (h[k] ^ w[i + k]) * 0x100000001B3ull;
    for (; i < n; ++i) h[0] = (h[0] ^ w[i]) * 0x100000001B3ull;
    return (h[0] * 31 + h[1]) * 31 + (h[2] * 31 + h[3]) + n;
}

// Device buffers of the stage functions live per host thread and only ever grow: a call costs transfers and
// launches, not cudaMalloc / cudaFree.
struct StageWorkspace {
    int device = -1;
    TrackTables tt;
    uint64_t tt_key = 0;   // hash of (win, n_fft, window, filterbank) the tables were built from; 0 = none
    bool tt_valid = false;
    DevBuf<float> in, out, grey;
    DevBuf<StftTrack> desc;
    DevBuf<RenderTrack> rdesc;
    DevBuf<uint8_t> pix;
    AxisTableDev v, h;
};
static StageWorkspace &stage_workspace()
{
    thread_local StageWorkspace ws;
    const int d = current_device();
    if (ws.device != d) { ws = StageWorkspace(); ws.device = d; }
    return ws;
}

StageOut stage_stft(int mode, const float *input, size_t n, size_t win, size_t hop, size_t n_fft,
                    const float *window, const float *mel_fb, size_t n_mel, float *out, size_t cap_elems)
{
    DeviceCtx &ctx = DeviceCtx::get(current_device());
    long T = 0;
    check_stft_args(n, win, hop, n_fft, &T);
    if (mode == MODE_MEL_DB && (!mel_fb || n_mel == 0)) throw Error(SGX_ERR_BAD_ARG, "mel filterbank missing (mel.rs:51 assert_ne!(n_mel, 0))");
    const size_t B = n_fft / 2 + 1;
    const size_t n_out = mode == MODE_MEL_DB ? n_mel : B;
    const size_t elems = (size_t)T * n_out * (mode == MODE_COMPLEX ? 2 : 1);
    StageOut so{(size_t)T, n_out};
    if (!out) return so;
    if (cap_elems < elems) throw Error(SGX_ERR_BUFFER, "output buffer too small");
    cudaStream_t s = 0;
    FftPlan &pl = ctx.plan(n_fft);
    std::vector<float> wbuf;
    if (!window) { wbuf.resize(win); calc_window(win, n_fft, wbuf.data()); window = wbuf.data(); } // lib.rs:403-408
    StageWorkspace &ws = stage_workspace();
    TrackTables &tt = ws.tt;
    {
        const bool use_mel = mode == MODE_MEL_DB;
        uint64_t key = hash_floats(window, win, (uint64_t)win * 1315423911u + n_fft);
        if (use_mel) key = hash_floats(mel_fb, B * n_mel, key ^ (uint64_t)n_mel);
        key ^= use_mel ? 0x5bd1e995u : 0u;
        if (!ws.tt_valid || ws.tt_key != key) {
            ws.tt_valid = false;
            fill_tables(tt, win, n_fft, window, use_mel ? mel_fb : nullptr, n_mel, pl.cfg, s);
            ws.tt_key = key; ws.tt_valid = true;
        }
    }
    DevBuf<float> &d_in = ws.in, &d_out = ws.out;
    d_in.ensure(n + 16); d_out.ensure(elems);
    SGX_CUDA(cudaMemcpyAsync(d_in.p, input, n * sizeof(float), cudaMemcpyHostToDevice, s));
    StftTrack d = make_desc(d_in.p, PCM_F32, n, 1, win, hop, n_fft, (size_t)T, tt, d_out.p, n_out, nullptr);
    const StftTiling tl = plan_stft_tiles(pl.cfg, (int)hop, mode != MODE_MEL_DB ? 0 : tt.bank_floats(pl.cfg.fused), 1, mode != MODE_MEL_DB || tt.seg_nblk > 0);
    DevBuf<StftTrack> &dd = ws.desc; dd.upload(&d, 1, s);
    StftLaunch L{};
    L.tracks = dd.p; L.n_tracks = 1;
    L.n_tiles = (int)(((size_t)T + tl.frames_per_tile - 1) / tl.frames_per_tile);
    L.mode = mode; L.frames_per_tile = tl.frames_per_tile; L.staged = tl.staged; L.tile_floats = tl.tile_floats;
    L.bank_floats = tl.bank_floats; L.warp2 = tl.warp2; L.warp1 = tl.warp1;
    L.tw = pl.tw.p; L.split = pl.split.p; L.twr = pl.twr.p;
    SGX_CUDA(launch_stft(pl.cfg, L, s));
    SGX_CUDA(cudaMemcpyAsync(out, d_out.p, elems * sizeof(float), cudaMemcpyDeviceToHost, s));
    SGX_CUDA(cudaStreamSynchronize(s));
    return so;
}

void stage_amp_to_db(float *x, size_t n)
{
    DeviceCtx::get(current_device());
    if (n == 0) return;
    DevBuf<float> d; d.alloc(n);
    DevBuf<int> flag; flag.alloc(1);
    SGX_CUDA(cudaMemcpyAsync(d.p, x, n * sizeof(float), cudaMemcpyHostToDevice, 0));
    SGX_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), 0));
    SGX_CUDA(launch_amp_to_db(d.p, n, flag.p, 0));
    int bad = 0;
    SGX_CUDA(cudaMemcpy(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) throw Error(SGX_ERR_BAD_ARG, "amp_to_db: negative or NaN input (assert at decibel.rs:34)");
    SGX_CUDA(cudaMemcpy(x, d.p, n * sizeof(float), cudaMemcpyDeviceToHost));
}

uint32_t stage_spec_to_grey(const float *spec, size_t T, size_t n_out, float up_ratio, float max_db,
                            float min_db, float *grey, size_t cap)
{
    DeviceCtx::get(current_device());
    const uint32_t height = grey_height(n_out, up_ratio);
    if (!grey) return height;
    if (height < n_out) throw Error(SGX_ERR_BAD_ARG, "up_ratio < 1: u32 underflow at display.rs:47");
    const size_t elems = (size_t)T * height;
    if (cap < elems) throw Error(SGX_ERR_BUFFER, "grey buffer too small");
    if (elems == 0) return height;
    DevBuf<float> ds, dg; ds.alloc(T * n_out); dg.alloc(elems);
    SGX_CUDA(cudaMemcpyAsync(ds.p, spec, T * n_out * sizeof(float), cudaMemcpyHostToDevice, 0));
    SGX_CUDA(launch_spec_to_grey(ds.p, (int)T, (int)n_out, (int)height, max_db, min_db, dg.p, 0));
    SGX_CUDA(cudaMemcpy(grey, dg.p, elems * sizeof(float), cudaMemcpyDeviceToHost));
    return height;
}

void stage_grey_to_rgb(const float *grey, uint32_t width, uint32_t height, uint32_t nwidth, uint32_t nheight,
                       int channels, uint8_t *out, size_t cap)
{
    DeviceCtx::get(current_device());
    if (channels != 3 && channels != 4) throw Error(SGX_ERR_BAD_ARG, "channels must be 3 or 4");
    if (!width || !height || !nwidth || !nheight) throw Error(SGX_ERR_BAD_ARG, "empty image");
    const size_t need = (size_t)nwidth * nheight * channels;
    if (cap < need) throw Error(SGX_ERR_BUFFER, "output buffer too small");
    cudaStream_t s = 0;
    StageWorkspace &ws = stage_workspace();
    DevBuf<float> &dg = ws.grey; dg.ensure((size_t)width * height);
    DevBuf<uint8_t> &dout = ws.pix; dout.ensure(need);
    SGX_CUDA(cudaMemcpyAsync(dg.p, grey, (size_t)width * height * sizeof(float), cudaMemcpyHostToDevice, s));
    AxisTableDev &v = ws.v, &h = ws.h;
    v.taps = (std::max(16, (int)lanczos3_max_taps(height, nheight)) + 3) & ~3;
    h.taps = (std::max(16, (int)lanczos3_max_taps(width, nwidth)) + 3) & ~3;
    v.left.ensure(nheight); v.cnt.ensure(nheight); v.sum.ensure(nheight); v.w.ensure((size_t)nheight * v.taps);
    h.left.ensure(nwidth); h.cnt.ensure(nwidth); h.sum.ensure(nwidth); h.w.ensure((size_t)nwidth * h.taps);
    SGX_CUDA(launch_build_axis_table((int)height, (int)nheight, v.taps, false, v.left.p, v.cnt.p, v.sum.p, v.w.p, s));
    SGX_CUDA(launch_build_axis_table((int)width, (int)nwidth, h.taps, true, h.left.p, h.cnt.p, h.sum.p, h.w.p, s));
    RenderTrack r{};
    r.src = dg.p; r.width = (int)width; r.n_out = (int)height; r.height = (int)height;
    r.frame0 = 0; r.src_frames = (int)width; r.ox_begin = 0; r.ox_count = (int)nwidth;
    r.nwidth = (int)nwidth; r.nheight = (int)nheight; r.out = dout.p;
    r.v_left = v.left.p; r.v_cnt = v.cnt.p; r.v_sum = v.sum.p; r.v_w = v.w.p; r.v_taps = v.taps;
    r.h_left = h.left.p; r.h_cnt = h.cnt.p; r.h_sum = h.sum.p; r.h_w = h.w.p; r.h_taps = h.taps;
    DevBuf<RenderTrack> &dr = ws.rdesc; dr.upload(&r, 1, s);
    const RenderTiling tl = plan_render_tiles((int)width, (int)height, (int)nwidth, (int)nheight);
    RenderLaunch L{};
    L.tracks = dr.p; L.n_tracks = 1; L.from_db = 0; L.range = nullptr; L.channels = channels;
    L.px = tl.px; L.py = tl.py; L.fc = tl.fc; L.rv_max = tl.rv_max;
    SGX_CUDA(launch_render(L, (int)nwidth, (int)nheight, tl.smem_bytes, tl.fast, s));
    SGX_CUDA(cudaMemcpyAsync(out, dout.p, need, cudaMemcpyDeviceToHost, s));
    SGX_CUDA(cudaStreamSynchronize(s));
}

void stage_wav_to_image(const float *wav, size_t n, uint32_t nwidth, uint32_t nheight, float amp_min,
                        float amp_max, uint8_t *out, size_t cap)
{
    DeviceCtx::get(current_device());
    const size_t need = (size_t)nwidth * nheight * 4;
    if (cap < need) throw Error(SGX_ERR_BUFFER, "output buffer too small");
    if (need == 0) return;
    if (n == 0) throw Error(SGX_ERR_BAD_ARG, "empty waveform");
    DevBuf<float> dw; dw.alloc(n);
    DevBuf<uint8_t> d; d.alloc(need);
    DevBuf<int> flag; flag.alloc(1);
    SGX_CUDA(cudaMemcpyAsync(dw.p, wav, n * sizeof(float), cudaMemcpyHostToDevice, 0));
    SGX_CUDA(cudaMemsetAsync(d.p, 0, need, 0));
    SGX_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), 0));
    SGX_CUDA(launch_wav_image(dw.p, PCM_F32, 1, (long long)n, (int)nwidth, (int)nheight, amp_min, amp_max, d.p, flag.p, 0));
    int bad = 0;
    SGX_CUDA(cudaMemcpy(&bad, flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) throw Error(SGX_ERR_BAD_ARG, "wav_to_image: empty sample slice for a pixel column (the reference panics, display.rs:95)");
    SGX_CUDA(cudaMemcpy(out, d.p, need, cudaMemcpyDeviceToHost));
}

} // namespace sgx
