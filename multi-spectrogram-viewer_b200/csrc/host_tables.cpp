// host_tables.cpp -- see host_tables.h.  All arithmetic is f32 in the reference's operation
// order (build with -ffp-contract=off); tests/test_host_tables.py checks every function
// bit-for-bit against the CPU oracle.
#include "host_tables.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace sgx {

static const double kPi = 3.14159265358979323846264338327950288;

size_t calc_proper_n_fft(size_t win_length)
{
    // utils.rs:18: 2usize.pow((win_length as f32).log2().ceil() as u32)
    float e = std::ceil(std::log2((float)win_length));
    if (!(e > 0.0f)) return 1;
    return (size_t)1 << (unsigned)e;
}

void hann(size_t size, bool symmetric, float *out)
{
    // windows.rs:7-19 with (a,b,c,d) = (0.5,0.5,0,0): x = pi*i/(size2-1);
    // (a - b*cos(2x)) + (c*cos(4x) - d*cos(6x))
    const float pi = (float)kPi;
    const size_t size2 = symmetric ? size : size + 1;
    const float denom = (float)(size2 - 1);
    for (size_t i = 0; i < size; ++i) {
        float x = pi * (float)i / denom;
        float t1 = 0.5f * std::cos(2.0f * x);
        float t2 = 0.0f * std::cos(4.0f * x);
        float t3 = 0.0f * std::cos(6.0f * x);
        out[i] = (0.5f - t1) + (t2 - t3);
    }
}

void calc_window(size_t win_length, size_t n_fft, float *out)
{
    hann(win_length, false, out);
    const float d = (float)n_fft;
    for (size_t i = 0; i < win_length; ++i) out[i] = out[i] / d;
}

// mel.rs:8-11
static const float kMinLogMel = 15.0f;
static const float kMinLogHz = (float)1000.0;
static const float kLogStep = (float)0.06875177742094912;
static const float kLinearScale = (float)(200.0 / 3.0);

float mel_to_hz(float mel)
{
    if (mel < kMinLogMel) return kLinearScale * mel;
    return kMinLogHz * std::exp(kLogStep * (mel - kMinLogMel));
}

float hz_to_mel(float hz)
{
    if (hz < kMinLogHz) return hz / kLinearScale;
    return kMinLogMel + std::log(hz / kMinLogHz) / kLogStep;
}

// ndarray Array::linspace(a,b,n): a + i*((b-a)/(n-1))
static void linspace(float a, float b, size_t n, std::vector<float> &v)
{
    v.resize(n);
    const float step = n > 1 ? (b - a) / (float)(n - 1) : 0.0f;
    for (size_t i = 0; i < n; ++i) v[i] = a + step * (float)i;
}

void calc_mel_fb(uint32_t sr, size_t n_fft, size_t n_mel, float fmin, float fmax, bool do_norm,
                 float *out)
{
    const float nyq = (float)sr / 2.0f;
    if (fmax < 0.0f) fmax = nyq;
    const size_t n_freq = n_fft / 2 + 1;
    std::vector<float> lin, edges;
    linspace(0.0f, nyq, n_freq, lin);
    linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mel + 2, edges);
    for (float &e : edges) e = mel_to_hz(e);
    std::memset(out, 0, sizeof(float) * n_freq * n_mel);
    for (size_t m = 0; m < n_mel; ++m) {
        const float f0 = edges[m], f1 = edges[m + 1], f2 = edges[m + 2];
        for (size_t b = 0; b < n_freq; ++b) { // mel.rs:67-79
            const float f = lin[b];
            if (f <= f0) continue;
            if (f0 < f && f < f1) out[b * n_mel + m] = (f - f0) / (f1 - f0);
            else if (f == f1) out[b * n_mel + m] = 1.0f;
            else if (f1 < f && f < f2) out[b * n_mel + m] = (f2 - f) / (f2 - f1);
            else break;
        }
        if (do_norm) { // mel.rs:80-82
            float s = 0.0f;
            for (size_t b = 0; b < n_freq; ++b) s += out[b * n_mel + m];
            const float eps = 1.1920929e-07f;
            const float d = s > eps ? s : eps;
            for (size_t b = 0; b < n_freq; ++b) out[b * n_mel + m] = out[b * n_mel + m] / d;
        }
    }
}

size_t calc_mel_fb_default(uint32_t sr, size_t n_fft, std::vector<float> &fb)
{
    const size_t n_freq = n_fft / 2 + 1;
    const float guess =
        2.0f * hz_to_mel((float)sr / 2.0f) / hz_to_mel((float)sr / (float)n_fft) - 1.0f;
    size_t n_mel = guess > 0.0f ? (size_t)guess : 0; // `as usize`: truncating, saturating
    n_mel = std::min(n_mel, n_freq);
    for (;;) {
        fb.assign(n_freq * std::max<size_t>(n_mel, 1), 0.0f);
        if (n_mel == 0) break; // the reference would assert (mel.rs:51); callers treat 0 as error
        calc_mel_fb(sr, n_fft, n_mel, 0.0f, -1.0f, true, fb.data());
        bool all_pos = true;
        for (size_t m = 0; m < n_mel && all_pos; ++m) {
            float s = 0.0f;
            for (size_t b = 0; b < n_freq; ++b) s += fb[b * n_mel + m];
            all_pos = s > 0.0f;
        }
        if (all_pos) break;
        --n_mel;
    }
    fb.resize(n_freq * n_mel);
    return n_mel;
}

static size_t count_windows(size_t len, size_t w, size_t hop)
{
    return len < w ? 0 : (len - w) / hop + 1; // ndarray .windows(w).step_by(hop)
}

long stft_num_frames(size_t n, size_t win, size_t hop)
{
    if (win < 2 || hop < 1 || n < win) return -1;
    const size_t half = win / 2;
    if (half + 1 > n || half + 1 > win - 1) return -1; // reflect pads need x[1..=half]
    // lib.rs:412-418  front: input[..win-1] left-padded by half
    const size_t n_front = count_windows(win - 1 + half, win, hop);
    if (n_front * hop < half) return -1; // usize underflow at lib.rs:420
    size_t first = n_front * hop - half;
    if (first > n) return -1;
    // lib.rs:420-421  middle
    const size_t n_mid = count_windows(n - first, win, hop);
    first += n_mid * hop;
    // lib.rs:423-433  back
    const size_t back_start = std::min(first, n - half - 1);
    const size_t back_len = (n - back_start) + half - (first - back_start);
    const size_t n_back = count_windows(back_len, win, hop);
    return (long)(n_front + n_mid + n_back);
}

uint32_t calc_nwidth(float px_per_sec, size_t n, uint32_t sr)
{
    const float v = px_per_sec * (float)n / (float)sr;
    if (!(v > 0.0f)) return 0;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

float calc_up_ratio(uint32_t max_sr, uint32_t sr, bool mel)
{
    if (!mel) return (float)max_sr / (float)sr;
    return hz_to_mel((float)max_sr / 2.0f) / hz_to_mel((float)sr / 2.0f);
}

uint32_t grey_height(size_t n_out, float up_ratio)
{
    return (uint32_t)std::round((float)n_out * up_ratio);
}

void lanczos3_span(uint32_t n_in, uint32_t n_out, uint32_t o, uint32_t *left, uint32_t *right)
{
    const float ratio = (float)n_in / (float)n_out;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = 3.0f * sratio;
    const float inputx = ((float)o + 0.5f) * ratio;
    long long l = (long long)std::floor(inputx - support);
    l = l < 0 ? 0 : (l > (long long)n_in - 1 ? (long long)n_in - 1 : l);
    long long r = (long long)std::ceil(inputx + support);
    r = r < l + 1 ? l + 1 : (r > (long long)n_in ? (long long)n_in : r);
    *left = (uint32_t)l; *right = (uint32_t)r;
}

// Segment form of a mel-like bank (MelBands::seg).  A triangular bank touches every bin with at most two
// NEIGHBOURING filters: the falling side of filter s-1 and the rising side of filter s.  "Segment s" is the run of
// bins for which that holds; a lane that walks segment s reads each magnitude ONCE and feeds two accumulators,
// U_s (rising weights, -> filter s) and D_s (falling weights, -> filter s-1); filter m is U_m + D_(m+1).  Against the
// filter-major form (every magnitude read twice, every filter as long as two segments) this halves both the
// shared-memory reads and the tap iterations of the projection.  Returns false (and leaves mb.seg empty) when the
// bank does not have that structure; the kernels then use the banded form.
static bool build_mel_segments(MelBands &mb, const float *fb, size_t n_freq, size_t n_mel, int warps, int vec)
{
    mb.seg.clear(); mb.seg_nwq = mb.seg_nblk = 0; mb.seg_log2p = 0;
    if (n_mel == 0 || n_mel >= 0xfffe || n_freq > 0xffff) return false;
    const size_t S = n_mel + 1; // segments
    // ---- bin -> segment ---------------------------------------------------------------------------------
    std::vector<int> seg_of(n_freq, -1);
    for (size_t k = 0; k < n_freq; ++k) {
        int f0 = -1, nf = 0;
        for (size_t m = 0; m < n_mel; ++m)
            if (fb[k * n_mel + m] != 0.0f) { if (nf == 0) f0 = (int)m; else if ((int)m != f0 + 1 || nf > 1) return false; ++nf; }
        if (nf == 2) seg_of[k] = f0 + 1;
        else if (nf == 1) { // rising side of f0 up to and including its peak, falling side after it
            const int lo = mb.lo[f0], c = mb.cnt[f0];
            int peak = lo;
            for (int j = 1; j < c; ++j) if (mb.w[mb.off[f0] + j] > mb.w[mb.off[f0] + peak - lo]) peak = lo + j;
            seg_of[k] = (int)k <= peak ? f0 : f0 + 1;
        }
    }
    std::vector<int> first(S, -1), len(S, 0);
    int prev = -1;
    for (size_t k = 0; k < n_freq; ++k) {
        const int sg = seg_of[k];
        if (sg < 0) continue;
        if (sg < prev) return false;                                   // segments must rise with the bins
        if (len[sg] > 0 && first[sg] + len[sg] != (int)k) return false; // and be contiguous
        if (len[sg] == 0) first[sg] = (int)k;
        ++len[sg];
        prev = sg;
    }
    { // empty segments sit where their neighbours meet (their weights are zero anyway)
        int at = 0;
        for (size_t sg = 0; sg < S; ++sg) { if (len[sg] == 0) first[sg] = at; else at = first[sg] + len[sg]; }
    }
    // ---- lanes per segment: tap iterations + fixed cost per block of 32 lanes -----------------------------
    int best_lg = 0; double best_cost = 1e300;
    for (int lg = 0; lg <= 4; ++lg) {
        const int P = 1 << lg, U = 32 / P, outs = U - 1;
        double cost = 0.0;
        for (size_t s0 = 0; s0 < n_mel; s0 += outs) {
            int longest = 0;
            for (size_t sg = s0; sg < std::min(S, s0 + U); ++sg) longest = std::max(longest, (len[sg] + P - 1) / P);
            cost += std::max(2, (longest + 1) & ~1) + 5.0 + 1.0 * lg;
        }
        if (cost < best_cost) { best_cost = cost; best_lg = lg; }
    }
    const int lg = best_lg, P = 1 << lg, U = 32 / P, outs = U - 1;
    const size_t n_blocks = (n_mel + outs - 1) / outs;
    // magnitudes sit at their bin index as vectors of `vec` floats: a shared-memory phase serves 32 / vec lanes,
    // which collide when their bins agree modulo that number
    const int phase = std::max(1, 32 / std::max(1, vec)), mod = phase;
    const int limit = (int)n_freq; // every read stays inside [0, n_freq)
    struct Blk { int off, nj; std::vector<int> q; };
    std::vector<Blk> blks(n_blocks);
    size_t nwq = 0;
    for (size_t b = 0; b < n_blocks; ++b) {
        Blk &B = blks[b];
        int nj = 2;
        for (int u = 0; u < U; ++u) { const size_t sg = b * outs + u; if (sg < S) nj = std::max(nj, ((len[sg] + P - 1) / P + 1) & ~1); }
        for (;; nj += 2) { // shifts: a window may start q P bins early on zero weights (free while it stays within nj taps)
            if ((nj + 1) * P > (int)n_freq) return false; // degenerate (a handful of bins): the banded form serves
            B.q.assign(U, 0);
            bool ok = true;
            std::vector<unsigned long long> occ((U * P + phase - 1) / phase, 0ull); // residues taken per phase
            for (int u = 0; u < U && ok; ++u) {
                const size_t sg = b * outs + u;
                const int f = sg < S ? first[sg] : limit - 1, need = sg < S ? (len[sg] + P - 1) / P : 0;
                const int q_hi = std::min(nj - need, f / P);
                const int over = f + (P - 1) + (nj - 1) * P - (limit - 1);
                const int q_lo = over > 0 ? (over + P - 1) / P : 0;
                if (q_lo > q_hi) { ok = false; break; }
                int best_q = q_lo, best_c = 1 << 30;
                for (int q = q_lo; q <= q_hi; ++q) {
                    int c = 0;
                    for (int pl = 0; pl < P; ++pl) {
                        const int lane = u * P + pl;
                        if ((occ[lane / phase] >> ((f + pl - q * P) % mod)) & 1ull) ++c;
                    }
                    c = c * 64 + (q - q_lo);
                    if (c < best_c) { best_c = c; best_q = q; }
                }
                B.q[u] = best_q;
                for (int pl = 0; pl < P; ++pl) { const int lane = u * P + pl; occ[lane / phase] |= 1ull << ((f + pl - best_q * P) % mod); }
            }
            if (ok) break;
        }
        B.off = (int)nwq; B.nj = nj;
        nwq += (size_t)32 * nj;
    }
    // ---- schedule: blocks longest-first onto the warps of a thread group ------------------------------------
    std::vector<std::pair<int, int>> cost(n_blocks);
    for (size_t b = 0; b < n_blocks; ++b) cost[b] = {blks[b].nj + 6, (int)b};
    std::sort(cost.begin(), cost.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first > b.first; });
    std::vector<std::vector<int>> lists(warps);
    std::vector<long> load(warps, 0);
    for (auto &c : cost) {
        const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        lists[w].push_back(c.second); load[w] += c.first;
    }
    size_t slots = 0;
    for (auto &l : lists) slots = std::max(slots, l.size());
    // ---- words: weight pairs | lane descriptors | block descriptors | schedule ------------------------------
    mb.seg.assign(2 * nwq + 32 * n_blocks + 2 * n_blocks + 1 + slots * warps, 0);
    float *wq = reinterpret_cast<float *>(mb.seg.data());
    int *lo_item = mb.seg.data() + 2 * nwq;
    int *desc = lo_item + 32 * n_blocks;
    int *sched = desc + 2 * n_blocks;
    sched[0] = (int)slots;
    for (size_t i = 0; i < slots * warps; ++i) sched[1 + i] = -1;
    for (int w = 0; w < warps; ++w)
        for (size_t i = 0; i < lists[w].size(); ++i) sched[1 + i * warps + w] = lists[w][i];
    for (size_t b = 0; b < n_blocks; ++b) {
        const Blk &B = blks[b];
        for (int u = 0; u < U; ++u) {
            const size_t sg = b * outs + u;
            const int f = sg < S ? first[sg] : limit - 1, c = sg < S ? len[sg] : 0, q = B.q[u];
            const int filt = (u < outs && sg < n_mel) ? (int)sg : 0xffff; // the last unit of a block only lends its D
            for (int pl = 0; pl < P; ++pl) {
                const size_t lane = (size_t)u * P + pl;
                lo_item[b * 32 + lane] = (f + pl - q * P) | (filt << 16);
                for (int j = 0; pl + j * P < c; ++j) {
                    const size_t k = (size_t)f + pl + (size_t)j * P;
                    const size_t at = 2 * ((size_t)B.off + 32 * (size_t)(j + q) + lane);
                    wq[at] = sg < n_mel ? fb[k * n_mel + sg] : 0.0f;      // rising side of filter sg
                    wq[at + 1] = sg >= 1 ? fb[k * n_mel + sg - 1] : 0.0f; // falling side of filter sg - 1
                }
            }
        }
        desc[2 * b] = B.off; desc[2 * b + 1] = B.nj;
    }
    mb.seg_nwq = (int)nwq; mb.seg_nblk = (int)n_blocks; mb.seg_log2p = lg; mb.seg_slots = (int)slots;
    return true;
}

MelBands make_mel_bands(const float *fb, size_t n_freq, size_t n_mel, int threads_per_group,
                        size_t stage_capacity_floats, int vec)
{
    MelBands mb;
    mb.lo.resize(n_mel); mb.cnt.resize(n_mel); mb.off.resize(n_mel);
    for (size_t m = 0; m < n_mel; ++m) {
        long lo = -1, hi = -1;
        for (size_t b = 0; b < n_freq; ++b)
            if (fb[b * n_mel + m] != 0.0f) { if (lo < 0) lo = (long)b; hi = (long)b; }
        mb.off[m] = (int)mb.w.size();
        if (lo < 0) { mb.lo[m] = 0; mb.cnt[m] = 0; continue; }
        mb.lo[m] = (int)lo;
        mb.cnt[m] = (int)(hi - lo + 1);
        mb.max_cnt = std::max(mb.max_cnt, mb.cnt[m]);
        for (long b = lo; b <= hi; ++b) mb.w.push_back(fb[(size_t)b * n_mel + m]);
    }
    if (mb.w.empty()) mb.w.push_back(0.0f);
    // lanes per filter: minimise   sum over rounds of (longest band in the round / P)  + shuffles
    int best = 0; double best_cost = 1e300;
    for (int lg = 0; lg <= 5; ++lg) {
        const int P = 1 << lg;
        const size_t per_round = (size_t)std::max(1, threads_per_group / P);
        double cost = 0.0;
        for (size_t m0 = 0; m0 < n_mel; m0 += per_round) {
            int longest = 0;
            for (size_t m = m0; m < std::min(n_mel, m0 + per_round); ++m)
                longest = std::max(longest, mb.cnt[m]);
            cost += 2.0 * std::ceil((double)longest / P) + 1.5 * lg + 3.0;
        }
        if (cost < best_cost) { best_cost = cost; best = lg; }
    }
    mb.log2_split = best;
    // block schedule: 32 consecutive work items (filter, lane-of-filter) per block, blocks packed
    // longest-first onto the warps of a thread group
    // cost of a block in tap iterations (~7 instructions each): its longest lane plus the fixed part --
    // descriptor fetch, lane reduction, dB and the stores -- which weighs about ten of them
    int kMelBlockFixed = 10;
    if (const char *e = getenv("SGX_MEL_FIXED")) kMelBlockFixed = atoi(e);
    const int P = 1 << best, warps = std::max(1, threads_per_group / 32);
    const size_t n_items = n_mel * (size_t)P, n_blocks = (n_items + 31) / 32;
    std::vector<std::pair<int, int>> cost(n_blocks); // (cost, block)
    for (size_t b = 0; b < n_blocks; ++b) {
        int longest = 0;
        for (size_t it = b * 32; it < std::min(n_items, (b + 1) * 32); ++it)
            longest = std::max(longest, (mb.cnt[it / P] + P - 1) / P);
        cost[b] = {longest + kMelBlockFixed, (int)b};
    }
    std::sort(cost.begin(), cost.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first > b.first; });
    std::vector<std::vector<int>> lists(warps);
    std::vector<long> load(warps, 0);
    for (auto &c : cost) {
        const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        lists[w].push_back(c.second); load[w] += c.first;
    }
    size_t slots = 0;
    for (auto &l : lists) slots = std::max(slots, l.size());
    const size_t nnz = mb.w.size();
    const bool staged = ((nnz + 3) & ~(size_t)3) + 4 * n_mel <= stage_capacity_floats;
    mb.sched.assign(4 + slots * warps, -1);
    mb.sched[0] = (int)slots; mb.sched[1] = (int)nnz; mb.sched[2] = staged ? 1 : 0; mb.sched[3] = 0;
    for (int w = 0; w < warps; ++w)
        for (size_t i = 0; i < lists[w].size(); ++i) mb.sched[4 + i * warps + w] = lists[w][i];
    build_mel_segments(mb, fb, n_freq, n_mel, warps, vec);
    return mb;
}

} // namespace sgx
