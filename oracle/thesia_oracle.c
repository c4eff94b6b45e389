/*
 * thesia_oracle.c -- CPU restatement of the reference's spectrogram -> pixels path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under multi-spectrogram-viewer_b200/ may link, import
 * or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU arm.
 *
 * Parity status: the reference is Rust and no Rust toolchain exists in this image, so the
 * reference itself cannot be run here.  This restatement is pinned by the reference's own
 * known-answer tests (tests/test_oracle_kats.py): stft_works (lib.rs:491-514),
 * hann_window_works (windows.rs:35-38), pad_works (utils.rs:125-140), rfft_wrapper_works
 * (utils.rs:117-123), real_to_complex (realfft.rs:253-272), mel_hz_convert (mel.rs:107-113),
 * mel_default_works (mel.rs:135-165).  The stages the reference never asserts on (dense mel
 * projection order, dB, global range, grey, `image` 0.23 Lanczos3 resize, colormap) are
 * restated from the source text (and, for `image::imageops::resize`, from the published
 * algorithm of image 0.23.x, which is NOT vendored in /root/reference): PARITY UNPINNED there.
 *
 * All arithmetic is f32 with the reference's operation order; build with -ffp-contract=off so
 * gcc does not fuse a*b+c (rustc never contracts).  *_f64 entry points are the "truth" twin.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static const double PI64 = 3.14159265358979323846264338327950288;

/* ------------------------------------------------------------------------------------------
 * windows.rs:7-30  cosine_window / hann ;  lib.rs:138-140 calc_window
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_hann(size_t size, int symmetric, float *out)
{
    /* windows.rs:9-18: pi = f32(PI_f64); size2 = size or size+1;
     * x = pi * i / (size2-1); (a - b*cos(2x)) + (c*cos(4x) - d*cos(6x)) with a=b=.5, c=d=0 */
    const float pi = (float)PI64;
    size_t size2 = symmetric ? size : size + 1;
    for (size_t i = 0; i < size; ++i) {
        float x = pi * (float)i / (float)(size2 - 1);
        float b_ = 0.5f * cosf(2.0f * x);
        float c_ = 0.0f * cosf(4.0f * x);
        float d_ = 0.0f * cosf(6.0f * x);
        out[i] = (0.5f - b_) + (c_ - d_);
    }
}

ORC_API void orc_hann_f64(size_t size, int symmetric, double *out)
{
    size_t size2 = symmetric ? size : size + 1;
    for (size_t i = 0; i < size; ++i) {
        double x = PI64 * (double)i / (double)(size2 - 1);
        out[i] = (0.5 - 0.5 * cos(2.0 * x)) + (0.0 * cos(4.0 * x) - 0.0 * cos(6.0 * x));
    }
}

ORC_API void orc_calc_window(size_t win_length, size_t n_fft, float *out)
{
    /* lib.rs:138-140: hann(win_length, false) / n_fft as f32 */
    orc_hann(win_length, 0, out);
    float d = (float)n_fft;
    for (size_t i = 0; i < win_length; ++i) out[i] = out[i] / d;
}

/* utils.rs:17-19 calc_proper_n_fft: 2^(ceil(log2(win as f32))) */
ORC_API size_t orc_calc_proper_n_fft(size_t win_length)
{
    float l = ceilf(log2f((float)win_length));
    return (size_t)1 << (unsigned)l;
}

/* lib.rs:43-46 AudioTrack::new parameter derivation (win_ms, t_overlap, f_overlap). */
ORC_API void orc_track_params(uint32_t sr, float win_ms, size_t t_overlap, size_t f_overlap,
                              size_t *win_length, size_t *hop_length, size_t *n_fft)
{
    float w0 = win_ms * (float)sr / 1000.0f;
    size_t hop = (size_t)roundf(w0 / (float)t_overlap); /* Rust round = half away from zero */
    size_t win = hop * t_overlap;
    *hop_length = hop;
    *win_length = win;
    *n_fft = orc_calc_proper_n_fft(win) * f_overlap;
}

/* ------------------------------------------------------------------------------------------
 * utils.rs:59-87 pad (1-D only; the path only pads along the sample axis)
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_pad_reflect(const float *x, size_t n, size_t left, size_t right, float *out)
{
    /* utils.rs:79-85: left = x[1..=left] reversed, right = x[n-1-right .. n-1] reversed */
    if (left + 1 > n || right + 1 > n) return -1; /* the reference panics on the slice */
    for (size_t i = 0; i < left; ++i) out[i] = x[left - i];
    memcpy(out + left, x, n * sizeof(float));
    for (size_t i = 0; i < right; ++i) out[left + n + i] = x[n - 2 - i];
    return 0;
}

ORC_API void orc_pad_constant(const float *x, size_t n, size_t left, size_t right, float c,
                              float *out)
{
    /* utils.rs:70-78 */
    for (size_t i = 0; i < left; ++i) out[i] = c;
    memcpy(out + left, x, n * sizeof(float));
    for (size_t i = 0; i < right; ++i) out[left + n + i] = c;
}

/* ------------------------------------------------------------------------------------------
 * realfft.rs:80-159 RealFFT::new / process, on top of a power-of-two complex FFT standing in
 * for rustfft 4.0 `Radix4` (third-party, not vendored; twiddles computed in f64 and cast, as
 * rustfft does).  Output is the plain forward DFT up to f32 round-off.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float re, im; } cf32;
typedef struct { double re, im; } cf64;

typedef struct {
    size_t length;   /* real length F */
    size_t h;        /* F/2 */
    float *sin_;     /* realfft.rs:88-93, evaluated in f32 */
    float *cos_;
    cf32 *tw;        /* W_h^k, k in [0,h), f64-derived */
    cf32 *buf;       /* realfft.rs:46 buffer_out (h+1) */
    cf32 *scratch;
} orc_rfft_plan;

ORC_API orc_rfft_plan *orc_rfft_plan_new(size_t length)
{
    if (length < 2 || (length & (length - 1)) != 0) return NULL; /* Radix4: power of two */
    orc_rfft_plan *p = (orc_rfft_plan *)calloc(1, sizeof(*p));
    size_t h = length / 2;
    p->length = length;
    p->h = h;
    p->sin_ = (float *)malloc(sizeof(float) * h);
    p->cos_ = (float *)malloc(sizeof(float) * h);
    p->tw = (cf32 *)malloc(sizeof(cf32) * (h ? h : 1));
    p->buf = (cf32 *)malloc(sizeof(cf32) * (h + 1));
    p->scratch = (cf32 *)malloc(sizeof(cf32) * (h ? h : 1));
    /* realfft.rs:86-93: pi = f32(PI), halflength = f32(h); sin(k*pi/halflength) in f32 */
    const float pi = (float)PI64;
    const float halflength = (float)h;
    for (size_t k = 0; k < h; ++k) {
        float kf = (float)k;
        p->sin_[k] = sinf(kf * pi / halflength);
        p->cos_[k] = cosf(kf * pi / halflength);
    }
    for (size_t k = 0; k < h; ++k) {
        double a = -2.0 * PI64 * (double)k / (double)h;
        p->tw[k].re = (float)cos(a);
        p->tw[k].im = (float)sin(a);
    }
    return p;
}

ORC_API void orc_rfft_plan_free(orc_rfft_plan *p)
{
    if (!p) return;
    free(p->sin_); free(p->cos_); free(p->tw); free(p->buf); free(p->scratch);
    free(p);
}

/* forward complex FFT of length h (power of two), in: x[h] -> out[h]; f32 arithmetic */
static void cfft_f32(const orc_rfft_plan *p, const cf32 *x, cf32 *out)
{
    size_t h = p->h;
    if (h == 1) { out[0] = x[0]; return; }
    unsigned lg = 0;
    while (((size_t)1 << lg) < h) ++lg;
    /* bit-reversed load; DIT stages then produce natural order.  When lg is odd one radix-2
     * stage runs first, the rest are radix-4 stages (two radix-2 levels merged). */
    for (size_t i = 0; i < h; ++i) {
        size_t r = 0, v = i;
        for (unsigned b = 0; b < lg; ++b) { r = (r << 1) | (v & 1); v >>= 1; }
        out[r] = x[i];
    }
    size_t len = 1; /* current transformed sub-length */
    if (lg & 1) {
        for (size_t i = 0; i < h; i += 2) {
            cf32 a = out[i], b = out[i + 1];
            out[i].re = a.re + b.re; out[i].im = a.im + b.im;
            out[i + 1].re = a.re - b.re; out[i + 1].im = a.im - b.im;
        }
        len = 2;
    }
    /* radix-4 DIT stages expressed on bit-reversed data: a radix-4 butterfly over sub-length
     * `len` combines 4 sub-DFTs E0,E1,E2,E3 stored (because of bit reversal) in the order
     * n mod 4 = 0,2,1,3. */
    while (len < h) {
        size_t m = len * 4;
        size_t tstep = h / m;
        for (size_t base = 0; base < h; base += m) {
            for (size_t k = 0; k < len; ++k) {
                cf32 w1 = p->tw[k * tstep];
                cf32 w2 = p->tw[2 * k * tstep];
                cf32 w3 = p->tw[3 * k * tstep]; /* 3k*tstep < 3h/4 */
                cf32 a0 = out[base + k];
                cf32 a2 = out[base + k + len];       /* sub-DFT of n = 2 mod 4 */
                cf32 a1 = out[base + k + 2 * len];   /* sub-DFT of n = 1 mod 4 */
                cf32 a3 = out[base + k + 3 * len];   /* sub-DFT of n = 3 mod 4 */
                cf32 b1, b2, b3;
                b1.re = a1.re * w1.re - a1.im * w1.im; b1.im = a1.re * w1.im + a1.im * w1.re;
                b2.re = a2.re * w2.re - a2.im * w2.im; b2.im = a2.re * w2.im + a2.im * w2.re;
                b3.re = a3.re * w3.re - a3.im * w3.im; b3.im = a3.re * w3.im + a3.im * w3.re;
                cf32 s02, d02, s13, d13;
                s02.re = a0.re + b2.re; s02.im = a0.im + b2.im;
                d02.re = a0.re - b2.re; d02.im = a0.im - b2.im;
                s13.re = b1.re + b3.re; s13.im = b1.im + b3.im;
                d13.re = b1.re - b3.re; d13.im = b1.im - b3.im;
                /* forward transform: -i * d13 = (d13.im, -d13.re) */
                out[base + k].re = s02.re + s13.re;           out[base + k].im = s02.im + s13.im;
                out[base + k + len].re = d02.re + d13.im;     out[base + k + len].im = d02.im - d13.re;
                out[base + k + 2 * len].re = s02.re - s13.re; out[base + k + 2 * len].im = s02.im - s13.im;
                out[base + k + 3 * len].re = d02.re - d13.im; out[base + k + 3 * len].im = d02.im + d13.re;
            }
        }
        len = m;
    }
}

/* realfft.rs:105-159.  `input` (length F) is consumed (the reference uses it as scratch). */
ORC_API void orc_rfft_process(orc_rfft_plan *p, float *input, float *output /* (h+1)*2 */)
{
    size_t h = p->h;
    cf32 *out = (cf32 *)output;
    cf32 *buf = p->buf;
    /* :130-138 reinterpret pairs as complex, FFT of length h into buffer_out[0..h] */
    cfft_f32(p, (const cf32 *)input, buf);
    /* :140 */
    buf[h] = buf[0];
    /* :142-156 */
    for (size_t k = 0; k < h; ++k) {
        cf32 a = buf[k], b = buf[h - k];
        float s = p->sin_[k], c = p->cos_[k];
        float xr = 0.5f * (((a.re + b.re) + c * (a.im + b.im)) - s * (a.re - b.re));
        float xi = 0.5f * (((a.im - b.im) - s * (a.im + b.im)) - c * (a.re - b.re));
        out[k].re = xr;
        out[k].im = xi;
    }
    /* :157 */
    out[h].re = buf[0].re - buf[0].im;
    out[h].im = 0.0f;
}

/* f64 truth: direct split-free rfft via a double complex FFT of length F (simple radix-2) */
static void cfft_f64(cf64 *a, size_t n)
{
    unsigned lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    for (size_t i = 0; i < n; ++i) {
        size_t r = 0, v = i;
        for (unsigned b = 0; b < lg; ++b) { r = (r << 1) | (v & 1); v >>= 1; }
        if (r > i) { cf64 t = a[i]; a[i] = a[r]; a[r] = t; }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t base = 0; base < n; base += len) {
            for (size_t k = 0; k < len / 2; ++k) {
                double ang = -2.0 * PI64 * (double)k / (double)len;
                double wr = cos(ang), wi = sin(ang);
                cf64 u = a[base + k], v = a[base + k + len / 2];
                double tr = v.re * wr - v.im * wi, ti = v.re * wi + v.im * wr;
                a[base + k].re = u.re + tr; a[base + k].im = u.im + ti;
                a[base + k + len / 2].re = u.re - tr; a[base + k + len / 2].im = u.im - ti;
            }
        }
    }
}

/* f64 RealFFT used by the real_to_complex KAT (realfft.rs:253-272 runs RealFFT::<f64>):
 * same packing + split algorithm in double. */
ORC_API int orc_rfft_f64(const double *input, size_t length, double *output /* (h+1)*2 */)
{
    if (length < 2 || (length & (length - 1)) != 0) return -1;
    size_t h = length / 2;
    cf64 *buf = (cf64 *)malloc(sizeof(cf64) * (h + 1));
    for (size_t i = 0; i < h; ++i) { buf[i].re = input[2 * i]; buf[i].im = input[2 * i + 1]; }
    if (h > 1) cfft_f64(buf, h);
    buf[h] = buf[0];
    cf64 *out = (cf64 *)output;
    for (size_t k = 0; k < h; ++k) {
        cf64 a = buf[k], b = buf[h - k];
        double s = sin((double)k * PI64 / (double)h), c = cos((double)k * PI64 / (double)h);
        out[k].re = 0.5 * (((a.re + b.re) + c * (a.im + b.im)) - s * (a.re - b.re));
        out[k].im = 0.5 * (((a.im - b.im) - s * (a.im + b.im)) - c * (a.re - b.re));
    }
    out[h].re = buf[0].re - buf[0].im;
    out[h].im = 0.0;
    free(buf);
    return 0;
}

/* plain complex f64 DFT via FFT for cross-checks (the "rustfft planner" side of the KAT) */
ORC_API int orc_cfft_f64(double *inout /* n*2 */, size_t n)
{
    if (n < 1 || (n & (n - 1)) != 0) return -1;
    if (n > 1) cfft_f64((cf64 *)inout, n);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * lib.rs:367-471  to_windowed_frames / perform_stft -- literal three-list construction.
 * ---------------------------------------------------------------------------------------- */
/* number of length-`w` windows at stride `hop` over `len` samples (ndarray .windows().step_by):
 * zero when len < w */
static size_t n_windows(size_t len, size_t w, size_t hop)
{
    if (len < w) return 0;
    return (len - w) / hop + 1;
}

ORC_API long orc_stft_n_frames(size_t n, size_t win, size_t hop)
{
    /* front: input[..win-1] left-reflect-padded by win/2   (lib.rs:412-418)
     * mid:   input[first_idx..]                            (lib.rs:420-421)
     * back:  right-reflect-padded tail                     (lib.rs:423-433) */
    if (win < 2 || hop < 1 || n < win || win / 2 + 1 > n) return -1;
    if (win / 2 + 1 > win - 1) return -1; /* front reflect needs x[1..=win/2] inside x[..win-1] */
    size_t n_front = n_windows(win - 1 + win / 2, win, hop);
    if (n_front * hop < win / 2) return -1; /* usize underflow in the reference */
    size_t first_idx = n_front * hop - win / 2;
    if (first_idx > n) return -1;
    size_t n_mid = n_windows(n - first_idx, win, hop);
    first_idx += n_mid * hop;
    size_t back_start = first_idx < n - win / 2 - 1 ? first_idx : n - win / 2 - 1;
    size_t back_len = (n - back_start) + win / 2 - (first_idx - back_start);
    size_t n_back = n_windows(back_len, win, hop);
    return (long)(n_front + n_mid + n_back);
}

/* Writes frame t (already windowed and centre-zero-padded to n_fft) into g[n_fft].
 * Literal: builds the three padded lists lazily per frame index. */
typedef struct {
    const float *x; size_t n, win, hop, n_fft, pad_l, pad_r;
    size_t n_front, n_mid, n_back, mid_first, back_first, back_start;
    float *front_wav; size_t front_len;
    float *back_wav; size_t back_len; /* after slice_collapse */
    float *back_alloc;
} framer;

static int framer_init(framer *f, const float *x, size_t n, size_t win, size_t hop, size_t n_fft)
{
    memset(f, 0, sizeof(*f));
    f->x = x; f->n = n; f->win = win; f->hop = hop; f->n_fft = n_fft;
    f->pad_l = (n_fft - win) / 2;                                  /* lib.rs:400 */
    f->pad_r = (size_t)ceilf(((float)(n_fft - win)) / 2.0f);       /* lib.rs:401 */
    if (orc_stft_n_frames(n, win, hop) < 0) return -1;
    /* lib.rs:412-418 */
    f->front_len = win - 1 + win / 2;
    f->front_wav = (float *)malloc(sizeof(float) * f->front_len);
    if (orc_pad_reflect(x, win - 1, win / 2, 0, f->front_wav)) return -1;
    f->n_front = n_windows(f->front_len, win, hop);
    /* lib.rs:420-421 */
    f->mid_first = f->n_front * hop - win / 2;
    f->n_mid = n_windows(n - f->mid_first, win, hop);
    /* lib.rs:423-433 */
    size_t first_idx = f->mid_first + f->n_mid * hop;
    size_t back_start = first_idx < n - win / 2 - 1 ? first_idx : n - win / 2 - 1;
    size_t blen = n - back_start + win / 2;
    f->back_alloc = (float *)malloc(sizeof(float) * blen);
    if (orc_pad_reflect(x + back_start, n - back_start, 0, win / 2, f->back_alloc)) return -1;
    size_t skip = first_idx - back_start;
    f->back_wav = f->back_alloc + skip;
    f->back_len = blen - skip;
    f->n_back = n_windows(f->back_len, win, hop);
    return 0;
}

static void framer_free(framer *f) { free(f->front_wav); free(f->back_alloc); }

static void framer_frame(const framer *f, const float *window, size_t t, float *g)
{
    const float *src;
    if (t < f->n_front) src = f->front_wav + t * f->hop;
    else if (t < f->n_front + f->n_mid) src = f->x + f->mid_first + (t - f->n_front) * f->hop;
    else src = f->back_wav + (t - f->n_front - f->n_mid) * f->hop;
    /* lib.rs:377-384: pad(x * window, (n_pad_left, n_pad_right), Constant(0)) */
    for (size_t i = 0; i < f->pad_l; ++i) g[i] = 0.0f;
    for (size_t j = 0; j < f->win; ++j) g[f->pad_l + j] = src[j] * window[j];
    for (size_t i = 0; i < f->pad_r; ++i) g[f->pad_l + f->win + i] = 0.0f;
}

/* perform_stft: out is [T][n_fft/2+1] complex (re,im interleaved), C order (lib.rs:435-440).
 * window == NULL -> hann(win,false)/n_fft (lib.rs:403-408).
 * plan_per_frame != 0 reproduces the `parallel` branch that builds a fresh RealFFT per frame
 * (lib.rs:449-458); results are identical, only the cost differs.  Returns T or <0. */
ORC_API long orc_perform_stft(const float *input, size_t n, size_t win, size_t hop, size_t n_fft,
                              const float *window, float *out, int parallel)
{
    if (n_fft < win || n_fft < 2 || (n_fft & (n_fft - 1))) return -1;
    framer f;
    if (framer_init(&f, input, n, win, hop, n_fft)) return -1;
    float *wbuf = NULL;
    if (!window) {
        wbuf = (float *)malloc(sizeof(float) * win);
        orc_calc_window(win, n_fft, wbuf);
        window = wbuf;
    }
    size_t T = f.n_front + f.n_mid + f.n_back;
    size_t B = n_fft / 2 + 1;
    if (parallel) {
#pragma omp parallel
        {
            float *g = (float *)malloc(sizeof(float) * n_fft);
#pragma omp for schedule(static)
            for (long t = 0; t < (long)T; ++t) {
                orc_rfft_plan *p = orc_rfft_plan_new(n_fft); /* lib.rs:455 */
                framer_frame(&f, window, (size_t)t, g);
                orc_rfft_process(p, g, out + (size_t)t * B * 2);
                orc_rfft_plan_free(p);
            }
            free(g);
        }
    } else {
        orc_rfft_plan *p = orc_rfft_plan_new(n_fft);
        float *g = (float *)malloc(sizeof(float) * n_fft);
        for (size_t t = 0; t < T; ++t) {
            framer_frame(&f, window, t, g);
            orc_rfft_process(p, g, out + t * B * 2);
        }
        free(g);
        orc_rfft_plan_free(p);
    }
    free(wbuf);
    framer_free(&f);
    return (long)T;
}

/* f64 truth STFT magnitude: frames by the closed form (reflect index), double FFT of length
 * n_fft.  out_mag [T][B] double.  window is the f32 table promoted to double (the table is an
 * INPUT of the path, so truth uses the same table). */
ORC_API long orc_stft_mag_f64(const float *input, size_t n, size_t win, size_t hop, size_t n_fft,
                              const float *window, double *out_mag)
{
    long Tl = orc_stft_n_frames(n, win, hop);
    if (Tl < 0) return -1;
    size_t T = (size_t)Tl, B = n_fft / 2 + 1, pad_l = (n_fft - win) / 2;
    float *wbuf = NULL;
    if (!window) {
        wbuf = (float *)malloc(sizeof(float) * win);
        orc_calc_window(win, n_fft, wbuf);
        window = wbuf;
    }
#pragma omp parallel
    {
        cf64 *a = (cf64 *)malloc(sizeof(cf64) * n_fft);
#pragma omp for schedule(static)
        for (long t = 0; t < (long)T; ++t) {
            for (size_t i = 0; i < n_fft; ++i) { a[i].re = 0; a[i].im = 0; }
            for (size_t j = 0; j < win; ++j) {
                long i = (long)t * (long)hop + (long)j - (long)(win / 2);
                if (i < 0) i = -i;
                if (i >= (long)n) i = 2 * ((long)n - 1) - i;
                a[pad_l + j].re = (double)input[i] * (double)window[j];
            }
            cfft_f64(a, n_fft);
            for (size_t k = 0; k < B; ++k)
                out_mag[(size_t)t * B + k] = hypot(a[k].re, a[k].im);
        }
        free(a);
    }
    free(wbuf);
    return Tl;
}

/* ------------------------------------------------------------------------------------------
 * mel.rs:8-99
 * ---------------------------------------------------------------------------------------- */
#define MIN_LOG_MEL 15
static const double MIN_LOG_HZ = 1000.0;
static const double LOGSTEP = 0.06875177742094912;
static const double LINEARSCALE = 200.0 / 3.0;

ORC_API float orc_mel_to_hz(float mel)
{   /* mel.rs:14-21, A = f32: constants are cast to f32 first */
    float min_log_mel = (float)MIN_LOG_MEL;
    if (mel < min_log_mel) return (float)LINEARSCALE * mel;
    return (float)MIN_LOG_HZ * expf((float)LOGSTEP * (mel - min_log_mel));
}
ORC_API float orc_hz_to_mel(float freq)
{   /* mel.rs:24-31 */
    float min_log_hz = (float)MIN_LOG_HZ;
    if (freq < min_log_hz) return freq / (float)LINEARSCALE;
    return (float)MIN_LOG_MEL + logf(freq / min_log_hz) / (float)LOGSTEP;
}
ORC_API double orc_mel_to_hz_f64(double mel)
{
    if (mel < (double)MIN_LOG_MEL) return LINEARSCALE * mel;
    return MIN_LOG_HZ * exp(LOGSTEP * (mel - (double)MIN_LOG_MEL));
}
ORC_API double orc_hz_to_mel_f64(double freq)
{
    if (freq < MIN_LOG_HZ) return freq / LINEARSCALE;
    return (double)MIN_LOG_MEL + log(freq / MIN_LOG_HZ) / LOGSTEP;
}

/* ndarray Array::linspace(a, b, n): step = (b-a)/(n-1); element i = a + step*i  (f32) */
static void linspace_f32(float a, float b, size_t n, float *out)
{
    float step = n > 1 ? (b - a) / (float)(n - 1) : 0.0f;
    for (size_t i = 0; i < n; ++i) out[i] = a + step * (float)i;
}

/* mel.rs:33-85 calc_mel_fb::<f32>; fmax < 0 means None.  out is [n_freq][n_mel] C order. */
ORC_API void orc_calc_mel_fb(uint32_t sr, size_t n_fft, size_t n_mel, float fmin, float fmax,
                             int do_norm, float *out)
{
    float f_nyquist = ((float)sr) / 2.0f;
    if (fmax < 0.0f) fmax = f_nyquist;
    size_t n_freq = n_fft / 2 + 1;
    float min_mel = orc_hz_to_mel(fmin), max_mel = orc_hz_to_mel(fmax);
    float *lin = (float *)malloc(sizeof(float) * n_freq);
    float *mf = (float *)malloc(sizeof(float) * (n_mel + 2));
    linspace_f32(0.0f, f_nyquist, n_freq, lin);
    linspace_f32(min_mel, max_mel, n_mel + 2, mf);
    for (size_t i = 0; i < n_mel + 2; ++i) mf[i] = orc_mel_to_hz(mf[i]);
    memset(out, 0, sizeof(float) * n_freq * n_mel);
    for (size_t m = 0; m < n_mel; ++m) {
        for (size_t i = 0; i < n_freq; ++i) {
            float f = lin[i];
            if (f <= mf[m]) continue;
            else if (mf[m] < f && f < mf[m + 1]) out[i * n_mel + m] = (f - mf[m]) / (mf[m + 1] - mf[m]);
            else if (f == mf[m + 1]) out[i * n_mel + m] = 1.0f;
            else if (mf[m + 1] < f && f < mf[m + 2]) out[i * n_mel + m] = (mf[m + 2] - f) / (mf[m + 2] - mf[m + 1]);
            else break;
        }
        if (do_norm) {
            /* mel.rs:80-82: w /= w.sum().max(EPSILON); ndarray sum of a strided column:
             * sequential f32 accumulation is the restatement (unrolled order is unpinned) */
            float s = 0.0f;
            for (size_t i = 0; i < n_freq; ++i) s += out[i * n_mel + m];
            float d = s > 1.1920929e-07f ? s : 1.1920929e-07f;
            for (size_t i = 0; i < n_freq; ++i) out[i * n_mel + m] = out[i * n_mel + m] / d;
        }
    }
    free(lin); free(mf);
}

/* mel.rs:87-99.  Returns n_mel; writes the bank if out != NULL (needs cap >= n_freq*n_mel). */
ORC_API size_t orc_calc_mel_fb_default(uint32_t sr, size_t n_fft, float *out, size_t cap)
{
    size_t n_freq = n_fft / 2 + 1;
    float v = 2.0f * orc_hz_to_mel((float)sr / 2.0f) / orc_hz_to_mel((float)sr / (float)n_fft) - 1.0f;
    size_t n_mel = v <= 0.0f ? 0 : (size_t)v; /* `as usize` truncates, saturates at 0 */
    if (n_mel > n_freq) n_mel = n_freq;
    float *fb = (float *)malloc(sizeof(float) * n_freq * (n_mel ? n_mel : 1));
    for (;;) {
        orc_calc_mel_fb(sr, n_fft, n_mel, 0.0f, -1.0f, 1, fb);
        int ok = 1;
        for (size_t m = 0; m < n_mel && ok; ++m) {
            float s = 0.0f;
            for (size_t i = 0; i < n_freq; ++i) s += fb[i * n_mel + m];
            if (!(s > 0.0f)) ok = 0;
        }
        if (ok) break;
        n_mel -= 1;
    }
    if (out && cap >= n_freq * n_mel) memcpy(out, fb, sizeof(float) * n_freq * n_mel);
    free(fb);
    return n_mel;
}

/* ------------------------------------------------------------------------------------------
 * decibel.rs:33-88 amp_to_db_default (ref = 1, amin = 1e-18): two passes, two roundings.
 * Returns -1 if any element is negative/NaN (the reference asserts, decibel.rs:34).
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_amp_to_db_default(float *x, size_t n)
{
    const float amin = 1e-18f, refv = 1.0f;
    for (size_t i = 0; i < n; ++i) if (!(x[i] >= 0.0f)) return -1;
    float log_amin = log10f(amin);
    float log_ref = refv > amin ? log10f(refv) : log_amin;
    for (size_t i = 0; i < n; ++i) x[i] = x[i] > amin ? log10f(x[i]) - log_ref : log_amin - log_ref;
    for (size_t i = 0; i < n; ++i) x[i] = 20.0f * x[i];
    return 0;
}

/* lib.rs:112-136 calc_spec_of.  mel_fb == NULL -> FreqScale::Linear.  out [T][n_out].
 * dense != 0: the dense `.dot` of lib.rs:131 (every zero multiplied); else skip zero weights
 * (same sums of the same non-zero terms in the same order; identical results because adding
 * +0.0 products of non-negative magnitudes never changes an f32 partial sum). */
ORC_API long orc_calc_spec(const float *wav, size_t n, size_t win, size_t hop, size_t n_fft,
                           const float *window, const float *mel_fb, size_t n_mel,
                           float *out, int parallel, int dense)
{
    long Tl = orc_stft_n_frames(n, win, hop);
    if (Tl < 0) return -1;
    size_t T = (size_t)Tl, B = n_fft / 2 + 1;
    float *stft = (float *)malloc(sizeof(float) * T * B * 2);
    if (orc_perform_stft(wav, n, win, hop, n_fft, window, stft, parallel) != Tl) { free(stft); return -1; }
    /* lib.rs:124 mapv(norm) -> hypot */
    float *lin = (float *)malloc(sizeof(float) * T * B);
#pragma omp parallel for schedule(static) if (parallel)
    for (long i = 0; i < (long)(T * B); ++i) lin[i] = hypotf(stft[2 * i], stft[2 * i + 1]);
    free(stft);
    if (!mel_fb) {
        memcpy(out, lin, sizeof(float) * T * B);
        free(lin);
        return orc_amp_to_db_default(out, T * B) ? -1 : Tl;
    }
    /* lib.rs:131 linspec.dot(mel_fb): single-threaded sgemm in the reference (matrixmultiply,
     * no rayon inside); k-sequential accumulation here (order unpinned, third-party). */
    int *lo = NULL, *hi = NULL;
    if (!dense) {
        lo = (int *)malloc(sizeof(int) * n_mel); hi = (int *)malloc(sizeof(int) * n_mel);
        for (size_t m = 0; m < n_mel; ++m) {
            int l = (int)B, h2 = -1;
            for (size_t k = 0; k < B; ++k) if (mel_fb[k * n_mel + m] != 0.0f) { if ((int)k < l) l = (int)k; h2 = (int)k; }
            lo[m] = l; hi[m] = h2;
        }
    }
#pragma omp parallel for schedule(static) if (parallel)
    for (long t = 0; t < (long)T; ++t) {
        const float *row = lin + (size_t)t * B;
        float *o = out + (size_t)t * n_mel;
        if (dense) {
            for (size_t m = 0; m < n_mel; ++m) o[m] = 0.0f;
            for (size_t k = 0; k < B; ++k) {
                float a = row[k];
                const float *w = mel_fb + k * n_mel;
                for (size_t m = 0; m < n_mel; ++m) o[m] += a * w[m];
            }
        } else {
            for (size_t m = 0; m < n_mel; ++m) {
                float s = 0.0f;
                for (int k = lo[m]; k <= hi[m]; ++k) s += row[k] * mel_fb[(size_t)k * n_mel + m];
                o[m] = s;
            }
        }
    }
    free(lo); free(hi); free(lin);
    return orc_amp_to_db_default(out, T * n_mel) ? -1 : Tl;
}

/* f64 truth of calc_spec (dB): window/filterbank tables are the same f32 inputs. */
ORC_API long orc_calc_spec_f64(const float *wav, size_t n, size_t win, size_t hop, size_t n_fft,
                               const float *window, const float *mel_fb, size_t n_mel, double *out)
{
    long Tl = orc_stft_n_frames(n, win, hop);
    if (Tl < 0) return -1;
    size_t T = (size_t)Tl, B = n_fft / 2 + 1;
    double *mag = (double *)malloc(sizeof(double) * T * B);
    orc_stft_mag_f64(wav, n, win, hop, n_fft, window, mag);
    size_t n_out = mel_fb ? n_mel : B;
#pragma omp parallel for schedule(static)
    for (long t = 0; t < (long)T; ++t) {
        for (size_t m = 0; m < n_out; ++m) {
            double s;
            if (mel_fb) {
                s = 0.0;
                for (size_t k = 0; k < B; ++k) s += mag[(size_t)t * B + k] * (double)mel_fb[k * n_mel + m];
            } else s = mag[(size_t)t * B + m];
            out[(size_t)t * n_out + m] = s > 1e-18 ? 20.0 * log10(s) : -360.0;
        }
    }
    free(mag);
    return Tl;
}

/* ------------------------------------------------------------------------------------------
 * lib.rs:193-209 global range: per-spec max/min, reduce, clamp.  NaN elements: ndarray-stats
 * max()/min() return Err -> unwrap_or(-inf/+inf) (lib.rs:198-199).
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_spec_max_min(const float *spec, size_t n, float *mx, float *mn)
{
    float a = -INFINITY, b = INFINITY;
    int nan = 0;
    for (size_t i = 0; i < n; ++i) {
        float v = spec[i];
        if (v != v) { nan = 1; break; }
        if (v > a) a = v;
        if (v < b) b = v;
    }
    if (nan || n == 0) { a = -INFINITY; b = INFINITY; }
    *mx = a; *mn = b;
}

ORC_API void orc_clamp_range(float gmax, float gmin, float db_range, float *max_db, float *min_db)
{
    /* lib.rs:208-209: max = max.min(0.); min = min.max(max - db_range) */
    float mx = fminf(gmax, 0.0f);
    float mn = fmaxf(gmin, mx - db_range);
    *max_db = mx; *min_db = mn;
}

/* ------------------------------------------------------------------------------------------
 * display.rs:44-54 spec_to_grey.  spec [T][n_out]; grey is width=T, height rows, row-major.
 * Returns height.
 * ---------------------------------------------------------------------------------------- */
ORC_API uint32_t orc_grey_height(size_t n_out, float up_ratio)
{
    return (uint32_t)roundf((float)n_out * up_ratio); /* display.rs:45 */
}

ORC_API uint32_t orc_spec_to_grey(const float *spec, size_t T, size_t n_out, float up_ratio,
                                  float max, float min, float *grey)
{
    uint32_t height = orc_grey_height(n_out, up_ratio);
    for (uint32_t y = 0; y < height; ++y) {
        for (size_t x = 0; x < T; ++x) {
            float g;
            /* display.rs:47: y >= height - n_out as u32 (u32 arithmetic; height >= n_out when
             * up_ratio >= 1; if it were smaller the reference would overflow/panic) */
            if ((long)y >= (long)height - (long)n_out) {
                long col = (long)height - 1 - (long)y;
                float db = spec[x * n_out + (size_t)col];
                g = fminf(fmaxf((db - min) / (max - min), 0.0f), 1.0f);
            } else g = 0.0f;
            grey[(size_t)y * T + x] = g;
        }
    }
    return height;
}

/* ------------------------------------------------------------------------------------------
 * image 0.23.x imageops::resize(.., FilterType::Lanczos3) on ImageBuffer<Luma<f32>>
 * (called at display.rs:57; third-party `image = "0.23.12"`, Cargo.toml:20, NOT vendored:
 * restated from the published 0.23 algorithm -- vertical_sample then horizontal_sample,
 * per-pass clamp to [0, f32::MAX], weights summed then divided).  PARITY UNPINNED.
 * ---------------------------------------------------------------------------------------- */
static float sinc_f32(float t)
{
    float a = t * (float)PI64;
    if (t == 0.0f) return 1.0f;
    return sinf(a) / a;
}
static float lanczos3_kernel(float x)
{
    if (fabsf(x) < 3.0f) return sinc_f32(x) * sinc_f32(x / 3.0f);
    return 0.0f;
}

/* Geometry + weights of one output index along an axis: shared by both passes. */
ORC_API void orc_lanczos3_taps(uint32_t n_in, uint32_t n_out, uint32_t o, uint32_t *left_out,
                               uint32_t *right_out, float *ws /* >= right-left */, float *sum_out)
{
    float ratio = (float)n_in / (float)n_out;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float src_support = 3.0f * sratio;
    float inputx = ((float)o + 0.5f) * ratio;
    int64_t left = (int64_t)floorf(inputx - src_support);
    if (left < 0) left = 0;
    if (left > (int64_t)n_in - 1) left = (int64_t)n_in - 1;
    int64_t right = (int64_t)ceilf(inputx + src_support);
    if (right < left + 1) right = left + 1;
    if (right > (int64_t)n_in) right = (int64_t)n_in;
    inputx = inputx - 0.5f;
    float sum = 0.0f;
    for (int64_t i = left; i < right; ++i) {
        float w = lanczos3_kernel(((float)i - inputx) / sratio);
        if (ws) ws[i - left] = w;
        sum += w;
    }
    *left_out = (uint32_t)left; *right_out = (uint32_t)right; *sum_out = sum;
}

ORC_API uint32_t orc_lanczos3_max_taps(uint32_t n_in, uint32_t n_out)
{
    float ratio = (float)n_in / (float)n_out;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    return (uint32_t)(2.0f * 3.0f * sratio) + 3;
}

static void clamp_store(float t, float sum, float *dst)
{
    float v = t / sum;
    /* clamp(v, 0, f32::MAX): image's clamp is `if a < min {min} else if a > max {max} else {a}` */
    if (v < 0.0f) v = 0.0f; else if (v > 3.4028235e38f) v = 3.4028235e38f;
    *dst = v;
}

/* in: width x height (row-major) -> out: nwidth x nheight.  threads>1 parallelises columns /
 * rows with OpenMP (results are order-independent per pixel); the reference is single-threaded. */
ORC_API int orc_resize_lanczos3(const float *in, uint32_t width, uint32_t height, uint32_t nwidth,
                                uint32_t nheight, float *out, int parallel)
{
    if (!width || !height || !nwidth || !nheight) return -1;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)width * nheight);
    uint32_t mt_v = orc_lanczos3_max_taps(height, nheight), mt_h = orc_lanczos3_max_taps(width, nwidth);
    /* vertical_sample: for each output row, weights once, then every x */
#pragma omp parallel if (parallel)
    {
        float *ws = (float *)malloc(sizeof(float) * (mt_v > mt_h ? mt_v : mt_h));
#pragma omp for schedule(static)
        for (long oy = 0; oy < (long)nheight; ++oy) {
            uint32_t left, right; float sum;
            orc_lanczos3_taps(height, nheight, (uint32_t)oy, &left, &right, ws, &sum);
            for (uint32_t x = 0; x < width; ++x) {
                float t = 0.0f;
                for (uint32_t i = 0; i < right - left; ++i)
                    t += in[(size_t)(left + i) * width + x] * ws[i];
                clamp_store(t, sum, &tmp[(size_t)oy * width + x]);
            }
        }
#pragma omp for schedule(static)
        for (long ox = 0; ox < (long)nwidth; ++ox) {
            uint32_t left, right; float sum;
            orc_lanczos3_taps(width, nwidth, (uint32_t)ox, &left, &right, ws, &sum);
            for (uint32_t y = 0; y < nheight; ++y) {
                float t = 0.0f;
                for (uint32_t i = 0; i < right - left; ++i)
                    t += tmp[(size_t)y * width + left + i] * ws[i];
                clamp_store(t, sum, &out[(size_t)y * nwidth + ox]);
            }
        }
        free(ws);
    }
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * display.rs:10-42 COLORMAP / convert_grey_to_color ; display.rs:56-61 grey_to_rgb
 * ---------------------------------------------------------------------------------------- */
static const uint8_t COLORMAP[10][3] = {
    {0, 0, 4}, {27, 12, 65}, {74, 12, 107}, {120, 28, 109}, {165, 44, 96},
    {207, 68, 70}, {237, 105, 37}, {251, 155, 6}, {247, 209, 61}, {252, 255, 164}};

ORC_API void orc_get_colormap(uint8_t *out30) { memcpy(out30, COLORMAP, 30); } /* lib.rs:473-480 */

ORC_API int orc_convert_grey_to_color(float x, uint8_t *rgb)
{
    if (!(x >= 0.0f)) return -1; /* display.rs:25 assert */
    float position = 10.0f * x;
    float fl = floorf(position);
    /* `as usize` saturates; anything >= 9 takes the last colour */
    if (fl >= 9.0f) { memcpy(rgb, COLORMAP[9], 3); return 0; }
    size_t index = (size_t)fl;
    float ratio = position - (float)index;
    for (int c = 0; c < 3; ++c) {
        float a = (float)COLORMAP[index][c], b = (float)COLORMAP[index + 1][c];
        float v = roundf(ratio * b + (1.0f - ratio) * a); /* f32::round: half away from zero */
        rgb[c] = v <= 0.0f ? 0 : (v >= 255.0f ? 255 : (uint8_t)v);
    }
    return 0;
}

/* grey (width x height) -> RGB (channels=3) or RGBA (channels=4, A=255) nwidth x nheight */
ORC_API int orc_grey_to_rgb(const float *grey, uint32_t width, uint32_t height, uint32_t nwidth,
                            uint32_t nheight, int channels, uint8_t *out, int parallel)
{
    if (!nwidth || !nheight) return 0; /* an empty image: RgbImage::from_fn(0, h, ..) has no pixels */
    float *res = (float *)malloc(sizeof(float) * (size_t)nwidth * nheight);
    if (orc_resize_lanczos3(grey, width, height, nwidth, nheight, res, parallel)) { free(res); return -1; }
    int bad = 0;
#pragma omp parallel for schedule(static) if (parallel)
    for (long i = 0; i < (long)((size_t)nwidth * nheight); ++i) {
        uint8_t rgb[3];
        if (orc_convert_grey_to_color(res[i], rgb)) { bad = 1; continue; }
        uint8_t *o = out + (size_t)i * channels;
        o[0] = rgb[0]; o[1] = rgb[1]; o[2] = rgb[2];
        if (channels == 4) o[3] = 255;
    }
    free(res);
    return bad ? -1 : 0;
}

/* lib.rs:296 nwidth = (px_per_sec * len as f32 / sr as f32) as u32 */
ORC_API uint32_t orc_calc_nwidth(float px_per_sec, size_t n, uint32_t sr)
{
    float v = px_per_sec * (float)n / (float)sr;
    if (!(v > 0.0f)) return 0;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

/* lib.rs:231-248 up_ratio */
ORC_API float orc_up_ratio(uint32_t max_sr, uint32_t sr, int mel)
{
    if (!mel) return (float)max_sr / (float)sr;
    return orc_hz_to_mel((float)max_sr / 2.0f) / orc_hz_to_mel((float)sr / 2.0f);
}

/* ------------------------------------------------------------------------------------------
 * display.rs:63-115 wav_to_image (RGBA), the "next" row n2.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_wav_to_image(const float *wav_in, size_t n_in, uint32_t nwidth, uint32_t nheight,
                             float amp_min, float amp_max, uint8_t *out)
{
    static const uint8_t WAVECOLOR[4] = {200, 21, 103, 255};
    memset(out, 0, (size_t)nwidth * nheight * 4);
    if (!nwidth || !nheight) return 0;
    float samples_per_px = (float)n_in / (float)nwidth;
    const float *wav = wav_in;
    size_t n = n_in;
    float *up = NULL;
    if (samples_per_px < 1.0f) {
        size_t factor = (size_t)ceilf(1.0f / samples_per_px);
        n = factor * n_in;
        up = (float *)malloc(sizeof(float) * n);
        for (size_t i = 0; i < n; ++i) {
            float b = (i / factor + 1 < n_in) ? wav_in[i / factor + 1] : 0.0f;
            float fr = (float)(i % factor) / (float)factor;
            up[i] = b * fr + wav_in[i / factor] * (1.0f - fr);
        }
        wav = up;
    }
    for (int32_t i_px = 0; i_px < (int32_t)nwidth; ++i_px) {
        float fs = fmaxf(roundf(((float)i_px - 1.5f) * samples_per_px), 0.0f);
        size_t i_start = (size_t)fs;
        float fe = roundf(((float)i_px + 1.5f) * samples_per_px);
        size_t i_end = fe <= 0.0f ? 0 : (size_t)fe;
        if (i_end > n) i_end = n;
        if (i_start >= i_end) { free(up); return -1; } /* reference: max() of empty -> panic */
        float mx = wav[i_start], mn = wav[i_start];
        for (size_t i = i_start; i < i_end; ++i) { if (wav[i] > mx) mx = wav[i]; if (wav[i] < mn) mn = wav[i]; }
        long top = (long)roundf((amp_max - mx) * (float)nheight / (amp_max - amp_min));
        long bottom = (long)roundf((amp_max - mn) * (float)nheight / (amp_max - amp_min));
        if (bottom - top < 3) {
            long pad_bottom = (long)ceilf((float)(3 - bottom + top) / 2.0f);
            long pad_top = (long)floorf((float)(3 - bottom + top) / 2.0f);
            top -= pad_top; bottom += pad_bottom;
        }
        size_t t = top < 0 ? 0 : (size_t)top;
        long bl = bottom < (long)nheight ? bottom : (long)nheight;
        /* arr.slice_mut(s![top..bottom+1, ..]) : bottom+1 > nheight would panic in ndarray */
        if (bl + 1 > (long)nheight) bl = (long)nheight - 1;
        for (long y = (long)t; y <= bl; ++y) memcpy(out + ((size_t)y * nwidth + (size_t)i_px) * 4, WAVECOLOR, 4);
    }
    free(up);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Whole-pipeline driver used as the CPU baseline: mirrors MultiTrack::add_tracks +
 * get_spec_image for n tracks already decoded to mono f32 (lib.rs:142-191, 193-263, 294-298).
 * Parallel structure = the reference's rayon usage: over tracks when n_tracks > 1
 * (calc_spec_of(id, parallel = id_list.len()==1), lib.rs:161-166), else over frames; render is
 * single-threaded per track in the reference (display.rs has no rayon) -- tracks still render
 * independently here only when `parallel_render` is set.
 * out_images[i] must hold nwidth_i * nheight * channels bytes.  Returns 0 / <0.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_pipeline(size_t n_tracks, const float *const *wavs, const size_t *n_samples,
                         const uint32_t *srs, const size_t *wins, const size_t *hops,
                         const size_t *n_ffts, const float *const *windows,
                         const float *const *mel_fbs, const size_t *n_mels, int mel_scale,
                         float db_range, float px_per_sec, uint32_t nheight, int channels,
                         int dense_mel, int parallel_render, uint8_t *const *out_images,
                         float *out_max_db, float *out_min_db)
{
    float **specs = (float **)calloc(n_tracks, sizeof(float *));
    long *Ts = (long *)calloc(n_tracks, sizeof(long));
    size_t *n_outs = (size_t *)calloc(n_tracks, sizeof(size_t));
    int err = 0;
    int par_tracks = n_tracks > 1;
#pragma omp parallel for schedule(dynamic, 1) if (par_tracks)
    for (long i = 0; i < (long)n_tracks; ++i) {
        long T = orc_stft_n_frames(n_samples[i], wins[i], hops[i]);
        if (T < 0) { err = 1; continue; }
        size_t n_out = mel_fbs && mel_fbs[i] ? n_mels[i] : n_ffts[i] / 2 + 1;
        specs[i] = (float *)malloc(sizeof(float) * (size_t)T * n_out);
        Ts[i] = orc_calc_spec(wavs[i], n_samples[i], wins[i], hops[i], n_ffts[i], windows[i],
                              mel_fbs ? mel_fbs[i] : NULL, n_mels ? n_mels[i] : 0, specs[i],
                              !par_tracks, dense_mel);
        n_outs[i] = n_out;
        if (Ts[i] < 0) err = 1;
    }
    if (err) goto done;
    {
        float gmax = -INFINITY, gmin = INFINITY;
        uint32_t max_sr = 0;
        for (size_t i = 0; i < n_tracks; ++i) {
            float a, b;
            orc_spec_max_min(specs[i], (size_t)Ts[i] * n_outs[i], &a, &b);
            gmax = fmaxf(gmax, a); gmin = fminf(gmin, b);
            if (srs[i] > max_sr) max_sr = srs[i];
        }
        float max_db, min_db;
        orc_clamp_range(gmax, gmin, db_range, &max_db, &min_db);
        if (out_max_db) *out_max_db = max_db;
        if (out_min_db) *out_min_db = min_db;
#pragma omp parallel for schedule(dynamic, 1) if (par_tracks || parallel_render)
        for (long i = 0; i < (long)n_tracks; ++i) {
            float up = orc_up_ratio(max_sr, srs[i], mel_scale);
            uint32_t height = orc_grey_height(n_outs[i], up);
            float *grey = (float *)malloc(sizeof(float) * (size_t)Ts[i] * height);
            orc_spec_to_grey(specs[i], (size_t)Ts[i], n_outs[i], up, max_db, min_db, grey);
            if (out_images && out_images[i]) {
                uint32_t nwidth = orc_calc_nwidth(px_per_sec, n_samples[i], srs[i]);
                if (orc_grey_to_rgb(grey, (uint32_t)Ts[i], height, nwidth, nheight, channels,
                                    out_images[i], parallel_render && !par_tracks))
                    err = 1;
            }
            free(grey);
        }
    }
done:
    for (size_t i = 0; i < n_tracks; ++i) free(specs[i]);
    free(specs); free(Ts); free(n_outs);
    return err ? -1 : 0;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU arm of bench.py asks for every host core explicitly. */
ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
