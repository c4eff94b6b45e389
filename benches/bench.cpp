// benches/bench.cpp -- the reference's four criterion benches (benches/bench.rs:32-95), same names and bodies,
// against the C ABI of include/sgx.h.  The reference loads samples/sample.wav (a file that does not exist in its
// tree); here a WAV path may be given, else 44.03 s of synthetic 48 kHz audio is used.
//
//   "get mel spectrogram"        bench.rs:62-77   1 s of audio, W=1920 hop=480 n_fft=2048, default mel bank, dB
//   "draw spectrogram"           bench.rs:79-95   grey -> RGB, 100 px/s x 500
//   "add track"                  bench.rs:32-45   6 x the same track through MultiTrack::add_tracks
//   "multitrack get spec image"  bench.rs:47-60   get_spec_image(0, 100., 500)
// and, beyond the reference's four, the multi-track batch of BASELINE.json configs[4] on every visible GPU with no
// Python anywhere (sgx_mt_new_sharded: track t on GPU t mod G, the dB-range exchange inside the library):
//   "add track x256 (all GPUs)"  256 ten-minute 48 kHz tracks through ONE add_tracks call (`--c5 [tracks] [seconds]`)
//
// build: g++ -O2 -std=c++17 -I include benches/bench.cpp -L multi-spectrogram-viewer_b200 -lsgx -Wl,-rpath,'$ORIGIN/../multi-spectrogram-viewer_b200' -o benches/bench
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

#include "sgx.h"

static void check(int code, const char *what)
{
    if (code != SGX_OK) { std::fprintf(stderr, "%s failed: [%d] %s\n", what, code, sgx_last_error()); std::exit(1); }
}

static void bench_function(const char *name, const std::function<void()> &body)
{
    using clk = std::chrono::steady_clock;
    for (int i = 0; i < 3; ++i) body(); // warm-up
    std::vector<double> us;
    const auto t_end = clk::now() + std::chrono::milliseconds(1500);
    while (clk::now() < t_end || us.size() < 10) {
        const auto t0 = clk::now();
        body();
        us.push_back(std::chrono::duration<double, std::micro>(clk::now() - t0).count());
        if (us.size() >= 2000) break;
    }
    std::sort(us.begin(), us.end());
    std::printf("%-28s time: [%10.2f us %10.2f us %10.2f us]  (%zu samples)\n", name, us.front(), us[us.size() / 2], us.back(), us.size());
}

// 256 (or argv) synthetic ten-minute tracks through ONE MultiTrack over all visible GPUs: host PCM in, host RGBA out.
static int bench_c5(int n_tracks, int seconds)
{
    const uint32_t sr = 48000;
    const size_t n = (size_t)seconds * sr;
    std::vector<float> base(n);
    unsigned s = 5005u;
    for (size_t i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        const double t = (double)i / sr;
        base[i] = (float)(0.1 * std::sin(6.283185307179586 * 440.0 * t) + 0.05 * std::sin(6.283185307179586 * 3520.0 * t * (1.0 + 0.1 * t / seconds)) +
                          0.01 * ((double)(s >> 8) / 8388608.0 - 1.0));
    }
    sgx_multitrack *mt = nullptr;
    check(sgx_mt_new_sharded(nullptr, nullptr, 0, &mt), "MultiTrack over all GPUs");
    int n_dev = 0;
    check(sgx_mt_get_device_count(mt, &n_dev, nullptr, nullptr), "device count");
    // every track is the same clip at its own gain (the loudest one decides the range of all of them)
    std::vector<std::vector<float>> tracks((size_t)n_tracks);
    std::vector<size_t> ids((size_t)n_tracks), ns((size_t)n_tracks, n);
    std::vector<const float *> pcm((size_t)n_tracks);
    std::vector<uint32_t> srs((size_t)n_tracks, sr), chs((size_t)n_tracks, 1);
    const int distinct = std::min(n_tracks, 8); // host memory: 8 distinct gains, re-used
    for (int t = 0; t < distinct; ++t) {
        tracks[(size_t)t].resize(n);
        const float g = (128.0f + (float)((37 * t) % 128)) / 256.0f;
        for (size_t i = 0; i < n; ++i) tracks[(size_t)t][i] = base[i] * g;
    }
    for (int t = 0; t < n_tracks; ++t) { ids[(size_t)t] = (size_t)t; pcm[(size_t)t] = tracks[(size_t)(t % distinct)].data(); }
    for (int t = 0; t < distinct; ++t) check(sgx_host_pin(tracks[(size_t)t].data(), n * sizeof(float)), "pin PCM");
    int changed = 0;
    using clk = std::chrono::steady_clock;
    check(sgx_mt_add_tracks_pcm(mt, ids.data(), ids.size(), pcm.data(), ns.data(), srs.data(), chs.data(), &changed), "add_tracks (warm-up)");
    std::vector<size_t> need((size_t)n_tracks, 0);
    check(sgx_mt_get_spec_images(mt, ids.data(), ids.size(), 100.0f, 500, 4, nullptr, nullptr, need.data()), "image sizes");
    std::vector<std::vector<uint8_t>> imgs((size_t)n_tracks);
    std::vector<uint8_t *> outs((size_t)n_tracks);
    for (int t = 0; t < n_tracks; ++t) {
        imgs[(size_t)t].resize(need[(size_t)t]); outs[(size_t)t] = imgs[(size_t)t].data();
        check(sgx_host_pin(outs[(size_t)t], need[(size_t)t]), "pin image");
    }
    double best_add = 1e30, best_all = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        const auto t0 = clk::now();
        check(sgx_mt_add_tracks_pcm(mt, ids.data(), ids.size(), pcm.data(), ns.data(), srs.data(), chs.data(), &changed), "add_tracks");
        const auto t1 = clk::now();
        check(sgx_mt_get_spec_images(mt, ids.data(), ids.size(), 100.0f, 500, 4, outs.data(), need.data(), need.data()), "get_spec_images");
        const auto t2 = clk::now();
        best_add = std::min(best_add, std::chrono::duration<double, std::milli>(t1 - t0).count());
        best_all = std::min(best_all, std::chrono::duration<double, std::milli>(t2 - t0).count());
    }
    float mx = 0, mn = 0;
    check(sgx_mt_get_max_db(mt, &mx), "max_db"); check(sgx_mt_get_min_db(mt, &mn), "min_db");
    const double audio_s = (double)n_tracks * seconds;
    std::printf("%-28s %d tracks x %d s on %d GPU(s), pinned host buffers: add_tracks %.1f ms, + get_spec_images %.1f ms  "
                "(%.0f audio-s/s end to end)  range [%.2f, %.2f] dB  kernels launched %llu\n",
                "add track x256 (all GPUs)", n_tracks, seconds, n_dev, best_add, best_all, audio_s / (best_all * 1e-3), mx, mn,
                (unsigned long long)sgx_kernel_launch_count());
    for (int t = 0; t < distinct; ++t) sgx_host_unpin(tracks[(size_t)t].data());
    for (int t = 0; t < n_tracks; ++t) sgx_host_unpin(outs[(size_t)t]);
    sgx_mt_free(mt);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 1 && std::string(argv[1]) == "--c5")
        return bench_c5(argc > 2 ? std::atoi(argv[2]) : 256, argc > 3 ? std::atoi(argv[3]) : 600);
    std::vector<float> wav;
    uint32_t sr = 48000;
    if (argc > 1) { // audio::open_audio_file + sum_axis(Axis(0)), bench.rs:63-64
        size_t n = 0; uint32_t ch = 0;
        check(sgx_open_wav(argv[1], nullptr, 0, &n, &ch, &sr), "open_wav");
        std::vector<float> inter(n * ch);
        check(sgx_open_wav(argv[1], inter.data(), inter.size(), &n, &ch, &sr), "open_wav");
        wav.assign(n, 0.0f);
        for (size_t i = 0; i < n; ++i) for (uint32_t c = 0; c < ch; ++c) wav[i] += inter[i * ch + c];
    } else {
        wav.resize(2113529);
        unsigned s = 12345u;
        for (size_t i = 0; i < wav.size(); ++i) {
            s = s * 1664525u + 1013904223u;
            const double t = (double)i / sr;
            wav[i] = (float)(0.1 * std::sin(6.283185307179586 * 440.0 * t) + 0.05 * std::sin(6.283185307179586 * 3520.0 * t) +
                             0.01 * ((double)(s >> 8) / 8388608.0 - 1.0));
        }
    }
    const size_t one_sec = std::min<size_t>(sr, wav.size());
    // windows::hann(1920, false) / 2048., mel::calc_mel_fb_default(sr, 2048)   bench.rs:66-67
    std::vector<float> window(1920);
    check(sgx_calc_window(1920, 2048, window.data()), "calc_window");
    size_t n_mel = 0;
    check(sgx_calc_mel_fb_default(sr, 2048, nullptr, 0, &n_mel), "calc_mel_fb_default");
    std::vector<float> mel_fb(1025 * n_mel);
    check(sgx_calc_mel_fb_default(sr, 2048, mel_fb.data(), mel_fb.size(), &n_mel), "calc_mel_fb_default");
    size_t T = 0;
    check(sgx_melspectrogram_db(wav.data(), one_sec, 1920, 480, 2048, window.data(), mel_fb.data(), n_mel, nullptr, 0, &T), "melspectrogram size");
    std::vector<float> spec(T * n_mel);

    bench_function("get mel spectrogram", [&] {
        check(sgx_melspectrogram_db(wav.data(), one_sec, 1920, 480, 2048, window.data(), mel_fb.data(), n_mel, spec.data(), spec.size(), &T),
              "melspectrogram_db");
    });

    // display::spec_to_grey(spec, up_ratio = 1, max, min)   bench.rs:86 (with the 4-argument signature of display.rs:44)
    const float mx = *std::max_element(spec.begin(), spec.end()), mn = *std::min_element(spec.begin(), spec.end());
    uint32_t height = 0;
    check(sgx_spec_to_grey(spec.data(), T, n_mel, 1.0f, mx, mn, nullptr, 0, &height), "spec_to_grey size");
    std::vector<float> grey((size_t)T * height);
    check(sgx_spec_to_grey(spec.data(), T, n_mel, 1.0f, mx, mn, grey.data(), grey.size(), &height), "spec_to_grey");
    const uint32_t nwidth = (uint32_t)(100 * (uint32_t)one_sec / sr); // bench.rs:91
    std::vector<uint8_t> rgb((size_t)nwidth * 500 * 3);
    bench_function("draw spectrogram", [&] {
        check(sgx_grey_to_rgb(grey.data(), (uint32_t)T, height, nwidth, 500, 3, rgb.data(), rgb.size()), "grey_to_rgb");
    });

    sgx_multitrack *mt = nullptr;
    check(sgx_mt_new(&mt), "MultiTrack::new");
    const size_t ids[6] = {0, 1, 2, 3, 4, 5};
    const float *pcm[6]; size_t ns[6]; uint32_t srs[6], chs[6];
    for (int i = 0; i < 6; ++i) { pcm[i] = wav.data(); ns[i] = wav.size(); srs[i] = sr; chs[i] = 1; }
    int changed = 0;
    bench_function("add track", [&] { // bench.rs:35-44 (decoded PCM instead of re-reading the file six times)
        check(sgx_mt_add_tracks_pcm(mt, ids, 6, pcm, ns, srs, chs, &changed), "add_tracks");
    });
    size_t need = 0;
    check(sgx_mt_get_spec_image(mt, 0, 100.0f, 500, nullptr, 0, &need), "get_spec_image size");
    std::vector<uint8_t> img(need);
    bench_function("multitrack get spec image", [&] { // bench.rs:55-59
        check(sgx_mt_get_spec_image(mt, 0, 100.0f, 500, img.data(), img.size(), &need), "get_spec_image");
    });
    sgx_mt_free(mt);
    return 0;
}
