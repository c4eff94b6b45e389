/*
 * sgx.h -- C ABI of the B200-native spectrogram engine (libsgx.so).
 *
 * This is the drop-in boundary for the reference's per-file analysis + render path
 * (Sytronik/multi-spectrogram-viewer, crate `thesia`).  Every entry point names the reference
 * item it replaces (paths relative to the reference root).  Two surfaces are mirrored:
 *
 *   surface 1  the `#[wasm_bindgen] impl MultiTrack` methods     src_rust/lib.rs:87-365, 473-480
 *   surface 2  the rlib functions `benches/bench.rs` calls        benches/bench.rs:5-30
 *
 * Conventions
 *   - plain C types only: pointers, sizes, floats.  No torch / C++ types cross this line.
 *   - every function returns an `int` status (SGX_OK == 0); the message of the last failure on
 *     the calling thread is available from sgx_last_error().  Nothing unwinds across the ABI.
 *     Where the reference panics (unknown id, window length mismatch, ...) a status is returned.
 *   - output buffers are caller-allocated: pass (out, cap) and receive `*written`; calling with
 *     out == NULL only reports the required size.
 *   - "host" pointers are ordinary CPU memory; "_device" variants take CUDA device pointers of
 *     the handle's device and enqueue work on the handle's stream without synchronising.
 *   - a handle is not thread-safe (the reference's mutators take `&mut self`).  sgx_mt_new / _ex bind it to
 *     one CUDA device and one stream; sgx_mt_new_sharded spreads the tracks of ONE handle over several
 *     devices of this process; sgx_mt_attach_nccl joins the handles of several processes (one per GPU).
 *   - there is no CPU fallback: if no CUDA device is usable the constructors fail with
 *     SGX_ERR_CUDA.
 */
#ifndef SGX_H_
#define SGX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SGX_API __attribute__((visibility("default")))
#else
#define SGX_API
#endif

/* ---------------------------------------------------------------------------------------------
 * status codes
 * ------------------------------------------------------------------------------------------- */
enum {
    SGX_OK = 0,
    SGX_ERR_IO = 1,         /* file could not be opened / decoded  (lib.rs:174-177 Err(JsValue)) */
    SGX_ERR_UNKNOWN_ID = 2, /* the reference unwrap()-panics       (lib.rs:113,266,295,316,337)  */
    SGX_ERR_BAD_ARG = 3,    /* the reference assert!-panics        (lib.rs:404, mel.rs:50-51)    */
    SGX_ERR_CUDA = 4,       /* CUDA runtime / no device / kernel image missing                    */
    SGX_ERR_STATE = 5,      /* call order (e.g. image requested before any track exists)          */
    SGX_ERR_NOMEM = 6,
    SGX_ERR_BUFFER = 7,     /* caller buffer too small; *written holds the required size          */
    SGX_ERR_NCCL = 8        /* libnccl.so.2 missing or a collective failed                         */
};

/* Message of the last error raised on this thread ("" if none).  Never NULL. */
SGX_API const char *sgx_last_error(void);

/* Library / device facts: returns SGX_OK and fills what is non-NULL. */
SGX_API int sgx_device_info(int device, int *sm_count, int *cc_major, int *cc_minor,
                            size_t *total_mem);

/* Page-locks (pins) an existing host allocation so that the uploads of sgx_mt_add_tracks_pcm* and the downloads of
 * sgx_mt_get_spec_images* are true DMA transfers that overlap kernels and each other (pageable memory is staged
 * through a driver buffer, several times slower).  A host written in Rust / C needs no CUDA binding for this. */
SGX_API int sgx_host_pin(void *ptr, size_t bytes);
SGX_API int sgx_host_unpin(void *ptr);

/* Number of kernels this library has launched in the calling process (all handles). */
SGX_API uint64_t sgx_kernel_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * settings  ==  struct SpecSetting                                         src_rust/lib.rs:64-70
 * The reference hard-codes them in MultiTrack::new (lib.rs:93-99) and has no setter; the four
 * override fields (0 = derive exactly like AudioTrack::new, lib.rs:43-46) exist because
 * BASELINE.json's configs name explicit n_fft / hop / n_mel values that bench.rs reaches by
 * calling the stage functions directly.
 * ------------------------------------------------------------------------------------------- */
enum { SGX_FREQ_LINEAR = 0, SGX_FREQ_MEL = 1 }; /* enum FreqScale, lib.rs:25-28 */

typedef struct sgx_settings {
    float win_ms;      /* 40.0 */
    size_t t_overlap;  /* 4    */
    size_t f_overlap;  /* 1    */
    int freq_scale;    /* SGX_FREQ_MEL */
    float db_range;    /* 120.0 */
    size_t win_length; /* override; 0 = hop*t_overlap from win_ms        (lib.rs:43-45) */
    size_t hop_length; /* override; 0 = round(win_ms*sr/1000/t_overlap)  (lib.rs:44)    */
    size_t n_fft;      /* override; 0 = calc_proper_n_fft(win)*f_overlap (lib.rs:46)    */
    size_t n_mel;      /* override; 0 = mel::calc_mel_fb_default          (mel.rs:87-99) */
} sgx_settings;

SGX_API void sgx_settings_default(sgx_settings *s); /* the values of lib.rs:93-99 */

/* ---------------------------------------------------------------------------------------------
 * surface 1:  MultiTrack                                                   src_rust/lib.rs:72-365
 * ------------------------------------------------------------------------------------------- */
typedef struct sgx_multitrack sgx_multitrack;

/* MultiTrack::new()  lib.rs:90-110.  Device 0, a private stream, default settings. */
SGX_API int sgx_mt_new(sgx_multitrack **out);
/* Same with explicit settings (NULL = default), CUDA device ordinal and an optional existing
 * CUDA stream (cudaStream_t as void*; NULL = create a private non-blocking stream). */
SGX_API int sgx_mt_new_ex(const sgx_settings *settings, int device, void *cuda_stream,
                          sgx_multitrack **out);
SGX_API void sgx_mt_free(sgx_multitrack *mt);

/* --- multi-GPU inside the library -------------------------------------------------------------
 * The reference parallelises add_tracks over tracks (rayon, lib.rs:161-166) and reduces the dB range over all
 * of them (lib.rs:194-209).  Here track id t lives on shard t mod G; PCM, dB and pixels never leave their GPU;
 * the one exchange of the path -- {max, -min, max_sr, max_sec}, 16 bytes, all-reduce(MAX) -- is an
 * ncclAllReduce enqueued by the library on the stream that carries the analysis and the render, between the
 * two: no host round trip.  NCCL is loaded at run time (dlopen of libnccl.so.2): libsgx.so links without it
 * and only these entry points can return SGX_ERR_NCCL.
 *
 * One process, several GPUs: a handle with one engine per listed device (devices == NULL or n_devices == 0: every
 * visible device).  Every call of surface 1 works on it unchanged: ids route to their device, device-pointer
 * arguments of track t must live on device devices[t mod G], time slices (n3) need single-device handles. */
SGX_API int sgx_mt_new_sharded(const sgx_settings *settings, const int *devices, size_t n_devices,
                               sgx_multitrack **out);
/* One process per GPU (torchrun / MPI style): rank 0 draws an id, ships the 128 bytes to every rank by any host
 * channel, and every rank attaches its own single-device handle (a collective call).  From then on the handle is
 * one shard of a global MultiTrack: sgx_mt_add_tracks* take the WHOLE id list on every rank and keep the tracks
 * with id mod world == rank (entries of other ranks may carry NULL pointers; a rank only opens its own files),
 * sgx_mt_remove_track must be called on every rank, and get_max_db / get_min_db / get_max_sec / image geometry
 * reflect all ranks.  Per-track calls answer for owned ids only (SGX_ERR_UNKNOWN_ID otherwise). */
SGX_API int sgx_nccl_unique_id(uint8_t out[128]);
SGX_API int sgx_mt_attach_nccl(sgx_multitrack *mt, const uint8_t unique_id[128], int rank, int world);
/* engines inside the handle; rank / world of an attached single-device handle (0 / 1 otherwise) */
SGX_API int sgx_mt_get_device_count(sgx_multitrack *mt, int *n_devices, int *rank, int *world);

/* MultiTrack::add_tracks(&mut self, id_list: &[usize], path_list: &str) -> Result<bool, JsValue>
 * lib.rs:171-191.  path_list is '\n'-joined (lib.rs:173).  *changed receives the returned bool
 * (update_spec_greys, lib.rs:193-263).  WAV files only (audio.rs:9-21); unlike the reference the
 * call is atomic: on SGX_ERR_IO no track of the batch has been inserted. */
SGX_API int sgx_mt_add_tracks(sgx_multitrack *mt, const size_t *id_list, size_t n_ids,
                              const char *path_list, int *changed);

/* The same call with decoded PCM already in memory (what bench.rs:63-67 does to keep file I/O out
 * of the timed region).  pcm[i] is interleaved f32 [n_samples[i]][channels[i]] exactly as
 * audio::open_audio_file lays it out (audio.rs:33-35); channels are SUMMED (lib.rs:42). */
SGX_API int sgx_mt_add_tracks_pcm(sgx_multitrack *mt, const size_t *id_list, size_t n_ids,
                                  const float *const *pcm, const size_t *n_samples,
                                  const uint32_t *sr, const uint32_t *channels, int *changed);
/* 16-bit integer PCM as stored in the WAV fixtures: the /32768 of audio.rs:16-19 is applied on the
 * GPU while loading (exact), halving the host->device bytes. */
SGX_API int sgx_mt_add_tracks_pcm_i16(sgx_multitrack *mt, const size_t *id_list, size_t n_ids,
                                      const int16_t *const *pcm, const size_t *n_samples,
                                      const uint32_t *sr, const uint32_t *channels, int *changed);
/* PCM already resident in device memory (borrowed for the duration of the call, not copied; must
 * stay valid until the track is removed because get_wav_image reads it).  Work is enqueued on the
 * handle's stream; if `changed` is NULL the call is DEFERRED: it does not synchronise, only the
 * local extrema are reduced on the device and nothing is committed -- the driver all-reduces
 * them in-stream (sgx_mt_range_device_ptr) and then calls sgx_mt_commit_range_device. */
SGX_API int sgx_mt_add_tracks_pcm_device(sgx_multitrack *mt, const size_t *id_list, size_t n_ids,
                                         const float *const *d_pcm, const size_t *n_samples,
                                         const uint32_t *sr, const uint32_t *channels,
                                         int *changed);

/* MultiTrack::remove_track(&mut self, id) -> bool   lib.rs:265-292.  changed == NULL: deferred like
 * sgx_mt_add_tracks_pcm_device (local extrema recomputed, nothing committed or synchronised). */
SGX_API int sgx_mt_remove_track(sgx_multitrack *mt, size_t id, int *changed);

/* MultiTrack::get_spec_image(&self, id, px_per_sec: f32, nheight: u32) -> Vec<u8>  lib.rs:294-298
 * RGB, 3 bytes per pixel, nwidth = (px_per_sec * len / sr) as u32, row 0 = highest frequency
 * (display.rs:44-61).  *written = nwidth * nheight * 3. */
SGX_API int sgx_mt_get_spec_image(sgx_multitrack *mt, size_t id, float px_per_sec,
                                  uint32_t nheight, uint8_t *out, size_t cap, size_t *written);
/* Same pixels as RGBA (A = 255): the layout BASELINE.json's metric is quoted on. */
SGX_API int sgx_mt_get_spec_image_rgba(sgx_multitrack *mt, size_t id, float px_per_sec,
                                       uint32_t nheight, uint8_t *out, size_t cap,
                                       size_t *written);
/* Device-resident output, asynchronous on the handle's stream.  channels = 3 or 4. */
SGX_API int sgx_mt_get_spec_image_device(sgx_multitrack *mt, size_t id, float px_per_sec,
                                         uint32_t nheight, int channels, uint8_t *d_out,
                                         size_t cap, size_t *written);
/* Batched form of the above: one launch sequence for n_ids tracks (d_out[i] has cap[i] bytes). */
SGX_API int sgx_mt_get_spec_images_device(sgx_multitrack *mt, const size_t *id_list, size_t n_ids,
                                          float px_per_sec, uint32_t nheight, int channels,
                                          uint8_t *const *d_out, const size_t *cap,
                                          size_t *written);

/* Host-buffer form of the batched call: what a viewer does after add_tracks -- get_spec_image for every track
 * (lib.rs:294-298; the bench loop of benches/bench.rs:47-60 over a list of ids).  All renders are enqueued at
 * once into per-image device staging and the device->host copies run on a second stream underneath them; one
 * synchronisation at the end.  out[i] == NULL skips an image, out == NULL only reports sizes.  Pinned (page-locked)
 * host buffers make the copies truly asynchronous. */
SGX_API int sgx_mt_get_spec_images(sgx_multitrack *mt, const size_t *id_list, size_t n_ids,
                                   float px_per_sec, uint32_t nheight, int channels,
                                   uint8_t *const *out, const size_t *cap, size_t *written);
/* The same without the final wait: returns once everything is enqueued.  The compute stream is free as soon as
 * the last render has run, so the NEXT sgx_mt_add_tracks* (uploads on a third stream, then analysis) overlaps the
 * downloads still on the wire -- the two directions of the link are independent.  The host buffers are complete
 * after sgx_mt_wait_images; a new request waits for the previous one first. */
SGX_API int sgx_mt_get_spec_images_async(sgx_multitrack *mt, const size_t *id_list, size_t n_ids,
                                         float px_per_sec, uint32_t nheight, int channels,
                                         uint8_t *const *out, const size_t *cap, size_t *written);
SGX_API int sgx_mt_wait_images(sgx_multitrack *mt);

/* MultiTrack::get_wav_image(&self, id, px_per_sec, nheight, amp_min, amp_max) -> Vec<u8> (RGBA)
 * lib.rs:300-313, display.rs:63-115. */
SGX_API int sgx_mt_get_wav_image(sgx_multitrack *mt, size_t id, float px_per_sec, uint32_t nheight,
                                 float amp_min, float amp_max, uint8_t *out, size_t cap,
                                 size_t *written);

/* getters, lib.rs:315-364 */
SGX_API int sgx_mt_get_frequency_hz(sgx_multitrack *mt, size_t id, float relative_freq,
                                    float *out);                         /* lib.rs:315-322 */
SGX_API int sgx_mt_get_max_db(sgx_multitrack *mt, float *out);           /* lib.rs:324-326 */
SGX_API int sgx_mt_get_min_db(sgx_multitrack *mt, float *out);           /* lib.rs:328-330 */
SGX_API int sgx_mt_get_max_sec(sgx_multitrack *mt, float *out);          /* lib.rs:332-334 */
SGX_API int sgx_mt_get_sec(sgx_multitrack *mt, size_t id, float *out);   /* lib.rs:336-339 */
SGX_API int sgx_mt_get_sr(sgx_multitrack *mt, size_t id, uint32_t *out); /* lib.rs:341-343 */
SGX_API int sgx_mt_get_path(sgx_multitrack *mt, size_t id, char *out, size_t cap,
                            size_t *written);                            /* lib.rs:345-353 */
SGX_API int sgx_mt_get_filename(sgx_multitrack *mt, size_t id, char *out, size_t cap,
                                size_t *written);                        /* lib.rs:355-364 */

/* get_colormap() -> Vec<u8> (30 bytes)   lib.rs:473-480, display.rs:10-21 */
SGX_API int sgx_get_colormap(uint8_t out[30]);

/* --- engine-side extensions of surface 1 (no reference counterpart) -------------------------- */

/* Shape of the cached dB spectrogram of a track: frames T and rows n_out (mel bands or bins). */
SGX_API int sgx_mt_get_spec_shape(sgx_multitrack *mt, size_t id, size_t *n_frames, size_t *n_out);
/* Copies the cached dB spectrogram [T][n_out] f32 (== MultiTrack.specs[id], lib.rs:78) to host. */
SGX_API int sgx_mt_get_spec_db(sgx_multitrack *mt, size_t id, float *out, size_t cap_elems,
                               size_t *written_elems);
/* Image width get_spec_image will produce (lib.rs:296). */
SGX_API int sgx_mt_get_image_width(sgx_multitrack *mt, size_t id, float px_per_sec,
                                   uint32_t *nwidth);

/* Multi-GPU hook for drivers that bring their own collective (sgx_mt_new_sharded / sgx_mt_attach_nccl do this
 * inside the library): device pointer to four floats {max, -min, max_sr, max_sec} holding this handle's LOCAL
 * un-clamped dB extrema (lib.rs:194-207) and metadata maxima.  A driver runs all_reduce(MAX) on them on the
 * handle's stream, then calls sgx_mt_commit_range_device, which applies lib.rs:208-209 on the device.
 * Neither call synchronises.  Single-device handles without an attached communicator only. */
SGX_API int sgx_mt_range_device_ptr(sgx_multitrack *mt, float **d_max_negmin);
SGX_API int sgx_mt_commit_range_device(sgx_multitrack *mt);
/* Time-sharding ONE long track over several GPUs.  Every GPU owns a strip of output columns:
 * sgx_slice_plan tells which frames that strip needs (its Lanczos taps) and which samples those frames read
 * (with the reflection at the true ends of the track); the driver uploads just those samples, registers them
 * with sgx_mt_add_track_slice_device (deferred like sgx_mt_add_tracks_pcm_device: all-reduce the range, then
 * sgx_mt_commit_range_device) and renders the strip with sgx_mt_get_spec_image_slice_device into a
 * [nheight][ox_count][channels] buffer.  Strips of all GPUs side by side are bit-identical to the image of
 * the whole track. */
SGX_API int sgx_slice_plan(size_t n_total, uint32_t sr, const sgx_settings *settings, float px_per_sec,
                           uint32_t ox_begin, uint32_t ox_count, size_t *frame_begin, size_t *frame_count,
                           size_t *sample_begin, size_t *sample_count);
SGX_API int sgx_mt_add_track_slice_device(sgx_multitrack *mt, size_t id, const float *d_pcm,
                                          size_t chunk_offset, size_t chunk_len, size_t n_total, uint32_t sr,
                                          uint32_t channels, size_t frame_begin, size_t frame_count);
SGX_API int sgx_mt_get_spec_image_slice_device(sgx_multitrack *mt, size_t id, float px_per_sec,
                                               uint32_t nheight, int channels, uint32_t ox_begin,
                                               uint32_t ox_count, uint8_t *d_out, size_t cap, size_t *written);
/* max sample rate across ALL shards (lib.rs:220-224 is metadata-only; the driver max-reduces it
 * on the host before the first render).  0 = use the local maximum. */
SGX_API int sgx_mt_set_global_max_sr(sgx_multitrack *mt, uint32_t max_sr);
/* Stage timing for bench.py's roofline: when enabled, CUDA events bracket the K1 (analysis) launches
 * of the most recent add_tracks call and the K3 (render) launches of the most recent image call on
 * the handle's stream.  get_stage_times synchronises; -1 = nothing recorded. */
SGX_API int sgx_mt_set_profiling(sgx_multitrack *mt, int on);
SGX_API int sgx_mt_get_stage_times(sgx_multitrack *mt, float *analysis_ms, float *render_ms);
/* Blocks until everything enqueued on the handle's stream has finished; refreshes the host
 * copies of max_db/min_db and reports `changed` (lib.rs:210-229) if non-NULL. */
SGX_API int sgx_mt_synchronize(sgx_multitrack *mt, int *changed);

/* ---------------------------------------------------------------------------------------------
 * surface 2:  stage functions used by benches/bench.rs  (host buffers in, host buffers out;
 * each runs on device 0, default stream of the library, and synchronises before returning)
 * ------------------------------------------------------------------------------------------- */

/* utils::calc_proper_n_fft   utils.rs:17-19 */
SGX_API size_t sgx_calc_proper_n_fft(size_t win_length);
/* AudioTrack::new parameter derivation   lib.rs:43-46 */
SGX_API int sgx_track_params(uint32_t sr, const sgx_settings *s, size_t *win_length,
                             size_t *hop_length, size_t *n_fft);
/* windows::hann(size, symmetric) -> Array1<f32>   windows.rs:21-30 (host table) */
SGX_API int sgx_hann(size_t size, int symmetric, float *out);
/* MultiTrack::calc_window = hann(win,false)/n_fft   lib.rs:138-140 (host table) */
SGX_API int sgx_calc_window(size_t win_length, size_t n_fft, float *out);
/* mel::hz_to_mel / mel_to_hz   mel.rs:14-31 */
SGX_API float sgx_hz_to_mel(float hz);
SGX_API float sgx_mel_to_hz(float mel);
/* mel::calc_mel_fb::<f32>(sr, n_fft, n_mel, fmin, fmax, do_norm) -> Array2 [n_fft/2+1][n_mel]
 * mel.rs:33-85 (host table).  fmax < 0 means None. */
SGX_API int sgx_calc_mel_fb(uint32_t sr, size_t n_fft, size_t n_mel, float fmin, float fmax,
                            int do_norm, float *out);
/* mel::calc_mel_fb_default(sr, n_fft)   mel.rs:87-99.  *n_mel is always set; the bank is written
 * when out != NULL and cap_elems >= (n_fft/2+1) * n_mel. */
SGX_API int sgx_calc_mel_fb_default(uint32_t sr, size_t n_fft, float *out, size_t cap_elems,
                                    size_t *n_mel);
/* Number of frames perform_stft produces (front + middle + back lists, lib.rs:412-435);
 * negative when the reference would panic on its slicing. */
SGX_API long sgx_stft_num_frames(size_t n, size_t win_length, size_t hop_length);

/* perform_stft(input, win_length, hop_length, n_fft, window, fft_module, parallel)
 *   -> Array2<Complex<f32>> [T][n_fft/2+1]                                   lib.rs:388-471
 * window == NULL -> hann(win,false)/n_fft (lib.rs:403-408).  out holds interleaved (re,im);
 * cap_elems counts floats.  The fft_module / parallel arguments have no meaning on the GPU. */
SGX_API int sgx_perform_stft(const float *input, size_t n, size_t win_length, size_t hop_length,
                             size_t n_fft, const float *window, float *out, size_t cap_elems,
                             size_t *n_frames);
/* stft.mapv(|x| x.norm())   lib.rs:124 / bench.rs:21 : magnitudes [T][n_fft/2+1] */
SGX_API int sgx_stft_magnitude(const float *input, size_t n, size_t win_length, size_t hop_length,
                               size_t n_fft, const float *window, float *out, size_t cap_elems,
                               size_t *n_frames);
/* DeciBelInplace::amp_to_db_default on a host array   decibel.rs:79-88.  SGX_ERR_BAD_ARG if any
 * element is negative or NaN (the assert of decibel.rs:34). */
SGX_API int sgx_amp_to_db_default(float *x, size_t n);
/* bench.rs:7-25 get_melspectrogram: perform_stft -> norm -> .dot(mel_fb) -> amp_to_db_default.
 * mel_fb is [n_fft/2+1][n_mel] C order; mel_fb == NULL selects the linear-frequency dB
 * spectrogram (lib.rs:126-129) with n_out = n_fft/2+1.  out is [T][n_out]. */
SGX_API int sgx_melspectrogram_db(const float *input, size_t n, size_t win_length,
                                  size_t hop_length, size_t n_fft, const float *window,
                                  const float *mel_fb, size_t n_mel, float *out, size_t cap_elems,
                                  size_t *n_frames);
/* display::spec_to_grey(spec [T][n_out], up_ratio, max, min) -> GreyF32Image (width T, height
 * round(n_out*up_ratio), row-major)   display.rs:44-54 */
SGX_API int sgx_spec_to_grey(const float *spec, size_t n_frames, size_t n_out, float up_ratio,
                             float max_db, float min_db, float *grey, size_t cap_elems,
                             uint32_t *height);
/* display::grey_to_rgb(&grey, nwidth, nheight) -> RgbImage   display.rs:56-61
 * (image::imageops::resize Lanczos3 + convert_grey_to_color).  channels = 3 (reference) or 4. */
SGX_API int sgx_grey_to_rgb(const float *grey, uint32_t width, uint32_t height, uint32_t nwidth,
                            uint32_t nheight, int channels, uint8_t *out, size_t cap);
/* display::wav_to_image(wav, nwidth, nheight, (amp_min, amp_max)) -> RgbaImage  display.rs:63-115 */
SGX_API int sgx_wav_to_image(const float *wav, size_t n, uint32_t nwidth, uint32_t nheight,
                             float amp_min, float amp_max, uint8_t *out, size_t cap);
/* audio::open_audio_file(path) -> (Array2<f32> [ch][n], sr)   audio.rs:9-37, WAV branch only.
 * Host-side decode; out is interleaved [n][ch] (the memory layout behind audio.rs:33-35). */
SGX_API int sgx_open_wav(const char *path, float *out, size_t cap_elems, size_t *n_samples,
                         uint32_t *channels, uint32_t *sr);

#ifdef __cplusplus
}
#endif
#endif /* SGX_H_ */
