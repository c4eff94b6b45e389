//! `MultiTrack` of src_rust/lib.rs:72-365 on the B200 engine.  Same method names, argument meaning and panics
//! (an unknown id panics, as `HashMap::get(..).unwrap()` does in the reference); I/O errors of `add_tracks`
//! come back as `Err(String)` -- the forwarding lib.rs turns them into `JsValue` exactly like lib.rs:174-189.
//! UNCOMPILED here (no Rust toolchain in the build image).
use sgx_sys::*;
use std::ffi::{CStr, CString};
use std::os::raw::c_int;

fn check(code: c_int) -> Result<(), String> {
    if code == 0 {
        Ok(())
    } else {
        Err(unsafe { CStr::from_ptr(sgx_last_error()) }.to_string_lossy().into_owned())
    }
}

/// Two-call pattern of the byte-returning entry points: size query, then fill.
fn bytes<F: Fn(*mut u8, usize, *mut usize) -> c_int>(f: F) -> Vec<u8> {
    let mut need = 0usize;
    check(f(std::ptr::null_mut(), 0, &mut need)).unwrap();
    let mut v = vec![0u8; need];
    if need > 0 {
        check(f(v.as_mut_ptr(), need, &mut need)).unwrap();
    }
    v
}

pub struct MultiTrack {
    h: *mut SgxMultiTrack,
}

impl MultiTrack {
    /// lib.rs:90-107 (the hard-coded SpecSetting of lib.rs:93-99)
    pub fn new() -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { sgx_mt_new(&mut h) }).expect("no CUDA device");
        MultiTrack { h }
    }

    /// lib.rs:171-191.  `path_list` is '\n'-joined; the returned bool is update_spec_greys' "range changed".
    pub fn add_tracks(&mut self, id_list: &[usize], path_list: &str) -> Result<bool, String> {
        let c = CString::new(path_list).map_err(|e| e.to_string())?;
        let mut changed: c_int = 0;
        check(unsafe { sgx_mt_add_tracks(self.h, id_list.as_ptr(), id_list.len(), c.as_ptr(), &mut changed) })?;
        Ok(changed != 0)
    }

    /// bench.rs:63-67 keeps file I/O out of its timed closure: the same call on decoded, interleaved PCM.
    pub fn add_tracks_pcm(&mut self, id_list: &[usize], pcm: &[&[f32]], sr: &[u32], channels: &[u32]) -> Result<bool, String> {
        assert!(pcm.len() == id_list.len() && sr.len() == id_list.len() && channels.len() == id_list.len());
        let ptrs: Vec<*const f32> = pcm.iter().map(|p| p.as_ptr()).collect();
        let lens: Vec<usize> = pcm.iter().zip(channels).map(|(p, &c)| p.len() / c as usize).collect();
        let mut changed: c_int = 0;
        check(unsafe {
            sgx_mt_add_tracks_pcm(self.h, id_list.as_ptr(), id_list.len(), ptrs.as_ptr(), lens.as_ptr(), sr.as_ptr(), channels.as_ptr(), &mut changed)
        })?;
        Ok(changed != 0)
    }

    /// lib.rs:265-292
    pub fn remove_track(&mut self, id: usize) -> bool {
        let mut changed: c_int = 0;
        check(unsafe { sgx_mt_remove_track(self.h, id, &mut changed) }).unwrap();
        changed != 0
    }

    /// lib.rs:294-298: RGB, 3 bytes per pixel, row 0 = highest frequency
    pub fn get_spec_image(&self, id: usize, px_per_sec: f32, nheight: u32) -> Vec<u8> {
        let h = self.h;
        bytes(|out, cap, need| unsafe { sgx_mt_get_spec_image(h, id, px_per_sec, nheight, out, cap, need) })
    }

    /// lib.rs:300-313: RGBA waveform image
    pub fn get_wav_image(&self, id: usize, px_per_sec: f32, nheight: u32, amp_min: f32, amp_max: f32) -> Vec<u8> {
        let h = self.h;
        bytes(|out, cap, need| unsafe { sgx_mt_get_wav_image(h, id, px_per_sec, nheight, amp_min, amp_max, out, cap, need) })
    }

    /// lib.rs:315-322
    pub fn get_frequency_hz(&self, id: usize, relative_freq: f32) -> f32 {
        let mut x = 0f32;
        check(unsafe { sgx_mt_get_frequency_hz(self.h, id, relative_freq, &mut x) }).unwrap();
        x
    }

    /// lib.rs:324-326
    pub fn get_max_db(&self) -> f32 {
        let mut x = 0f32;
        check(unsafe { sgx_mt_get_max_db(self.h, &mut x) }).unwrap();
        x
    }

    /// lib.rs:328-330
    pub fn get_min_db(&self) -> f32 {
        let mut x = 0f32;
        check(unsafe { sgx_mt_get_min_db(self.h, &mut x) }).unwrap();
        x
    }

    /// lib.rs:332-334
    pub fn get_max_sec(&self) -> f32 {
        let mut x = 0f32;
        check(unsafe { sgx_mt_get_max_sec(self.h, &mut x) }).unwrap();
        x
    }

    /// lib.rs:336-339
    pub fn get_sec(&self, id: usize) -> f32 {
        let mut x = 0f32;
        check(unsafe { sgx_mt_get_sec(self.h, id, &mut x) }).unwrap();
        x
    }

    /// lib.rs:341-343
    pub fn get_sr(&self, id: usize) -> u32 {
        let mut x = 0u32;
        check(unsafe { sgx_mt_get_sr(self.h, id, &mut x) }).unwrap();
        x
    }

    fn text<F: Fn(*mut std::os::raw::c_char, usize, *mut usize) -> c_int>(f: F) -> String {
        let mut need = 0usize;
        check(f(std::ptr::null_mut(), 0, &mut need)).unwrap();
        let mut buf = vec![0u8; need]; // `written` counts the terminating NUL
        check(f(buf.as_mut_ptr() as *mut std::os::raw::c_char, buf.len(), &mut need)).unwrap();
        buf.truncate(need.saturating_sub(1));
        String::from_utf8_lossy(&buf).into_owned()
    }

    /// lib.rs:345-353
    pub fn get_path(&self, id: usize) -> String {
        let h = self.h;
        Self::text(|out, cap, need| unsafe { sgx_mt_get_path(h, id, out, cap, need) })
    }

    /// lib.rs:355-364
    pub fn get_filename(&self, id: usize) -> String {
        let h = self.h;
        Self::text(|out, cap, need| unsafe { sgx_mt_get_filename(h, id, out, cap, need) })
    }
}

impl Drop for MultiTrack {
    fn drop(&mut self) {
        unsafe { sgx_mt_free(self.h) }
    }
}

/// lib.rs:473-480, display.rs:10-21: the ten RGB stops of the colour map
pub fn get_colormap() -> Vec<u8> {
    let mut v = vec![0u8; 30];
    check(unsafe { sgx_get_colormap(v.as_mut_ptr()) }).unwrap();
    v
}

/// bench.rs:7-25 `get_melspectrogram`: perform_stft -> norm -> dot(mel_fb) -> amp_to_db_default in one fused launch.
/// `mel_fb` is row-major [n_fft/2+1][n_mel]; returns (frames, row-major [frames][n_mel]).
pub fn melspectrogram_db(input: &[f32], win_length: usize, hop_length: usize, n_fft: usize, mel_fb: &[f32], n_mel: usize) -> (usize, Vec<f32>) {
    let mut t = 0usize;
    check(unsafe {
        sgx_melspectrogram_db(input.as_ptr(), input.len(), win_length, hop_length, n_fft, std::ptr::null(), mel_fb.as_ptr(), n_mel, std::ptr::null_mut(), 0, &mut t)
    })
    .unwrap();
    let mut out = vec![0f32; t * n_mel];
    check(unsafe {
        sgx_melspectrogram_db(input.as_ptr(), input.len(), win_length, hop_length, n_fft, std::ptr::null(), mel_fb.as_ptr(), n_mel, out.as_mut_ptr(), out.len(), &mut t)
    })
    .unwrap();
    (t, out)
}

/// display.rs:56-61 `grey_to_rgb` (image 0.23 Lanczos3 resize + colour map); `grey` is row-major [height][width].
pub fn grey_to_rgb(grey: &[f32], width: u32, height: u32, nwidth: u32, nheight: u32) -> Vec<u8> {
    let mut out = vec![0u8; nwidth as usize * nheight as usize * 3];
    check(unsafe { sgx_grey_to_rgb(grey.as_ptr(), width, height, nwidth, nheight, 3, out.as_mut_ptr(), out.len()) }).unwrap();
    out
}
