"""numpy_twin.py -- a SECOND, independent restatement of the render half of the path (test infrastructure).

Written from the reference's `src_rust/display.rs` and from the published algorithm of `image` 0.23
(`imageops::sample`: `resize` = `vertical_sample` then `horizontal_sample`, `lanczos3_kernel`, `sinc`), WITHOUT
consulting oracle/thesia_oracle.c.  It exists because the reference's own tests assert on no pixel
("parity unpinned" for resize / clamp / colour map): with a single restatement a misreading shared by
the oracle and the kernels would pass every test; two independent readings that agree bit for bit on content
where the per-pass clamp bites make that much less likely.  tests/test_numpy_twin.py compares this file with
the C oracle; only tests/ may import it (same rule as the rest of oracle/).

Every operation is float32 and applied in the reference's order (numpy scalars / arrays of dtype float32; no
float64 intermediates), taps are accumulated one at a time in ascending order like the Rust loops.
"""
import numpy as np

F = np.float32

# display.rs:10-21
COLORMAP = np.array([[0, 0, 4], [27, 12, 65], [74, 12, 107], [120, 28, 109], [165, 44, 96], [207, 68, 70], [237, 105, 37],
                     [251, 155, 6], [247, 209, 61], [252, 255, 164]], np.uint8)


def spec_to_grey(spec, up_ratio, mx, mn):
    """display.rs:44-54.  spec [T][n_out] -> grey [height][T] (image rows; row 0 = top = highest frequency)."""
    spec = np.asarray(spec, F)
    T, n_out = spec.shape
    height = int(np.round(F(n_out) * F(up_ratio)))           # (shape[1] as f32 * up_ratio).round() as u32
    grey = np.zeros((height, T), F)
    y = np.arange(height - n_out, height)                     # rows that hold data; the rows above stay 0
    db = spec[:, height - 1 - y].T                            # spec[[x, height - 1 - y]]
    g = (db - F(mn)) / (F(mx) - F(mn))
    grey[y, :] = np.minimum(np.maximum(g, F(0)), F(1))        # .max(0.).min(1.)
    return grey


def _sinc(t):
    """image 0.23 `sinc`: sin(pi t) / (pi t), 1 at 0 -- f32 throughout."""
    t = np.asarray(t, F)
    a = t * F(np.pi)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.sin(a, dtype=F) / a
    return np.where(t == 0, F(1), r).astype(F)


def _lanczos3(x):
    """image 0.23 `lanczos3_kernel` = `lanczos(x, 3.0)`: sinc(x) * sinc(x / 3) inside |x| < 3, else 0."""
    x = np.asarray(x, F)
    return np.where(np.abs(x) < F(3), _sinc(x) * _sinc(x / F(3)), F(0)).astype(F)


def _sample_axis(img, n_out):
    """One pass of image 0.23's `vertical_sample` / `horizontal_sample` along axis 0 of `img` [n_in][other]:
        ratio = n_in / n_out; sratio = max(ratio, 1); src_support = 3 * sratio
        per output o: inputx = (o + 0.5) * ratio; left = clamp(floor(inputx - src_support), 0, n_in - 1);
                      right = clamp(ceil(inputx + src_support), left + 1, n_in); inputx -= 0.5
                      w_i = lanczos3((i - inputx) / sratio), i in [left, right); t = sum_i w_i * p_i; t /= sum_i w_i
                      result = clamp(t, 0, f32::MAX)                       <- the per-pass clamp
    """
    img = np.asarray(img, F)
    n_in = img.shape[0]
    ratio = F(n_in) / F(n_out)
    sratio = ratio if ratio > F(1) else F(1)
    support = F(3) * sratio
    o = np.arange(n_out, dtype=F)
    inputx = (o + F(0.5)) * ratio
    left = np.clip(np.floor(inputx - support).astype(np.int64), 0, n_in - 1)
    right = np.clip(np.ceil(inputx + support).astype(np.int64), left + 1, n_in)
    inputx = inputx - F(0.5)
    taps = int((right - left).max())
    out = np.zeros((n_out,) + img.shape[1:], F)
    wsum = np.zeros(n_out, F)
    for j in range(taps):                                      # ascending taps, one f32 accumulation each
        i = left + j
        on = i < right
        w = np.where(on, _lanczos3((i.astype(F) - inputx) / sratio), F(0)).astype(F)
        src = img[np.minimum(i, n_in - 1)]
        out = (out + w.reshape((-1,) + (1,) * (img.ndim - 1)) * src).astype(F)
        wsum = (wsum + w).astype(F)
    out = out / wsum.reshape((-1,) + (1,) * (img.ndim - 1))
    return np.maximum(out, F(0)).astype(F)                     # clamp(t, 0, max): NaN-free input assumed


def resize_lanczos3(grey, nwidth, nheight):
    """image::imageops::resize(grey, nwidth, nheight, Lanczos3): rows first, then columns, a clamp after each."""
    tmp = _sample_axis(np.asarray(grey, F), nheight)           # vertical_sample: [height][W] -> [nheight][W]
    return _sample_axis(tmp.T.copy(), nwidth).T.copy()         # horizontal_sample: -> [nheight][nwidth]


def convert_grey_to_color(x):
    """display.rs:24-42 on an array: position = 10 x; index = floor; last colour from index >= 9; else
    round(ratio * b + (1 - ratio) * a) per channel, round half away from zero (f32::round)."""
    x = np.asarray(x, F)
    assert (x >= 0).all()
    position = F(10) * x
    fl = np.floor(position)
    index = np.minimum(fl, F(9)).astype(np.int64)
    ratio = (position - fl).astype(F)
    a = COLORMAP[np.minimum(index, 8)].astype(F)
    b = COLORMAP[np.minimum(index + 1, 9)].astype(F)
    r = ratio[..., None]
    v = (r * b + (F(1) - r) * a).astype(F)
    fl_v = np.floor(v)
    rounded = fl_v + ((v - fl_v) >= F(0.5))                   # f32::round: half away from zero (v >= 0; v - floor(v) is exact)
    out = rounded.astype(np.uint8)
    out[index >= 9] = COLORMAP[9]
    return out


def grey_to_rgb(grey, nwidth, nheight):
    """display.rs:56-61."""
    return convert_grey_to_color(resize_lanczos3(grey, nwidth, nheight))
