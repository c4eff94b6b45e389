"""Numpy prototype of the K1 FFT schedule (index math only) -- design aid, not shipped code.

Emulates, thread by thread, the Stockham autosort passes the CUDA kernel runs
(8 points per thread, radix-8 passes then one radix-4/2 pass), the padded shared-memory
indexing, the real-FFT split and checks against numpy.fft.  Also counts bank conflicts of
every shared-memory access pattern for vector widths V=1,2,4.
"""
import numpy as np, sys

def radices(h):
    r = []
    n = h
    while n % 8 == 0 and n >= 8:
        r.append(8); n //= 8
    if n > 1:
        r.append(n)
    assert np.prod(r) == h, (h, r)
    return r

def pad(i):
    return i + (i >> 3)

def dft_small(v):
    R = len(v)
    k = np.arange(R)
    W = np.exp(-2j * np.pi * np.outer(k, k) / R)
    return W @ v

def bank_conflicts(idxs, V):
    """idxs: padded element indices accessed by 32 lanes; element = V floats (4V bytes).
    returns max wavefronts per phase beyond the ideal"""
    lanes_per_phase = 32 // V
    worst = 1
    for p in range(0, 32, lanes_per_phase):
        ph = idxs[p:p + lanes_per_phase]
        # bank groups of 4V bytes: 128/(4V) groups
        groups = {}
        for a in ph:
            g = a % (32 // V)
            groups.setdefault(g, set()).add(a)
        worst = max(worst, max(len(s) for s in groups.values()))
    return worst

def run(h, V=4, check_banks=True):
    rs = radices(h)
    nthr = h // 8
    rng = np.random.default_rng(h)
    z = rng.standard_normal(h) + 1j * rng.standard_normal(h)
    buf = np.zeros(pad(h) + 8, complex)
    Ns = 1
    report = []
    cur = None
    for pi, R in enumerate(rs):
        nb = 8 // R                       # butterflies per thread
        regs = {}
        for tid in range(nthr):
            for q in range(nb):
                j = tid + q * nthr        # butterfly id in [0, h/R)
                k = j % Ns
                ang = -2 * np.pi * k / (Ns * R)
                v = np.empty(R, complex)
                for r in range(R):
                    src = j + r * (h // R)
                    x = z[src] if pi == 0 else buf[pad(src)]
                    v[r] = x * np.exp(1j * ang * r)
                regs[(tid, q)] = dft_small(v)
        if check_banks and pi > 0 and nthr >= 32:
            for q in range(nb):
                for r in range(R):
                    idxs = [pad((t + q * nthr) + r * (h // R)) for t in range(32)]
                    c = bank_conflicts(idxs, V)
                    if c > 1: report.append(("read", pi, R, q, r, c))
        # (sync) then write
        newbuf = np.zeros_like(buf)
        for tid in range(nthr):
            for q in range(nb):
                j = tid + q * nthr
                k = j % Ns
                d = (j // Ns) * Ns * R + k
                for r in range(R):
                    newbuf[pad(d + r * Ns)] = regs[(tid, q)][r]
        if check_banks and nthr >= 32:
            for q in range(nb):
                for r in range(R):
                    idxs = []
                    for t in range(32):
                        j = t + q * nthr
                        d = (j // Ns) * Ns * R + (j % Ns)
                        idxs.append(pad(d + r * Ns))
                    c = bank_conflicts(idxs, V)
                    if c > 1: report.append(("write", pi, R, q, r, c))
        buf = newbuf
        Ns *= R
    Z = np.array([buf[pad(i)] for i in range(h)])
    err = np.abs(Z - np.fft.fft(z)).max()
    return err, rs, report

def split_check(h):
    """real FFT of length 2h from the packed complex FFT, pair-wise (k, h-k) as the kernel does"""
    rng = np.random.default_rng(1)
    g = rng.standard_normal(2 * h)
    z = g[0::2] + 1j * g[1::2]
    Z = np.fft.fft(z)
    X = np.zeros(h + 1, complex)
    for k in range(0, h // 2 + 1):
        a = Z[k]; b = Z[(h - k) % h]
        s = np.sin(k * np.pi / h); c = np.cos(k * np.pi / h)
        # reference formula for k
        xr = 0.5 * ((a.real + b.real) + c * (a.imag + b.imag) - s * (a.real - b.real))
        xi = 0.5 * ((a.imag - b.imag) - s * (a.imag + b.imag) - c * (a.real - b.real))
        X[k] = xr + 1j * xi
        # for h-k: a'=b, b'=a, s'=s, c'=-c
        xr2 = 0.5 * ((b.real + a.real) - c * (b.imag + a.imag) - s * (b.real - a.real))
        xi2 = 0.5 * ((b.imag - a.imag) - s * (b.imag + a.imag) + c * (b.real - a.real))
        if k != 0:
            X[h - k] = xr2 + 1j * xi2
    X[h] = Z[0].real - Z[0].imag
    return np.abs(X - np.fft.rfft(g)).max()

if __name__ == "__main__":
    for h in [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]:
        for V in ([4] if h > 1024 else [1, 2, 4]):
            err, rs, rep = run(h, V, check_banks=(h <= 4096))
            print(f"h={h} V={V} radices={rs} maxerr={err:.2e} conflicts={rep[:6]}{'...' if len(rep)>6 else ''} n={len(rep)}")
    for h in [4, 8, 256, 1024]:
        print("split", h, split_check(h))
