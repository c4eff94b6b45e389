"""Numpy prototype of the warp-per-frame schedule (K1W): 1024-point complex FFT as 32 x 32 with one
transpose, followed by the real-FFT split done with lane <-> 32-lane partner exchanges."""
import numpy as np

H = 1024
rng = np.random.default_rng(0)
g = rng.standard_normal(2 * H)
z = g[0::2] + 1j * g[1::2]
# pass 1: lane = m2, register = m1 ; DFT over m1
regs = np.zeros((32, 32), complex)  # [lane][reg]
for lane in range(32):
    col = np.array([z[32 * m1 + lane] for m1 in range(32)])
    A = np.fft.fft(col)                       # A[k1]
    A = A * np.exp(-2j * np.pi * lane * np.arange(32) / H)   # twiddle W_1024^(lane*k1)
    regs[lane] = A
# transpose through buf[k1*33 + lane]
buf = np.zeros(32 * 33, complex)
for lane in range(32):
    for k1 in range(32):
        buf[k1 * 33 + lane] = regs[lane][k1]
regs2 = np.zeros((32, 32), complex)
for lane in range(32):           # lane = k1
    row = np.array([buf[lane * 33 + m2] for m2 in range(32)])
    regs2[lane] = np.fft.fft(row)             # Z[lane + 32*k2], reg = k2
Z = np.zeros(H, complex)
for lane in range(32):
    for k2 in range(32):
        Z[lane + 32 * k2] = regs2[lane][k2]
print("fft err", np.abs(Z - np.fft.fft(z)).max())
# split with shuffles
X = np.zeros(H + 1, complex)
src = np.zeros((32, 32), complex)
for lane in range(32):
    for j in range(32):
        src[lane][j] = regs2[lane][(j + 1) & 31] if lane == 0 else regs2[lane][j]
for lane in range(32):
    pl = (32 - lane) & 31
    for j in range(32):
        k = lane + 32 * j
        a = regs2[lane][j]
        b = src[pl][31 - j]
        c, s = np.cos(k * np.pi / H), np.sin(k * np.pi / H)
        xr = 0.5 * ((a.real + b.real) + c * (a.imag + b.imag) - s * (a.real - b.real))
        xi = 0.5 * ((a.imag - b.imag) - s * (a.imag + b.imag) - c * (a.real - b.real))
        X[k] = xr + 1j * xi
X[H] = regs2[0][0].real - regs2[0][0].imag
print("rfft err", np.abs(X - np.fft.rfft(g)).max())
# bank conflicts of the transpose read: lane stride 33 float2 -> 64-bit accesses, half-warp phases
for half in (0, 16):
    banks = [((lane * 33 + 5) * 2) % 32 for lane in range(half, half + 16)]
    assert len(set(banks)) == 16
print("transpose reads conflict free")
