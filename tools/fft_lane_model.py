"""Lane-level numpy model of the feature kernel's 512-point FFT (16 x 16 x 2 over 32 lanes) -- the index arithmetic
the CUDA code in csrc/feat_kernel.cu mirrors, checked against numpy.fft.  No GPU needed.

  n = 32 a + m  (lane m = 0..31 holds a = 0..15)      stage 1: DFT16 over a, twiddle W512^(m r)  (sign flipped for m % 4 == 3)
  exchange through shared memory: lane (r, c) gathers m = 2 b + c, b = 0..15
  stage 2: DFT16 over b  ->  V[t]  (c = 0: E_r[t];  c = 1: O_r[(t + 8) % 16], the sign flip rotates the outputs by 8)
  radix-2 across the lane pair (r, 0) <-> (r, 1): each sends V[8 + i], i = 0..7
        c = 0:  Z[r + 16 i]       = V[i] + W32^i      * R[i]
        c = 1:  Z[384 + r + 16 i] = R[i] + i * W32^i  * V[i]
  unpack: lane (r, 0) <-> lane ((16 - r) % 16, 1) exchange their upper four values; each forms four bins' power pairs.
"""
import numpy as np

W = lambda N, e: np.exp(-2j * np.pi * (e % N) / N)


def lane_of(r, c):
    return r + 16 * c


def fft_pair_power(x0, x1, win):
    """x0, x1: two real frames of 400 samples; returns dict bin -> (P0, P1) as the lanes would store them."""
    z = np.zeros(512, complex)
    z[:400] = win * (x0 + 1j * x1)
    # stage 1
    Y = np.zeros((32, 16), complex)
    for m in range(32):
        a = np.arange(16)
        for r in range(16):
            Y[m, r] = np.sum(z[32 * a + m] * W(16, a * r))
        for r in range(16):
            tw = W(512, m * r) * (-1 if m % 4 == 3 else 1)
            Y[m, r] *= tw
    # exchange + stage 2
    V = np.zeros((32, 16), complex)
    for r in range(16):
        for c in range(2):
            b = np.arange(16)
            inp = Y[2 * b + c, r]
            for t in range(16):
                V[lane_of(r, c), t] = np.sum(inp * W(16, b * t))
    # radix-2 across lane pairs (xor 16)
    Zl = np.zeros((32, 8), complex)
    for lane in range(32):
        r, c = lane % 16, lane // 16
        partner = lane ^ 16
        for i in range(8):
            R = V[partner, 8 + i]
            if c == 0:
                Zl[lane, i] = V[lane, i] + W(32, i) * R
            else:
                Zl[lane, i] = R + 1j * W(32, i) * V[lane, i]
    # check against the definition
    Zfull = np.fft.fft(z)
    for lane in range(32):
        r, c = lane % 16, lane // 16
        for i in range(8):
            k = r + 16 * i + (384 if c else 0)
            assert abs(Zl[lane, i] - Zfull[k]) < 1e-9 * (1 + abs(Zfull[k])), (lane, i, k)
    # unpack: partner lane and register pairing
    out = {}
    for lane in range(32):
        r, c = lane % 16, lane // 16
        partner = lane_of((16 - r) % 16, 1 - c)
        for q in range(4):
            # sender: Zl[.][4 + q] in general; the two r == 0 lanes send rotated: Zl[.][4 + (q + 1) % 4]
            src_idx = 4 + ((q + 1) % 4 if r == 0 else q)
            rq = Zl[partner, src_idx]
            own_j = 3 - q
            own = Zl[lane, own_j]
            if lane == 0 and own_j == 0:
                own = Zl[lane, 4]                      # bin 64 rides in lane 0's unused slot
            k_own = r + 16 * own_j + (384 if c else 0)
            if lane == 0 and own_j == 0:
                k_own = 64
            if c == 1 and r == 0 and own_j == 0:
                continue                               # bin 128: not needed
            k = k_own if c == 0 else 512 - k_own
            S, D = own + rq, own - rq
            p0 = (S.real ** 2 + D.imag ** 2) / 4
            p1 = (S.imag ** 2 + D.real ** 2) / 4
            assert k not in out, k
            out[k] = (p0, p1)
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)
    x0, x1 = rng.standard_normal(400), rng.standard_normal(400)
    got = fft_pair_power(x0, x1, win)
    f0 = np.zeros(512); f0[:400] = win * x0
    f1 = np.zeros(512); f1[:400] = win * x1
    P0, P1 = np.abs(np.fft.fft(f0)) ** 2, np.abs(np.fft.fft(f1)) ** 2
    ks = sorted(got)
    assert set(range(4, 128)) <= set(ks), sorted(set(range(4, 128)) - set(ks))
    for k in range(1, 128):
        assert abs(got[k][0] - P0[k]) < 1e-8 * P0.max() and abs(got[k][1] - P1[k]) < 1e-8 * P1.max(), k
    print("ok: bins", ks[0], "..", ks[-1], len(ks), "bins")
