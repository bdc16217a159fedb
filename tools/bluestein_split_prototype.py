"""Prototype (numpy) of the split chirp-z transform planned for the polar-cap rings whose Bluestein work length is
16384 (nph = 4100 .. 8188 at nside 2048): the length-n chirp convolution is cut into input halves and output halves so
that every piece is a circular convolution of length 8192, which fits in the shared memory of one CTA.

    x_j = c_j sum_{k<n} (X_k c_k) conj(c_{j-k}),   c_l = exp(i pi l^2 / n)            (DFT with the + sign)

With H = ceil(n/2), inputs I0 = [0,H), I1 = [H,n) and outputs J0, J1 alike, and the three filter tables
    A_l = conj(c_l), B_l = conj(c_{l+H}), C_l = conj(c_{l-H}),  |l| < H,
    out[J0] = IFFT(U0 .* FFT(A) + U1 .* FFT(C)),   out[J1] = IFFT(U0 .* FFT(B) + U1 .* FFT(A)),   U_b = FFT(u[I_b] zero-padded),
all transforms of length M' = 8192 >= 2H - 1.  On the GPU: a 2-CTA cluster per (ring pair, component); CTA b folds and
transforms input half b, the spectra are combined through distributed shared memory, CTA a inverse-transforms and
scatters output half a -- the same work per CTA as one M = 8192 ring of the fused kernel.

Run: python tools/bluestein_split_prototype.py   (checks the identity against a direct DFT for several ring lengths)"""
import numpy as np


def chirp(l, n):
    l = np.asarray(l, dtype=np.int64)
    return np.exp(1j * np.pi * ((l * l) % (2 * n)) / n)


def split_chirpz(X, Mp=8192):
    n = X.size
    H = (n + 1) // 2
    assert 2 * H - 1 <= Mp
    u = X * chirp(np.arange(n), n)

    def table(shift):                      # g_l = conj(c_{l+shift}) for |l| < H, wrapped onto [0, Mp)
        g = np.zeros(Mp, dtype=complex)
        l = np.arange(-(H - 1), H)
        g[l % Mp] = np.conj(chirp(l + shift, n))
        return np.fft.fft(g)
    A, B, C = table(0), table(H), table(-H)
    U0 = np.fft.fft(np.concatenate([u[:H], np.zeros(Mp - H)]))
    U1 = np.fft.fft(np.concatenate([u[H:], np.zeros(Mp - (n - H))]))
    o0 = np.fft.ifft(U0 * A + U1 * C)[:H]
    o1 = np.fft.ifft(U0 * B + U1 * A)[:n - H]
    return np.concatenate([o0, o1]) * chirp(np.arange(n), n)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (4100, 5000, 6148, 8187, 8188, 12, 13):
        X = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        ref = np.fft.ifft(X) * n           # sum_k X_k exp(+2 pi i j k / n)
        got = split_chirpz(X, 8192 if n > 64 else 16)
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print(f"n = {n:5d}: relative L2 error {err:.2e}")
        assert err < 1e-11
