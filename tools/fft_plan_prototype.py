#!/usr/bin/env python3
"""numpy prototype of the K2 FFT plan (index math only; not product code).

Real N-point frame -> M=N/2 complex (z[m] = x[2m] + i x[2m+1]) -> F interleaved
sub-FFTs of length L = M/F, each as radix-16 Stockham DIT passes -> final pass =
radix-F combine + real-input untangle + Hermitian mirror.  Mirrors the thread
mapping of csrc/k2_fft.cu so the CUDA index arithmetic can be checked on the CPU.
"""
import numpy as np

def stockham16(sub, L):
    """sub: complex [L]; radix-16 DIT Stockham passes, natural order out."""
    a = sub.copy()
    ns = 1
    nb = L // 16
    while ns < L:
        out = np.empty_like(a)
        for j in range(nb):
            k = j % ns
            v = np.array([a[j + nb * r] for r in range(16)])
            v = v * np.exp(-2j * np.pi * np.arange(16) * k / (ns * 16))
            V = np.fft.fft(v)
            j0 = (j // ns) * ns * 16 + k
            for r in range(16):
                out[j0 + r * ns] = V[r]
        a = out
        ns *= 16
    return a

def fft_real_plan(x, F):
    N = len(x); M = N // 2; L = M // F
    z = x[0::2] + 1j * x[1::2]
    S = [stockham16(z[f::F], L) for f in range(F)]          # S_f = FFT_L(z[F m + f])
    X = np.zeros(N, dtype=complex)
    WN = lambda e: np.exp(-2j * np.pi * e / N)
    for k in range(L // 2 + 1):
        kk = (L - k) % L
        # radix-F combine at k and at L-k
        Zk = [sum(S[f][k] * np.exp(-2j*np.pi*f*(k + L*q)/M) for f in range(F)) for q in range(F)]
        Zm = [sum(S[f][kk] * np.exp(-2j*np.pi*f*((L-k) + L*q)/M) for f in range(F)) for q in range(F)]
        for q in range(F):
            j = k + L * q                      # partner M - j = (L-k) + L (F-1-q)
            A = Zk[q]; B = np.conj(Zm[F - 1 - q])
            Fe = 0.5 * (A + B); Fo = -0.5j * (A - B)
            T = WN(j) * Fo
            X[j] = Fe + T
            X[(M + j) % N] = Fe - T
            X[M - j] = np.conj(Fe - T)
            if j: X[N - j] = np.conj(Fe + T)
    return X

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N, F in [(1024, 2), (8192, 1), (16384, 2), (4096, 8)]:
        x = rng.standard_normal(N)
        e = np.linalg.norm(fft_real_plan(x, F) - np.fft.fft(x)) / np.linalg.norm(np.fft.fft(x))
        print(N, F, e)
