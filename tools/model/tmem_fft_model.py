"""Executable model of the 512-point FFT of k_warp_synth with its three lane<->register exchanges routed through tensor
memory (tcgen05.st .32x32b -> tcgen05.ld .16x256b). The thread <-> (lane, column) maps are the ones measured on B200 by
tools/probes/tmem_probe.cu. The model carries real data through the same register/lane movements as the kernel and checks
the result against numpy's FFT; it also emits the per-lane twiddle tables and the storage permutation of the D array.
Used to derive (and regression-check) the index maps hard-coded in kernel_warp.cu.  python tools/model/tmem_fft_model.py
"""
import numpy as np

Q = 512
LANES = 32


def st_32x32b(regs):
    """regs[lane][col] -> tmem[lane][col] (thread t <-> TMEM lane t, register i <-> column i)"""
    return regs.copy()


def ld_16x256b(tmem, ncols=32):
    """tcgen05.ld.16x256b.x(ncols/8) twice (lane base 0 and 16). Thread t, instruction I, register 4k+2h+b holds
    tmem[16 I + t/4 + 8 h][8 k + 2 (t%4) + b]; returns out[t][16 I + 4k+2h+b]."""
    out = np.zeros((LANES, ncols), tmem.dtype)
    for t in range(LANES):
        for I in range(2):
            for k in range(ncols // 8):
                for h in range(2):
                    for b in range(2):
                        out[t, 16 * I + 4 * k + 2 * h + b] = tmem[16 * I + t // 4 + 8 * h, 8 * k + 2 * (t % 4) + b]
    return out


def bits(v, n):
    return [(v >> i) & 1 for i in range(n)]


def run(x):
    """x: complex[512] in natural order (the buffer T the spectral stage leaves in shared memory). Returns (X[k] per lane /
    register, the k of every (lane, reg)), registers being complex (z is implicit: a complex register = two columns)."""
    # ---- pass 1: lane l owns points j = 64 m + 2 l + j0, m < 8 (radix-8 digit), j0 in {0,1}; complex register index = 2 m + j0
    #      (as float columns: 4 m + 2 j0 + z). Two radix-8 butterflies per lane, loaded as eight float4.
    lab = np.zeros((LANES, 16), int)          # index label carried along: after pass 1 the "m" digit means k0
    val = np.zeros((LANES, 16), complex)
    for l in range(LANES):
        for m in range(8):
            for j0 in range(2):
                val[l, 2 * m + j0] = x[64 * m + 2 * l + j0]
    W = lambda N, e: np.exp(-2j * np.pi * e / N)
    # radix-8 DIF over m, twiddle W_512^{(2l + j0) k0}
    out = np.zeros_like(val)
    for l in range(LANES):
        for j0 in range(2):
            a = np.array([val[l, 2 * m + j0] for m in range(8)])
            A = np.fft.fft(a)
            for k0 in range(8):
                out[l, 2 * k0 + j0] = A[k0] * W(512, (2 * l + j0) * k0)
    val = out
    # registers now: index 2 k0 + j0 ; lanes: (j5 j4 j3 j2 j1) = l
    # ---- exchange 1: columns chosen so that (c2, c1) [complex-register bits 1,0 of the *column pair index*] carry the bits
    #      that leave. Work in complex columns C = c >> 1 (4 bits: C3 C2 C1 C0), float column c = 2 C + z.
    #      ld 16x256b in complex terms: thread t gets rows 16 I + t/4 + 8 h, complex columns 4 k + (t % 4), k < 4
    #      -> complex register index 8 I + 2 k + h.   Leaving register bits = (C1, C0); entering: h <- L3, I <- L4.
    return val


def tm_exchange(val, colmap):
    """One forward trip on complex registers. colmap[r] = complex column that register r is stored to (a permutation of
    0..15: free, it is only a naming of registers). Returns new[t][8 I + 2 k + h] = old[16 I + t/4 + 8 h][reg stored at column 4 k + t % 4]."""
    inv = np.argsort(colmap)                  # column -> register
    new = np.zeros_like(val)
    src = np.zeros(val.shape + (2,), int)
    for t in range(LANES):
        for I in range(2):
            for k in range(4):
                for h in range(2):
                    row = 16 * I + t // 4 + 8 * h
                    col = 4 * k + (t % 4)
                    new[t, 8 * I + 2 * k + h] = val[row, inv[col]]
                    src[t, 8 * I + 2 * k + h] = (row, inv[col])
    return new, src


def fft512_model(x):
    W = lambda N, e: np.exp(-2j * np.pi * e / N)
    val = run(x)
    # label of every (lane, reg) as a dict of index bits, to keep the bookkeeping honest
    # after pass 1: reg = 2 k0 + j0, lane = (j5..j1)
    def lab1(l, r):
        return dict(k0=r >> 1, j0=r & 1, j51=l)            # j51 = bits j5..j1
    labels = [[lab1(l, r) for r in range(16)] for l in range(LANES)]

    def exchange(val, labels, colmap_fn):
        colmap = np.array([colmap_fn(r) for r in range(16)])
        assert sorted(colmap) == list(range(16)), colmap
        new, src = tm_exchange(val, colmap)
        nl = [[labels[src[t, r, 0]][src[t, r, 1]] for r in range(16)] for t in range(LANES)]
        return new, nl

    # ---- exchange 1: registers (k0_2 k0_1 k0_0 j0). Leaving: (k0_0 ^ r, k0_1 ^ r), r = k0_2. Kept: r, j0.
    #      complex column C = (C3 C2 | C1 C0) = (r, j0 | k0_0 ^ r, k0_1 ^ r): the C1 bit comes back into the registers at
    #      exchange 3 (so that k and k+1 end up in one lane), the C0 bit stays a lane bit to the end
    def cm1(reg):
        k0, j0 = reg >> 1, reg & 1
        r = k0 >> 2
        return (r << 3) | (j0 << 2) | (((k0 & 1) ^ r) << 1) | (((k0 >> 1) & 1) ^ r)
    val, labels = exchange(val, labels, cm1)
    # now reg = 8 I + 2 k + h with I = L4 = j5, h = L3 = j4, k = (C3 C2) = (r, j0); thread t = (j3 j2 j1 | c1 c0)
    # ---- pass 2: radix-4 over (j5 j4) = (I, h): butterflies indexed by k = (r, j0); twiddle W_64^{(j3..j0) k1}
    out = np.zeros_like(val)
    for t in range(LANES):
        for k in range(4):
            a = np.array([val[t, 8 * I + 2 * k + h] for I in range(2) for h in range(2)])     # digit d = 2 I + h = (j5 j4)
            lb = labels[t][2 * k]
            j30 = ((lb['j51'] & 7) << 1) | lb['j0']
            A = np.fft.fft(a)
            for k1 in range(4):
                out[t, 8 * (k1 >> 1) + 2 * k + (k1 & 1)] = A[k1] * W(64, j30 * k1)
                labels[t][8 * (k1 >> 1) + 2 * k + (k1 & 1)] = dict(lb, k1=k1, j51=lb['j51'] & 7)
    val = out
    # registers: 8 k1_1 + 2 (r j0) + k1_0. Leaving: (k1_1 ^ r, k1_0 ^ r); kept (r, j0)
    def cm2(reg):
        k1 = ((reg >> 3) << 1) | (reg & 1)
        r, j0 = (reg >> 2) & 1, (reg >> 1) & 1
        return (r << 3) | (j0 << 2) | (((k1 >> 1) ^ r) << 1) | ((k1 & 1) ^ r)
    val, labels = exchange(val, labels, cm2)
    # thread t' = (L2 L1 L0 | c1 c0) = (j1, e1a, e1b | k1_1^r, k1_0^r) ; entering I = L4 = j3, h = L3 = j2
    out = np.zeros_like(val)
    for t in range(LANES):
        for k in range(4):
            a = np.array([val[t, 8 * I + 2 * k + h] for I in range(2) for h in range(2)])     # digit (j3 j2)
            lb = labels[t][2 * k]
            j10 = ((lb['j51'] & 1) << 1) | lb['j0']
            A = np.fft.fft(a)
            for k2 in range(4):
                out[t, 8 * (k2 >> 1) + 2 * k + (k2 & 1)] = A[k2] * W(16, j10 * k2)
                labels[t][8 * (k2 >> 1) + 2 * k + (k2 & 1)] = dict(lb, k2=k2, j51=lb['j51'] & 1)
    val = out
    def cm3(reg):
        k2 = ((reg >> 3) << 1) | (reg & 1)
        r, j0 = (reg >> 2) & 1, (reg >> 1) & 1
        return (r << 3) | (j0 << 2) | (((k2 >> 1) ^ r) << 1) | ((k2 & 1) ^ r)
    val, labels = exchange(val, labels, cm3)
    # entering: I = L4 = j1, h = L3 = (k0_0 ^ r from exchange 1); registers 8 j1 + 2 (r j0) + h
    out = np.zeros_like(val)
    for t in range(LANES):
        for h in range(2):
            for r in range(2):
                a = np.array([val[t, 8 * j1 + 2 * (2 * r + j0) + h] for j1 in range(2) for j0 in range(2)])   # digit (j1 j0)
                lb = labels[t][2 * (2 * r) + h]
                A = np.fft.fft(a)
                for k3 in range(4):
                    out[t, 8 * (k3 >> 1) + 2 * (2 * r + (k3 & 1)) + h] = A[k3]
                    labels[t][8 * (k3 >> 1) + 2 * (2 * r + (k3 & 1)) + h] = dict(lb, k3=k3)
    val = out
    # final: k = k0 + 8 k1 + 32 k2 + 128 k3
    X = np.zeros(Q, complex)
    kmap = np.zeros((LANES, 16), int)
    for t in range(LANES):
        for rg in range(16):
            lb = labels[t][rg]
            k = lb['k0'] + 8 * lb['k1'] + 32 * lb['k2'] + 128 * lb['k3']
            X[k] = val[t, rg]
            kmap[t, rg] = k
    return X, kmap


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.normal(size=Q) + 1j * rng.normal(size=Q)
    X, kmap = fft512_model(x)
    ref = np.fft.fft(x)
    print("max err", np.abs(X - ref).max())
    assert np.abs(X - ref).max() < 1e-9
    assert sorted(kmap.ravel()) == list(range(Q))
    # complement closure: lane holding k also holds 511 - k
    for t in range(LANES):
        ks = set(kmap[t])
        assert all((511 - k) in ks for k in ks), t
    print("lane 0 k:", kmap[0]); print("lane 5 k:", kmap[5])

    def freq_of(lane, reg):          # the closed form used by host_tables.cpp: tm_fft_freq_of
        k3 = ((reg >> 3) << 1) | ((reg >> 1) & 1); r = (reg >> 2) & 1; h = reg & 1; x = r
        k0_0 = h ^ x; k0_1 = ((lane >> 4) & 1) ^ x; k0_2 = r
        k1 = ((lane >> 2) & 3) ^ (x * 3); k2 = (lane & 3) ^ (x * 3)
        return k0_0 | (k0_1 << 1) | (k0_2 << 2) | (k1 << 3) | (k2 << 5) | (k3 << 7)
    for t in range(LANES):
        for R in range(16):
            assert freq_of(t, R) == kmap[t, R], (t, R, freq_of(t, R), kmap[t, R])
    print("closed form of k(lane, register) verified")
    # storage of D2[k] inside a half (k < 256 / k >= 256): quad sq = [k3_0][k1_1][k0_2][k0_1][k1_0][k2_1][k2_0], pair element k0_0
    def storage_quad(k):
        b = lambda i: (k >> i) & 1
        return (b(7) << 6) | (b(4) << 5) | (b(2) << 4) | (b(1) << 3) | (b(3) << 2) | (b(6) << 1) | b(5)
    for R in range(0, 16, 2):          # one 128-bit store per register pair: quarter warps must hit 8 distinct 16-byte columns
        for qw in range(4):
            cols = {storage_quad(kmap[t, R] & 255) % 8 for t in range(8 * qw, 8 * qw + 8)}
            assert len(cols) == 8, (R, qw, cols)
            assert all((kmap[t, R] ^ kmap[t, R + 1]) == 1 for t in range(LANES))
    # overlap-add: lane L, iteration i reads storage quad 32 i + L  ->  true quad tq; every iteration covers whole 128-byte lines
    for i in range(4):
        tqs = []
        for L in range(LANES):
            sq = 32 * i + L
            tq = (sq & 0x44) | ((sq & 3) << 4) | ((sq & 0x20) >> 2) | ((sq & 0x18) >> 3)
            assert storage_quad(2 * tq) == sq
            tqs.append(tq)
        lines = {}
        for tq in tqs: lines.setdefault(tq >> 3, set()).add(tq & 7)
        assert all(len(v) == 8 for v in lines.values()) and len(lines) == 4
    print("storage permutation: conflict-free stores, whole-line PCM stores verified")
