"""CPU: numpy model of the CTA engine (rustfft_b200/csrc/engine.h) -- the thread-ownership pattern,
the per-stage twiddle / scatter formula, the four-step split, and the shared-memory layout's bank
behaviour.  These are the derivations the kernel comments point to."""
import numpy as np
import pytest


def engine_model(x, radices, E):
    L = len(x)
    T = L // E
    assert int(np.prod(radices)) == L
    cur = x.astype(np.complex128).copy()
    p = 1
    for s, R in enumerate(radices):
        assert E % R == 0
        nxt = np.zeros(L, complex)
        for j in range(T):
            v = [cur[j + T * q] for q in range(E)]  # thread j owns j + T*q on the read side
            for u in range(E // R):
                i = j + u * T
                k = i % p
                base = (i - k) * R + k
                a = np.array([v[u + r * (E // R)] * np.exp(-2j * np.pi * k * r / (p * R)) for r in range(R)])
                out = np.fft.fft(a)
                for m in range(R):
                    nxt[base + m * p] = out[m]
                    if s == len(radices) - 1:  # last stage lands on the owner's own slots
                        assert base + m * p == j + T * (u + m * (E // R))
        cur = nxt
        p *= R
    return cur


@pytest.mark.parametrize("L,radices,E", [
    (4096, [16, 16, 16], 16), (1024, [4, 16, 16], 16), (2048, [8, 16, 16], 16), (512, [2, 16, 16], 16),
    (128, [8, 16], 16), (32, [4, 8], 8), (64, [8, 8], 8), (16, [4, 4], 4), (8, [8], 8),
    (1024, [2, 8, 8, 8], 8), (360, [3, 4, 5, 6], 60),
])
def test_stockham_ownership_model(L, radices, E):
    rng = np.random.default_rng(L)
    x = rng.standard_normal(L) + 1j * rng.standard_normal(L)
    assert np.abs(engine_model(x, radices, E) - np.fft.fft(x)).max() < 1e-11


def test_four_step_split():
    # n = N2*n1 + n2, k = k1 + N1*k2 (kernels.h: LoadCols / StoreColsTw / StoreTransposed)
    N1, N2 = 16, 32
    N = N1 * N2
    rng = np.random.default_rng(1)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    w = np.fft.fft(x.reshape(N1, N2), axis=0)
    w *= np.exp(-2j * np.pi * np.outer(np.arange(N1), np.arange(N2)) / N)
    w = np.fft.fft(w, axis=1)
    X = np.zeros(N, complex)
    for k1 in range(N1):
        X[k1 + N1 * np.arange(N2)] = w[k1]
    assert np.abs(X - np.fft.fft(x)).max() < 1e-11


def _degree(addrs, elem_words):
    group = 32 // elem_words  # threads per shared-memory transaction (8-byte: 16, 16-byte: 8)
    worst = 0
    for g in range(0, 32, group):
        banks = {}
        for a in addrs[g:g + group]:
            for w in range(elem_words):
                banks.setdefault((a + w) % 32, set()).add(a + w)
        worst = max(worst, max(len(s) for s in banks.values()))
    return worst


def _worst_conflict(L, radices, E, F, maps, PS=4, elem_words=2):
    T = L // E
    NT = F * T
    unit = 32 // elem_words  # elements per 128-byte bank sweep
    LP = L + (L >> PS)
    if F > 1:  # Geo::LP in engine.h
        LPR = (LP + unit - 1) // unit * unit
        LP = LPR + 1 if F >= unit else LPR + unit // F
    sidx = lambda f, e: (f * LP + e + (e >> PS)) * elem_words
    worst, p = 1, 1
    for s, R in enumerate(radices):
        fj = (lambda t: (t // T, t % T)) if maps[s] == "JF" else (lambda t: (t % F, t // F))
        for warp in range(0, NT - 31, 32):
            if s > 0:
                for q in range(E):
                    worst = max(worst, _degree([sidx(fj(t)[0], fj(t)[1] + T * q) for t in range(warp, warp + 32)], elem_words))
            if s < len(radices) - 1:
                for u in range(E // R):
                    for m in range(R):
                        addrs = []
                        for t in range(warp, warp + 32):
                            f, j = fj(t)
                            i = j + u * T
                            k = i % p
                            addrs.append(sidx(f, (i - k) * R + k + m * p))
                        worst = max(worst, _degree(addrs, elem_words))
        p *= R
    return worst


@pytest.mark.parametrize("L,radices,E,F,maps", [
    (4096, [16, 16, 16], 16, 1, ["JF"] * 3), (2048, [8, 16, 16], 16, 2, ["JF"] * 3),
    (1024, [4, 16, 16], 16, 4, ["JF"] * 3), (512, [2, 16, 16], 16, 8, ["JF"] * 3), (256, [16, 16], 16, 8, ["JF"] * 2),
    (256, [16, 16], 16, 16, ["FF"] * 2), (512, [2, 16, 16], 16, 16, ["FF"] * 3),          # four-step pass A tiles
    (256, [16, 16], 16, 16, ["JF", "FF"]), (512, [2, 16, 16], 16, 16, ["JF", "FF", "FF"]),  # pass B tiles
    (1024, [16, 16, 4], 16, 8, ["FF"] * 3), (1024, [16, 16, 4], 16, 8, ["JF", "FF", "FF"]),  # 8-wide 1024-point tiles
])
def test_smem_layout_is_conflict_free_f32(L, radices, E, F, maps):
    assert _worst_conflict(L, radices, E, F, maps) == 1


def test_radix32_tiles_bank_model():
    """The radix-32 geometries: 512 = 16*32 (16 columns) is conflict free with the 1-in-16 pad; 1024 = 32*32 (8 columns) is
    2-way with it and conflict free with a 1-in-32 pad (impl.inl: B2_TILE1024_PS, queued for a timed A/B)."""
    assert _worst_conflict(512, [16, 32], 32, 16, ["FF"] * 2) == 1
    assert _worst_conflict(512, [16, 32], 32, 16, ["JF", "FF"]) == 1
    assert _worst_conflict(1024, [32, 32], 32, 8, ["FF"] * 2, PS=4) == 2
    assert _worst_conflict(1024, [32, 32], 32, 8, ["JF", "FF"], PS=4) == 2
    assert _worst_conflict(1024, [32, 32], 32, 8, ["FF"] * 2, PS=5) == 1
    assert _worst_conflict(1024, [32, 32], 32, 8, ["JF", "FF"], PS=5) == 1
