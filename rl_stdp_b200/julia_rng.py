"""Julia's MersenneTwister(seed) restated (dSFMT-19937, Saito & Matsumoto 2009) — just enough to replay
the reference's `rng_action = Random.MersenneTwister(0)` tie-breaker (Actors/cartpole_fitness.jl:55,86;
Actors/mountcar_fitness.jl:51,88): `rand(rng, 0:1)` / `rand(rng, pool)` with a two-element pool takes the
lowest mantissa bit of the next double in [1, 2) (Random's SamplerRangeFast with mask 1, no rejection).

TEST INFRASTRUCTURE.  Pinned by the published first doubles of `rand(MersenneTwister(0))`
(0.8236475079774124, 0.9103565379264364, 0.16456579813368521, ...), see tests/test_oracle_cpu.py."""
from __future__ import annotations

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF
N = 191            # (19937 - 128) / 104 + 1 128-bit words, + 1 "lung"
POS1, SL1, SR = 117, 19, 12
MSK1, MSK2 = 0x000FFAFFFFFFFB3F, 0x000FFDFFFC90FFFD
FIX1, FIX2 = 0x90014964B32F4329, 0x3B8D12AC548A7C7A
PCV1, PCV2 = 0x3D84E1AC0DC82880, 0x0000000000000001
LOW_MASK, HIGH_CONST = 0x000FFFFFFFFFFFFF, 0x3FF0000000000000


def _f1(x):
    return ((x ^ (x >> 27)) * 1664525) & M32


def _f2(x):
    return ((x ^ (x >> 27)) * 1566083941) & M32


class DSFMT:
    def __init__(self, key):
        """dsfmt_init_by_array with 32-bit keys (Julia: make_seed(n) -> the 32-bit limbs of n, [0] for 0)."""
        size = (N + 1) * 4
        lag, mid = 11, ((N + 1) * 4 - 11) // 2
        p = [0x8B8B8B8B] * size
        klen = len(key)
        count = max(klen + 1, size)
        r = _f1(p[0] ^ p[mid % size] ^ p[(size - 1) % size])
        p[mid % size] = (p[mid % size] + r) & M32
        r = (r + klen) & M32
        p[(mid + lag) % size] = (p[(mid + lag) % size] + r) & M32
        p[0] = r
        count -= 1
        i, j = 1, 0
        while j < count:
            r = _f1(p[i] ^ p[(i + mid) % size] ^ p[(i + size - 1) % size])
            p[(i + mid) % size] = (p[(i + mid) % size] + r) & M32
            r = (r + (key[j] if j < klen else 0) + i) & M32
            p[(i + mid + lag) % size] = (p[(i + mid + lag) % size] + r) & M32
            p[i] = r
            i = (i + 1) % size
            j += 1
        for j in range(size):
            r = _f2((p[i] + p[(i + mid) % size] + p[(i + size - 1) % size]) & M32)
            p[(i + mid) % size] ^= r
            r = (r - i) & M32
            p[(i + mid + lag) % size] ^= r
            p[i] = r
            i = (i + 1) % size
        s = [p[2 * k] | (p[2 * k + 1] << 32) for k in range((N + 1) * 2)]   # little-endian 64-bit view
        for k in range(N * 2):                                               # initial_mask
            s[k] = (s[k] & LOW_MASK) | HIGH_CONST
        inner = ((s[2 * N] ^ FIX1) & PCV1) ^ ((s[2 * N + 1] ^ FIX2) & PCV2)  # period_certification
        sh = 32
        while sh:
            inner ^= inner >> sh
            sh >>= 1
        if not inner & 1:
            s[2 * N + 1] ^= 1
        self.s = s
        self.idx = 2 * N

    def _gen_all(self):
        s = self.s
        l0, l1 = s[2 * N], s[2 * N + 1]
        for i in range(N):
            b = i + POS1 if i + POS1 < N else i + POS1 - N
            t0, t1 = s[2 * i], s[2 * i + 1]
            n0 = ((t0 << SL1) & M64) ^ (l1 >> 32) ^ ((l1 << 32) & M64) ^ s[2 * b]
            n1 = ((t1 << SL1) & M64) ^ (l0 >> 32) ^ ((l0 << 32) & M64) ^ s[2 * b + 1]
            l0, l1 = n0, n1
            s[2 * i] = (l0 >> SR) ^ (l0 & MSK1) ^ t0
            s[2 * i + 1] = (l1 >> SR) ^ (l1 & MSK2) ^ t1
        s[2 * N], s[2 * N + 1] = l0, l1
        self.idx = 0

    def raw52(self):
        """the 64-bit pattern of the next double in [1, 2)"""
        if self.idx >= 2 * N:
            self._gen_all()
        v = self.s[self.idx]
        self.idx += 1
        return v

    def rand(self):
        import struct
        return struct.unpack("<d", struct.pack("<Q", self.raw52()))[0] - 1.0


def julia_mt_pick_bits(seed: int, n: int):
    """The first n results of `rand(MersenneTwister(seed), 0:1)` (0 = first element of a two-element pool)."""
    key = []
    while True:
        key.append(seed & M32)
        seed >>= 32
        if seed == 0:
            break
    g = DSFMT(key)
    return [g.raw52() & 1 for _ in range(n)]
