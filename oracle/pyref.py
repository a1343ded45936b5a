"""Big-integer oracle for the BLS12-381 hot path (TEST INFRASTRUCTURE ONLY).

This module is the slow, obviously-correct restatement that pins the C oracle
(`oracle/oracle.c`) and, through it, the CUDA path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import it.

What it restates (reference file:line, relative to /root/reference):
  * field/curve constants          bls12-381/include/bls12_381_constants.h:66-224
    (every constant below is DERIVED from p, r and the generator 7, then
     compared with the spec literals the reference's KAT test pins,
     bls12-381/tests/test_known_answer_vectors.cu:60-200 -- see
     tests/test_oracle_constants.py)
  * Montgomery encodings           bls12-381/include/field.cuh:906-928
  * affine/Jacobian conventions    bls12-381/include/point.cuh:286-318, :455-525
  * MSM semantics                  bls12-381/src/backend/icicle_curve_api.cu:243-407
  * NTT semantics (= best_fft)     core/ntt.rs:1488-1603, bls12-381/include/ntt.cuh:123-183
  * signed-digit window scheme     bls12-381/src/curve/msm_kernels.cu:69-143

The CPU arithmetic the reference's MIDNIGHT_DEVICE=cpu path really runs lives
in third-party crates that are NOT under /root/reference (midnight-curves 0.2.0
over blst; Cargo.toml:19); what is restated here is their published
mathematical contract, which fixes every output bit because outputs are
canonical field elements / canonical affine points.
"""
from __future__ import annotations

# --------------------------------------------------------------------------
# constants (derived, not transcribed)
# --------------------------------------------------------------------------
# BLS parameter x = -0xd201000000010000 ; r = x^4 - x^2 + 1 ; p = (x-1)^2 r / 3 + x
BLS_X = -0xD201000000010000
R_MOD = BLS_X**4 - BLS_X**2 + 1                       # scalar field order r (255 bit)
P_MOD = ((BLS_X - 1) ** 2 * R_MOD) // 3 + BLS_X       # base field order p (381 bit)

FR_BITS, FQ_BITS = 256, 384
FR_R = (1 << FR_BITS) % R_MOD                          # Montgomery one
FQ_R = (1 << FQ_BITS) % P_MOD
FR_R2 = (FR_R * FR_R) % R_MOD
FQ_R2 = (FQ_R * FQ_R) % P_MOD
FR_RINV = pow(FR_R, -1, R_MOD)
FQ_RINV = pow(FQ_R, -1, P_MOD)
FR_INV64 = (-pow(R_MOD, -1, 1 << 64)) % (1 << 64)      # -r^-1 mod 2^64
FQ_INV64 = (-pow(P_MOD, -1, 1 << 64)) % (1 << 64)
FR_INV32 = FR_INV64 & 0xFFFFFFFF
FQ_INV32 = FQ_INV64 & 0xFFFFFFFF

FR_TWO_ADICITY = 32
FR_GENERATOR = 7                                       # multiplicative generator used by ff/halo2/ICICLE
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_TWO_ADICITY, R_MOD)   # order 2^32

# G1 generator (standard form) -- the IETF/zkcrypto generator
G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
# G2 generator (standard form), x = x0 + x1 u, y = y0 + y1 u
G2_X0 = 0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8
G2_X1 = 0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E
G2_Y0 = 0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801
G2_Y1 = 0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE


def fr_omega(log_n: int) -> int:
    """omega_k = ROOT_OF_UNITY^(2^(32-k))  (core/ntt.rs:1488-1494)."""
    assert 0 <= log_n <= FR_TWO_ADICITY
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_TWO_ADICITY - log_n), R_MOD)


# --------------------------------------------------------------------------
# byte / limb layouts (core/types.rs:148-270: little-endian u64 limbs)
# --------------------------------------------------------------------------
def to_limbs(v: int, n: int) -> list[int]:
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs(l) -> int:
    v = 0
    for i, x in enumerate(l):
        v |= int(x) << (64 * i)
    return v


def fr_to_mont(v: int) -> int:
    return (v * FR_R) % R_MOD


def fr_from_mont(v: int) -> int:
    return (v * FR_RINV) % R_MOD


def fq_to_mont(v: int) -> int:
    return (v * FQ_R) % P_MOD


def fq_from_mont(v: int) -> int:
    return (v * FQ_RINV) % P_MOD


def fr_bytes(v: int) -> bytes:
    return int(v).to_bytes(32, "little")


def fq_bytes(v: int) -> bytes:
    return int(v).to_bytes(48, "little")


# --------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2+1)   (bls12-381/include/point.cuh:81-225)
# --------------------------------------------------------------------------
class Fq2:
    __slots__ = ("c0", "c1")

    def __init__(self, c0=0, c1=0):
        self.c0 = c0 % P_MOD
        self.c1 = c1 % P_MOD

    def __add__(self, o):
        return Fq2(self.c0 + o.c0, self.c1 + o.c1)

    def __sub__(self, o):
        return Fq2(self.c0 - o.c0, self.c1 - o.c1)

    def __neg__(self):
        return Fq2(-self.c0, -self.c1)

    def __mul__(self, o):
        if isinstance(o, int):
            return Fq2(self.c0 * o, self.c1 * o)
        return Fq2(self.c0 * o.c0 - self.c1 * o.c1, self.c0 * o.c1 + self.c1 * o.c0)

    __rmul__ = __mul__

    def __eq__(self, o):
        if isinstance(o, int):
            return self.c0 == o % P_MOD and self.c1 == 0
        return self.c0 == o.c0 and self.c1 == o.c1

    def __hash__(self):
        return hash((self.c0, self.c1))

    def inv(self):
        n = pow(self.c0 * self.c0 + self.c1 * self.c1, -1, P_MOD)
        return Fq2(self.c0 * n, -self.c1 * n)

    def is_zero(self):
        return self.c0 == 0 and self.c1 == 0

    def __repr__(self):
        return f"Fq2({hex(self.c0)}, {hex(self.c1)})"


# --------------------------------------------------------------------------
# generic short-Weierstrass y^2 = x^3 + b, affine with None = infinity
# --------------------------------------------------------------------------
class _Fq:
    """int wrapper so G1 and G2 share the same curve code."""
    __slots__ = ("v",)

    def __init__(self, v=0):
        self.v = v % P_MOD

    def __add__(self, o):
        return _Fq(self.v + o.v)

    def __sub__(self, o):
        return _Fq(self.v - o.v)

    def __neg__(self):
        return _Fq(-self.v)

    def __mul__(self, o):
        if isinstance(o, int):
            return _Fq(self.v * o)
        return _Fq(self.v * o.v)

    __rmul__ = __mul__

    def __eq__(self, o):
        return self.v == (o.v if isinstance(o, _Fq) else o % P_MOD)

    def __hash__(self):
        return hash(self.v)

    def inv(self):
        return _Fq(pow(self.v, -1, P_MOD))

    def is_zero(self):
        return self.v == 0


def _aff_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2).is_zero():
            return None
        lam = (x1 * x1 * 3) * (y1 * 2).inv()
    else:
        lam = (y2 - y1) * (x2 - x1).inv()
    x3 = lam * lam - x1 - x2
    y3 = lam * (x1 - x3) - y1
    return (x3, y3)


def _aff_neg(P):
    return None if P is None else (P[0], -P[1])


def _aff_mul(k: int, P):
    k %= R_MOD
    R = None
    Q = P
    while k:
        if k & 1:
            R = _aff_add(R, Q)
        Q = _aff_add(Q, Q)
        k >>= 1
    return R


# ---- G1: points are (x:int, y:int) standard form, None = infinity ---------
G1_GEN = (G1_X, G1_Y)


def g1_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - 4) % P_MOD == 0


def _w1(P):
    return None if P is None else (_Fq(P[0]), _Fq(P[1]))


def _u1(P):
    return None if P is None else (P[0].v, P[1].v)


def g1_add(P, Q):
    return _u1(_aff_add(_w1(P), _w1(Q)))


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % P_MOD)


def g1_mul(k: int, P):
    return _u1(_aff_mul(k, _w1(P)))


def g1_msm(scalars, points):
    """sum_i s_i * P_i, naive (for small n)."""
    acc = None
    for s, P in zip(scalars, points):
        acc = _aff_add(acc, _aff_mul(s, _w1(P)))
    return _u1(acc)


# ---- G2: points are (Fq2, Fq2) -----------------------------------------
G2_GEN = (Fq2(G2_X0, G2_X1), Fq2(G2_Y0, G2_Y1))
G2_B = Fq2(4, 4)


def g2_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - G2_B).is_zero()


def g2_add(P, Q):
    return _aff_add(P, Q)


def g2_neg(P):
    return _aff_neg(P)


def g2_mul(k: int, P):
    return _aff_mul(k, P)


def g2_msm(scalars, points):
    acc = None
    for s, P in zip(scalars, points):
        acc = _aff_add(acc, _aff_mul(s, P))
    return acc


# --------------------------------------------------------------------------
# endomorphisms and subgroup membership (ground truth for csrc/glv.cuh).  The reference has only the constants
# (bls12-381/src/curve/point_ops.cu:103-142) and TODOs for the checks (include/point.cuh:419-448).
# --------------------------------------------------------------------------
def curve_mul_unreduced(k: int, P):
    """[k] P for ANY point of the curve (k is NOT reduced mod r: the point need not have order r).  P is a pair of
    _Fq or Fq2 coordinates or None."""
    R = None
    Q = P
    while k:
        if k & 1:
            R = _aff_add(R, Q)
        Q = _aff_add(Q, Q)
        k >>= 1
    return R


GLV_LAMBDA = BLS_X * BLS_X - 1                          # eigenvalue of phi on G1; r = lambda^2 + lambda + 1
assert GLV_LAMBDA * GLV_LAMBDA + GLV_LAMBDA + 1 == R_MOD


def _glv_beta() -> int:
    g = 2
    while pow(g, (P_MOD - 1) // 3, P_MOD) == 1:
        g += 1
    b = pow(g, (P_MOD - 1) // 3, P_MOD)
    lg = _u1(curve_mul_unreduced(GLV_LAMBDA, _w1(G1_GEN)))
    for cand in (b, b * b % P_MOD):
        if (cand * G1_X % P_MOD, G1_Y) == lg:
            return cand
    raise AssertionError("no cube root of unity matches lambda")


GLV_BETA = _glv_beta()                                  # phi(x, y) = (beta x, y) = [lambda](x, y) on G1


def glv_decompose(k: int):
    """k = k1 + k2 * lambda with 0 <= k1 < lambda, 0 <= k2 <= lambda + 1."""
    k2, k1 = divmod(k % R_MOD, GLV_LAMBDA)
    return k1, k2


def g1_in_subgroup(P) -> bool:
    """ground truth: [r] P = O"""
    return P is None or curve_mul_unreduced(R_MOD, _w1(P)) is None


def g2_in_subgroup(P) -> bool:
    return P is None or curve_mul_unreduced(R_MOD, P) is None


def fq2_pow(a: "Fq2", e: int) -> "Fq2":
    r = Fq2(1, 0)
    while e:
        if e & 1:
            r = r * a
        a = a * a
        e >>= 1
    return r


PSI_CX = fq2_pow(Fq2(1, 1), (P_MOD - 1) // 3).inv()       # psi(x, y) = (conj(x) PSI_CX, conj(y) PSI_CY)
PSI_CY = fq2_pow(Fq2(1, 1), (P_MOD - 1) // 2).inv()


def g2_psi(P):
    if P is None:
        return None
    x, y = P
    return (Fq2(x.c0, -x.c1) * PSI_CX, Fq2(y.c0, -y.c1) * PSI_CY)


def fq_sqrt(a: int):
    """square root in Fq (p = 3 mod 4) or None"""
    a %= P_MOD
    y = pow(a, (P_MOD + 1) // 4, P_MOD)
    return y if y * y % P_MOD == a else None


def fq2_sqrt(a: "Fq2"):
    """square root in Fq2 = Fq[u]/(u^2+1), p = 3 mod 4 (complex method), or None"""
    if a.is_zero():
        return a
    n = fq_sqrt(a.c0 * a.c0 + a.c1 * a.c1)              # norm must be a square in Fq
    if n is None:
        return None
    inv2 = pow(2, -1, P_MOD)
    for nn in (n, -n):
        d = fq_sqrt((a.c0 + nn) * inv2)
        if d is None or d == 0:
            continue
        c = Fq2(d, a.c1 * pow(2 * d, -1, P_MOD))
        if c * c == a:
            return c
    return None


def g1_curve_point(seed: int):
    """a point of E(Fq): y^2 = x^3 + 4 that is (with overwhelming probability) NOT in the order-r subgroup"""
    x = seed
    while True:
        y = fq_sqrt(x * x * x + 4)
        if y is not None:
            return (x % P_MOD, y)
        x += 1


def g2_curve_point(seed: int):
    x = Fq2(seed, 1)
    while True:
        y = fq2_sqrt(x * x * x + G2_B)
        if y is not None:
            return (x, y)
        x = x + Fq2(1, 0)


# --------------------------------------------------------------------------
# wire encodings of points
# --------------------------------------------------------------------------
def g1_affine_mont_bytes(P) -> bytes:
    """96-byte Montgomery affine, infinity = (0,0)  (point.cuh:295-302)."""
    if P is None:
        return bytes(96)
    return fq_bytes(fq_to_mont(P[0])) + fq_bytes(fq_to_mont(P[1]))


def g2_affine_mont_bytes(P) -> bytes:
    if P is None:
        return bytes(192)
    x, y = P
    return b"".join(fq_bytes(fq_to_mont(c)) for c in (x.c0, x.c1, y.c0, y.c1))


def g1_result_std_bytes(P) -> bytes:
    """ICICLE result: (x, y, 1) standard form, infinity = (0, 1, 0)
    (icicle_curve_api.cu:134-179)."""
    if P is None:
        return fq_bytes(0) + fq_bytes(1) + fq_bytes(0)
    return fq_bytes(P[0]) + fq_bytes(P[1]) + fq_bytes(1)


def g2_result_std_bytes(P) -> bytes:
    if P is None:
        return fq_bytes(0) * 2 + fq_bytes(1) + fq_bytes(0) + fq_bytes(0) * 2
    x, y = P
    return b"".join(fq_bytes(c) for c in (x.c0, x.c1, y.c0, y.c1, 1, 0))


# --------------------------------------------------------------------------
# NTT (natural in / natural out), = halo2/midnight-curves best_fft semantics
# --------------------------------------------------------------------------
def bit_reverse(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def ntt_naive(a, omega):
    n = len(a)
    return [sum(a[j] * pow(omega, (i * j) % n, R_MOD) for j in range(n)) % R_MOD for i in range(n)]


def ntt(a, omega=None, inverse=False):
    """Radix-2 NTT on canonical ints; forward uses omega_k, inverse omega_k^-1 and n^-1."""
    n = len(a)
    if n == 0:
        return []
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    if omega is None:
        omega = fr_omega(log_n)
    if inverse:
        omega = pow(omega, -1, R_MOD)
    a = [a[bit_reverse(i, log_n)] for i in range(n)]
    m = 1
    while m < n:
        wm = pow(omega, n // (2 * m), R_MOD)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % R_MOD
                u = a[k + j]
                a[k + j] = (u + t) % R_MOD
                a[k + j + m] = (u - t) % R_MOD
                w = w * wm % R_MOD
        m *= 2
    if inverse:
        ninv = pow(n, -1, R_MOD)
        a = [x * ninv % R_MOD for x in a]
    return a


def coset_ntt(a, g, inverse=False):
    """forward: x[i]*g^i then NTT; inverse: iNTT then * g^-i (ntt.cuh:123-183)."""
    n = len(a)
    if not inverse:
        return ntt([x * pow(g, i, R_MOD) % R_MOD for i, x in enumerate(a)])
    y = ntt(a, inverse=True)
    gi = pow(g, -1, R_MOD)
    return [x * pow(gi, i, R_MOD) % R_MOD for i, x in enumerate(y)]


def apply_ordering(vec, ordering: str, which: str):
    """ICICLE orderings kNN/kNR/kRN/kRR (icicle_types.cuh:89-96): first letter = input
    order, second = output order; R = bit-reversed index."""
    n = len(vec)
    bits = n.bit_length() - 1
    letter = ordering[0] if which == "in" else ordering[1]
    if letter == "N":
        return list(vec)
    return [vec[bit_reverse(i, bits)] for i in range(n)]


# --------------------------------------------------------------------------
# signed-digit window decomposition (msm_kernels.cu:69-143), used to pin the
# CUDA digit kernel and the C oracle's Pippenger
# --------------------------------------------------------------------------
def signed_digits(s: int, c: int, num_windows: int):
    """digits d_w in [-(2^(c-1)-1) .. 2^(c-1)] hmm -- returns list with sum d_w 2^(cw) == s."""
    out = []
    carry = 0
    half = 1 << (c - 1)
    for w in range(num_windows):
        d = ((s >> (c * w)) & ((1 << c) - 1)) + carry
        carry = 0
        if d > half:
            d -= 1 << c
            carry = 1
        out.append(d)
    return out, carry


# --------------------------------------------------------------------------
# deterministic inputs: SplitMix64 -> 4 limbs, top limb &= 2^63-1, reject >= r
# (same acceptance rule as bls12-381/tests/security_audit_tests.cuh:400-416)
# --------------------------------------------------------------------------
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def fr(self) -> int:
        while True:
            l = [self.next() for _ in range(4)]
            l[3] &= 0x7FFFFFFFFFFFFFFF
            v = from_limbs(l)
            if v < R_MOD:
                return v
