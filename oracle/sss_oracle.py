"""Oracle restatement of Shamir share-combine and the threshold-signature combine steps.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Follows the reference:
  crypto/sss/sss.go:23-47      Distribute (polynomial evaluation; coefficients are an input here)
  crypto/sss/sss.go:81-92      SSSProcess.calculateSecret
  crypto/sss/sss.go:94-107     Lagrange  (numerator/denominator built as UNREDUCED integers,
                               big.Int.ModInverse of a possibly negative b, one final Mod)
  crypto/threshold/dsa/dsa_core.go:375-387   formatDSA
  crypto/threshold/dsa/dsa_core.go:389-403   calculateS
  crypto/threshold/dsa/dsa.go:33-52          dsaGroupOperations.CalculateR
  crypto/threshold/rsa/rsa.go:318-329        calculateSignature (product of partial signatures)
  crypto/threshold/rsa/rsa.go:345-378        hashPrefixes / emsaEncode
  crypto/threshold/rsa/rsa.go:380-393        I2OS
"""
from typing import List, Sequence, Tuple

SHA256_PREFIX = bytes([0x30, 0x31, 0x30, 0x0d, 0x06, 0x09, 0x60, 0x86, 0x48, 0x01,
                       0x65, 0x03, 0x04, 0x02, 0x01, 0x05, 0x00, 0x04, 0x20])   # rsa.go:350


def distribute(secret: int, coeffs: Sequence[int], n: int, m: int) -> List[Tuple[int, int]]:
    """sss.go:23-47 with the random coefficients poly[1..k-1] passed in."""
    poly = [secret] + list(coeffs)
    res = []
    for i in range(n):
        x0 = i + 1
        x = x0
        f = poly[0]
        for j in range(1, len(poly)):
            f = (f + poly[j] * x) % m
            x *= x0
        res.append((i + 1, f))
    return res


def lagrange(x: int, xs: Sequence[int], m: int) -> int:      # sss.go:94-107
    a, b = 1, 1
    for r in xs:
        if r == x:
            continue
        a *= r
        b *= (r - x)
    # Go >= 1.11 ModInverse reduces a negative operand mod m first; returns nil when not
    # invertible (the reference then panics) — mirrored by Python's ValueError.
    binv = pow(b % m, -1, m)
    return (a * binv) % m


def calculate_secret(shares: Sequence[Tuple[int, int]], m: int) -> int:   # sss.go:81-92
    xs = [x for x, _ in shares]
    s = 0
    for x, y in shares:
        s = (s + lagrange(x, xs, m) * y) % m
    return s


def calculate_s(shares: Sequence[Tuple[int, int]], q: int) -> int:        # dsa_core.go:389-403
    xs = [x for x, _ in shares]
    s = 0
    for x, y in shares:
        t = (y * lagrange(x, xs, q)) % q
        s = (s + t) % q
    return s


def format_dsa(r: int, s: int, q: int) -> bytes:                           # dsa_core.go:375-387
    n = (q.bit_length() + 7) // 8
    return r.to_bytes(n, "big") + s.to_bytes(n, "big")


def dsa_calculate_r(rs: Sequence[Tuple[int, int, int]], p: int, q: int) -> int:   # dsa.go:33-52
    """rs: (x_i, R_i, v_i).  r = (prod R_i^l_i mod p)^((sum v_i l_i)^-1 mod q) mod p mod q."""
    xs = [x for x, _, _ in rs]
    r, v = 1, 0
    for x, ri, vi in rs:
        l = lagrange(x, xs, q)
        r = (r * pow(ri, l, p)) % p
        v = (v + (vi * l) % q) % q
    v = pow(v, -1, q)
    return pow(r, v, p) % q


def i2os(b: int, sz: int) -> bytes:                                        # rsa.go:380-393
    c = b.to_bytes((b.bit_length() + 7) // 8, "big")
    return c if len(c) >= sz else b"\x00" * (sz - len(c)) + c


def emsa_encode(prefix: bytes, dgst: bytes, n: int) -> int:                # rsa.go:356-378
    emlen = (n.bit_length() + 7) // 8
    mlen = len(prefix) + len(dgst)
    padlen = emlen - mlen
    if padlen < 3:
        raise ValueError("invalid input")
    em = b"\x00\x01" + b"\xff" * (padlen - 3) + b"\x00" + prefix + dgst
    return int.from_bytes(em, "big")


def rsa_combine(psigs: Sequence[int], n: int) -> bytes:                    # rsa.go:244-251,318-329
    """Product of the partial signatures at the leaves of a completed signature tree."""
    s = 1
    for p in psigs:
        s = (s * p) % n
    return i2os(s, (n.bit_length() + 7) // 8)


# ---- P-256 (crypto/elliptic) and ecdsaGroupOperations.CalculateR ---------------------------------
P256_P = 2 ** 256 - 2 ** 224 + 2 ** 192 + 2 ** 96 - 1
P256_B = 0x5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b
P256_N = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
P256_G = (0x6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296,
          0x4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5)


def p256_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    p = P256_P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % p == 0:
            return None
        l = (3 * P[0] * P[0] - 3) * pow(2 * P[1], -1, p) % p
    else:
        l = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
    x = (l * l - P[0] - Q[0]) % p
    return x, (l * (P[0] - x) - P[1]) % p


def p256_mul(k, P):
    R = None
    while k:
        if k & 1:
            R = p256_add(R, P)
        P = p256_add(P, P)
        k >>= 1
    return R


def p256_marshal(P) -> bytes:                       # elliptic.Marshal
    return b"\x04" + P[0].to_bytes(32, "big") + P[1].to_bytes(32, "big")


def ecdsa_calculate_r(rs: Sequence[Tuple[int, Tuple[int, int], int]]) -> int:
    """crypto/threshold/ecdsa/ecdsa.go:36-59.  rs: (x_i, R_i (affine point), v_i)."""
    xs = [x for x, _, _ in rs]
    acc, v = None, 0
    for x, ri, vi in rs:
        l = lagrange(x, xs, P256_N)
        acc = p256_add(acc, p256_mul(l, ri))
        v = (v + (vi * l) % P256_N) % P256_N
    v = pow(v, -1, P256_N)
    R = p256_mul(v, acc)
    return R[0] % P256_N
