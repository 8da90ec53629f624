#!/usr/bin/env python3
"""Prints the curve constants of bftkv_b200/csrc/ed25519.cuh as radix-2^25.5 limb arrays
(10 limbs: 26,25,26,25,... bits), computed from the RFC 8032 definitions."""
p = 2**255 - 19
d = (-121665 * pow(121666, -1, p)) % p
sqrtm1 = pow(2, (p - 1) // 4, p)
By = (4 * pow(5, -1, p)) % p
u, v = (By * By - 1) % p, (d * By * By + 1) % p
Bx = pow(u * pow(v, -1, p) % p, (p + 3) // 8, p)
if (Bx * Bx - u * pow(v, -1, p)) % p != 0:
    Bx = Bx * sqrtm1 % p
if Bx & 1:
    Bx = p - Bx
assert (-Bx * Bx + By * By - 1 - d * Bx * Bx * By * By) % p == 0
L = 2**252 + 27742317777372353535851937790883648493


def limbs(x):
    out, sh = [], 0
    for i in range(10):
        b = 26 if i % 2 == 0 else 25
        out.append((x >> sh) & ((1 << b) - 1))
        sh += b
    return "{" + ", ".join(str(v) for v in out) + "}"


for name, val in [("kD", d), ("kD2", 2 * d % p), ("kSqrtM1", sqrtm1), ("kBx", Bx), ("kBy", By), ("kBt", Bx * By % p)]:
    print("BFTQ_ED_CONST int32_t %s[10] = %s;" % (name, limbs(val)))
print("// L =", hex(L))
print("BFTQ_ED_CONST uint32_t kL[8] = {" + ", ".join(hex((L >> (32 * i)) & 0xFFFFFFFF) for i in range(8)) + "};")
