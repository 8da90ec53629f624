#!/usr/bin/env python3
"""Deterministic RSA-2048 test keys (e = 65537) from seed 0xBF7C0001.
Seeded Mersenne-Twister candidates + Miller-Rabin; 33 keys so configs 3 (R=16) and 5 (R=31) can
give each replica its own key (+ 2 for a non-member and an outsider in the n = 31 tests).  Output committed as tests/golden/rsa_keys_bf7c0001.json.
TEST/BENCH FIXTURE ONLY — these private keys are public."""
import json, os, random
SEED, NKEYS = 0xBF7C0001, 33
rng = random.Random(SEED)
SMALL = [p for p in range(3, 2000, 2) if all(p % q for q in range(3, int(p ** 0.5) + 1, 2))]


def is_prime(n):
    for p in SMALL:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2; s += 1
    for _ in range(8):
        a = rng.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = pow(x, 2, n)
            if x == n - 1:
                break
        else:
            return False
    return True


def gen_prime(bits):
    while True:
        c = rng.getrandbits(bits) | (3 << (bits - 2)) | 1
        if (c - 1) % 65537 and is_prime(c):
            return c


keys = []
for _ in range(NKEYS):
    p, q = gen_prime(1024), gen_prime(1024)
    assert p != q and (p * q).bit_length() == 2048
    keys.append({"p": "%x" % p, "q": "%x" % q})
out = os.path.join(os.path.dirname(__file__), "rsa_keys_bf7c0001.json")
json.dump({"seed": "0xBF7C0001", "e": 65537, "keys": keys}, open(out, "w"))
print("wrote", out)
