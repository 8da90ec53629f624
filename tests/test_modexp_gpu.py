"""GPU parity for K5 (Lagrange in the exponent): modexp, auth.calculateSharedSecret
(crypto/auth/auth.go:386-399) and dsa.CalculateR (crypto/threshold/dsa/dsa.go:33-52) against Python
big-int restatements (oracle/sss_oracle.py) on the reference's own group parameters."""
import random

import numpy as np
import pytest

from oracle import sss_oracle as sss

pytestmark = pytest.mark.gpu

# crypto/auth/auth.go:81-115 (== crypto/sss/sss_test.go:15-45): 2048-bit safe prime p = 2q + 1
AUTH_P = int("b0a67d9f5cebc0ffe81690e7b2670ab05f9fa4c2e73639f660c0408a2d9a4a8b454a9893fd7d4e8fa399cfc9c9ba05b080f903e33bcdcbef"
             "aed40915e51d46f58d1a5bd204db20fa3fe9db71f0b8e0aa87b5771406f25fad59e7f10fe5255644758872ea2dec1f6dcd11be905de59a04"
             "4f6c2ea3982b2235acc9021a196fc4ce0b19f6b312ee9cfc5997dc5f7ce2f386131294a56ba93a41a3b60e27e03956039f51ae73b89c795c"
             "5ae7d841e9b455c37341c052404e8fe9fe4f0d52bc162a41f1eeb9ef292c66a9d6a619aa548807eb1187ee22bd62e20e26c3c08c22ecef12"
             "d3b2304a010ed1f50a68e0261afe1a0bdddf7ab8a61774d3af3f1cce2b95dad3", 16)
AUTH_Q = (AUTH_P - 1) // 2


def test_modexp_matches_pow(engine, golden):
    rng = random.Random(5)
    dsa_p = int(golden["dsa_test"]["P"], 16)
    for m in (AUTH_P, dsa_p, (1 << 2047) | rng.getrandbits(2047) | 1):
        mlen = (m.bit_length() + 7) // 8
        bases = [rng.getrandbits(8 * mlen) for _ in range(60)] + [0, 1, m - 1, m, m + 5 if m + 5 < 1 << (8 * mlen) else 2]
        exps = [rng.getrandbits(rng.choice([1, 8, 160, 256])) for _ in range(60)] + [0, 1, 2, 3, 65537]
        got = engine.modexp_batch(m, bases, exps, elen=32)
        assert got == [pow(b, e, m) for b, e in zip(bases, exps)]
    # full-length exponents (auth: lambda mod q is 2047 bit)
    bases = [rng.getrandbits(2048) for _ in range(12)]
    exps = [rng.getrandbits(2047) for _ in range(12)]
    assert engine.modexp_batch(AUTH_P, bases, exps, elen=256) == [pow(b, e, AUTH_P) for b, e in zip(bases, exps)]


def test_auth_shared_secret(engine):
    """TestAuth's core (crypto/auth/auth_test.go:103): k of n servers' g^{y_i} combine to g^s."""
    rng = random.Random(9)
    n, k, g = 10, 7, 4                                          # 4 = 2^2 generates the order-q subgroup of the safe prime
    xs, ys, exp = [], [], []
    for _ in range(24):
        s = rng.randrange(AUTH_Q)
        shares = sss.distribute(s, [rng.randrange(AUTH_Q) for _ in range(k - 1)], n, AUTH_Q)
        pick = rng.sample(shares, k)
        xs.append([x for x, _ in pick])
        ys.append([pow(g, y, AUTH_P) for _, y in pick])
        gs = 1
        for x, y in pick:                                         # auth.go:386-399 restated
            gs = (gs * pow(pow(g, y, AUTH_P), sss.lagrange(x, [a for a, _ in pick], AUTH_Q), AUTH_P)) % AUTH_P
        assert gs == pow(g, s, AUTH_P)
        exp.append(gs)
    got, st = engine.lagrange_exp_product_batch(AUTH_P, AUTH_Q, np.array(xs, np.int32), ys)
    assert not st.any() and got == exp


def test_dsa_calculate_r(engine, golden):
    """dsa.CalculateR on the parameters of dsa_test.go:26-28 (n=10, 2t=8 partial results)."""
    p, q, g = (int(golden["dsa_test"][k], 16) for k in "PQG")
    rng = random.Random(17)
    n, t2 = 10, 8
    xs, ris, vis, exp = [], [], [], []
    for _ in range(40):
        a, kk = rng.randrange(1, q), rng.randrange(1, q)
        sa = sss.distribute(a, [rng.randrange(q) for _ in range(t2 // 2 - 1)], n, q)
        sk = sss.distribute(kk, [rng.randrange(q) for _ in range(t2 // 2 - 1)], n, q)
        idx = rng.sample(range(n), t2)
        rs = [(sa[i][0], pow(g, sa[i][1], p), (sa[i][1] * sk[i][1]) % q) for i in idx]    # (x_i, R_i = g^a_i, v_i = a_i k_i)
        r = sss.dsa_calculate_r(rs, p, q)
        assert r == pow(g, pow(kk, -1, q), p) % q                 # the property dsa_test.go:286-319 checks: r = g^(k^-1) mod p mod q
        xs.append([x for x, _, _ in rs]); ris.append([ri for _, ri, _ in rs]); vis.append([vi for _, _, vi in rs])
        exp.append(r)
    got, st = engine.dsa_calculate_r_batch(p, q, np.array(xs, np.int32), ris, vis)
    assert not st.any() and got == exp


def test_threshold_rsa_combine(engine):
    """crypto/threshold/rsa: the signature is the product of the partial signatures m^{d_i} mod N over an
    additive split of d (rsa.go:318-329; TestCombine rsa_test.go:165-206 checks it equals the plain
    PKCS#1 v1.5 signature).  Product on the GPU (K5), result verified by K1."""
    import hashlib
    from bftkv_b200 import workload
    rng = random.Random(3)
    keys = workload.load_keys(3)
    vals, exp, kidx, digs = [], [], [], []
    for i in range(48):
        key = keys[i % 3]
        d = hashlib.sha256(b"tbs%d" % i).digest()
        em = workload.em_for_digest(d)
        parts = [rng.randrange(key["d"]) for _ in range(6)]
        parts.append(key["d"] - sum(parts))                       # additive split; the last share may be negative
        psigs = [pow(em, p, key["n"]) if p >= 0 else pow(pow(em, -p, key["n"]), -1, key["n"]) for p in parts]   # rsa.go:140-178
        vals.append(psigs)
        exp.append(pow(em, key["d"], key["n"]).to_bytes(256, "big"))
        kidx.append(i % 3); digs.append(d)
    for n_, rows in [(keys[j]["n"], [v for i, v in enumerate(vals) if i % 3 == j]) for j in range(3)]:
        got = engine.modprod_batch(n_, rows)
        assert got == [e for i, e in enumerate(exp) if keys[i % 3]["n"] == n_]
    first = engine.register_rsa_keys([k["n"] for k in keys], [65537] * 3)
    st = engine.rsa_verify_batch(np.array(kidx, np.uint32) + first, np.frombuffer(b"".join(exp), np.uint8).reshape(-1, 256).copy(),
                                 np.frombuffer(b"".join(digs), np.uint8).reshape(-1, 32).copy())
    assert not st.any()
