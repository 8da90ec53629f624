"""ECDSA (P-256) and DSA signature verification — OpenPGP public-key algorithms 19 and 17, the two non-RSA
arms of packet.PublicKey.VerifySignature reached from crypto/pgp/crypto_pgp.go:319-344.

CPU part: the oracle's restatement of Go crypto/ecdsa.Verify / crypto/dsa.Verify is pinned on GnuPG-made keys
and signatures (tests/golden/golden_algos.json) and on OpenSSL-made ones; the P-256 verification core that the
kernel runs is compiled for the host and checked against OpenSSL.  GPU part: the flat entry points and the
packer against the oracle."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

from oracle import pgp_oracle as pgp, sss_oracle as so

HERE = os.path.dirname(__file__)
N = so.P256_N


@pytest.fixture(scope="module")
def algos():
    return json.load(open(os.path.join(HERE, "golden", "golden_algos.json")))


@pytest.fixture(scope="module")
def ring(algos):
    ents = []
    for k in algos["keys"].values():
        ents += pgp.read_entities(bytes.fromhex(k["pub"]))
    return ents


def _openssl_ecdsa(n, seed):
    from cryptography.hazmat.primitives import hashes, serialization
    from cryptography.hazmat.primitives.asymmetric import ec, utils
    rng = random.Random(seed)
    out = []
    keys = [ec.generate_private_key(ec.SECP256R1()) for _ in range(4)]
    for i in range(n):
        sk = keys[i % 4]
        nums = sk.public_key().public_numbers()
        halg, hl = [(hashes.SHA256(), 32), (hashes.SHA512(), 64), (hashes.SHA1(), 20), (hashes.SHA384(), 48)][i % 4]
        d = bytes(rng.randrange(256) for _ in range(hl))
        r, s = utils.decode_dss_signature(sk.sign(d, ec.ECDSA(utils.Prehashed(halg))))
        out.append(((nums.x, nums.y), d, r, s))
    return out


def _mutations(q, d, r, s, rng):
    """(key, digest, r, s, expected-by-construction or None)"""
    d2 = bytearray(d); d2[rng.randrange(min(len(d), 32))] ^= 1 << rng.randrange(8)
    return [(q, d, r, s, True), (q, bytes(d2), r, s, False), (q, d, r, N - s, True), (q, d, (r + 1) % N or 1, s, False),
            (q, d, 0, s, False), (q, d, r, 0, False), (q, d, N, s, False), (q, d, r, N, False), (q, d, r + N if r + N < 2 ** 256 else r, s, None)]


def test_oracle_pinned_on_gpg_ecdsa_dsa(algos, ring):
    seen = set()
    for c in algos["cases"]:
        tbs, sig = bytes.fromhex(c["tbs"]), bytes.fromhex(c["sig"])
        assert c["gpg_ok"] and pgp.signature_verify(ring, tbs, sig) is None
        assert pgp.signature_verify(ring, tbs + b"x", sig) is not None
        seen.add((algos["keys"][c["signer"]]["algo"], c["hash"]))
    assert {a for a, _ in seen} == {17, 19} and len(seen) >= 7


def test_oracle_ecdsa_matches_openssl():
    rng = random.Random(5)
    for q, d, r, s in _openssl_ecdsa(24, 1):
        for (qq, dd, rr, ss, want) in _mutations(q, d, r, s, rng):
            got = pgp.ecdsa_p256_verify(qq, dd, rr, ss)
            if want is not None:
                assert got == want
    assert not pgp.ecdsa_p256_verify((1, 1), b"\x01" * 32, 5, 7)         # key off the curve


def test_oracle_dsa_matches_openssl():
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import dsa, utils
    for bits, halg, hl in [(1024, hashes.SHA1(), 20), (2048, hashes.SHA256(), 32), (2048, hashes.SHA512(), 64)]:
        sk = dsa.generate_private_key(bits)
        pn = sk.public_key().public_numbers()
        p, q, g, y = pn.parameter_numbers.p, pn.parameter_numbers.q, pn.parameter_numbers.g, pn.y
        for i in range(4):
            d = os.urandom(hl)
            r, s = utils.decode_dss_signature(sk.sign(d, utils.Prehashed(halg)))
            # OpenSSL signs over the leftmost |q| bits, which is what x/crypto + dsa.Verify check
            assert pgp.dsa_verify(p, q, g, y, d, r, s)
            assert not pgp.dsa_verify(p, q, g, y, d[:-1] + bytes([d[-1] ^ 1]) if hl * 8 <= q.bit_length() else bytes([d[0] ^ 1]) + d[1:], r, s)
            assert not pgp.dsa_verify(p, q, g, y, d, r, q) and not pgp.dsa_verify(p, q, g, y, d, 0, s)


def test_p256_verify_core_on_host(tmp_path):
    """The __host__ __device__ verification core of csrc/p256.cuh, compiled with g++."""
    import subprocess
    so_path = tmp_path / "libp256host.so"
    root = os.path.dirname(HERE)
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", str(so_path), os.path.join(HERE, "harness", "p256_host.cpp"),
                    "-I", os.path.join(root, "bftkv_b200", "csrc")], check=True)
    lib = ctypes.CDLL(str(so_path))
    rng = random.Random(9)
    n = 0
    for q, d, r, s in _openssl_ecdsa(16, 2):
        pub = b"\x04" + q[0].to_bytes(32, "big") + q[1].to_bytes(32, "big")
        for (qq, dd, rr, ss, want) in _mutations(q, d, r, s, rng):
            if rr >= 2 ** 256 or ss >= 2 ** 256:
                continue
            got = lib.p256_ecdsa_verify_host(pub, rr.to_bytes(32, "big"), ss.to_bytes(32, "big"), dd, len(dd))
            assert got == int(pgp.ecdsa_p256_verify(qq, dd, rr, ss))
            n += 1
    assert n > 100


# ---- GPU ------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def engine(built):
    from bftkv_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_ecdsa_flat_api_vs_oracle(engine):
    rng = random.Random(11)
    items = []
    for q, d, r, s in _openssl_ecdsa(64, 3):
        if len(d) != 32:
            continue
        items += [(qq, dd, rr % 2 ** 256, ss) for (qq, dd, rr, ss, _) in _mutations(q, d, r, s, rng)]
    keys = sorted({q for q, _, _, _ in items})
    pub = np.frombuffer(b"".join(x.to_bytes(32, "big") + y.to_bytes(32, "big") for x, y in keys), np.uint8).reshape(-1, 64)
    bad_key = np.frombuffer((1).to_bytes(32, "big") + (1).to_bytes(32, "big"), np.uint8).reshape(1, 64)
    pub = np.concatenate([pub, bad_key])
    kidx = np.array([keys.index(q) for q, _, _, _ in items] + [len(keys), len(keys) + 1], np.uint32)
    rb = np.frombuffer(b"".join(r.to_bytes(32, "big") for _, _, r, _ in items) + b"\x01" * 64, np.uint8).reshape(-1, 32)
    sb = np.frombuffer(b"".join(s.to_bytes(32, "big") for _, _, _, s in items) + b"\x01" * 64, np.uint8).reshape(-1, 32)
    dg = np.frombuffer(b"".join(d for _, d, _, _ in items) + b"\x02" * 64, np.uint8).reshape(-1, 32)
    st = engine.ecdsa_p256_verify_batch(pub, kidx, rb, sb, dg)
    ref = [0 if pgp.ecdsa_p256_verify(q, d, r, s) else 1 for q, d, r, s in items] + [3, 4]
    assert st.tolist() == ref
    assert 0 < sum(x == 0 for x in ref) < len(ref)


@pytest.mark.gpu
@pytest.mark.parametrize("hl", [20, 48, 64])
def test_ecdsa_flat_api_digest_lengths(engine, hl):
    items = [(q, d, r, s) for q, d, r, s in _openssl_ecdsa(32, 4) if len(d) == hl]
    assert items
    pub = np.frombuffer(b"".join(q[0].to_bytes(32, "big") + q[1].to_bytes(32, "big") for q, _, _, _ in items), np.uint8).reshape(-1, 64)
    rb = np.frombuffer(b"".join(r.to_bytes(32, "big") for _, _, r, _ in items), np.uint8).reshape(-1, 32)
    sb = np.frombuffer(b"".join(s.to_bytes(32, "big") for _, _, _, s in items), np.uint8).reshape(-1, 32)
    dg = np.frombuffer(b"".join(d for _, d, _, _ in items), np.uint8).reshape(-1, hl).copy()
    kidx = np.arange(len(items), dtype=np.uint32)
    assert engine.ecdsa_p256_verify_batch(pub, kidx, rb, sb, dg).tolist() == [0] * len(items)
    dg[:, min(hl, 32) - 1] ^= 1                       # last byte that still counts
    assert engine.ecdsa_p256_verify_batch(pub, kidx, rb, sb, dg).tolist() == [1] * len(items)
    if hl > 32:
        dg[:, min(hl, 32) - 1] ^= 1
        dg[:, 32:] ^= 0x55                            # bytes beyond the order size are ignored (hashToInt)
        assert engine.ecdsa_p256_verify_batch(pub, kidx, rb, sb, dg).tolist() == [0] * len(items)


@pytest.mark.gpu
def test_dsa_flat_api_vs_oracle(engine, algos, ring):
    rng = random.Random(13)
    for name in ("dsa1024", "dsa2048"):
        ent = [e for e in ring if e.primary_key.key_id == int(algos["keys"][name]["key_id"], 16)][0]
        p, q, g, y = ent.primary_key.dsa
        items = []
        for c in algos["cases"]:
            if c["signer"] != name:
                continue
            sig = pgp.parse_signature(pgp.read_packet(pgp.Reader(bytes.fromhex(c["sig"])))[1])
            d = pgp.signature_digest(bytes.fromhex(c["tbs"]), sig)
            r, s = sig.sig_r, sig.sig_s
            d2 = bytearray(d); d2[rng.randrange((q.bit_length() + 7) // 8)] ^= 4
            for (dd, rr, ss) in [(d, r, s), (bytes(d2), r, s), (d, s, r), (d, 0, s), (d, r, 0), (d, q, s), (d, r, q), (d, r, q - s), (d, q - r, s)]:
                items.append((len(d), dd, rr, ss))
        for dlen in sorted({i[0] for i in items}):
            sel = [i for i in items if i[0] == dlen]
            rb = np.frombuffer(b"".join(r.to_bytes(32, "big") for _, _, r, _ in sel), np.uint8).reshape(-1, 32)
            sb = np.frombuffer(b"".join(s.to_bytes(32, "big") for _, _, _, s in sel), np.uint8).reshape(-1, 32)
            dg = np.frombuffer(b"".join(d for _, d, _, _ in sel), np.uint8).reshape(-1, dlen)
            st = engine.dsa_verify_batch(p, q, g, y, rb, sb, dg)
            ref = [0 if pgp.dsa_verify(p, q, g, y, d, r, s) else 1 for _, d, r, s in sel]
            assert st.tolist() == ref
            assert 0 < sum(x == 0 for x in ref) < len(ref)


@pytest.mark.gpu
def test_dsa_flat_api_random_domain(engine):
    """A synthetic DSA domain (1024-bit p, 160-bit q) with many signatures, signed here with known x."""
    from cryptography.hazmat.primitives.asymmetric import dsa
    rng = random.Random(17)
    pn = dsa.generate_parameters(1024).parameter_numbers()
    p, q, g = pn.p, pn.q, pn.g
    x = rng.randrange(1, q); y = pow(g, x, p)
    rs, ss, ds, ref = [], [], [], []
    for i in range(300):
        d = bytes(rng.randrange(256) for _ in range(32))
        z = int.from_bytes(d[:20], "big")
        k = rng.randrange(1, q)
        r = pow(g, k, p) % q
        s = pow(k, -1, q) * (z + x * r) % q
        if i % 3 == 1:
            s = (s + 1) % q
        if i % 7 == 3:
            d = bytes([d[0] ^ 0x80]) + d[1:]
        rs.append(r.to_bytes(32, "big")); ss.append(s.to_bytes(32, "big")); ds.append(d)
        ref.append(0 if pgp.dsa_verify(p, q, g, y, d, r, s) else 1)
    st = engine.dsa_verify_batch(p, q, g, y, np.frombuffer(b"".join(rs), np.uint8).reshape(-1, 32),
                                 np.frombuffer(b"".join(ss), np.uint8).reshape(-1, 32), np.frombuffer(b"".join(ds), np.uint8).reshape(-1, 32))
    assert st.tolist() == ref and 100 < sum(x == 0 for x in ref) < 250


@pytest.mark.gpu
def test_packer_mixed_algorithms(engine, algos, ring, golden):
    """One keyring holding RSA, ECDSA and DSA keys; Signature.Verify over the GnuPG fixtures of all three.
    DSA-3072 is outside K5's modulus classes: the packer must say UNSUPPORTED (an error), never 'valid'."""
    from bftkv_b200.crypto_gpu import ErrNotBuilt, Keyring, Signature
    kr = Keyring(engine)
    ents = list(ring)
    u0 = engine.stats()["unsupported_items"]
    for k in algos["keys"].values():
        assert kr.register(bytes.fromhex(k["pub"])) == 1
    for n in ("a01", "a02"):
        blob = bytes.fromhex(golden["keys"][n]["pub"])
        kr.register(blob)
        ents += pgp.read_entities(blob)
    cases = list(algos["cases"]) + [c for c in golden["cases"] if c["signer"] in ("a01", "a02")][:8]
    tbs = [bytes.fromhex(c["tbs"]) for c in cases] + [bytes.fromhex(c["tbs"]) + b"x" for c in cases]
    sig = [bytes.fromhex(c["sig"]) for c in cases] * 2
    # a signature stream concatenating an ECDSA and a DSA signature over the same tbs: both must verify
    by = {}
    for c in algos["cases"]:
        by.setdefault(c["tbs"], {}).setdefault(c["signer"], c["sig"])
    t0 = [t for t, v in by.items() if "p256a" in v and "dsa2048" in v][0]
    tbs += [bytes.fromhex(t0)] * 2
    sig += [bytes.fromhex(by[t0]["p256a"] + by[t0]["dsa2048"]), bytes.fromhex(by[t0]["dsa1024"] + by[t0]["p256b"])]
    got = Signature(kr).verify_batch(tbs, sig)
    ref = [pgp.signature_verify(ents, t, s) for t, s in zip(tbs, sig)]
    big = {i + k for i, c in enumerate(cases) if c["signer"] == "dsa3072" for k in (0, len(cases))}       # genuine and tampered copies alike
    for i, (a, b) in enumerate(zip(got, ref)):
        if i in big:
            assert a == ErrNotBuilt                    # 3072-bit DSA domain is not built: reported as such (BFTQ_ERR_UNSUPPORTED), the shim re-runs it on crypto/pgp
        else:
            assert a == b, (i, a, b)
    assert sum(r is None for r in ref) == len(cases) + 2
    assert engine.stats()["unsupported_items"] - u0 == len(big) > 0        # counted for the operator (bftq_stats)
    kr.close()
