"""GPU parity for the RSA key-size classes beyond 2048 bits (k = 128/192/384/512 bytes): flat K1 API
against the Python oracle's rsa.VerifyPKCS1v15 restatement, and GnuPG-made RSA-3072/4096 signatures
through the packer (Signature.Verify)."""
import json
import os

import numpy as np
import pytest

from bftkv_b200 import Engine, workload
from bftkv_b200.crypto_gpu import ErrInvalidSignature, Keyring, Signature
from oracle import pgp_oracle as pgp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def em(digest: bytes, k: int) -> int:
    t = workload.SHA256_PREFIX + digest
    return int.from_bytes(b"\x00\x01" + b"\xff" * (k - len(t) - 3) + b"\x00" + t, "big")


@pytest.mark.parametrize("bits", [1024, 1536, 3072, 4096, 2048, 1529])
def test_flat_api_key_classes(built, bits):
    from cryptography.hazmat.primitives.asymmetric import rsa
    keys = []
    for _ in range(3):
        pn = rsa.generate_private_key(65537, bits).private_numbers()
        keys.append({"n": pn.public_numbers.n, "e": 65537, "d": pn.d})
    assert keys[0]["n"].bit_length() == bits
    k = (bits + 7) // 8
    kb = {128: 128, 192: 192, 256: 256, 384: 384, 512: 512}.get(k)
    e = Engine(0)
    if kb is None:                       # e.g. 3065 bits -> k = 384 is a class; 1017 -> k = 128
        pytest.skip("not a size class")
    e.register_rsa_keys([x["n"] for x in keys], [65537] * 3)
    rng = np.random.default_rng(bits)
    N = 200
    kidx = rng.integers(0, 3, N).astype(np.uint32)
    dig = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    sig = np.zeros((N, kb), np.uint8)
    for i in range(N):
        key = keys[kidx[i]]
        s = pow(em(dig[i].tobytes(), kb), key["d"], key["n"])
        sig[i] = np.frombuffer(s.to_bytes(kb, "big"), np.uint8)
    flip = rng.random(N) < 0.3
    for i in np.nonzero(flip)[0]:
        sig[i, int(rng.integers(1, kb))] ^= 0x08
    extra = {0: 0, 1: 1, 2: keys[kidx[2]]["n"] - 1, 3: keys[kidx[3]]["n"]}           # edge values of s
    for i, v in extra.items():
        sig[i] = np.frombuffer(int(v).to_bytes(kb, "big"), np.uint8)
    got = e.rsa_verify_batch(kidx, sig, dig, key_bytes=kb)
    for i in range(N):
        key = keys[kidx[i]]
        ok = pgp.rsa_verify_pkcs1v15(key["n"], 65537, 8, dig[i].tobytes(), sig[i].tobytes())
        assert (got[i] == 0) == ok, (bits, i)
    assert (got == 0).sum() > 100
    # a key of another class in the same batch: len(sig) != k  ->  rejected
    other = rsa.generate_private_key(65537, 2048 if kb != 256 else 3072).private_numbers().public_numbers.n
    first = e.register_rsa_keys([other], [65537])
    st = e.rsa_verify_batch(np.array([first], np.uint32), sig[:1].copy(), dig[:1].copy(), key_bytes=kb)
    assert st.tolist() == [1]
    e.close()


def test_gpg_rsa3072_rsa4096_through_packer(built):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_sizes.json")))
    e = Engine(0)
    kr = Keyring(e)
    ents = []
    for k in g["keys"].values():
        blob = bytes.fromhex(k["pub"])
        kr.register(blob)
        ents += pgp.read_entities(blob)
    sig = Signature(kr)
    tbs = [bytes.fromhex(c["tbs"]) for c in g["cases"]]
    sd = [bytes.fromhex(c["sig"]) for c in g["cases"]]
    got = sig.verify_batch(tbs + [t + b"x" for t in tbs], sd * 2)
    ref = [pgp.signature_verify(ents, t, s) for t, s in zip(tbs + [t + b"x" for t in tbs], sd * 2)]
    assert got == ref
    assert got[:len(tbs)] == [None] * len(tbs) and set(got[len(tbs):]) == {ErrInvalidSignature}
    # mixed classes in ONE collective-style stream: 3072 + 4096 + 2048 signatures over the same message
    same = [c for c in g["cases"] if c["tbs"] == g["cases"][0]["tbs"] and c["hash"] == "SHA256"]
    assert len(same) == 3
    stream = b"".join(bytes.fromhex(c["sig"]) for c in same)
    assert sig.verify(bytes.fromhex(same[0]["tbs"]), stream) is None
    assert len(sig.signers(stream)) == 3
    kr.close()
    e.close()


def _rsa_key(bits):
    """An RSA key whose modulus has exactly `bits` bits (bits need not be a multiple of 8)."""
    import random
    from cryptography.hazmat.primitives.asymmetric import rsa
    if bits % 8 == 0 and bits >= 1024:
        pn = rsa.generate_private_key(65537, bits).private_numbers()
        return {"n": pn.public_numbers.n, "e": 65537, "d": pn.d, "p": pn.p, "q": pn.q}
    rnd = random.Random(bits)

    def prime(b):
        while True:
            c = rnd.getrandbits(b) | (1 << (b - 1)) | 1
            if c % 65537 != 1 and all(pow(a, c - 1, c) == 1 for a in (2, 3, 5, 7, 11, 13)):
                return c
    while True:
        p, q = prime(bits // 2), prime(bits - bits // 2)
        n = p * q
        if n.bit_length() == bits and p != q:
            return {"n": n, "e": 65537, "d": pow(65537, -1, (p - 1) * (q - 1)), "p": p, "q": q}


@pytest.mark.parametrize("bits", [2056, 2047, 2041, 1000, 3000, 1100, 4000])
def test_odd_key_sizes_travel_in_the_next_class(built, bits):
    """Moduli whose byte length is not a class size (RSA-2056: k = 257 -> the 384-byte class; RSA-1000: k = 125 ->
    128): EM is built for the key's own k and the signature is left-padded to the class stride, exactly what
    rsa.VerifyPKCS1v15 computes for such a key after x/crypto's padToKeySize."""
    keys = [_rsa_key(bits) for _ in range(2)]
    k = (bits + 7) // 8
    kb = next(c for c in (128, 192, 256, 384, 512) if c >= k)
    e = Engine(0)
    e.register_rsa_keys([x["n"] for x in keys], [65537] * 2)
    rng = np.random.default_rng(bits)
    N = 120
    kidx = rng.integers(0, 2, N).astype(np.uint32)
    dig = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    sig = np.zeros((N, kb), np.uint8)
    for i in range(N):
        key = keys[kidx[i]]
        s = pow(em(dig[i].tobytes(), k), key["d"], key["n"])
        sig[i] = np.frombuffer(s.to_bytes(kb, "big"), np.uint8)
    flip = rng.random(N) < 0.3
    for i in np.nonzero(flip)[0]:
        sig[i, int(rng.integers(kb - k + 1, kb))] ^= 0x20
    for i, v in {0: 0, 1: 1, 2: keys[kidx[2]]["n"] - 1, 3: keys[kidx[3]]["n"], 4: keys[kidx[4]]["n"] + 5}.items():
        sig[i] = np.frombuffer(int(v).to_bytes(kb, "big"), np.uint8)
    got = e.rsa_verify_batch(kidx, sig, dig, key_bytes=kb)
    for i in range(N):
        key = keys[kidx[i]]
        ok = pgp.rsa_verify_pkcs1v15(key["n"], 65537, 8, dig[i].tobytes(), sig[i, kb - k:].tobytes())
        assert (got[i] == 0) == ok, (bits, i)
    assert (got == 0).sum() > 50
    if kb > k:      # a signature LONGER than k bytes (non-zero above the key's length): len(sig) != k -> rejected
        j = int(np.nonzero(got == 0)[0][0])
        s2 = sig[j:j + 1].copy()
        n_j = keys[kidx[j]]["n"]
        v = int.from_bytes(s2[0].tobytes(), "big") + n_j * (1 << (8 * k - n_j.bit_length() + 1))   # same residue, needs k+1 bytes or more
        if v.bit_length() <= 8 * kb and v.bit_length() > 8 * k:
            s2[0] = np.frombuffer(v.to_bytes(kb, "big"), np.uint8)
            assert e.rsa_verify_batch(kidx[j:j + 1].copy(), s2, dig[j:j + 1].copy(), key_bytes=kb).tolist() == [1]
    e.close()


def test_odd_key_sizes_through_packer(built):
    """Signature.Verify with OpenPGP keys of 2056 / 2047 / 1000 bits next to a 2048-bit one (hand-built v4 key
    blocks and signature packets; the 2048-bit pair is the generator gpg accepted in the CPU suite)."""
    keys = [_rsa_key(b) for b in (2056, 2047, 1000)] + workload.load_keys(1)
    privs = [workload._private_key(k) for k in keys]
    e = Engine(0)
    kr = Keyring(e)
    blocks, kids = [], []
    for i, k in enumerate(keys):
        b, kid = workload.pgp_public_key_block(k, privs[i], b"odd-%d" % i)
        blocks.append(b); kids.append(kid)
        assert kr.register(b) == 1
    ents = pgp.read_entities(b"".join(blocks))
    sig = Signature(kr)
    tbs, sigs = [], []
    for j in range(40):
        i = j % 4
        m = b"message %d" % j
        p = bytearray(workload._v4_sig_packet(privs[i], kids[i], 0x00, m, 0x5F000000 + j))
        if j % 5 == 4:
            p[-3] ^= 0x40
        tbs.append(m); sigs.append(bytes(p))
    got = sig.verify_batch(tbs, sigs)
    n_ok = 0
    for j in range(40):
        ref = pgp.signature_verify(ents, tbs[j], sigs[j])
        assert (got[j] is None) == (ref is None), j
        n_ok += ref is None
    assert n_ok == 32
    kr.close()
    e.close()
