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
