import sys, time, hashlib, numpy as np
sys.path.insert(0, "/root/repo")
from cryptography.hazmat.primitives.asymmetric import rsa, padding
from cryptography.hazmat.primitives.asymmetric.utils import Prehashed
from cryptography.hazmat.primitives import hashes
from bftkv_b200 import Engine
K, N = 4, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
keys = [rsa.generate_private_key(65537, 2048) for _ in range(K)]
ns = [k.public_key().public_numbers().n for k in keys]
rng = np.random.default_rng(1)
kidx = rng.integers(0, K, N).astype(np.uint32)
dig = np.zeros((N, 32), np.uint8); sig = np.zeros((N, 256), np.uint8)
t0 = time.time()
usig = {}
for i in range(N):
    d = hashlib.sha256(b"msg%d" % (i % 512)).digest()
    dig[i] = np.frombuffer(d, np.uint8)
    key = (int(kidx[i]), i % 512)
    if key not in usig:
        usig[key] = keys[kidx[i]].sign(d, padding.PKCS1v15(), Prehashed(hashes.SHA256()))
    sig[i] = np.frombuffer(usig[key], np.uint8)
print("gen", time.time() - t0)
exp = np.zeros(N, np.uint8)
bad = rng.random(N) < 0.1
for i in np.nonzero(bad)[0]:
    sig[i, rng.integers(0, 256)] ^= 1 << rng.integers(0, 8); exp[i] = 1
unk = rng.random(N) < 0.01
kidx[unk] = 1000; exp[unk] = 4
# check expectations with python pow
for i in range(min(N, 300)):
    if kidx[i] >= K: continue
    m = pow(int.from_bytes(sig[i].tobytes(), "big"), 65537, ns[kidx[i]])
    em = b"\x00\x01" + b"\xff" * 202 + b"\x00" + bytes.fromhex("3031300d060960864801650304020105000420") + dig[i].tobytes()
    assert (m == int.from_bytes(em, "big")) == (exp[i] == 0), i
import os
for T in (4, 8):
    os.environ["BFTQ_RSA_T"] = str(T)
    e = Engine(0)
    e.register_rsa_keys(ns, [65537] * K)
    st = e.rsa_verify_batch(kidx, sig, dig)
    print("T", T, "mismatches", int((st != exp).sum()), "ok", int((st == 0).sum()), "bad", int((st == 1).sum()), "unk", int((st == 4).sum()))
    import torch
    t0 = time.time(); 
    for _ in range(3): st = e.rsa_verify_batch(kidx, sig, dig)
    dt = (time.time() - t0) / 3
    print("  e2e verifies/s", N / dt)
    print("  int peak T mac/s", e.measure_int_peak() / 1e12)
    e.close()
