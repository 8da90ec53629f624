"""P-256 arithmetic (row a12, ECDSA CalculateR).  CPU: the kernels' own __host__ __device__ code
compiled for the host against big-int EC math; GPU: ecdsa.CalculateR through the C ABI against the
oracle's restatement of crypto/threshold/ecdsa/ecdsa.go:36-59 and the property ecdsa_test.go:36-92
checks (sum lambda_i (y_i G) == f(0) G)."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import sss_oracle as sss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, G = sss.P256_N, sss.P256_G


@pytest.fixture(scope="module")
def host():
    so = os.path.join(ROOT, "tests", "harness", "libp256host.so")
    src = os.path.join(ROOT, "tests", "harness", "p256_host.cpp")
    hdr = os.path.join(ROOT, "bftkv_b200", "csrc", "p256.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def test_host_scalar_mul_and_add(host):
    rng = random.Random(4)
    out = ctypes.create_string_buffer(65)
    for t in range(40):
        k = [1, 2, 3, N - 1, N + 1, 5][t] if t < 6 else rng.randrange(1, N)
        P = sss.p256_mul(rng.randrange(1, N), G)
        rc = host.p256_mul_host(sss.p256_marshal(P), (k % 2 ** 256).to_bytes(32, "big"), out)
        exp = sss.p256_mul(k, P)
        assert (rc == 2 and exp is None) or (rc == 1 and out.raw == sss.p256_marshal(exp))
        Q = sss.p256_mul(rng.randrange(1, N), G) if t % 5 else (P if t % 10 else (P[0], sss.P256_P - P[1]))
        rc = host.p256_add_host(sss.p256_marshal(P), sss.p256_marshal(Q), out)
        e2 = sss.p256_add(P, Q)
        assert (rc == 2 and e2 is None) or (rc == 1 and out.raw == sss.p256_marshal(e2))
    assert host.p256_mul_host(sss.p256_marshal(G), N.to_bytes(32, "big"), out) == 2          # N * G = infinity
    bad = bytearray(sss.p256_marshal(G)); bad[40] ^= 1
    assert host.p256_mul_host(bytes(bad), (5).to_bytes(32, "big"), out) == 0                  # off the curve


@pytest.mark.gpu
def test_ecdsa_calculate_r_gpu(engine):
    rng = random.Random(21)
    n, t2 = 15, 10                                           # BASELINE config 4: t = 5 -> 2t = 10 of n = 15
    xs, ris, vis, exp = [], [], [], []
    for _ in range(24):
        a, kk = rng.randrange(1, N), rng.randrange(1, N)
        sa = sss.distribute(a, [rng.randrange(N) for _ in range(t2 // 2 - 1)], n, N)
        sk = sss.distribute(kk, [rng.randrange(N) for _ in range(t2 // 2 - 1)], n, N)
        idx = rng.sample(range(n), t2)
        rs = [(sa[i][0], sss.p256_mul(sa[i][1], G), (sa[i][1] * sk[i][1]) % N) for i in idx]
        # ecdsa_test.go:36-92 property: sum lambda_i (a_i G) == a G
        acc = None
        for x, ri, _ in rs:
            acc = sss.p256_add(acc, sss.p256_mul(sss.lagrange(x, [q[0] for q in rs], N), ri))
        assert acc == sss.p256_mul(a, G)
        r = sss.ecdsa_calculate_r(rs)
        assert r == sss.p256_mul(pow(kk, -1, N), G)[0] % N   # r = (k^-1 G).x mod N
        xs.append([x for x, _, _ in rs]); ris.append([sss.p256_marshal(ri) for _, ri, _ in rs]); vis.append([vi for _, _, vi in rs])
        exp.append(r)
    got, st = engine.ecdsa_p256_calculate_r_batch(np.array(xs, np.int32), ris, vis)
    assert not st.any() and got == exp
    # a point that is not on the curve -> MALFORMED (the reference dereferences the nil Unmarshal returns)
    badrow = list(ris[0]); b = bytearray(badrow[3]); b[50] ^= 4; badrow[3] = bytes(b)
    got2, st2 = engine.ecdsa_p256_calculate_r_batch(np.array(xs[:1], np.int32), [badrow], vis[:1])
    assert st2.tolist() == [3]
