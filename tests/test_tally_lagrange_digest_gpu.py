"""GPU parity suite for K2 (wotqs tally), K3 (Lagrange combine), K4 (OpenPGP v4 digest) and the
fused verify+tally call, through the C ABI, against the oracle."""
import hashlib
import random

import numpy as np
import pytest

from bftkv_b200 import workload
from bftkv_b200.engine import NO_WINNER, ST_MISSING
from oracle import c_oracle, pgp_oracle as pgp, sss_oracle as sss, wotqs_oracle as wq
from oracle.wotqs_oracle import Node

pytestmark = pytest.mark.gpu


def random_ops(rng, n_ops, key_pool, max_r, status_pool):
    off, idx, st = [0], [], []
    for _ in range(n_ops):
        r = int(rng.integers(0, max_r + 1))
        idx += [int(x) for x in rng.choice(key_pool, r)]                       # duplicates allowed
        st += [int(x) for x in rng.choice(status_pool, r)]
        off.append(len(idx))
    return np.array(off, np.uint32), np.array(idx or [0], np.uint32)[:len(idx)], np.array(st or [0], np.uint8)[:len(st)]


QUORUMS = [
    [],                                                                          # empty quorum (wotqs.go:178: Reject = true)
    [(1, 4, 3, 3, [0, 1, 2, 3])],                                               # config 1: 4-node clique, AUTH
    [(5, 16, 6, 11, list(range(16)))],                                          # config 3: n=16 READ
    [(10, 31, 11, 21, list(range(31)))],                                        # config 5: n=31 READ
    [(3, 10, 7, 7, list(range(10))), (3, 10, 7, 0, list(range(20, 30))), (0, 0, 0, 0, [40, 41])],   # two cliques + WRITE complement
    [(1, 4, 2, 3, [5, 6, 7, 8]), (1, 4, 2, 3, [8, 9, 10, 11])],                 # overlapping member
]


@pytest.mark.parametrize("qi", range(len(QUORUMS)))
def test_tally_matches_oracle(engine, qi):
    qcs = QUORUMS[qi]
    rng = np.random.default_rng(100 + qi)
    off, idx, st = random_ops(rng, 3000, list(range(0, 45)), 40, [0, 0, 0, 0, 1, 4, 6])
    q = engine.quorum_create(qcs)
    got = engine.tally_batch(q, off, idx, st)
    ref = c_oracle.tally_batch(qcs, off, idx.astype(np.uint64) if len(idx) else [0], st if len(st) else [0])
    assert np.array_equal(got, ref)
    # and against the Python restatement of wotqs.go on a sample
    quorum = wq.Quorum([wq.QC([Node(m) for m in mem], f, mn, th, sf) for f, mn, th, sf, mem in qcs])
    for i in range(0, 3000, 37):
        ok = [Node(int(idx[p])) for p in range(off[i], off[i + 1]) if st[p] == 0]
        bad = [Node(int(idx[p])) for p in range(off[i], off[i + 1]) if st[p] != 0]
        exp = quorum.is_quorum(ok) | (quorum.is_threshold(ok) << 1) | (quorum.is_sufficient(ok) << 2) | (quorum.reject(bad) << 3)
        assert got[i] == exp
    engine.quorum_destroy(q)


def test_read_tally_matches_max_timestamped_value(engine):
    """protocol/client.go:189-205 incl. the quirk that only the max-t bucket set is inspected."""
    qcs = [(5, 16, 6, 11, list(range(16)))]
    quorum = wq.Quorum([wq.QC([Node(m) for m in range(16)], 5, 16, 6, 11)])
    rng = np.random.default_rng(77)
    n_ops = 4000
    off, idx, st, ts, vid = [0], [], [], [], []
    for _ in range(n_ops):
        r = int(rng.integers(0, 33))
        for j in range(r):
            idx.append(int(rng.integers(0, 20)))
            st.append(int(rng.choice([0, 1, 6], p=[0.9, 0.05, 0.05])))
            ts.append(int(rng.choice([5, 4, 9, 2 ** 63 - 1], p=[0.93, 0.05, 0.015, 0.005])))
            vid.append(int(rng.choice([0, 1, 2], p=[0.85, 0.1, 0.05])))
        off.append(len(idx))
    off, idx, st = np.array(off, np.uint32), np.array(idx, np.uint32), np.array(st, np.uint8)
    ts, vid = np.array(ts, np.uint64), np.array(vid, np.uint32)
    q = engine.quorum_create(qcs)
    win, bits = engine.read_tally_batch(q, off, idx, st, ts, vid)
    n_win = 0
    for i in range(n_ops):
        m = {}
        first = {}
        for p in range(off[i], off[i + 1]):
            if st[p] != 0:
                continue
            m.setdefault(int(ts[p]), {}).setdefault(int(vid[p]), []).append(Node(int(idx[p])))
            first.setdefault((int(ts[p]), int(vid[p])), p - off[i])
        r = wq.max_timestamped_value(m, quorum)
        if r is None:
            assert win[i] == NO_WINNER, i
        else:
            # the reference may return ANY qualifying value of the max-t bucket (Go map order);
            # the kernel returns the first in responder order
            maxt = max(m)
            qualifying = {v for v, l in m[maxt].items() if quorum.is_threshold(l)}
            assert win[i] != NO_WINNER
            p = off[i] + win[i]
            assert int(ts[p]) == maxt and int(vid[p]) in qualifying and st[p] == 0
            assert win[i] == min(first[(maxt, v)] for v in qualifying)
            n_win += 1
        bad = [Node(int(idx[p])) for p in range(off[i], off[i + 1]) if st[p] != 0]
        assert bool(bits[i] & 8) == quorum.reject(bad)
        assert bool(bits[i] & 2) == (r is not None)
    assert 200 < n_win < n_ops
    engine.quorum_destroy(q)


def test_verify_tally_fused_config3_scaled(built):
    """BASELINE config 3 scaled to 2048 ops x 16 replicas: verify + tally in one call."""
    from bftkv_b200 import Engine
    R, M = 16, 2048
    pool = workload.make_verify_batch(4096, n_keys=R, corrupt_rate=0.0, unknown_rate=0.0)
    w = workload.make_read_ops(pool, M, R, seed=0xBF7C0004)
    e = Engine(0)
    e.register_rsa_keys([k["n"] for k in pool["keys"]], [k["e"] for k in pool["keys"]])
    qcs = [(5, 16, 6, 11, list(range(16)))]
    q = e.quorum_create(qcs)
    st, bits, win = e.verify_tally_batch(q, w["op_off"], w["key_idx"], w["sig"], w["digest"], pre_status=w["pre_status"],
                                         ts=w["ts"], value_id=w["value_id"])
    ns, es = [k["n"] for k in pool["keys"]], [k["e"] for k in pool["keys"]]
    ref_st = c_oracle.rsa_verify_batch(ns, es, w["key_idx"], w["sig"], w["digest"], threads=8)
    ref_st[w["pre_status"] != 0] = w["pre_status"][w["pre_status"] != 0]
    assert np.array_equal(st, ref_st)
    assert np.array_equal(st, w["expect_status"])
    quorum = wq.Quorum([wq.QC([Node(m) for m in range(16)], 5, 16, 6, 11)])
    acc = 0
    for i in range(M):
        m = {}
        for p in range(w["op_off"][i], w["op_off"][i + 1]):
            if ref_st[p] == 0:
                m.setdefault(int(w["ts"][p]), {}).setdefault(int(w["value_id"][p]), []).append(Node(int(w["key_idx"][p])))
        r = wq.max_timestamped_value(m, quorum)
        assert (win[i] != NO_WINNER) == (r is not None), i
        acc += r is not None
    assert 0.5 * M < acc <= M
    # plain (non-read) fused form gives the four predicate bits
    st2, bits2, _ = e.verify_tally_batch(q, w["op_off"], w["key_idx"], w["sig"], w["digest"], pre_status=w["pre_status"])
    ref_bits = c_oracle.tally_batch(qcs, w["op_off"], w["key_idx"].astype(np.uint64), ref_st)
    assert np.array_equal(st2, ref_st) and np.array_equal(bits2, ref_bits)
    e.quorum_destroy(q)
    e.close()


P256_N = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
DSA_Q = 0xE950511EAB424B9A19A2AEB4E159B7844C589C4F          # 160-bit prime (DSA L1024N160 style)


@pytest.mark.parametrize("m,k,n", [(1237, 4, 6), (P256_N, 10, 15), (P256_N, 5, 15), (DSA_Q, 10, 15), (2 ** 127 - 1, 3, 5)])
def test_lagrange_matches_oracle(engine, m, k, n):
    if m == DSA_Q:
        assert pow(2, m - 1, m) == 1                                         # make sure the fixture really is prime
    rng = random.Random(k * 1000 + n)
    B = 700
    mlen = (m.bit_length() + 7) // 8
    x = np.empty((B, k), np.int32)
    y = np.empty((B, k, mlen), np.uint8)
    exp = []
    for b in range(B):
        secret = rng.randrange(m)
        shares = sss.distribute(secret, [rng.randrange(m) for _ in range(k - 1)], n, m)
        pick = rng.sample(shares, k)
        for j, (xx, yy) in enumerate(pick):
            x[b, j] = xx
            y[b, j] = np.frombuffer(yy.to_bytes(mlen, "big"), np.uint8)
        s = sss.calculate_secret(pick, m)
        assert s == secret and sss.calculate_s(pick, m) == secret
        exp.append(s.to_bytes(mlen, "big"))
    out, st = engine.lagrange_combine_batch(m, x, y)
    assert not st.any()
    assert [bytes(o) for o in out] == exp


def test_lagrange_kats_and_edges(engine, golden):
    a = golden["sss"]["auth_test"]                                           # crypto/auth/auth_test.go:121-155: S = 1234
    picked = [tuple(s) for s in a["shares"] if s[0] in a["xs"]]
    x = np.array([[p[0] for p in picked]], np.int32)
    y = np.array([[list(p[1].to_bytes(2, "big")) for p in picked]], np.uint8)
    out, st = engine.lagrange_combine_batch(a["q"], x, y)
    assert st.tolist() == [0] and int.from_bytes(bytes(out[0]), "big") == a["S"] == 1234
    # crypto/sss/sss_test.go:15-75: 2048-bit modulus, "secret", n=10, k=7
    m = int("b0a67d9f5cebc0ffe81690e7b2670ab05f9fa4c2e73639f660c0408a2d9a4a8b454a9893fd7d4e8fa399cfc9c9ba05b080f903e33bcdcbef"
            "aed40915e51d46f58d1a5bd204db20fa3fe9db71f0b8e0aa87b5771406f25fad59e7f10fe5255644758872ea2dec1f6dcd11be905de59a04"
            "4f6c2ea3982b2235acc9021a196fc4ce0b19f6b312ee9cfc5997dc5f7ce2f386131294a56ba93a41a3b60e27e03956039f51ae73b89c795c"
            "5ae7d841e9b455c37341c052404e8fe9fe4f0d52bc162a41f1eeb9ef292c66a9d6a619aa548807eb1187ee22bd62e20e26c3c08c22ecef12"
            "d3b2304a010ed1f50a68e0261afe1a0bdddf7ab8a61774d3af3f1cce2b95dad3", 16)
    rng = random.Random(1)
    secret = int.from_bytes(b"secret", "big")
    shares = sss.distribute(secret, [rng.randrange(m) for _ in range(6)], 10, m)
    B = 40
    x = np.empty((B, 7), np.int32)
    y = np.empty((B, 7, 256), np.uint8)
    for b in range(B):
        for j, (xx, yy) in enumerate(rng.sample(shares, 7)):
            x[b, j] = xx
            y[b, j] = np.frombuffer(yy.to_bytes(256, "big"), np.uint8)
    out, st = engine.lagrange_combine_batch(m, x, y)
    assert not st.any()
    assert all(bytes(o).lstrip(b"\0") == b"secret" for o in out)
    # edges: non-invertible difference (m = 15, x difference 5) -> MALFORMED; duplicate x skipped like sss.go:100-102
    out, st = engine.lagrange_combine_batch(15, np.array([[1, 6], [1, 3]], np.int32), np.array([[[2], [3]], [[2], [3]]], np.uint8))
    assert st.tolist() == [3, 0]
    assert int(out[1][0]) == sss.calculate_secret([(1, 2), (3, 3)], 15)
    dup = [(2, 10), (2, 10), (5, 77)]
    out, st = engine.lagrange_combine_batch(1237, np.array([[p[0] for p in dup]], np.int32),
                                            np.array([[list(p[1].to_bytes(2, "big")) for p in dup]], np.uint8))
    assert st.tolist() == [0] and int.from_bytes(bytes(out[0]), "big") == sss.calculate_secret(dup, 1237)


def test_pgp_digest_matches_hashlib_and_gpg(engine, golden):
    rng = random.Random(3)
    datas = [bytes(rng.randrange(256) for _ in range(ln)) for ln in [0, 1, 31, 46, 47, 48, 55, 56, 63, 64, 65, 119, 120, 300, 1000]]
    sufs, didx = [], []
    for i in range(400):
        didx.append(rng.randrange(len(datas)))
        sufs.append(bytes(rng.randrange(256) for _ in range(rng.choice([0, 6, 12, 23, 41, 64, 100]))))
    got = engine.pgp_digest_batch(datas, sufs, didx)
    for i in range(400):
        assert bytes(got[i]) == hashlib.sha256(datas[didx[i]] + sufs[i]).digest()
    # every device hash against hashlib
    for hid, name in [(1, "md5"), (2, "sha1"), (8, "sha256"), (9, "sha384"), (10, "sha512"), (11, "sha224")]:
        got = engine.pgp_digest_batch(datas, sufs, didx, hash_alg=hid)
        for i in range(0, 400, 3):
            assert bytes(got[i]) == hashlib.new(name, datas[didx[i]] + sufs[i]).digest(), (name, i)
    # GnuPG-made SHA-256 signatures: the digest's first two bytes must equal the packet's hash tag
    datas, sufs, tags, ref = [], [], [], []
    for c in golden["cases"]:
        if c["hash"] != "SHA256":
            continue
        tag, body = pgp.read_packet(pgp.Reader(bytes.fromhex(c["sig"])))
        s = pgp.parse_signature(body)
        if s.sig_type != 0:
            continue                                  # text-mode canonicalisation is the packer's job
        datas.append(bytes.fromhex(c["tbs"])); sufs.append(s.hash_suffix); tags.append(s.hash_tag)
        ref.append(pgp.signature_digest(bytes.fromhex(c["tbs"]), s))
    got = engine.pgp_digest_batch(datas, sufs)
    assert len(datas) > 20
    for i in range(len(datas)):
        assert bytes(got[i]) == ref[i] and bytes(got[i][:2]) == tags[i]
