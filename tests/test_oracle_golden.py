"""CPU suite: the oracle against every golden vector / KAT the reference's tests hold for the
path and against GnuPG-made signatures.  No GPU, no libbftq compute calls."""
import hashlib
import os

import numpy as np
import pytest

from oracle import c_oracle, packet_oracle as pk, pgp_oracle as pgp, sss_oracle as sss, wotqs_oracle as wq
from oracle.wotqs_oracle import AUTH, CERT, PEER, READ, WRITE, Graph, Node, WotQS

RING = ["a01", "a02", "a03", "a04", "u01"]


def keyring(golden, names=RING):
    ents = []
    for n in names:
        ents += pgp.read_entities(bytes.fromhex(golden["keys"][n]["pub"]))
    return ents


def test_gpg_signatures_verify(golden, built):
    ring = keyring(golden)
    seen = set()
    for c in golden["cases"]:
        tbs, sig = bytes.fromhex(c["tbs"]), bytes.fromhex(c["sig"])
        err = pgp.signature_verify(ring, tbs, sig)
        if c["signer"] == "x99":
            assert err == pgp.ERR_INVALID_SIGNATURE          # unknown issuer
        else:
            assert err is None, c
        assert pgp.signature_verify(ring, tbs + b"\x00", sig) == pgp.ERR_INVALID_SIGNATURE
        seen.add(c["hash"])
    assert seen == {"SHA256", "SHA512", "SHA1"}


def test_verify_edge_cases(golden, built):
    ring = keyring(golden)
    good = [c for c in golden["cases"] if c["signer"] == "a01" and c["hash"] == "SHA256"][0]
    unk = [c for c in golden["cases"] if c["signer"] == "x99" and c["tbs"] == good["tbs"]][0]
    tbs, s1, sx = bytes.fromhex(good["tbs"]), bytes.fromhex(good["sig"]), bytes.fromhex(unk["sig"])
    assert pgp.signature_verify(ring, tbs, b"") == pgp.ERR_INVALID_SIGNATURE          # crypto_pgp.go:322,329
    assert pgp.signature_verify(ring, tbs, s1 + s1) is None                            # every packet must verify
    assert pgp.signature_verify(ring, tbs, sx + s1) is None                            # unknown issuer skipped
    assert pgp.signature_verify(ring, tbs, s1 + sx) == pgp.ERR_INVALID_SIGNATURE       # trailing unknown issuer
    bad = bytearray(s1); bad[-5] ^= 1
    assert pgp.signature_verify(ring, tbs, bytes(bad)) == pgp.ERR_INVALID_SIGNATURE
    assert pgp.signature_verify(ring, tbs, s1 + bytes(bad)) == pgp.ERR_INVALID_SIGNATURE
    assert pgp.signature_verify(ring, tbs, s1[:-7]) == pgp.ERR_INVALID_SIGNATURE       # truncated packet
    uid = bytes([0xB4, 3]) + b"abc"                                                   # user-id packet: "non signature packet found"
    assert pgp.signature_verify(ring, tbs, uid + s1) == pgp.ERR_INVALID_SIGNATURE
    assert pgp.signature_verify(ring, tbs, s1 + bytes([0xC0 | 60, 2, 1, 2])) == pgp.ERR_INVALID_SIGNATURE  # unknown tag then EOF
    assert pgp.signature_verify(ring, tbs, bytes([0xC0 | 60, 2, 1, 2]) + s1) is None   # unknown tag skipped
    # VerifyWithCertificate: keyring is the single issuer entity
    cert = bytes.fromhex(golden["keys"]["a01"]["pub"])
    assert pgp.signature_verify_with_certificate(cert, tbs, s1) is None
    cert2 = bytes.fromhex(golden["keys"]["a02"]["pub"])
    assert pgp.signature_verify_with_certificate(cert2, tbs, s1) == pgp.ERR_INVALID_SIGNATURE


def test_reference_rsa_kat(golden, built):
    """crypto/threshold/rsa/rsa_test.go:165-206 (TestCombine) on the reference-owned key."""
    k = golden["ref_rsa_kat"]
    n, sig, dig = int(k["n"], 16), bytes.fromhex(k["sig"]), bytes.fromhex(k["digest"])
    assert hashlib.sha256(sig).hexdigest() == "1d2cef7b44c674e771fdac4fb0f278c75e7b68fe40a835cd39ba5c23cd127998"
    assert hashlib.sha256(b"tbs").digest() == dig
    assert pgp.rsa_verify_pkcs1v15(n, k["e"], 8, dig, sig)
    # emsaEncode (rsa.go:356-378) == the PKCS#1 v1.5 encoding the signature opens to (TestEMSA)
    assert sss.emsa_encode(sss.SHA256_PREFIX, dig, n) == pow(int.from_bytes(sig, "big"), k["e"], n)
    # C oracle agrees
    st = c_oracle.rsa_verify_batch([n], [k["e"]], np.zeros(2, np.uint32),
                                   np.frombuffer(sig + sig[:-1] + bytes([sig[-1] ^ 1]), np.uint8).reshape(2, 256),
                                   np.frombuffer(dig + dig, np.uint8).reshape(2, 32))
    assert st.tolist() == [0, 1]
    # threshold-RSA combine = product of partial signatures (rsa.go:318-329): split d additively
    # (the key tree of rsa.go is an additive split) and multiply the partials.
    assert sss.rsa_combine([int.from_bytes(sig, "big")], n) == sig


def test_sha256_c_oracle(built):
    for m in [b"", b"abc", b"a" * 55, b"a" * 56, b"a" * 64, os.urandom(1000)]:
        assert c_oracle.sha256(m) == hashlib.sha256(m).digest()


def test_c_oracle_matches_python_pow(built):
    from bftkv_b200 import workload
    w = workload.make_verify_batch(600, n_keys=3, corrupt_rate=0.2, unknown_rate=0.05)
    ns, es = [k["n"] for k in w["keys"]], [k["e"] for k in w["keys"]]
    st = c_oracle.rsa_verify_batch(ns, es, w["key_idx"], w["sig"], w["digest"], threads=2)
    assert np.array_equal(st, w["expect"])
    for i in range(600):
        if w["key_idx"][i] >= 3:
            continue
        ok = pgp.rsa_verify_pkcs1v15(ns[w["key_idx"][i]], 65537, 8, w["digest"][i].tobytes(), w["sig"][i].tobytes())
        assert ok == (st[i] == 0)
    # s >= n: Go 1.13 semantics reduce mod n (accept); strict mode rejects
    i = int(np.nonzero(w["expect"] == 0)[0][0])
    n = ns[w["key_idx"][i]]
    s2 = int.from_bytes(w["sig"][i].tobytes(), "big") + n
    if s2 < 2 ** 2048:
        sig2 = np.frombuffer(s2.to_bytes(256, "big"), np.uint8).reshape(1, 256)
        kw = dict(moduli=ns, exps=es, key_idx=w["key_idx"][i:i + 1], sig=sig2, digest=w["digest"][i:i + 1])
        assert c_oracle.rsa_verify_batch(**kw).tolist() == [0]
        assert c_oracle.rsa_verify_batch(strict_range=True, **kw).tolist() == [1]


def test_lagrange_kats(golden):
    a = golden["sss"]["auth_test"]                       # crypto/auth/auth_test.go:121-155
    assert sss.distribute(a["poly"][0], a["poly"][1:], 6, a["q"]) == [tuple(s) for s in a["shares"]]
    assert [sss.lagrange(x, a["xs"], a["q"]) for x in a["xs"]] == a["lambda"]
    picked = [tuple(s) for s in a["shares"] if s[0] in a["xs"]]
    assert sss.calculate_secret(picked, a["q"]) == a["S"]
    # crypto/sss/sss_test.go:15-75: 2048-bit modulus, secret "secret", n=10, k=7, any 7 shares
    import random
    m = int("b0a67d9f5cebc0ffe81690e7b2670ab05f9fa4c2e73639f660c0408a2d9a4a8b454a9893fd7d4e8fa399cfc9c9ba05b080f903e33bcdcbef"
            "aed40915e51d46f58d1a5bd204db20fa3fe9db71f0b8e0aa87b5771406f25fad59e7f10fe5255644758872ea2dec1f6dcd11be905de59a04"
            "4f6c2ea3982b2235acc9021a196fc4ce0b19f6b312ee9cfc5997dc5f7ce2f386131294a56ba93a41a3b60e27e03956039f51ae73b89c795c"
            "5ae7d841e9b455c37341c052404e8fe9fe4f0d52bc162a41f1eeb9ef292c66a9d6a619aa548807eb1187ee22bd62e20e26c3c08c22ecef12"
            "d3b2304a010ed1f50a68e0261afe1a0bdddf7ab8a61774d3af3f1cce2b95dad3", 16)
    rng = random.Random(5)
    secret = int.from_bytes(b"secret", "big")
    shares = sss.distribute(secret, [rng.randrange(m) for _ in range(6)], 10, m)
    for _ in range(5):
        assert sss.calculate_secret(rng.sample(shares, 7), m) == secret
        assert sss.calculate_s(rng.sample(shares, 7), m) == secret     # dsa_core.go:389-403, same math


def test_packet_roundtrip():
    sig = pk.SignaturePacket(1, 0, False, b"SIG", b"CERT")
    ss = pk.SignaturePacket(1, 0, True, b"SS", b"")
    p = pk.serialize(b"var", b"val", 7, sig, ss)
    assert pk.parse(p) == (b"var", b"val", 7, sig, ss, None)
    assert pk.tbs(p) == pk.serialize(b"var", b"val", 7)
    assert pk.tbss(p) == pk.serialize(b"var", b"val", 7, sig)
    assert pk.parse(pk.serialize(b"var")) == (b"var", None, 0, None, None, None)
    assert pk.parse(pk.serialize(b"var", None, 0, None, None))[3:5] == (None, None)    # nil sig = type 0


def clique_graph(n, self_id=100, client_signs=None, certify_client=False):
    """scripts/setup.sh-style topology: clique 1..n fully cross-signed; client `self_id` signs
    `client_signs` members; with certify_client the members also sign the client, which makes
    the client a bidirectional neighbour of every member, i.e. part of the clique itself."""
    g = Graph()
    members = list(range(1, n + 1))
    client_signs = members if client_signs is None else client_signs
    nodes = [Node(i, signers=[j for j in members if j != i] + ([self_id] if i in client_signs else [])) for i in members]
    me = Node(self_id, signers=members if certify_client else [])
    g.add_nodes(nodes + [me])
    g.set_self_nodes([me])
    return g, nodes, me


@pytest.mark.parametrize("n,f,mn,thr_auth,thr_read,suff", [(4, 1, 4, 3, 2, 3), (10, 3, 10, 7, 4, 7), (15, 4, 13, 9, 5, 10),
                                                            (16, 5, 16, 11, 6, 11), (31, 10, 31, 21, 11, 21)])
def test_wotqs_thresholds(n, f, mn, thr_auth, thr_read, suff):
    """wotqs.go:55-66 closed forms."""
    g, nodes, me = clique_graph(n)
    qs = WotQS(g)
    qa = qs.choose_quorum(AUTH)
    assert len(qa.qcs) == 1
    qc = qa.qcs[0]
    assert (qc.f, qc.min, qc.threshold, qc.suff) == (f, mn, thr_auth, suff)
    assert sorted(x.id for x in qc.nodes) == list(range(1, n + 1))
    qr = qs.choose_quorum(READ | AUTH)
    assert qr.qcs[0].threshold == thr_read
    assert qa.get_threshold() == thr_auth
    # predicates incl. duplicate-counting quirk (wotqs.go:195-206)
    assert not qa.is_threshold(nodes[:thr_auth - 1])
    assert qa.is_threshold(nodes[:thr_auth])
    assert qa.is_threshold([nodes[0]] * thr_auth)
    assert qa.is_sufficient(nodes[:suff]) and not qa.is_sufficient(nodes[:suff - 1])
    assert qa.is_quorum(nodes[:mn]) and not qa.is_quorum(nodes[:mn - 1])
    assert qa.reject(nodes[:f + 1]) and not qa.reject(nodes[:f])
    # weight <= n - suff  =>  suff = 0 (wotqs.go:63-65)
    g2, nodes2, _ = clique_graph(n, client_signs=list(range(1, n - suff + 1)))
    q2 = WotQS(g2).choose_quorum(AUTH)
    assert q2.qcs[0].suff == 0 and not q2.is_sufficient(nodes2)


def test_wotqs_client_inside_clique():
    """A client that signs all four nodes AND is certified by all four is a 5th clique member
    (graph.go:333-368 only looks at bidirectional edges): n=5 -> f=1, thr 3, suff 4; PEER drops self."""
    g, nodes, me = clique_graph(4, certify_client=True)
    qc = WotQS(g).choose_quorum(AUTH).qcs[0]
    assert (len(qc.nodes), qc.f, qc.min, qc.threshold, qc.suff) == (5, 1, 4, 3, 4)
    qp = WotQS(g).choose_quorum(AUTH | PEER).qcs[0]
    assert (len(qp.nodes), qp.f, qp.threshold, qp.suff) == (4, 1, 3, 3)
    qcert = WotQS(g).choose_quorum(AUTH | CERT).qcs[0]               # distance 0 still finds self's clique
    assert qcert.threshold == 2                                      # f + 1


def test_wotqs_empty_and_write():
    g = Graph()
    me = Node(1)
    g.add_nodes([me]); g.set_self_nodes([me])
    q = WotQS(g).choose_quorum(AUTH)
    assert q.qcs == [] and q.reject([]) and not q.is_threshold([]) and not q.is_quorum([]) and not q.is_sufficient([])
    g, nodes, me = clique_graph(4)
    qw = WotQS(g).choose_quorum(WRITE)
    assert all(qc.threshold == 0 for qc in qw.qcs) or qw.qcs == []
    qp = WotQS(g).choose_quorum(AUTH | PEER)
    assert len(qp.qcs) == 1 and len(qp.qcs[0].nodes) == 4          # self (client) is not in the clique
    qc = WotQS(g).choose_quorum(AUTH | CERT)                        # distance 0: client alone is no clique
    assert qc.qcs == []


def test_c_tally_matches_python(built):
    rng = np.random.default_rng(3)
    qcs_py = [wq.QC([Node(i) for i in range(1, 17)], 5, 16, 6, 11), wq.QC([Node(i) for i in range(40, 44)], 1, 4, 2, 0)]
    q = wq.Quorum(qcs_py)
    qcs_c = [(c.f, c.min, c.threshold, c.suff, [n.id for n in c.nodes]) for c in qcs_py]
    n_ops, off, ids, st = 300, [0], [], []
    for _ in range(n_ops):
        r = int(rng.integers(0, 24))
        ids += [int(x) for x in rng.choice(list(range(1, 20)) + list(range(40, 45)), r)]       # duplicates possible
        st += [int(x) for x in rng.choice([0, 0, 0, 1, 4, 6], r)]
        off.append(len(ids))
    got = c_oracle.tally_batch(qcs_c, off, ids if ids else [0], st if st else [0])
    for i in range(n_ops):
        ok = [Node(ids[p]) for p in range(off[i], off[i + 1]) if st[p] == 0]
        bad = [Node(ids[p]) for p in range(off[i], off[i + 1]) if st[p] != 0]
        exp = q.is_quorum(ok) | (q.is_threshold(ok) << 1) | (q.is_sufficient(ok) << 2) | (q.reject(bad) << 3)
        assert got[i] == exp, i


def test_digestinfo_prefixes_against_openssl():
    """Every DigestInfo prefix the oracle (and K1's constant table) carries — MD5, SHA-1, SHA-224/256/384/512 —
    against OpenSSL in both directions: OpenSSL-made signatures verify in the oracle, and signatures made by
    textbook exponentiation over the oracle's EM verify in OpenSSL."""
    import hashlib
    from cryptography.exceptions import InvalidSignature
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import padding
    from cryptography.hazmat.primitives.asymmetric.utils import Prehashed
    from bftkv_b200 import workload
    from oracle import pgp_oracle as po
    k = workload.load_keys(1)[0]
    priv = workload._private_key(k)
    pub = priv.public_key()
    algs = {1: hashes.MD5(), 2: hashes.SHA1(), 8: hashes.SHA256(), 9: hashes.SHA384(), 10: hashes.SHA512(), 11: hashes.SHA224()}
    for hid, alg in algs.items():
        d = hashlib.new(po.HASH_BY_ID[hid], b"prefix check %d" % hid).digest()
        try:
            s = priv.sign(d, padding.PKCS1v15(), Prehashed(alg))
        except Exception:                                   # an OpenSSL build that refuses to SIGN with MD5 / SHA-1
            s = None
        if s is not None:
            assert po.rsa_verify_pkcs1v15(k["n"], k["e"], hid, d, s)
            assert not po.rsa_verify_pkcs1v15(k["n"], k["e"], hid, d[:-1] + bytes([d[-1] ^ 1]), s)
        t = po.DIGEST_PREFIX[hid] + d
        em = b"\x00\x01" + b"\xff" * (256 - len(t) - 3) + b"\x00" + t
        raw = pow(int.from_bytes(em, "big"), k["d"], k["n"]).to_bytes(256, "big")
        assert po.rsa_verify_pkcs1v15(k["n"], k["e"], hid, d, raw)
        try:
            pub.verify(raw, d, padding.PKCS1v15(), Prehashed(alg))
        except InvalidSignature:
            raise AssertionError("OpenSSL rejects the oracle's EM for hash id %d" % hid)
        except Exception:
            pass                                            # verification with this digest disabled in this OpenSSL


def test_v3_and_v4_hand_built_packets_against_gpg(tmp_path):
    """The packet makers the GPU parity tests use (bftkv_b200/workload.py: v3 and v4 RSA signature packets, binary
    and text mode, SHA-1/SHA-2 digests) produce signatures GnuPG accepts, and the oracle agrees with GnuPG on every
    one of them, valid and corrupted — so the oracle's v3 digest rule (H(data || sig type || creation time),
    packet.SignatureV3 / VerifySignatureV3) and its text canonicalisation are pinned on an independent tool.
    MD5 is left to the DigestInfo test above: GnuPG 2.4 refuses MD5 signatures outright."""
    import os
    import subprocess
    from bftkv_b200 import workload
    from oracle import pgp_oracle as po
    k = workload.load_keys(1)[0]
    priv = workload._private_key(k)
    blk, kid = workload.pgp_public_key_block(k, priv, b"v3 pin <v3@bftq.test>")
    ents = po.read_entities(blk)
    home = str(tmp_path)
    os.chmod(home, 0o700)
    env = dict(os.environ, GNUPGHOME=home)
    (tmp_path / "k").write_bytes(blk)
    r = subprocess.run(["gpg", "--batch", "--import", str(tmp_path / "k")], env=env, capture_output=True, text=True)
    assert "imported: 1" in r.stderr, r.stderr
    n = 0
    for maker in (workload.sig_packet_v3, workload.sig_packet_v4):
        for hid in (2, 8, 9, 10, 11):
            for sig_type in (0, 1):
                for bad in (False, True):
                    data = b"line one\nline two\r\nlast line %d" % n
                    pkt = bytearray(maker(k, kid, hid, data, 0x5F000000 + n, sig_type))
                    if bad:
                        pkt[-7] ^= 0x10
                    (tmp_path / "m").write_bytes(data)
                    (tmp_path / "s").write_bytes(bytes(pkt))
                    g = subprocess.run(["gpg", "--batch", "--verify", str(tmp_path / "s"), str(tmp_path / "m")], env=env,
                                       capture_output=True, text=True)
                    good = "Good signature" in g.stderr
                    assert good == (not bad), (maker.__name__, hid, sig_type, bad, g.stderr[-300:])
                    assert (po.signature_verify(ents, data, bytes(pkt)) is None) == good
                    n += 1
    assert n == 40
