"""GPU suite for the reference-facing operator interface (bftkv_b200.crypto_gpu over the packer
entry points of libbftq): Signature.Verify / VerifyWithCertificate / Signers and
CollectiveSignature.Verify / Combine decisions must equal the oracle's restatement of
crypto/pgp/crypto_pgp.go:319-344,373-390,485-515 on GnuPG-made inputs."""
import os

import numpy as np
import pytest

from bftkv_b200 import Engine
from bftkv_b200.crypto_gpu import (BatchingSignature, CollectiveSignature, ErrInsufficientNumberOfSignatures, ErrInvalidSignature, Keyring, Quorum,
                                   Signature)
from oracle import packet_oracle as pk, pgp_oracle as pgp, wotqs_oracle as wq
from oracle.wotqs_oracle import Node

pytestmark = pytest.mark.gpu
RING = ["a01", "a02", "a03", "a04", "u01"]


@pytest.fixture(scope="module")
def env(golden, built):
    e = Engine(0)
    kr = Keyring(e)
    ents = []
    for n in RING:
        blob = bytes.fromhex(golden["keys"][n]["pub"])
        assert kr.register(blob) == 1
        ents += pgp.read_entities(blob)
    yield {"engine": e, "kr": kr, "sig": Signature(kr), "ents": ents, "golden": golden}
    kr.close()
    e.close()


def case(golden, signer, tbs_hex=None, algo="SHA256"):
    for c in golden["cases"]:
        if c["signer"] == signer and c["hash"] == algo and (tbs_hex is None or c["tbs"] == tbs_hex):
            return bytes.fromhex(c["tbs"]), bytes.fromhex(c["sig"])
    raise KeyError


def test_keyring_mirrors_reference(env, golden):
    ids = env["kr"].get_keyring()
    assert ids == [int(golden["keys"][n]["key_id"], 16) for n in RING]
    # Certificate.Signers (crypto_pgp.go:80-88): a01 is certified by a02..a04 and u01
    cert = env["kr"].certifiers(ids[0])
    assert sorted(cert) == sorted(int(golden["keys"][n]["key_id"], 16) for n in ["a02", "a03", "a04", "u01"])
    assert cert == [int(c, 16) for c in golden["keys"]["a01"]["certifiers"]]


def test_signature_verify_all_golden(env, golden):
    cases = [c for c in golden["cases"]]
    tbs = [bytes.fromhex(c["tbs"]) for c in cases] + [bytes.fromhex(c["tbs"]) + b"x" for c in cases]
    sig = [bytes.fromhex(c["sig"]) for c in cases] * 2
    got = env["sig"].verify_batch(tbs, sig)
    ref = [pgp.signature_verify(env["ents"], t, s) for t, s in zip(tbs, sig)]
    assert got == ref
    assert sum(r is None for r in ref) == 41 and {c["hash"] for c in cases} == {"SHA256", "SHA512", "SHA1"}


def test_signature_verify_stream_semantics(env, golden):
    g = golden
    tbs, s1 = case(g, "a01")
    _, s2 = case(g, "a02", tbs.hex())
    _, sx = case(g, "x99", tbs.hex())
    bad = bytearray(s1); bad[-5] ^= 1
    uid = bytes([0xB4, 3]) + b"abc"
    unk = bytes([0xC0 | 60, 2, 1, 2])
    streams = [b"", s1, s1 + s2, sx + s1, s1 + sx, bytes(bad), s1 + bytes(bad), s1[:-7], uid + s1, s1 + unk, unk + s1, s2 + s1 + s2,
               sx, sx + sx, s1 + s1[:10]]
    got = env["sig"].verify_batch([tbs] * len(streams), streams)
    ref = [pgp.signature_verify(env["ents"], tbs, s) for s in streams]
    assert got == ref
    assert ref[:5] == [ErrInvalidSignature, None, None, None, ErrInvalidSignature]
    # single-item entry points
    assert env["sig"].verify(tbs, s1) is None and env["sig"].verify(tbs + b"!", s1) == ErrInvalidSignature
    # VerifyWithCertificate: keyring = the first entity of the certificate
    c1, c2 = bytes.fromhex(g["keys"]["a01"]["pub"]), bytes.fromhex(g["keys"]["a02"]["pub"])
    assert env["sig"].verify_with_certificate(tbs, s1, c1) is None
    assert env["sig"].verify_with_certificate(tbs, s1, c2) == ErrInvalidSignature
    assert env["sig"].verify_with_certificate(tbs, s1, c2 + c1) == ErrInvalidSignature      # only the FIRST entity counts
    assert env["sig"].verify_with_certificate(tbs, s1, b"") == ErrInvalidSignature
    xc = bytes.fromhex(g["keys"]["x99"]["pub"])                                             # signer outside the keyring, cert supplied
    assert env["sig"].verify_with_certificate(tbs, sx, xc) is None
    # Signers: issuers found in the keyring, duplicates kept, unknown dropped
    ids = {n: int(g["keys"][n]["key_id"], 16) for n in g["keys"]}
    assert env["sig"].signers(s1 + s2 + sx + s1) == [ids["a01"], ids["a02"], ids["a01"]] == pgp.signers(env["ents"], s1 + s2 + sx + s1)


def test_collective_signature(env, golden):
    """Server-side write path (protocol/server.go:300): CollectiveSignature.Verify(tbss, ss, quorum)
    for the 4-node clique of BASELINE config 1 (f=1, threshold 3, suff 3)."""
    g = golden
    ids = {n: int(g["keys"][n]["key_id"], 16) for n in g["keys"]}
    clique = [ids[n] for n in ["a01", "a02", "a03", "a04"]]
    q = Quorum(env["engine"], [(1, 4, 3, 3, clique)])
    oq = wq.Quorum([wq.QC([Node(i) for i in clique], 1, 4, 3, 3)])
    cs = CollectiveSignature(env["sig"])
    tbs, s1 = case(g, "a01")
    sigs = {n: case(g, n, tbs.hex())[1] for n in ["a01", "a02", "a03", "a04", "u01", "x99"]}
    bad2 = bytearray(sigs["a02"]); bad2[-9] ^= 4
    streams = [b"", sigs["a01"], sigs["a01"] + sigs["a02"], sigs["a01"] + sigs["a02"] + sigs["a03"],
               sigs["a01"] + bytes(bad2) + sigs["a03"], sigs["a01"] + bytes(bad2) + sigs["a03"] + sigs["a04"],
               sigs["a01"] * 3,                                   # duplicates count (F8 / Appendix C.1)
               sigs["x99"] + sigs["u01"] + sigs["a01"] + sigs["a02"],   # non-members and unknown issuers do not count
               sigs["x99"] + sigs["a01"] + sigs["a02"] + sigs["a04"] + sigs["x99"],
               bytes([0xB4, 3]) + b"abc" + sigs["a01"] + sigs["a02"] + sigs["a03"]]
    got = cs.verify_batch([tbs] * len(streams), streams, q)
    ref = [pgp.collective_verify(env["ents"], tbs, s, oq)[0] for s in streams]
    assert got == ref
    assert ref == [ErrInsufficientNumberOfSignatures] * 3 + [None, ErrInsufficientNumberOfSignatures, None, None,
                                                              ErrInsufficientNumberOfSignatures, None, None]
    assert cs.verify(tbs, streams[3], q) == (None, True)
    # Combine (client side, protocol/client.go:153): parse-only sufficiency of ss ++ s
    acc_t, acc = 0, b""
    oks = []
    for n in ["a01", "a02", "x99", "a03"]:
        ok, acc_t, acc = cs.combine(acc_t, acc, pk.SIGNATURE_TYPE_PGP, sigs[n], q)
        ok_ref, _, _ = pgp.collective_combine(env["ents"], 0 if not oks else 1, acc[:-len(sigs[n])], 1, sigs[n], oq)
        assert ok == ok_ref
        oks.append(ok)
    assert oks == [False, False, False, True]
    assert cs.combine(1, b"", 2, sigs["a01"], q)[0] is False       # type mismatch
    # Quorum predicates through the GPU tally mirror quorum.Quorum
    assert q.is_threshold(clique[:3]) and not q.is_threshold(clique[:2]) and q.is_threshold([clique[0]] * 3)
    assert q.is_sufficient(clique[:3]) and q.is_quorum(clique) and not q.is_quorum(clique[:3])
    assert q.reject(clique[:2]) and not q.reject(clique[:1]) and q.get_threshold() == 3


def test_aggregator_coalesces_concurrent_verifies(env, golden):
    """64 threads x 8 single Verify calls (the goroutine-per-peer pattern of transport.Multicast) must
    give the reference's answers and end up in far fewer GPU batches than calls."""
    import threading
    cases = [c for c in golden["cases"] if c["hash"] == "SHA256"]
    bs = BatchingSignature(env["kr"], max_batch=256, max_wait_us=20000)
    results = {}

    def worker(t):
        for i in range(8):
            c = cases[(t * 8 + i) % len(cases)]
            tamper = (t + i) % 3 == 0
            tbs = bytes.fromhex(c["tbs"]) + (b"!" if tamper else b"")
            if i % 4 == 3:
                cert = bytes.fromhex(golden["keys"][c["signer"]]["pub"])
                got = bs.verify_with_certificate(tbs, bytes.fromhex(c["sig"]), cert)
                ref = pgp.signature_verify_with_certificate(cert, tbs, bytes.fromhex(c["sig"]))
            else:
                got = bs.verify(tbs, bytes.fromhex(c["sig"]))
                ref = pgp.signature_verify(env["ents"], tbs, bytes.fromhex(c["sig"]))
            results[(t, i)] = (got, ref)
    ths = [threading.Thread(target=worker, args=(t,)) for t in range(64)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert len(results) == 512 and all(g == r for g, r in results.values())
    assert sum(g is None for g, _ in results.values()) > 150
    st = bs.stats()
    assert st["items"] == 512 and st["batches"] <= 64, st
    bs.close()


# ---- the chunked, multi-threaded packer (bftq_signature_verify_batch over a big batch) ---------------

def _pgp_batch(n, **kw):
    from bftkv_b200 import workload
    return workload.make_pgp_verify_batch(n, **kw)


def test_signature_verify_big_batch_chunked_threads():
    """20 000 detached OpenPGP signatures through Signature.Verify's batch form: several chunks per
    worker thread, several threads; decisions equal the generator's ground truth and (on a sample) the
    oracle's restatement of crypto_pgp.go:319-330.  Also with chunk boundaries that split unevenly."""
    w = _pgp_batch(20000, n_keys=16)
    eng = Engine(0)
    kr = Keyring(eng)
    kr.register(w["keyring"])
    sig = Signature(kr)
    ring = pgp.read_entities(w["keyring"])
    for chunk, threads in ((None, None), ("777", "5"), ("30000", "1")):
        if chunk:
            os.environ["BFTQ_PLAN_CHUNK"], os.environ["BFTQ_HOST_THREADS"] = chunk, threads
        try:
            got = sig.verify_batch(w["tbs"], w["sigs"])
        finally:
            os.environ.pop("BFTQ_PLAN_CHUNK", None), os.environ.pop("BFTQ_HOST_THREADS", None)
        ok = np.array([g is None for g in got])
        assert np.array_equal(ok, w["expect_ok"]), (chunk, threads, int(np.sum(ok != w["expect_ok"])))
    for i in list(range(0, 20000, 397)) + [int(j) for j in np.nonzero(~w["expect_ok"])[0][:40]]:
        assert (pgp.signature_verify(ring, w["tbs"][i], w["sigs"][i]) is None) == bool(ok[i])
    # an outsider becomes a known issuer once its key block is registered
    bad = [int(j) for j in np.nonzero(~w["expect_ok"])[0]]
    kr.register(w["outsider_block"])
    got2 = sig.verify_batch([w["tbs"][j] for j in bad], [w["sigs"][j] for j in bad])
    ring2 = pgp.read_entities(w["keyring"] + w["outsider_block"])
    for j, g in zip(bad, got2):
        assert (g is None) == (pgp.signature_verify(ring2, w["tbs"][j], w["sigs"][j]) is None)
    assert any(g is None for g in got2)
    kr.close()
    eng.close()


# ---- v3 signature packets and MD5 (what foreign keys / old implementations may send) -----------------

def test_v3_signatures_and_md5():
    """packet.SignatureV3 (digest = H(data || sig type || creation time)) and MD5 / SHA-1 / SHA-2 digests through
    Signature.Verify, binary and text mode, valid and corrupted — decisions equal the oracle's restatement of
    CheckDetachedSignature -> VerifySignature / VerifySignatureV3."""
    from bftkv_b200 import workload
    keys = workload.load_keys(3)
    privs = [workload._private_key(k) for k in keys]
    e = Engine(0)
    kr = Keyring(e)
    blocks, kids = [], []
    for i, k in enumerate(keys):
        b, kid = workload.pgp_public_key_block(k, privs[i], b"v3-%d" % i)
        blocks.append(b); kids.append(kid)
        kr.register(b)
    ents = pgp.read_entities(b"".join(blocks))
    sig = Signature(kr)
    tbs, sigs = [], []
    j = 0
    for maker in (workload.sig_packet_v3, workload.sig_packet_v4):
        for hid in (1, 2, 8, 9, 10, 11):
            for sig_type in (0, 1):
                for bad in (False, True):
                    i = j % 3
                    data = b"line one\nline two\r\nvalue %d" % j
                    pkt = bytearray(maker(keys[i], kids[i], hid, data, 0x5F000000 + j, sig_type))
                    if bad:
                        pkt[-5] ^= 0x02
                    tbs.append(data); sigs.append(bytes(pkt))
                    j += 1
    # a v3 and a v4 packet in one stream (Verify: all must be valid), and a v3 packet with an unknown issuer first
    tbs.append(b"both"); sigs.append(workload.sig_packet_v3(keys[0], kids[0], 8, b"both", 1) + workload.sig_packet_v4(keys[1], kids[1], 1, b"both", 2))
    tbs.append(b"skip"); sigs.append(workload.sig_packet_v3(keys[0], 0x1122334455667788, 8, b"skip", 1) + workload.sig_packet_v3(keys[2], kids[2], 2, b"skip", 3))
    got = sig.verify_batch(tbs, sigs)
    n_ok = 0
    for t, s, g in zip(tbs, sigs, got):
        ref = pgp.signature_verify(ents, t, s)
        assert (g is None) == (ref is None), (t, s.hex()[:40])
        n_ok += ref is None
    assert n_ok == 24 + 2
    kr.close()
    e.close()


def test_concurrent_batch_callers_share_the_pool():
    """Four threads call Signature.Verify's batch form at once (what concurrent server handlers do): the calls
    share the engine's worker pool and staging slots; every caller must get exactly the single-caller answers,
    through K0 and through the host packer."""
    import threading
    w = _pgp_batch(6000, n_keys=8)
    e = Engine(0)
    kr = Keyring(e)
    kr.register(w["keyring"])
    sig = Signature(kr)
    for gpu_parse in ("1", "0"):
        os.environ["BFTQ_GPU_PARSE"] = gpu_parse
        try:
            outs, errs = [None] * 4, []

            def caller(c):
                try:
                    for _ in range(3):
                        got = sig.verify_batch(w["tbs"], w["sigs"])
                        ok = np.array([g is None for g in got])
                        assert np.array_equal(ok, w["expect_ok"])
                    outs[c] = ok
                except Exception as ex:                      # surfaces in the main thread below
                    errs.append(ex)
            ths = [threading.Thread(target=caller, args=(c,)) for c in range(4)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            assert not errs, errs
            assert all(o is not None for o in outs)
        finally:
            del os.environ["BFTQ_GPU_PARSE"]
    st = e.stats()
    assert st["packer_chunks"] > 0
    kr.close()
    e.close()
