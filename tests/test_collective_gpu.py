"""GPU parity suite of CollectiveSignature.Verify's batch form (crypto_pgp.go:485-500) at the quorum sizes of the BASELINE
configs — n = 10 / 16 / 31 cliques, with duplicated signers, unknown issuers, non-members, corrupted members, stray bytes
and packet forms the GPU parser hands to the host packer — against the oracle, with the K0 fast path on and off."""
import os
import random
import struct

import pytest

from bftkv_b200 import workload
from bftkv_b200.crypto_gpu import CollectiveSignature, Keyring, Quorum, Signature
from oracle import pgp_oracle as pgp, wotqs_oracle as wq
from oracle.wotqs_oracle import Node

pytestmark = pytest.mark.gpu
NK = 33                     # 31 quorum members + one keyring member outside every clique + one outsider


@pytest.fixture(scope="module")
def world(engine):
    keys = workload.load_keys(NK)
    blocks, kids = [], []
    for i, k in enumerate(keys):
        b, kid = workload.pgp_public_key_block(k, workload._private_key(k), b"a%02d (http://localhost:57%02d) <a%02d@bftq.test>" % (i, i, i))
        blocks.append(b); kids.append(kid)
    ring = b"".join(blocks[:NK - 1])                              # the last key is not in the keyring
    kr = Keyring(engine)
    kr.register(ring)
    tbs = [workload.tbs_packet(b"variable-%d" % j, bytes([j]) * 32, 100 + j) for j in range(4)]
    sigs = {(j, i): workload.sig_packet_v4(keys[i], kids[i], 8, tbs[j], 0x5F000000 + i) for j in range(4) for i in range(NK)}
    yield {"kr": kr, "ents": pgp.read_entities(ring), "keys": keys, "kids": kids, "tbs": tbs, "sigs": sigs, "engine": engine}
    kr.close()


def new_format(pkt: bytes) -> bytes:
    """The same signature packet body under a new-format header (what Go's serializer writes)."""
    body = pkt[3:]
    return bytes([0xC2, ((len(body) - 192) >> 8) + 192, (len(body) - 192) & 0xFF]) + body


def partial(pkt: bytes) -> bytes:
    """... and split into a 256-byte partial chunk + the rest (legal for any packet in x/crypto's reader)."""
    body = pkt[3:]
    rest = body[256:]
    return bytes([0xC2, 224 + 8]) + body[:256] + bytes([len(rest)]) + rest


@pytest.mark.parametrize("n,params", [(10, (3, 10, 7, 7)), (16, (5, 16, 11, 11)), (31, (10, 31, 21, 21))])
@pytest.mark.parametrize("gpu_parse", ["1", "0"])
def test_collective_verify_matches_oracle(world, n, params, gpu_parse):
    w = world
    f, mn, th, sf = params
    clique = w["kids"][:n]
    q = Quorum(w["engine"], [(f, mn, th, sf, clique)])
    oq = wq.Quorum([wq.QC([Node(i) for i in clique], f, mn, th, sf)])
    cs = CollectiveSignature(Signature(w["kr"]))
    rng = random.Random(1000 + n)
    tbs_list, streams = [], []
    for trial in range(260):
        j = rng.randrange(4)
        k = rng.choice([sf - 2, sf - 1, sf, sf, sf + 1, n, n])
        members = rng.sample(range(n), min(k, n))
        parts = []
        for i in members:
            s = w["sigs"][(j, i)]
            u = rng.random()
            if u < 0.06:
                b = bytearray(s); b[len(b) - 1 - rng.randrange(200)] ^= 1 << rng.randrange(8); s = bytes(b)      # corrupted member
            elif u < 0.10:
                s = w["sigs"][((j + 1) % 4, i)]                                                                     # signed something else
            elif u < 0.14:
                s = new_format(s)
            elif u < 0.16:
                s = partial(s)                                                                                      # host-packer shape
            parts.append(s)
        for _ in range(rng.choice([0, 0, 1, 2])):
            extra = rng.choice(["dup", "outsider", "nonmember", "junk", "stray", "userid"])
            pos = rng.randrange(len(parts) + 1)
            if extra == "dup" and members:
                parts.insert(pos, w["sigs"][(j, rng.choice(members))])                                             # duplicates count twice
            elif extra == "outsider":
                parts.insert(pos, w["sigs"][(j, NK - 1)])                                                           # issuer not in the keyring
            elif extra == "nonmember":
                parts.insert(pos, w["sigs"][(j, NK - 2)])                                                           # in the keyring, in no clique
            elif extra == "junk":
                parts.insert(pos, bytes(rng.randrange(128) for _ in range(rng.randint(1, 4))))                      # MSB clear: one byte each
            elif extra == "stray":
                parts.insert(pos, bytes([0xC0 | 61, 3]) + b"abc")                                                   # unknown packet type
            else:
                parts.insert(pos, bytes([0xB4, 5]) + b"alice")                                                      # a user-id packet
        tbs_list.append(w["tbs"][j]); streams.append(b"".join(parts))
    streams[0] = b""
    os.environ["BFTQ_GPU_PARSE"] = gpu_parse
    try:
        got = cs.verify_batch(tbs_list, streams, q)
    finally:
        del os.environ["BFTQ_GPU_PARSE"]
    ref = [pgp.collective_verify(w["ents"], t, s, oq)[0] for t, s in zip(tbs_list, streams)]
    assert got == ref
    assert 40 < sum(r is None for r in ref) < 220, sum(r is None for r in ref)


def test_strict_range_policy_through_the_packer(world):
    """s + n has the same residue as s: Go 1.13's rsa.VerifyPKCS1v15 (the version go.mod pins) accepts it, Go >= 1.20 rejects
    s >= n.  The packet-level entry points follow the engine's policy flag (bftq_engine_set_verify_flags)."""
    import ctypes as C
    from bftkv_b200 import _lib
    w = world
    sig_api = Signature(w["kr"])
    tbs = w["tbs"][0]
    items = []
    for i in range(NK - 1):
        p = workload.sig_packet_v4(w["keys"][i], w["kids"][i], 8, tbs, 0x5F000000 + i, plus_n=True)
        if p is not None:
            items.append(p)
    assert len(items) >= 4                                   # s + n < 2^2048 for a good part of the keys
    plain = [w["sigs"][(0, 0)]] * 2
    lib, h = w["engine"]._lib, w["engine"]._h
    try:
        assert sig_api.verify_batch([tbs] * (len(items) + 2), items + plain) == [None] * (len(items) + 2)
        _lib.check(lib.bftq_engine_set_verify_flags(h, 1))   # BFTQ_F_STRICT_RANGE
        got = sig_api.verify_batch([tbs] * (len(items) + 2), items + plain)
        assert got == ["crypto: invalid signature"] * len(items) + [None] * 2
    finally:
        _lib.check(lib.bftq_engine_set_verify_flags(h, 0))
