"""GPU parity suite of bftq_read_responses_batch: raw transport answers in, Client.Read's decision out
(protocol/client.go:250-268 over transport.Multicast's per-response Decrypt + nonce check and processResponse's
packet.Parse), against the composition of the oracle's restatements."""
import random

import numpy as np
import pytest

from bftkv_b200 import workload
from bftkv_b200.crypto_gpu import Keyring, read_responses_batch
from oracle import packet_oracle, pgp_oracle as pgp, wotqs_oracle as wq
from oracle.wotqs_oracle import Node

pytestmark = pytest.mark.gpu


def status_class(st):
    return {0: pgp.ST_OK, 8: pgp.ST_UNVERIFIED, 7: pgp.ST_NONCE, 1: pgp.ST_INVALID, 2: pgp.ST_INVALID}.get(int(st), pgp.ST_OTHER)


def test_read_responses_against_oracle(engine):
    R = 10
    keys = workload.load_keys(R + 1)
    blocks, kids = [], []
    for i, k in enumerate(keys):
        b, kid = workload.pgp_public_key_block(k, workload._private_key(k), b"a%02d (http://localhost:57%02d) <a%02d@x>" % (i, i, i))
        blocks.append(b); kids.append(kid)
    ring = b"".join(blocks[:R])                                   # key R is an outsider
    kr = Keyring(engine)
    kr.register(ring)
    ents = pgp.read_entities(ring)
    qcs = [(3, 10, 4, 7, kids[:R])]                               # n = 10: f = 3, READ threshold 4
    quorum = wq.Quorum([wq.QC([Node(i) for i in kids[:R]], 3, 10, 4, 7)])
    rng = random.Random(99)
    n_ops = 300
    x = b"the variable"
    op_off, peers, msgs, nonces, pre = [0], [], [], [], []
    for op in range(n_ops):
        order = list(range(R))
        rng.shuffle(order)
        order = order[:rng.randint(0, R)] if rng.random() < 0.1 else order
        cur_t = 7
        cur_v = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 32, 300])))
        for r_ in order:
            u = rng.random()
            nonce = bytes(rng.randrange(256) for _ in range(8))
            t, v, signer, bad_pre = cur_t, cur_v, r_, 0
            if u < 0.08:
                t, v = cur_t - 1, b"older value"
            elif u < 0.12:
                v = cur_v + b"!"                                  # an equivocating replica: same t, other bytes
            elif u < 0.15:
                signer = R                                        # signed by a key outside the keyring: accepted unverified
            elif u < 0.18:
                bad_pre = 6
            plain = packet_oracle.serialize(x, v, t, None, None) if rng.random() < 0.95 else (b"" if rng.random() < 0.5 else packet_oracle.serialize(x, v, t)[:-3])
            m = bytearray(workload.make_transport_message(keys[signer], kids[signer], plain, nonce, hash_id=8 if rng.random() < 0.93 else 10))
            w = rng.random()
            if w < 0.05:
                m[rng.randrange(len(m))] ^= 1 << rng.randrange(8)
            elif w < 0.07:
                nonce = bytes(8)                                  # the request carried another nonce
            elif w < 0.08:
                del m[rng.randrange(len(m)):]
            peers.append(kids[r_]); msgs.append(bytes(m)); nonces.append(nonce); pre.append(bad_pre)
        op_off.append(len(peers))
    op_off = np.array(op_off, np.uint32)
    s0 = engine.stats()
    got = read_responses_batch(kr, qcs, op_off, np.array(peers, np.uint64), msgs, np.frombuffer(b"".join(nonces), np.uint8).reshape(-1, 8),
                               pre_status=np.array(pre, np.uint8))
    s1 = engine.stats()
    on_gpu, on_host = s1["msg_gpu_items"] - s0["msg_gpu_items"], s1["msg_host_items"] - s0["msg_host_items"]
    assert on_gpu + on_host == len(msgs) and on_gpu > 0.8 * len(msgs) and on_host > 0, (on_gpu, on_host)     # K0m decides the common shape, the rest falls back
    kinds, decs = {}, {}
    for op in range(n_ops):
        resp = []
        for p in range(op_off[op], op_off[op + 1]):
            st, t, v = pgp.read_response_status(ents, msgs[p], nonces[p], pre[p])
            assert status_class(got["status"][p]) == st, (op, p, int(got["status"][p]), st)
            kinds[st] = kinds.get(st, 0) + 1
            good = st in (pgp.ST_OK, pgp.ST_UNVERIFIED)
            if good:
                assert int(got["ts"][p]) == t and int(got["value_len"][p]) == len(v), (op, p)
            resp.append((Node(peers[p]), not good, t, v))
        kind, at, value, t = wq.read_decide(resp, quorum)
        assert (int(got["decision"][op]), int(got["decided_at"][op])) == (kind, at), (op, kind, at)
        decs[kind] = decs.get(kind, 0) + 1
        if kind == wq.READ_VALUE:
            wi = int(got["winner"][op])
            assert resp[wi][2] == t and resp[wi][3] == value and not resp[wi][1]
            assert all(not (not r2[1] and r2[2] == t and r2[3] == value) for r2 in resp[:wi])
        else:
            assert got["winner"][op] == 0xFFFFFFFF
    assert set(kinds) >= {pgp.ST_OK, pgp.ST_UNVERIFIED, pgp.ST_INVALID, pgp.ST_NONCE, pgp.ST_OTHER}, kinds
    assert set(decs) == {0, 1, 2}, decs
    kr.close()
