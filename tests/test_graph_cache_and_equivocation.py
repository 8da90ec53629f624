"""CPU suite (host-only entry points, no GPU): the version-stamped quorum-descriptor cache of bftq_graph_* and the
batched equivocation scan (Client.revoke, protocol/client.go:304-346) against the oracle's restatements."""
import ctypes as C
import random

import numpy as np

from bftkv_b200 import _lib
from oracle import wotqs_oracle as wq


class QCIds(C.Structure):
    _fields_ = [("f", C.c_int32), ("min", C.c_int32), ("threshold", C.c_int32), ("suff", C.c_int32), ("member_off", C.c_uint32), ("member_cnt", C.c_uint32)]


def choose(lib, g, rw):
    qcs = (QCIds * 8)()
    mem = np.zeros(256, np.uint64)
    nq, nm = C.c_uint32(), C.c_uint32()
    _lib.check(lib.bftq_graph_choose_quorum(g, rw, C.cast(qcs, C.c_void_p), 8, C.byref(nq), C.c_void_p(mem.ctypes.data), 256, C.byref(nm)))
    return [(q.f, q.min, q.threshold, q.suff, [int(x) for x in mem[q.member_off:q.member_off + q.member_cnt]]) for q in qcs[:nq.value]]


def stats(lib, g):
    v, h, b = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _lib.check(lib.bftq_graph_version(g, C.byref(v), C.byref(h), C.byref(b)))
    return v.value, h.value, b.value


def test_descriptor_cache_follows_graph_version(built):
    lib = _lib.load()
    g = C.c_void_p()
    _lib.check(lib.bftq_graph_create(C.byref(g)))
    og = wq.Graph()
    ids = list(range(101, 108))                                  # a 7-clique + a client that signs / is signed by all

    def add(i, signers):
        a = np.array(signers or [0], np.uint64)
        _lib.check(lib.bftq_graph_add_node(g, i, C.c_void_p(a.ctypes.data), len(signers)))
        og.add_nodes([wq.Node(i, list(signers))])
    for i in ids:
        add(i, [j for j in ids if j != i] + [900])
    add(900, [])
    _lib.check(lib.bftq_graph_set_self(g, 900))
    og.set_self_nodes([wq.Node(900, [])])

    def oracle(rw):
        q = wq.WotQS(og).choose_quorum(rw)
        return [(c.f, c.min, c.threshold, c.suff, [n.id for n in c.nodes]) for c in q.qcs]
    v0, h0, b0 = stats(lib, g)
    for rw in (wq.READ, wq.AUTH | wq.PEER, wq.AUTH, wq.READ | wq.AUTH):
        assert choose(lib, g, rw) == oracle(rw)
    v1, h1, b1 = stats(lib, g)
    assert v1 == v0 and b1 - b0 == 4 and h1 == h0              # four descriptors built
    assert choose(lib, g, wq.AUTH)[0][:4] == (2, 7, 5, 5)        # the 7-clique: f = 2, min 7, threshold 5, suff 5
    for _ in range(49):
        assert choose(lib, g, wq.AUTH) == oracle(wq.AUTH)
    v2, h2, b2 = stats(lib, g)
    assert b2 == b1 and h2 - h1 == 50                           # ... and then served from the cache
    # revocation (graph.go:131-146) advances the version: the next call rebuilds, and the clique has shrunk
    _lib.check(lib.bftq_graph_revoke(g, 103))
    og.revoke(wq.Node(103))
    v3, _, _ = stats(lib, g)
    assert v3 > v2
    r = choose(lib, g, wq.AUTH)
    assert r == oracle(wq.AUTH) and r and 103 not in r[0][4] and len(r[0][4]) == 6
    assert stats(lib, g)[2] == b2 + 1
    lib.bftq_graph_destroy(g)


def test_equivocation_scan_matches_client_revoke(built):
    lib = _lib.load()
    rng = random.Random(31)
    n_ops = 400
    op_off, st, ts, vid, soff, sids, expect = [0], [], [], [], [0], [], []
    for _ in range(n_ops):
        m = {}
        resp = []
        for _ in range(rng.randint(0, 12)):
            good = rng.random() < 0.85
            t = rng.choice([0, 7, 7, 7, 8])
            v = rng.choice([0, 0, 0, 1, 2])
            signers = [rng.randrange(1, 9) for _ in range(rng.randint(0, 5))]
            resp.append((good, t, v, signers))
            st.append(0 if good else 1); ts.append(t); vid.append(v)
            sids += signers; soff.append(len(sids))
            if good:
                m.setdefault(t, {}).setdefault(v, []).append(signers)
        op_off.append(len(st))
        expect.append(wq.revoke_scan(m, lambda sv: sv))
    a = lambda x, t: np.array(x if x else [0], t)
    op_off, st, ts, vid, soff, sids = a(op_off, np.uint32), a(st, np.uint8), a(ts, np.uint64), a(vid, np.uint32), a(soff, np.uint32), a(sids, np.uint64)
    out_off = np.zeros(n_ops + 1, np.uint32)
    n = C.c_uint64()
    p = lambda x: C.c_void_p(x.ctypes.data)
    _lib.check(lib.bftq_equivocation_scan_batch(p(op_off), n_ops, p(st), p(ts), p(vid), p(soff), p(sids), p(out_off), None, 0, C.byref(n)))
    out = np.zeros(max(1, n.value), np.uint64)
    _lib.check(lib.bftq_equivocation_scan_batch(p(op_off), n_ops, p(st), p(ts), p(vid), p(soff), p(sids), p(out_off), p(out), n.value, C.byref(n)))
    got = [sorted(int(x) for x in out[out_off[i]:out_off[i + 1]]) for i in range(n_ops)]
    # the reference's order follows Go map iteration: compare as sets (each id at most once)
    assert got == [sorted(e) for e in expect]
    assert all(len(set(e)) == len(e) for e in expect) and sum(map(len, expect)) > 100
