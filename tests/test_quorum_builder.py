"""CPU suite: the host-side quorum-descriptor builder of libbftq (bftq_graph_*, no GPU involved)
against the oracle's restatement of node/graph/graph.go + quorum/wotqs/wotqs.go."""
import random

import pytest

from bftkv_b200.crypto_gpu import AUTH, CERT, PEER, READ, WRITE, QuorumSystem
from oracle.wotqs_oracle import Graph, Node, WotQS

RWS = [AUTH, READ, WRITE, READ | AUTH, WRITE | AUTH, AUTH | PEER, AUTH | CERT, CERT, READ | WRITE, READ | WRITE | AUTH, AUTH | PEER | READ]


def both(nodes, self_id, ops=()):
    """nodes: list of (id, [signers]) in insertion order."""
    g = Graph()
    qs = QuorumSystem()
    objs = {}
    for nid, signers in nodes:
        objs[nid] = Node(nid, signers=list(signers))
        g.add_nodes([objs[nid]])
        qs.add_node(nid, signers)
    g.set_self_nodes([objs[self_id]])
    qs.set_self(self_id)
    for op, nid in ops:
        if op == "remove":
            g.remove_nodes([objs[nid]]); qs.remove_node(nid)
        else:
            g.revoke(objs[nid]); qs.revoke(nid)
    return WotQS(g), qs


def desc(q):
    return [(c.f, c.min, c.threshold, c.suff, [n.id for n in c.nodes]) for c in q.qcs]


def setup_sh_topology():
    """scripts/setup.sh:17-50: cliques a01-a10 and b01-b10 fully cross-signed, rw01-06 sign all a*, b*;
    u01 signs a01-a06 + rw*; a07-a10 sign u01."""
    a = list(range(101, 111)); b = list(range(201, 211)); rw = list(range(301, 307)); u01 = 1
    nodes = []
    for x in a:
        nodes.append((x, [y for y in a if y != x] + rw + ([u01] if x <= 106 else [])))
    for x in b:
        nodes.append((x, [y for y in b if y != x] + rw))
    for x in rw:
        nodes.append((x, [u01]))
    nodes.append((u01, a[6:]))
    return nodes, u01, a, b, rw


def test_setup_sh_topology(built):
    nodes, u01, a, b, rw = setup_sh_topology()
    o, q = both(nodes, u01)
    for r in RWS:
        assert q.choose_quorum_desc(r) == desc(o.choose_quorum(r)), r
    d = q.choose_quorum_desc(AUTH | PEER)          # one clique of the 10 a* nodes, f=3, threshold 7, suff 7
    assert len(d) == 1 and d[0][:4] == (3, 10, 7, 7) and sorted(d[0][4]) == a
    q.close()


def test_random_graphs_match_oracle(built):
    rng = random.Random(11)
    for trial in range(60):
        n_cl = rng.randint(1, 3)
        ids, nodes, nxt = [], [], 10
        cliques = []
        for _ in range(n_cl):
            size = rng.randint(2, 9)
            cl = list(range(nxt, nxt + size)); nxt += size + 3
            cliques.append(cl)
        extra = list(range(nxt, nxt + rng.randint(0, 4)))
        me = 1
        for cl in cliques:
            for x in cl:
                signers = [y for y in cl if y != x]
                if rng.random() < 0.7:
                    signers.append(me)                                  # client trusts this node
                signers += [e for e in extra if rng.random() < 0.3]
                nodes.append((x, signers))
        for e in extra:
            nodes.append((e, [me] if rng.random() < 0.5 else []))
        cert = [x for cl in cliques for x in cl if rng.random() < 0.3]
        nodes.append((me, cert))
        rng.shuffle(nodes)
        ops = []
        if rng.random() < 0.4:
            victim = rng.choice([n for n, _ in nodes if n != me])
            ops.append((rng.choice(["remove", "revoke"]), victim))
        o, q = both(nodes, me, ops)
        for r in RWS:
            assert q.choose_quorum_desc(r) == desc(o.choose_quorum(r)), (trial, r)
        q.close()
