"""Oracle restatement of the web-of-trust quorum system.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Follows the reference:
  quorum/quorum.go:10-16                   rw flag constants
  node/graph/graph.go:46-75                AddNodes (edge signer -> signee)
  node/graph/graph.go:117-125              GetPeers
  node/graph/graph.go:279-319              GetReachableNodes / GetCliques
  node/graph/graph.go:333-393              findMaximalClique / bidirect / putWeight
  node/graph/graph.go:420-438              bfs
  quorum/wotqs/wotqs.go:36-206             newQC, complement, getQuorumFrom, ChooseQuorum,
                                           IsQuorum/IsThreshold/IsSufficient/Reject/GetThreshold,
                                           intersection (duplicates in the INPUT list are kept)
Go's map iteration order is random; this restatement iterates in insertion order, which
is one of the orders the reference may take.  For disjoint cliques (the only topology the
reference's fixtures build, scripts/setup.sh) the result is order-independent.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

READ, WRITE, AUTH, CERT, PEER = 0x01, 0x02, 0x04, 0x08, 0x10


@dataclass
class Node:
    """Just enough of node.Node (node/node.go) for the quorum logic."""
    id: int
    signers: List[int] = field(default_factory=list)
    address: str = "x"
    active: bool = True


@dataclass
class Vertex:
    instance: Optional[Node]
    edges: Dict[int, "Vertex"] = field(default_factory=dict)


class Graph:
    def __init__(self):
        self.vertices: Dict[int, Vertex] = {}
        self.revoked: Dict[int, Optional[Node]] = {}
        self.self_: List[Vertex] = []

    def add_nodes(self, nodes):                       # graph.go:46-75
        res = []
        for n in nodes:
            if n.id in self.revoked:
                continue
            me = self.vertices.get(n.id)
            if me is None:
                me = Vertex(n)
                self.vertices[n.id] = me
            else:
                me.instance = n
            for signer in n.signers:
                if signer in self.revoked:
                    continue
                v = self.vertices.get(signer)
                if v is None:
                    v = Vertex(None)
                    self.vertices[signer] = v
                v.edges[n.id] = me
            res.append(n)
        return res

    def set_self_nodes(self, nodes):                  # graph.go:77-88
        for n in nodes:
            v = self.vertices.get(n.id)
            if v is None or v.instance is None:
                self.add_nodes([n])
                v = self.vertices[n.id]
            self.self_.append(v)

    def remove_nodes(self, nodes):                    # graph.go:90-108
        for n in nodes:
            for v in self.vertices.values():
                v.edges.pop(n.id, None)
            self.vertices.pop(n.id, None)
            for i, s in enumerate(self.self_):
                if s.instance.id == n.id:
                    del self.self_[i]
                    break

    def revoke(self, n):                              # graph.go:131-140
        v = self.vertices.get(n.id)
        inst = None
        if v is not None:
            inst = v.instance
            if inst is not None:
                self.remove_nodes([inst])
        self.revoked[n.id] = inst

    def get_self_id(self) -> int:                     # graph.go:262-267
        if not self.self_ or self.self_[0].instance is None:
            return 0
        return self.self_[0].instance.id

    def get_peers(self):                              # graph.go:117-125
        sid = self.get_self_id()
        return [v.instance for v in self.vertices.values()
                if v.instance is not None and v.instance.id != sid]

    @staticmethod
    def _bfs(v: Vertex, proc):                        # graph.go:420-438
        seen = {v.instance.id}
        q = [(v, 0)]
        while q:
            vd = q.pop(0)
            if proc(vd):
                return
            for id_, e in vd[0].edges.items():
                if id_ not in seen:
                    q.append((e, vd[1] + 1))
                    seen.add(id_)

    def get_reachable_nodes(self, sid: int, distance: int):   # graph.go:279-295
        nodes = []
        v = self.vertices.get(sid)
        if v is None:
            return nodes

        def proc(vd):
            if distance >= 0 and vd[1] > distance:
                return True
            if vd[0].instance is not None:
                nodes.append(vd[0].instance)
            return False
        self._bfs(v, proc)
        return nodes

    @staticmethod
    def _bidirect(v: Vertex, clique: List[Vertex]) -> bool:   # graph.go:370-380
        for c in clique:
            if v.instance.id not in c.edges:
                return False
            if c.instance.id not in v.edges:
                return False
        return True

    def _find_maximal_clique(self, s: Vertex):        # graph.go:333-368
        clique = [s]
        for v in self.vertices.values():
            if v.instance is None or v is s:
                continue
            if self._bidirect(v, clique):
                clique.append(v)
        for v in self.vertices.values():
            if (v.instance is not None and v is not s
                    and not any(v is c for c in clique) and self._bidirect(v, [s])):
                return None          # "found more than one maximal cliques"
        return {"nodes": [c.instance for c in clique], "weight": 0}

    def get_cliques(self, sid: int, distance: int):   # graph.go:297-319
        cliques = []
        v = self.vertices.get(sid)
        if v is None or v.instance is None:
            return cliques

        def in_clique(e):
            return any(e.id == n.id for c in cliques for n in c["nodes"])

        def proc(vd):
            if distance >= 0 and vd[1] > distance:
                return True
            if vd[0].instance is not None and not in_clique(vd[0].instance):
                clique = self._find_maximal_clique(vd[0])
                if clique is not None:
                    for i in v.edges:                 # putWeight graph.go:385-393
                        for n in clique["nodes"]:
                            if n.id == i:
                                clique["weight"] += 1
                    cliques.append(clique)
            return False
        self._bfs(v, proc)
        return cliques


@dataclass
class QC:                                             # wotqs.go:16-22
    nodes: List[Node]
    f: int
    min: int
    threshold: int
    suff: int


def new_qc(self_id: int, clique_nodes, weight: int, rw: int) -> Optional[QC]:   # wotqs.go:36-70
    if rw & PEER:
        nodes = [n for n in clique_nodes if n.id != self_id]
    else:
        nodes = list(clique_nodes)
    n = len(nodes)
    if n == 0:
        return None
    if rw == WRITE:
        return QC(nodes, 0, 0, 0, 0)
    f = (n - 1) // 3
    if f >= 1:
        mn = 3 * f + 1
        threshold = 2 * f + 1
        suff = f + (n - f) // 2 + 1
        if rw & (CERT | READ):
            threshold = f + 1
        if weight <= n - suff:
            suff = 0
        return QC(nodes, f, mn, threshold, suff)
    return None


def intersection(s1, s2):                             # wotqs.go:195-206
    """Elements of s1 (duplicates kept!) whose id occurs in s2."""
    ret = []
    for n1 in s1:
        for n2 in s2:
            if n1.id == n2.id:
                ret.append(n1)
                break
    return ret


class Quorum:                                         # wotqs.go:24-26,132-193
    def __init__(self, qcs: List[QC]):
        self.qcs = qcs

    def nodes(self):
        return [n for qc in self.qcs for n in qc.nodes if n.active and n.address != ""]

    def is_quorum(self, nodes) -> bool:
        if not self.qcs:
            return False
        for qc in self.qcs:
            if qc.f > 0 and len(intersection(nodes, qc.nodes)) < qc.min:
                return False
        return True

    def is_threshold(self, nodes) -> bool:
        if not self.qcs:
            return False
        for qc in self.qcs:
            if qc.threshold > 0 and len(intersection(nodes, qc.nodes)) < qc.threshold:
                return False
        return True

    def is_sufficient(self, nodes) -> bool:
        for qc in self.qcs:
            if qc.suff > 0 and len(intersection(nodes, qc.nodes)) >= qc.suff:
                return True
        return False

    def reject(self, nodes) -> bool:
        for qc in self.qcs:
            if qc.f == 0 or len(intersection(nodes, qc.nodes)) <= qc.f:
                return False
        return True

    def get_threshold(self) -> int:
        return sum(qc.threshold for qc in self.qcs)


class WotQS:                                          # wotqs.go:12-14,72-127
    def __init__(self, g: Graph):
        self.g = g

    def _complement(self, u, c: List[QC], e: List[QC], rw: int) -> List[QC]:   # wotqs.go:72-93
        nodes = [n1 for n1 in u if not any(n1.id == n2.id for qc in c for n2 in qc.nodes)]
        q = new_qc(self.g.get_self_id(), nodes, 0, rw)
        if q is not None:
            e = e + [q]
        return e

    def _get_quorum_from(self, rw: int, s: int, distance: int) -> Quorum:      # wotqs.go:95-115
        qcs = []
        for c in self.g.get_cliques(s, distance):
            qc = new_qc(self.g.get_self_id(), c["nodes"], c["weight"], rw | AUTH)
            if qc is not None:
                qcs.append(qc)
        if rw & (READ | WRITE):
            out = list(qcs) if (rw & AUTH) else []
            out = self._complement(self.g.get_reachable_nodes(s, distance), qcs, out, READ)
            if rw & WRITE:
                out = self._complement(self.g.get_peers(), qcs + out, out, WRITE)
            qcs = out
        return Quorum(qcs)

    def choose_quorum(self, rw: int) -> Quorum:       # wotqs.go:117-127
        if rw & CERT:
            distance = 0
        elif rw & AUTH:
            distance = 1
        else:
            distance = 2
        return self._get_quorum_from(rw, self.g.get_self_id(), distance)


# ---- tally drivers (protocol/client.go:181-230) ---------------------------------------------

def max_timestamped_value(m: dict, q: Quorum):
    """protocol/client.go:189-205.  m: {t: {value_bytes: [Node, ...]}} in insertion order.
    Returns (value, t) or None ("errInProgress").  Only the max-t bucket is inspected; among
    several qualifying values at max t the reference's winner depends on Go map order — this
    restatement returns the first in insertion order."""
    maxt, maxvl = 0, None
    for t, vl in m.items():
        if t >= maxt:
            maxt, maxvl = t, vl
    if maxvl is None:
        return None
    for v, l in maxvl.items():
        if q.is_threshold(l):
            return v, maxt
    return None


READ_VALUE, READ_REJECTED, READ_EXHAUSTED = 0, 1, 2


def read_decide(responses, q: Quorum):
    """Client.Read's multicast callback, protocol/client.go:250-268, run over `responses` in the order given
    (= arrival order).  responses: list of (Node, err, t, value) — err truthy when transport.Multicast delivered an
    error for that peer (signature check of the transport message failed, nonce mismatch, no answer...; then t / value
    are ignored), else (t, value) are what processResponse (:207-230) parsed out of the answer.
    Returns (kind, decided_at, value, t): kind READ_VALUE when maxTimestampedValue first returned a value (decided_at =
    number of responses consumed, 1-based), READ_REJECTED when q.Reject(failure) first held (the reference then reports
    majorityError(errs, ErrInsufficientNumberOfValidResponses)), READ_EXHAUSTED when the multicast ended without either
    (ErrInsufficientNumberOfResponses, :267).  Once `ch` has been served the callback keeps collecting but decides
    nothing more, so the FIRST decisive response fixes the result."""
    m, failure = {}, []
    for k, (node, err, t, value) in enumerate(responses):
        if not err:
            m.setdefault(t, {}).setdefault(value, []).append(node)            # processResponse: m[t][string(val)]
            r = max_timestamped_value(m, q)
            if r is not None:
                return READ_VALUE, k + 1, r[0], r[1]
        else:
            failure.append(node)
            if q.reject(failure):
                return READ_REJECTED, k + 1, None, 0
    return READ_EXHAUSTED, len(responses), None, 0


def revoke_scan(m: dict, signers_of):
    """Client.revoke's scan, protocol/client.go:304-346.  m: {t: {value: [signedValue, ...]}} (insertion order =
    one of Go's legal map orders); signers_of(signedValue) -> [signer id, ...] (CollectiveSignature.Signers(ss)).
    Returns the ids the reference would revoke, in discovery order: a signer is remembered under the first value
    bucket ("round") it is seen in — dup_map[id] only ever holds that one round — and revoked the first time it
    shows up in another bucket of the same t; t == 0 is skipped."""
    revoked = []
    for t, vl in m.items():
        if t == 0:
            continue
        dup_map = {}
        rnd = 0
        for _, l in vl.items():
            for sv in l:
                for sid in signers_of(sv):
                    if sid in dup_map:
                        for it in dup_map[sid]:
                            if it != rnd:
                                if sid in revoked:
                                    break
                                revoked.append(sid)
                    else:
                        dup_map[sid] = [rnd]
            rnd += 1
    return revoked
