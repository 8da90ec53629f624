"""Client.Read's decision (protocol/client.go:250-268) in arrival order: the Python restatement (oracle/wotqs_oracle.read_decide, a
line-by-line walk of the callback) against the batch C oracle, and both against hand-worked cases.  CPU only."""
import numpy as np

from oracle import c_oracle, wotqs_oracle as wq
from oracle.wotqs_oracle import Node

QUORUMS = [
    [],
    [(1, 4, 2, 3, [0, 1, 2, 3])],                                              # config 1: 4-node clique, READ threshold 2
    [(5, 16, 6, 11, list(range(16)))],                                         # config 3
    [(10, 31, 11, 21, list(range(31)))],                                       # config 5
    [(1, 4, 2, 3, [0, 1, 2, 3]), (1, 4, 2, 0, [10, 11, 12, 13])],              # two cliques: IsThreshold needs both
]


def random_ops(rng, n_ops, id_pool, max_r):
    off, idx, st, ts, vid = [0], [], [], [], []
    for _ in range(n_ops):
        r = int(rng.integers(0, max_r + 1))
        for _ in range(r):
            idx.append(int(rng.choice(id_pool)))                                # duplicates allowed (wotqs.go:195-206 counts them)
            st.append(int(rng.choice([0, 1, 4, 6], p=[0.7, 0.12, 0.08, 0.1])))
            ts.append([7, 6, 9, 0, 2 ** 64 - 1][int(rng.choice(5, p=[0.7, 0.2, 0.05, 0.04, 0.01]))])
            vid.append(int(rng.choice([0, 1, 2], p=[0.75, 0.2, 0.05])))
        off.append(len(idx))
    z = lambda a, t: np.array(a if a else [0], t)[:len(a)]
    return np.array(off, np.uint32), z(idx, np.uint64), z(st, np.uint8), z(ts, np.uint64), z(vid, np.uint32)


def python_decide(qcs, off, idx, st, ts, vid, i):
    quorum = wq.Quorum([wq.QC([Node(m) for m in mem], f, mn, th, sf) for f, mn, th, sf, mem in qcs])
    resp = [(Node(int(idx[p])), st[p] != 0, int(ts[p]), int(vid[p])) for p in range(off[i], off[i + 1])]
    return wq.read_decide(resp, quorum)


def test_c_oracle_matches_python_restatement():
    for qi, qcs in enumerate(QUORUMS):
        rng = np.random.default_rng(500 + qi)
        off, idx, st, ts, vid = random_ops(rng, 1500, list(range(0, 34)), 32)
        dec, win, at = c_oracle.read_decide_batch(qcs, off, idx, st, ts, vid)
        seen = set()
        for i in range(1500):
            kind, k, value, t = python_decide(qcs, off, idx, st, ts, vid, i)
            assert (dec[i], at[i]) == (kind, k), (qi, i)
            if kind == wq.READ_VALUE:
                p = off[i] + win[i]
                assert st[p] == 0 and int(vid[p]) == value and int(ts[p]) == t
                # the winner is the FIRST responder of its bucket
                assert all(not (st[p2] == 0 and ts[p2] == ts[p] and vid[p2] == vid[p]) for p2 in range(off[i], p))
            else:
                assert win[i] == 0xFFFFFFFF
            seen.add(int(kind))
        if qcs:
            assert seen == {0, 1, 2}, (qi, seen)


def test_hand_worked_cases():
    q = [(1, 4, 2, 3, [0, 1, 2, 3])]                                            # f = 1, READ threshold 2

    def run(rows):
        off = np.array([0, len(rows)], np.uint32)
        idx = np.array([r[0] for r in rows], np.uint64)
        st = np.array([r[1] for r in rows], np.uint8)
        ts = np.array([r[2] for r in rows], np.uint64)
        vid = np.array([r[3] for r in rows], np.uint32)
        d, w, a = c_oracle.read_decide_batch(q, off, idx, st, ts, vid)
        return int(d[0]), int(w[0]), int(a[0])
    # two equal answers at t = 7: decided at the second response, winner = its bucket's first responder
    assert run([(0, 0, 7, 0), (1, 0, 7, 0), (2, 0, 7, 0)]) == (0, 0, 2)
    # the stale bucket reaches the threshold first: Read returns the STALE value although a newer one arrives later
    assert run([(0, 0, 6, 1), (1, 0, 6, 1), (2, 0, 7, 0), (3, 0, 7, 0)]) == (0, 0, 2)
    # a lone newer answer hides a full older bucket (only the max-t buckets are inspected, client.go:191-198)
    assert run([(0, 0, 7, 0), (1, 0, 6, 1), (2, 0, 6, 1), (3, 0, 6, 1)]) == (2, 0xFFFFFFFF, 4)
    # two failures out of f = 1: rejected at the second failure, later good answers change nothing
    assert run([(0, 1, 0, 0), (1, 6, 0, 0), (2, 0, 7, 0), (3, 0, 7, 0)]) == (1, 0xFFFFFFFF, 2)
    # ... but a threshold reached BEFORE the second failure stands
    assert run([(2, 0, 7, 0), (0, 1, 0, 0), (3, 0, 7, 0), (1, 6, 0, 0)]) == (0, 0, 3)
    # a duplicated responder counts twice (intersection keeps duplicates of the input list)
    assert run([(2, 0, 7, 0), (2, 0, 7, 0)]) == (0, 0, 2)
    # responders outside the quorum never count
    assert run([(8, 0, 7, 0), (9, 0, 7, 0), (0, 0, 7, 0)]) == (2, 0xFFFFFFFF, 3)
    # no responses at all
    assert run([]) == (2, 0xFFFFFFFF, 0)
