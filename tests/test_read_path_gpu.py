"""GPU parity suite of the read path at BASELINE sizes: K1 + the wotqs read decision (Client.Read, protocol/client.go:250-268)
through the host C ABI — chunked, pipelined, pageable and bftq_host_alloc inputs — against the oracle."""
import os

import numpy as np
import pytest

from bftkv_b200 import Engine, workload
from bftkv_b200.engine import NO_WINNER, READ_EXHAUSTED, READ_REJECTED, READ_VALUE
from oracle import c_oracle
from test_read_decide import QUORUMS, random_ops

pytestmark = pytest.mark.gpu
NCPU = max(1, min(32, os.cpu_count() or 1))


@pytest.mark.parametrize("qi", range(len(QUORUMS)))
def test_read_decide_kernel_matches_oracle(engine, qi):
    qcs = QUORUMS[qi]
    rng = np.random.default_rng(900 + qi)
    off, idx, st, ts, vid = random_ops(rng, 6000, list(range(0, 34)), 32)
    q = engine.quorum_create(qcs)
    dec, win, at = engine.read_decide_batch(q, off, idx.astype(np.uint32), st, ts, vid)
    rdec, rwin, rat = c_oracle.read_decide_batch(qcs, off, idx, st, ts, vid)
    assert np.array_equal(dec, rdec) and np.array_equal(win, rwin) and np.array_equal(at, rat)
    if qcs:
        assert set(np.unique(dec)) == {READ_VALUE, READ_REJECTED, READ_EXHAUSTED}
    engine.quorum_destroy(q)


def _oracle_read(pool, w, qcs):
    ns, es = [k["n"] for k in pool["keys"]], [k["e"] for k in pool["keys"]]
    st = c_oracle.rsa_verify_batch(ns, es, w["key_idx"], w["sig"], w["digest"], threads=NCPU)
    st[w["pre_status"] != 0] = w["pre_status"][w["pre_status"] != 0]
    dec, win, at = c_oracle.read_decide_batch(qcs, w["op_off"], w["key_idx"].astype(np.uint64), st, w["ts"], w["value_id"])
    return st, dec, win, at


def test_verify_read_fused_r31_hard_mix(built):
    """BASELINE configs[4]'s shape (31-replica quorum: f = 10, READ threshold 11) on 8192 operations whose responses
    arrive in random order, with enough invalid / missing / stale responders that every arm of the decision occurs."""
    R, M = 31, 8192
    pool = workload.make_verify_batch(8192, n_keys=R, seed=0xBF7C0010, corrupt_rate=0.0, unknown_rate=0.0)
    w = workload.make_read_ops(pool, M, R, seed=0xBF7C0006, mix=workload.HARD_MIX, shuffle_arrival=True)
    qcs = [(10, 31, 11, 21, list(range(31)))]
    ref_st, ref_dec, ref_win, ref_at = _oracle_read(pool, w, qcs)
    assert np.array_equal(ref_st, w["expect_status"])
    kinds = {int(k): int((ref_dec == k).sum()) for k in (0, 1, 2)}
    assert min(kinds.values()) > 100, kinds
    e = Engine(0)
    e.register_rsa_keys([k["n"] for k in pool["keys"]], [k["e"] for k in pool["keys"]])
    q = e.quorum_create(qcs)
    # pageable inputs (numpy): 253 952 tuples = 16 chunks through the staging ring
    st, dec, win, at = e.verify_read_batch(q, w["op_off"], w["key_idx"], w["sig"], w["digest"], w["ts"], w["value_id"], pre_status=w["pre_status"])
    assert np.array_equal(st, ref_st)
    assert np.array_equal(dec, ref_dec) and np.array_equal(win, ref_win) and np.array_equal(at, ref_at)
    # early decisions exist: Read answers before all 31 responses are in
    assert (at[dec == READ_VALUE] < R).any() and (win[dec != READ_VALUE] == NO_WINNER).all()
    # the same through page-locked blobs from bftq_host_alloc (DMA'd in place) and caller-provided outputs
    pin = {k: e.host_copy(w[k]) for k in ("op_off", "key_idx", "sig", "digest", "ts", "value_id", "pre_status")}
    outs = (e.host_alloc(M * R, np.uint8), e.host_alloc(M, np.uint8), e.host_alloc(M, np.uint32), e.host_alloc(M, np.uint32))
    e.verify_read_batch(q, pin["op_off"], pin["key_idx"], pin["sig"], pin["digest"], pin["ts"], pin["value_id"], pre_status=pin["pre_status"],
                        out_status=outs[0], out_decision=outs[1], out_winner=outs[2], out_decided_at=outs[3])
    assert np.array_equal(outs[0], ref_st) and np.array_equal(outs[1], ref_dec) and np.array_equal(outs[2], ref_win) and np.array_equal(outs[3], ref_at)
    # the four predicate bits over all responses (non-read fused form), chunked as well
    st2, bits2, _ = e.verify_tally_batch(q, w["op_off"], w["key_idx"], w["sig"], w["digest"], pre_status=w["pre_status"])
    assert np.array_equal(st2, ref_st)
    assert np.array_equal(bits2, c_oracle.tally_batch(qcs, w["op_off"], w["key_idx"].astype(np.uint64), ref_st))
    # ragged operations (0..31 responders) through the chunker: offsets rebased per chunk
    rng = np.random.default_rng(3)
    sizes = rng.integers(0, 32, 3000)
    off = np.zeros(3001, np.uint32)
    off[1:] = np.cumsum(sizes)
    n = int(off[-1])
    sel = rng.integers(0, M * R, n)
    rg = {k: np.ascontiguousarray(w[k][sel]) for k in ("key_idx", "sig", "digest", "ts", "value_id", "pre_status")}
    rg["op_off"] = off
    r_st, r_dec, r_win, r_at = _oracle_read(pool, rg, qcs)
    os.environ["BFTQ_HOST_CHUNK"] = "16384"
    st3, dec3, win3, at3 = e.verify_read_batch(q, off, rg["key_idx"], rg["sig"], rg["digest"], rg["ts"], rg["value_id"], pre_status=rg["pre_status"])
    assert np.array_equal(st3, r_st) and np.array_equal(dec3, r_dec) and np.array_equal(win3, r_win) and np.array_equal(at3, r_at)
    for a in list(pin.values()) + list(outs):
        e.host_free(a)
    e.quorum_destroy(q)
    e.close()


def test_config3_full_size_against_oracle(built):
    """BASELINE configs[2] at FULL size — 65 536 read ops x 16 replicas = 1 048 576 verifies — through the host call,
    every status and every decision compared with the oracle (not with the generator's expectation)."""
    R, M = 16, 65536
    pool = workload.make_verify_batch(16384, n_keys=R, seed=0xBF7C0011, corrupt_rate=0.0, unknown_rate=0.0)
    w = workload.make_read_ops(pool, M, R, seed=0xBF7C0004)
    qcs = [(5, 16, 6, 11, list(range(16)))]
    ref_st, ref_dec, ref_win, ref_at = _oracle_read(pool, w, qcs)
    e = Engine(0)
    e.register_rsa_keys([k["n"] for k in pool["keys"]], [k["e"] for k in pool["keys"]])
    q = e.quorum_create(qcs)
    st, dec, win, at = e.verify_read_batch(q, w["op_off"], w["key_idx"], w["sig"], w["digest"], w["ts"], w["value_id"], pre_status=w["pre_status"])
    assert np.array_equal(st, ref_st) and np.array_equal(st, w["expect_status"])
    assert np.array_equal(dec, ref_dec) and np.array_equal(win, ref_win) and np.array_equal(at, ref_at)
    assert (dec == READ_VALUE).sum() > 0.9 * M
    # all-responses form: IsThreshold|Reject bits and first-in-responder-order winner; the winner must be a good
    # response of the maximum t whose bucket passes IsThreshold (same rule as test_tally_lagrange_digest_gpu.py)
    st2, bits2, win2 = e.verify_tally_batch(q, w["op_off"], w["key_idx"], w["sig"], w["digest"], pre_status=w["pre_status"], ts=w["ts"],
                                            value_id=w["value_id"])
    assert np.array_equal(st2, ref_st)
    okm = (ref_st == 0).reshape(M, R)
    tsm, vm = w["ts"].reshape(M, R), w["value_id"].reshape(M, R)
    maxt = np.where(okm, tsm, 0).max(axis=1)
    for v in (0, 1):
        cnt = (okm & (tsm == maxt[:, None]) & (vm == v)).sum(axis=1)
        passes = cnt >= 6
        sel = (win2 != NO_WINNER) & (vm[np.arange(M), np.minimum(win2, R - 1)] == v)
        assert passes[sel].all()
    any_pass = np.zeros(M, bool)
    for v in (0, 1):
        any_pass |= (okm & (tsm == maxt[:, None]) & (vm == v)).sum(axis=1) >= 6
    assert np.array_equal(win2 != NO_WINNER, any_pass)
    assert np.array_equal((bits2 & 2) != 0, any_pass)
    assert np.array_equal((bits2 & 8) != 0, (~okm).sum(axis=1) > 5)
    e.quorum_destroy(q)
    e.close()
