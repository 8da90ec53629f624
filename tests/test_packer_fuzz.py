"""CPU suite: the C++ OpenPGP packer's stream walk (bftq_signature_parse / bftq_signature_signers,
host only) against the oracle's restatement on GnuPG-made streams and on thousands of mutated /
truncated / spliced ones — the parser eats untrusted network input, so it must neither crash nor
disagree with the reference's accept/skip/stop structure."""
import random

import pytest

from bftkv_b200.crypto_gpu import Keyring, Signature
from oracle import pgp_oracle as pgp

RING = ["a01", "a02", "a03", "a04", "u01"]


def oracle_walk(ents, data, collective):
    """(list of (issuer, hash id) per call that reached a known-issuer signature, failed)."""
    r = pgp.Reader(data)
    calls, failed = [], False
    while r.remaining() > 0:
        try:
            # replicate check_detached_signature's parse half
            while True:
                while True:
                    pk = pgp.read_packet(r)
                    if pk is None:
                        raise pgp.UnknownIssuer()
                    if pk[0] in pgp.KNOWN_TAGS:
                        break
                if pk[0] != 2:
                    raise pgp.StructuralError("non signature packet found")
                sig = pgp.parse_signature(pk[1])
                if sig.issuer_key_id is None:
                    raise pgp.StructuralError("no issuer")
                if pgp.keys_by_id_usage(ents, sig.issuer_key_id, pgp.KEY_FLAG_SIGN):
                    break
            calls.append((sig.issuer_key_id, sig.hash_id))
        except pgp.PGPError:
            if collective:
                continue
            failed = True
            break
    return calls, failed


@pytest.fixture(scope="module")
def env(golden, built):
    kr = Keyring(None)                       # parse-only keyring: no GPU involved
    ents = []
    for n in RING:
        blob = bytes.fromhex(golden["keys"][n]["pub"])
        kr.register(blob)
        ents += pgp.read_entities(blob)
    yield kr, Signature(kr), ents
    kr.close()


def test_keyring_parse_matches_oracle(env, golden):
    kr, sig, ents = env
    assert kr.get_keyring() == [e.primary_key.key_id for e in ents]
    for e in ents:
        assert kr.certifiers(e.primary_key.key_id) == e.certifier_ids


def test_walk_on_valid_and_mutated_streams(env, golden):
    kr, sig, ents = env
    rng = random.Random(2024)
    sigs = [bytes.fromhex(c["sig"]) for c in golden["cases"]]
    pubs = [bytes.fromhex(golden["keys"][n]["pub"]) for n in golden["keys"]]
    n_checked = 0
    for trial in range(3000):
        parts = [rng.choice(sigs) for _ in range(rng.randint(0, 4))]
        if rng.random() < 0.15:
            parts.insert(rng.randrange(len(parts) + 1), rng.choice(pubs)[:rng.randint(1, 400)])      # key-block fragments
        if rng.random() < 0.2:
            parts.insert(rng.randrange(len(parts) + 1), bytes([0xC0 | rng.randrange(64)]) + bytes([rng.randrange(192)]) + bytes(rng.randrange(256) for _ in range(rng.randrange(8))))
        data = bytearray(b"".join(parts))
        mode = rng.random()
        if data and mode < 0.45:                                    # byte flips
            for _ in range(rng.randint(1, 4)):
                data[rng.randrange(len(data))] ^= 1 << rng.randrange(8)
        elif data and mode < 0.6:                                   # truncation
            del data[rng.randrange(len(data)):]
        elif data and mode < 0.7:                                   # random insertion
            pos = rng.randrange(len(data))
            data[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 6)))
        data = bytes(data)
        for collective in (False, True):
            got = sig.parse(data, collective)
            ref = oracle_walk(ents, data, collective)
            assert got == ref, (trial, collective, data.hex())
            n_checked += 1
        try:
            ref_signers = pgp.signers(ents, data)
        except RuntimeError:                                        # the reference would dereference nil here
            ref_signers = None
        try:
            got_signers = sig.signers(data)
        except Exception:
            got_signers = None
        assert got_signers == ref_signers, (trial, data.hex())
    assert n_checked == 6000


def test_pure_garbage_never_crashes(env):
    kr, sig, ents = env
    rng = random.Random(7)
    for _ in range(2000):
        data = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 600)))
        for collective in (False, True):
            assert sig.parse(data, collective) == oracle_walk(ents, data, collective)


def test_plan_measure_counts_tuples_like_the_oracle(built):
    """Host half of the batch verify only (no GPU): the number of (signature packet, candidate key)
    tuples the chunked, multi-threaded packer composes equals the oracle's count of known-issuer
    packets, whatever the thread / chunk split."""
    import ctypes as C
    import os
    from bftkv_b200 import _lib, workload
    from bftkv_b200.crypto_gpu import _blob
    w = workload.make_pgp_verify_batch(3000, n_keys=5, corrupt_rate=0.02, unknown_rate=0.03)
    ents = pgp.read_entities(w["keyring"])
    expect = sum(len(oracle_walk(ents, s, False)[0]) for s in w["sigs"])
    assert 0 < expect < 3000                                   # some issuers are outside the keyring
    kr = Keyring(None)
    assert kr.register(w["keyring"]) == 5
    lib = _lib.load()
    tb, to = _blob(w["tbs"])
    sb, so = _blob(w["sigs"])
    p = lambda a: C.c_void_p(a.ctypes.data)
    for threads, chunk in ((1, "100000"), (3, "257"), (8, "64")):
        os.environ["BFTQ_PLAN_CHUNK"] = chunk
        try:
            nt, sec = C.c_uint64(), C.c_double()
            _lib.check(lib.bftq_signature_plan_measure(kr._h, p(tb), p(to), p(sb), p(so), len(w["tbs"]), threads, C.byref(nt), C.byref(sec)))
        finally:
            del os.environ["BFTQ_PLAN_CHUNK"]
        assert nt.value == expect, (threads, chunk, nt.value, expect)
    kr.close()


def test_gpu_fast_parser_agrees_with_host_parser(golden):
    """K0's packet parser (bftkv_b200/csrc/pgp_fastparse.hpp, the same source the kernel compiles) built for
    the host and fuzzed against pgp_host.hpp: over a million mutated / truncated / spliced streams, whenever
    it answers "fast" the reference-shaped parser reads the same stream to the same fields (a "fallback"
    answer is always allowed — those items go through the host packer)."""
    import os
    import struct
    import subprocess
    from bftkv_b200 import workload
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "harness", "fastparse_host")
    src = os.path.join(root, "tests", "harness", "fastparse_host.cpp")
    deps = [src] + [os.path.join(root, "bftkv_b200", "csrc", f) for f in ("pgp_fastparse.hpp", "pgp_host.hpp")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, src])
    w = workload.make_pgp_verify_batch(48, n_keys=4)
    seeds = list(w["sigs"]) + [bytes.fromhex(c["sig"]) for c in golden["cases"]]
    for s in list(seeds[:8]):                                   # the same bodies under new-format headers
        body = s[3:]
        seeds.append(bytes([0xC2, 0xFF]) + struct.pack(">I", len(body)) + body)
        seeds.append(bytes([0xC2, ((len(body) - 192) >> 8) + 192, (len(body) - 192) & 0xFF]) + body)
    path = os.path.join(root, "tests", "harness", "fastparse_seeds.bin")
    with open(path, "wb") as f:
        f.write(b"".join(struct.pack(">I", len(s)) + s for s in seeds))
    r = subprocess.run([exe, path, "1500000"], capture_output=True, text=True)
    os.remove(path)
    tried, fast, bad = (int(x) for x in r.stdout.split())
    assert r.returncode == 0 and bad == 0, r.stderr[-2000:]
    assert tried > 1500000 and fast > 100000                     # the fast path is actually exercised


def test_collective_walk_resynchronises_after_stray_bytes(env, golden):
    """CollectiveSignature.Verify ignores errors (crypto_pgp.go:489-497) and x/crypto's readHeader consumes ONE byte when
    the tag byte has no MSB, so stray low bytes a Byzantine responder slipped between two packets cost nothing: every
    honest signature around them is still reached.  Signature.Verify's strict walk fails at the first stray byte."""
    kr, sig, ents = env
    sigs = [bytes.fromhex(c["sig"]) for c in golden["cases"]]
    known = [s for s in sigs if oracle_walk(ents, s, False) == ([oracle_walk(ents, s, False)[0][0]] if oracle_walk(ents, s, False)[0] else [], False)
             and len(oracle_walk(ents, s, False)[0]) == 1][:3]
    assert len(known) == 3
    issuers = [oracle_walk(ents, s, False)[0][0] for s in known]
    for junk in (b"\x00", b"\x7f", b"\x01\x02\x03", b"\x00" * 17):
        data = known[0] + junk + known[1] + junk + known[2]
        assert oracle_walk(ents, data, True) == (issuers, False)
        assert sig.parse(data, True) == (issuers, False)
        assert sig.parse(data, False) == ([issuers[0]], True) == oracle_walk(ents, data, False)
        assert sig.parse(junk + known[0], True) == ([issuers[0]], False) == oracle_walk(ents, junk + known[0], True)
