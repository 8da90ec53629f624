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
