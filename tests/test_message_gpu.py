"""GPU parity suite of bftq_message_verify_batch — PGPMessage.Decrypt's signature half (crypto_pgp.go:453-471), the
per-response check of every multicast — through the C ABI against the oracle: GnuPG-written messages, Go-writer-shaped
messages, and thousands of mutated / truncated / spliced ones."""
import json
import os
import random

import pytest

from bftkv_b200 import Engine, workload
from bftkv_b200.crypto_gpu import Keyring, Message
from oracle import pgp_oracle as pgp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compare(got, ref, ctx):
    assert got["err"] == ref.err, (ctx, got["err"], ref.err)
    if ref.err in (None, pgp.ERR_INVALID_SIGNATURE):
        assert got["plain"] == ref.plain and got["nonce"] == ref.nonce, ctx
    if ref.err not in (pgp.ERR_DECRYPTION_FAILED, pgp.ERR_TRANSPORT_SECURITY, pgp.ERR_MESSAGE_UNSUPPORTED) and not (ref.err == pgp.ERR_MESSAGE_BODY and ref.signed_by_key_id == 0):
        assert got["signed_by_key_id"] == ref.signed_by_key_id and got["signer_known"] == ref.signer_known, ctx


def test_gnupg_messages(engine):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_messages.json")))
    kr = Keyring(engine)
    kr.register(bytes.fromhex(g["keyring"]))
    ents = pgp.read_entities(bytes.fromhex(g["keyring"]))
    msgs = [bytes.fromhex(c["msg"]) for c in g["cases"]]
    got = Message(kr).decrypt_verify_batch(msgs)
    for c, m, r in zip(g["cases"], msgs, got):
        compare(r, pgp.message_verify(ents, m), c["name"])
        if c["signer"] == "m01" and c["name"] not in ("compressed-default", "name-not-base64"):
            assert (r["err"] is None) == c["gpg_good"], c["name"]           # GnuPG's own verdict
    kr.close()


def test_go_shaped_and_mutated_messages(engine):
    keys = workload.load_keys(4)
    blocks, kids = [], []
    for i, k in enumerate(keys):
        b, kid = workload.pgp_public_key_block(k, workload._private_key(k), b"n%d <n%d@x>" % (i, i))
        blocks.append(b); kids.append(kid)
    ring = b"".join(blocks[:3])                                    # key 3 stays outside the keyring
    kr = Keyring(engine)
    kr.register(ring)
    ents = pgp.read_entities(ring)
    rng = random.Random(4242)
    msgs = []
    for i in range(700):
        ki = rng.randrange(4)
        n = rng.choice([0, 1, 7, 56, 100, 300, 1000, 5000, 17000])
        plain = bytes(rng.randrange(256) for _ in range(min(n, 64))) * (n // 64 + 1)
        plain = plain[:n]
        hash_id = rng.choice([8, 8, 8, 10, 2, 9, 11])
        m = bytearray(workload.make_transport_message(keys[ki], kids[ki], plain, bytes(rng.randrange(256) for _ in range(8)), hash_id=hash_id))
        mode = rng.random()
        if mode < 0.25 and len(m):                                 # byte flips anywhere (framing, one-pass, literal, signature)
            for _ in range(rng.randint(1, 3)):
                m[rng.randrange(len(m))] ^= 1 << rng.randrange(8)
        elif mode < 0.35:
            del m[rng.randrange(len(m)):]
        elif mode < 0.42:
            pos = rng.randrange(len(m))
            m[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 5)))
        elif mode < 0.47:                                          # a second one-pass packet (the last one wins) / a stray signature packet first
            extra = workload.one_pass_packet(0, 8, 1, kids[rng.randrange(4)]) if rng.random() < 0.5 else workload.go_signature_packet(keys[0], kids[0], 8, b"x", 5)
            m = bytearray(extra) + m
        elif mode < 0.5:                                           # text-mode one-pass over a binary signature
            m = bytearray(workload.one_pass_packet(1, hash_id, 1, kids[ki])) + m[15:]
        msgs.append(bytes(m))
    got = Message(kr).decrypt_verify_batch(msgs)
    kinds = {}
    for i, (m, r) in enumerate(zip(msgs, got)):
        ref = pgp.message_verify(ents, m)
        compare(r, ref, (i, m[:40].hex()))
        kinds[ref.err] = kinds.get(ref.err, 0) + 1
    assert kinds[None] > 300 and kinds[pgp.ERR_INVALID_SIGNATURE] > 50 and kinds[pgp.ERR_DECRYPTION_FAILED] >= 2 and kinds[pgp.ERR_MESSAGE_BODY] > 10, kinds
    kr.close()
