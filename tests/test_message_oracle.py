"""CPU suite: the oracle's restatement of PGPMessage.Decrypt's signature half (oracle/pgp_oracle.message_verify) against
GnuPG's verdict on GnuPG-written messages (tests/golden/golden_messages.json, made by make_golden_messages.py) and on
Go-writer-shaped messages (workload.make_transport_message: new-format headers, every Write a run of partial chunks)."""
import json
import os

from bftkv_b200 import workload
from oracle import pgp_oracle as pgp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_messages.json")))
    return g, pgp.read_entities(bytes.fromhex(g["keyring"]))


def test_oracle_agrees_with_gnupg_on_its_own_messages():
    g, ents = load()
    seen = set()
    for c in g["cases"]:
        r = pgp.message_verify(ents, bytes.fromhex(c["msg"]))
        if c["name"] == "compressed-default":
            assert r.err == pgp.ERR_MESSAGE_UNSUPPORTED                       # stated divergence: the reference would inflate
            continue
        if c["name"] == "name-not-base64":
            assert r.err == pgp.ERR_MESSAGE_BODY                              # base64.DecodeString(FileName) fails before SignatureError is looked at
            continue
        if c["signer"] != "m01":
            # signer outside the keyring: m.SignedBy == nil, the packet behind the literal data is never read, err == nil
            assert r.err is None and not r.signer_known and r.signed_by_key_id != 0
            continue
        assert (r.err is None) == c["gpg_good"], c["name"]
        assert r.signer_known and r.nonce == bytes.fromhex(g["nonce"])
        if c["gpg_good"] and "textmode" not in c["name"]:
            assert r.plain == bytes.fromhex(c["body"])
        seen.add(c["gpg_good"])
    assert seen == {True, False}


def test_go_writer_shaped_messages():
    keys = workload.load_keys(3)
    blocks, kids = [], []
    for i, k in enumerate(keys):
        b, kid = workload.pgp_public_key_block(k, workload._private_key(k), b"n%d <n%d@x>" % (i, i))
        blocks.append(b); kids.append(kid)
    ents = pgp.read_entities(b"".join(blocks[:2]))
    for n in (0, 1, 2, 3, 511, 512, 513, 16384, 40000):
        plain = bytes((i * 13 + n) & 0xFF for i in range(n))
        m = workload.make_transport_message(keys[1], kids[1], plain, b"\x09" * 8)
        assert m[0] == 0xC4 and 0xCB in m[:20]
        r = pgp.message_verify(ents, m)
        assert r.err is None and r.plain == plain and r.nonce == b"\x09" * 8 and r.signed_by_key_id == kids[1]
        if n:
            bad = bytearray(m)
            bad[m.index(b"\xcb") + 30 if n > 40 else len(m) - 5] ^= 4
            assert pgp.message_verify(ents, bytes(bad)).err is not None
    # the one-pass packet decides how the body is hashed: hash id or type that differ from the signature packet's fail
    plain = b"abc\ndef\n"
    sig = workload.go_signature_packet(keys[0], kids[0], 8, plain, 1)
    lit = workload.go_literal_packet(plain, b"AAAAAAAAAAA=")
    assert pgp.message_verify(ents, workload.one_pass_packet(0, 8, 1, kids[0]) + lit + sig).err is None
    assert pgp.message_verify(ents, workload.one_pass_packet(0, 10, 1, kids[0]) + lit + sig).err == pgp.ERR_INVALID_SIGNATURE
    assert pgp.message_verify(ents, workload.one_pass_packet(1, 8, 1, kids[0]) + lit + sig).err == pgp.ERR_INVALID_SIGNATURE   # text canonicalisation changes the digest
    assert pgp.message_verify(ents, workload.one_pass_packet(0, 8, 1, kids[0], is_last=0) + lit + sig).err == pgp.ERR_DECRYPTION_FAILED
    assert pgp.message_verify(ents, lit + sig).err == pgp.ERR_TRANSPORT_SECURITY
    assert pgp.message_verify(ents, workload.one_pass_packet(0, 8, 1, kids[0]) + lit).err == pgp.ERR_INVALID_SIGNATURE        # io.EOF where the signature should be
    assert pgp.message_verify(ents, workload.one_pass_packet(0, 8, 1, kids[0])).err == pgp.ERR_DECRYPTION_FAILED
    assert pgp.message_verify(ents, workload.one_pass_packet(0, 8, 1, kids[0]) + lit[:-4]).err == pgp.ERR_MESSAGE_BODY
    # an unknown signer is not an error — not even with garbage where the signature should be
    assert pgp.message_verify(ents, workload.one_pass_packet(0, 8, 1, kids[2]) + lit + b"\xff\xff").err is None
