"""Oracle restatement of bftkv's OpenPGP signature checks.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

bftkv side (in the reference tree):
  crypto/pgp/crypto_pgp.go:195-197   getKeyring = secring ++ keyring
  crypto/pgp/crypto_pgp.go:319-330   PGPSignature.Verify
  crypto/pgp/crypto_pgp.go:332-344   PGPSignature.VerifyWithCertificate
  crypto/pgp/crypto_pgp.go:373-390   PGPSignature.Signers
  crypto/pgp/crypto_pgp.go:485-500   PGPCollectiveSignature.Verify
  crypto/pgp/crypto_pgp.go:506-515   PGPCollectiveSignature.Combine

Third-party side, NOT in the reference tree (golang.org/x/crypto/openpgp @
v0.0.0-20191227163750-53104e6ec876, Go 1.13 crypto/rsa): restated from RFC 4880 and the module's
published behaviour — packet.Read/Reader.Next framing, Signature.parse (v4 only; v2/v3 are parsed
as SignatureV3), subpacket handling (critical unknown => error), EntityList.KeysByIdUsage,
CheckDetachedSignature's "skip unknown issuers" loop, hashForSignature / VerifySignature's
hash-suffix + 16-bit tag check, padToKeySize, rsa.VerifyPKCS1v15.  Parity for this layer is
"unpinned" by the reference (it ships no vectors); it is pinned here against GnuPG-made
signatures (tests/golden/).
"""
import hashlib
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

# ---- errors (all collapse to ErrInvalidSignature at the bftkv layer) --------------------------


class PGPError(Exception):
    pass


class UnknownIssuer(PGPError):
    pass


class StructuralError(PGPError):
    pass


class UnsupportedError(PGPError):
    pass


class SignatureError(PGPError):
    pass


ERR_INVALID_SIGNATURE = "crypto: invalid signature"                       # crypto/crypto.go
ERR_INSUFFICIENT = "crypto: insufficient number of signatures"

HASH_BY_ID = {1: "md5", 2: "sha1", 3: "ripemd160", 8: "sha256", 9: "sha384", 10: "sha512", 11: "sha224"}
DIGEST_PREFIX = {    # Go crypto/rsa hashPrefixes == crypto/threshold/rsa/rsa.go:345-354
    1: bytes.fromhex("3020300c06082a864886f70d020505000410"),
    2: bytes.fromhex("3021300906052b0e03021a05000414"),
    3: bytes.fromhex("30203008060628cf060300310414"),
    8: bytes.fromhex("3031300d060960864801650304020105000420"),
    9: bytes.fromhex("3041300d060960864801650304020205000430"),
    10: bytes.fromhex("3051300d060960864801650304020305000440"),
    11: bytes.fromhex("302d300d06096086480165030402040500041c"),
}
KNOWN_TAGS = {1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 13, 14, 17, 18}     # packet.Read's switch
KEY_FLAG_CERTIFY, KEY_FLAG_SIGN = 0x01, 0x02


# ---- packet framing (RFC 4880 §4.2; x/crypto packet.readHeader) -------------------------------

class Reader:
    """bytes.Reader stand-in shared between successive CheckDetachedSignature calls."""

    def __init__(self, data: bytes):
        self.data, self.pos = data, 0

    def remaining(self) -> int:
        return len(self.data) - self.pos


def read_packet(r: Reader) -> Optional[Tuple[int, bytes]]:
    """Returns (tag, body) or None at a clean EOF.  Raises StructuralError on bad framing: a truncated
    length or body exhausts the reader (x/crypto's readFull hits io.ErrUnexpectedEOF with everything
    consumed); a tag byte without the MSB consumes exactly that one byte (packet.readHeader reads ONE byte
    before returning StructuralError), so a caller that ignores errors resynchronises on the next byte."""
    d = r.data
    if r.pos >= len(d):
        return None
    hdr = d[r.pos]
    if hdr & 0x80 == 0:
        r.pos += 1
        raise StructuralError("tag byte does not have MSB set")
    try:
        if hdr & 0x40 == 0:                       # old format
            tag = (hdr & 0x3F) >> 2
            lt = hdr & 3
            p = r.pos + 1
            if lt == 3:                           # indeterminate length: runs to EOF
                body = d[p:]
                r.pos = len(d)
                return tag, body
            nlen = 1 << lt
            if p + nlen > len(d):
                raise IndexError
            ln = int.from_bytes(d[p:p + nlen], "big")
            p += nlen
        else:                                     # new format
            tag = hdr & 0x3F
            p = r.pos + 1
            body = b""
            while True:
                o = d[p]
                if o < 192:
                    ln = o; p += 1; partial = False
                elif o < 224:
                    ln = ((o - 192) << 8) + d[p + 1] + 192; p += 2; partial = False
                elif o == 255:
                    if p + 5 > len(d):
                        raise IndexError
                    ln = int.from_bytes(d[p + 1:p + 5], "big"); p += 5; partial = False
                else:
                    ln = 1 << (o & 0x1F); p += 1; partial = True
                if p + ln > len(d):
                    raise IndexError
                body += d[p:p + ln]
                p += ln
                if not partial:
                    r.pos = p
                    return tag, body
        if p + ln > len(d):
            raise IndexError
        r.pos = p + ln
        return tag, d[p:p + ln]
    except IndexError:
        r.pos = len(d)
        raise StructuralError("unexpected EOF in packet")


def read_mpi(b: bytes, p: int) -> Tuple[int, int, int]:
    """Returns (value, bit_length_field, new_pos)."""
    if p + 2 > len(b):
        raise StructuralError("mpi truncated")
    bits = int.from_bytes(b[p:p + 2], "big")
    n = (bits + 7) // 8
    if p + 2 + n > len(b):
        raise StructuralError("mpi truncated")
    return int.from_bytes(b[p + 2:p + 2 + n], "big"), bits, p + 2 + n


# ---- public keys -------------------------------------------------------------------------------

@dataclass
class PublicKey:
    algo: int
    n: int = 0
    e: int = 0
    key_id: int = 0
    fingerprint: bytes = b""
    is_subkey: bool = False
    body: bytes = b""
    ec_oid: bytes = b""
    ec_point: Optional[Tuple[int, int]] = None      # P-256 only
    dsa: Optional[Tuple[int, int, int, int]] = None  # p, q, g, y


def parse_public_key(body: bytes, is_subkey: bool = False) -> PublicKey:
    if len(body) < 6 or body[0] != 4:
        raise UnsupportedError("public key version")
    algo = body[5]
    pk = PublicKey(algo=algo, is_subkey=is_subkey, body=body)
    if algo in (1, 2, 3):
        pk.n, nbits, p = read_mpi(body, 6)
        pk.e, ebits, p = read_mpi(body, p)
        if (ebits + 7) // 8 > 3:
            raise UnsupportedError("large public exponent")
    elif algo == 19:
        # x/crypto packet/public_key.go parseECDSA + newECDSA: one length octet, the curve OID, then the
        # SEC1 uncompressed point as one MPI; elliptic.Unmarshal needs 04 || X || Y on the curve.
        ol = body[6] if len(body) > 6 else 0
        if ol in (0, 0xff) or 7 + ol > len(body):
            raise StructuralError("invalid oid length")
        pk.ec_oid = body[7:7 + ol]
        pt, bits, p = read_mpi(body, 7 + ol)
        raw = body[7 + ol + 2:p]
        if pk.ec_oid == bytes([0x2A, 0x86, 0x48, 0xCE, 0x3D, 0x03, 0x01, 0x07]) and len(raw) == 65 and raw[0] == 4:
            pk.ec_point = (int.from_bytes(raw[1:33], "big"), int.from_bytes(raw[33:], "big"))
    elif algo == 17:
        # parseDSA: p, q, g, y
        vals, p = [], 6
        for _ in range(4):
            v, bits, p = read_mpi(body, p)
            vals.append(v)
        pk.dsa = tuple(vals)
    elif algo in (16, 18):
        pass      # ElGamal / ECDH: parsed by x/crypto, never signature keys
    else:
        raise UnsupportedError("public key type: %d" % algo)
    fp = hashlib.sha1(b"\x99" + len(body).to_bytes(2, "big") + body).digest()
    pk.fingerprint = fp
    pk.key_id = int.from_bytes(fp[12:20], "big")
    return pk


# ---- signature packets -------------------------------------------------------------------------

@dataclass
class Signature:
    version: int
    sig_type: int
    pk_algo: int
    hash_id: int
    hash_suffix: bytes = b""          # bytes appended to the hashed data (incl. trailer)
    hash_tag: bytes = b""
    issuer_key_id: Optional[int] = None
    creation_time: Optional[int] = None
    flags_valid: bool = False
    flag_certify: bool = False
    flag_sign: bool = False
    is_primary_id: Optional[bool] = None
    revocation_reason: Optional[int] = None
    sig_r: int = 0                    # DSA / ECDSA
    sig_s: int = 0
    rsa_sig: int = 0
    rsa_sig_bytes: bytes = b""        # MPI bytes as stored (leading zeros stripped)
    raw: bytes = b""


def _parse_subpackets(area: bytes, sig: Signature, hashed: bool):
    p = 0
    while p < len(area):
        o = area[p]
        if o < 192:
            ln = o; p += 1
        elif o < 255:
            if p + 2 > len(area):
                raise StructuralError("subpacket truncated")
            ln = ((o - 192) << 8) + area[p + 1] + 192; p += 2
        else:
            if p + 5 > len(area):
                raise StructuralError("subpacket truncated")
            ln = int.from_bytes(area[p + 1:p + 5], "big"); p += 5
        if ln == 0 or p + ln > len(area):
            raise StructuralError("subpacket truncated")
        typ = area[p] & 0x7F
        critical = bool(area[p] & 0x80)
        sub = area[p + 1:p + ln]
        p += ln
        if typ == 2:                               # creation time (hashed only)
            if not hashed:
                continue
            if len(sub) != 4:
                raise StructuralError("signature creation time not four bytes")
            sig.creation_time = int.from_bytes(sub, "big")
        elif typ == 3 or typ == 9:                 # sig / key lifetime
            if hashed and len(sub) != 4:
                raise StructuralError("expiry subpacket with bad length")
        elif typ == 16:                            # issuer — accepted from either area
            if len(sub) != 8:
                raise StructuralError("issuer subpacket with bad length")
            sig.issuer_key_id = int.from_bytes(sub, "big")
        elif typ == 27:                            # key flags (hashed only)
            if not hashed:
                continue
            if len(sub) == 0:
                raise StructuralError("empty key flags subpacket")
            sig.flags_valid = True
            sig.flag_certify = bool(sub[0] & KEY_FLAG_CERTIFY)
            sig.flag_sign = bool(sub[0] & KEY_FLAG_SIGN)
        elif typ == 25:                            # primary user id
            if hashed:
                if len(sub) != 1:
                    raise StructuralError("primary user id subpacket with bad length")
                sig.is_primary_id = sub[0] > 0
        elif typ == 29:                            # reason for revocation
            if hashed:
                if len(sub) == 0:
                    raise StructuralError("empty revocation reason subpacket")
                sig.revocation_reason = sub[0]
        elif typ in (11, 21, 22, 30, 32):           # prefs / features / embedded sig: accepted
            pass
        else:
            if critical:
                raise UnsupportedError("unknown critical signature subpacket type %d" % typ)


def parse_signature(body: bytes) -> Signature:
    """packet.Signature.parse (v4).  v2/v3 bodies become version-3 Signature objects."""
    if len(body) < 1:
        raise StructuralError("empty signature packet")
    ver = body[0]
    if ver < 4:
        return _parse_signature_v3(body)
    if ver != 4:
        raise UnsupportedError("signature packet version %d" % ver)
    if len(body) < 6:
        raise StructuralError("signature truncated")
    sig = Signature(version=4, sig_type=body[1], pk_algo=body[2], hash_id=body[3], raw=body)
    if sig.pk_algo not in (1, 3, 17, 19):
        raise UnsupportedError("public key algorithm %d" % sig.pk_algo)
    if sig.hash_id not in HASH_BY_ID:
        raise UnsupportedError("hash function %d" % sig.hash_id)
    hl = int.from_bytes(body[4:6], "big")
    if 6 + hl + 2 > len(body):
        raise StructuralError("signature truncated")
    hashed = body[6:6 + hl]
    ln = 6 + hl
    sig.hash_suffix = body[:ln] + b"\x04\xff" + ln.to_bytes(4, "big")
    _parse_subpackets(hashed, sig, True)
    p = 6 + hl
    ul = int.from_bytes(body[p:p + 2], "big")
    p += 2
    if p + ul + 2 > len(body):
        raise StructuralError("signature truncated")
    _parse_subpackets(body[p:p + ul], sig, False)
    p += ul
    if sig.creation_time is None:
        raise StructuralError("no creation time in signature")
    sig.hash_tag = body[p:p + 2]
    p += 2
    if sig.pk_algo in (1, 3):
        sig.rsa_sig, bits, q = read_mpi(body, p)
        sig.rsa_sig_bytes = body[p + 2:q]
    else:
        sig.sig_r, bits, q = read_mpi(body, p)      # DSASigR / ECDSASigR
        sig.sig_s, bits, q = read_mpi(body, q)
    return sig


def _parse_signature_v3(body: bytes) -> Signature:
    if len(body) < 19 or body[0] not in (2, 3) or body[1] != 5:
        raise UnsupportedError("signature packet version")
    sig = Signature(version=3, sig_type=body[2], pk_algo=body[15], hash_id=body[16], raw=body)
    sig.creation_time = int.from_bytes(body[3:7], "big")
    sig.issuer_key_id = int.from_bytes(body[7:15], "big")
    if sig.pk_algo not in (1, 3, 17):
        raise UnsupportedError("public key algorithm")
    if sig.hash_id not in HASH_BY_ID:
        raise UnsupportedError("hash function")
    sig.hash_suffix = body[2:7]
    sig.hash_tag = body[17:19]
    if sig.pk_algo in (1, 3):
        sig.rsa_sig, bits, q = read_mpi(body, 19)
        sig.rsa_sig_bytes = body[21:q]
    else:                                           # SignatureV3.parse: DSASigR, DSASigS
        sig.sig_r, bits, q = read_mpi(body, 19)
        sig.sig_s, bits, q = read_mpi(body, q)
    return sig


# ---- entities / keyring ------------------------------------------------------------------------

@dataclass
class Subkey:
    public_key: PublicKey
    sig: Optional[Signature]


@dataclass
class Entity:
    primary_key: PublicKey
    self_signature: Optional[Signature] = None       # of the (first / primary) identity
    revocations: List[Signature] = field(default_factory=list)
    subkeys: List[Subkey] = field(default_factory=list)
    identities: List[str] = field(default_factory=list)
    certifier_ids: List[int] = field(default_factory=list)   # issuers of third-party certifications


def read_entities(data: bytes) -> List[Entity]:
    """Simplified openpgp.ReadKeyRing: groups packets into entities, keeps the first identity's
    self-signature (or the one flagged primary), third-party certification issuers
    (crypto_pgp.go:80-88 reads them via id.Signatures) and subkey binding signatures.
    Self-signature CRYPTO checks done by x/crypto at load time are outside the per-signature hot
    path and are not restated (key material here comes from trusted fixtures)."""
    r = Reader(data)
    ents: List[Entity] = []
    cur: Optional[Entity] = None
    last = None           # ("uid", idx) | ("sub", Subkey) | ("key",)
    while True:
        pk = read_packet(r)
        if pk is None:
            break
        tag, body = pk
        if tag == 6:
            try:
                cur = Entity(parse_public_key(body))
                ents.append(cur)
                last = ("key",)
            except UnsupportedError:
                cur = None
        elif cur is None:
            continue
        elif tag == 13:
            cur.identities.append(body.decode("utf-8", "replace"))
            last = ("uid", len(cur.identities) - 1)
        elif tag == 14:
            try:
                sk = Subkey(parse_public_key(body, True), None)
                cur.subkeys.append(sk)
                last = ("sub", sk)
            except UnsupportedError:
                last = None
        elif tag == 2:
            try:
                sig = parse_signature(body)
            except PGPError:
                continue
            if last is None:
                continue
            if last[0] == "key":
                if sig.sig_type == 0x20:
                    cur.revocations.append(sig)
            elif last[0] == "uid":
                if sig.sig_type in (0x10, 0x11, 0x12, 0x13):
                    if sig.issuer_key_id == cur.primary_key.key_id:
                        if cur.self_signature is None or (sig.is_primary_id and last[1] > 0):
                            cur.self_signature = sig
                    elif sig.issuer_key_id is not None:
                        cur.certifier_ids.append(sig.issuer_key_id)
            elif last[0] == "sub":
                if sig.sig_type == 0x18:
                    last[1].sig = sig
    return ents


@dataclass
class Key:
    entity: Entity
    public_key: PublicKey
    self_signature: Optional[Signature]


def keys_by_id_usage(keyring: List[Entity], key_id: int, usage: int) -> List[Key]:
    """EntityList.KeysByIdUsage (x/crypto keys.go): primary key and subkeys whose id matches;
    skip revoked entities / keys; if the self (binding) signature carries key flags, require
    `usage`."""
    out = []
    for e in keyring:
        cands = []
        if e.primary_key.key_id == key_id:
            cands.append(Key(e, e.primary_key, e.self_signature))
        for sk in e.subkeys:
            if sk.public_key.key_id == key_id:
                cands.append(Key(e, sk.public_key, sk.sig))
        for k in cands:
            if e.revocations:
                continue
            ss = k.self_signature
            if ss is not None and ss.revocation_reason is not None:
                continue
            if ss is not None and ss.flags_valid and usage:
                have = (KEY_FLAG_CERTIFY if ss.flag_certify else 0) | (KEY_FLAG_SIGN if ss.flag_sign else 0)
                if have & usage != usage:
                    continue
            out.append(k)
    return out


# ---- the verify primitive ----------------------------------------------------------------------

def signature_digest(signed: bytes, sig: Signature) -> bytes:
    """hashForSignature + VerifySignature's suffix step.  sig type 0x00: raw bytes; 0x01: text
    with line endings canonicalised to CRLF."""
    h = hashlib.new(HASH_BY_ID[sig.hash_id])
    if sig.sig_type == 0x01:
        out, i = bytearray(), 0
        while i < len(signed):
            c = signed[i]
            if c == 0x0D and i + 1 < len(signed) and signed[i + 1] == 0x0A:
                out += b"\r\n"; i += 2
            elif c == 0x0A:
                out += b"\r\n"; i += 1
            else:
                out.append(c); i += 1
        signed = bytes(out)
    elif sig.sig_type != 0x00:
        raise UnsupportedError("signature type %d" % sig.sig_type)
    h.update(signed)
    h.update(sig.hash_suffix)
    return h.digest()


def rsa_verify_pkcs1v15(n: int, e: int, hash_id: int, digest: bytes, sig_bytes: bytes,
                        strict_range: bool = False) -> bool:
    """Go 1.13 rsa.VerifyPKCS1v15 after x/crypto's padToKeySize."""
    k = (n.bit_length() + 7) // 8
    prefix = DIGEST_PREFIX[hash_id]
    tlen = len(prefix) + len(digest)
    if k < tlen + 11:
        return False
    if len(sig_bytes) < k:                         # padToKeySize
        sig_bytes = b"\x00" * (k - len(sig_bytes)) + sig_bytes
    if len(sig_bytes) != k:
        return False
    s = int.from_bytes(sig_bytes, "big")
    if strict_range and s >= n:
        return False
    m = pow(s, e, n)
    em = m.to_bytes(k, "big")
    return em == b"\x00\x01" + b"\xff" * (k - tlen - 3) + b"\x00" + prefix + digest


def ecdsa_p256_verify(q: Tuple[int, int], digest: bytes, r: int, s: int) -> bool:
    """Go crypto/ecdsa.Verify on P-256 (ecdsa.go Verify + hashToInt): no low-s rule; the digest is cut
    to its leftmost 32 bytes."""
    from . import sss_oracle as so
    n = so.P256_N
    if r <= 0 or s <= 0 or r >= n or s >= n:
        return False
    x, y = q
    if (y * y - (x * x * x - 3 * x + so.P256_B)) % so.P256_P != 0 or x >= so.P256_P or y >= so.P256_P:
        return False                                # elliptic.Unmarshal would have refused the key
    e = int.from_bytes(digest[:32], "big")
    w = pow(s, -1, n)
    pt = so.p256_add(so.p256_mul(e * w % n, so.P256_G), so.p256_mul(r * w % n, q))
    if pt is None:
        return False
    return pt[0] % n == r


def dsa_verify(p: int, q: int, g: int, y: int, digest: bytes, r: int, s: int) -> bool:
    """x/crypto VerifySignature's DSA arm (digest cut to the subgroup size) + Go crypto/dsa.Verify."""
    sub = (q.bit_length() + 7) // 8
    digest = digest[:sub]
    if p == 0:
        return False
    if r < 1 or r >= q or s < 1 or s >= q:
        return False
    try:
        w = pow(s, -1, q)
    except ValueError:
        return False
    if q.bit_length() & 7:
        return False
    z = int.from_bytes(digest, "big")
    u1, u2 = z * w % q, r * w % q
    v = pow(g, u1, p) * pow(y, u2, p) % p % q
    return v == r


def verify_signature(pk: PublicKey, signed: bytes, sig: Signature):
    """packet.PublicKey.VerifySignature.  Raises SignatureError / UnsupportedError."""
    digest = signature_digest(signed, sig)
    if digest[:2] != sig.hash_tag:
        raise SignatureError("hash tag doesn't match")
    if pk.algo != sig.pk_algo:
        raise SignatureError("public key and signature use different algorithms")
    if pk.algo in (1, 3):
        if not rsa_verify_pkcs1v15(pk.n, pk.e, sig.hash_id, digest, sig.rsa_sig_bytes):
            raise SignatureError("RSA verification failure")
        return
    if pk.algo == 19 and pk.ec_point is not None:
        if not ecdsa_p256_verify(pk.ec_point, digest, sig.sig_r, sig.sig_s):
            raise SignatureError("ECDSA verification failure")
        return
    if pk.algo == 17 and pk.dsa is not None:
        if not dsa_verify(*pk.dsa, digest, sig.sig_r, sig.sig_s):
            raise SignatureError("DSA verification failure")
        return
    raise UnsupportedError("oracle: public key algorithm %d not restated" % pk.algo)


def check_detached_signature(keyring: List[Entity], signed: bytes, r: Reader) -> Entity:
    """openpgp.CheckDetachedSignature on a shared reader."""
    sig, keys = None, []
    while True:
        while True:                                # Reader.Next: skip unknown packet types
            pk = read_packet(r)
            if pk is None:
                raise UnknownIssuer()
            if pk[0] in KNOWN_TAGS:
                break
        tag, body = pk
        if tag != 2:
            raise StructuralError("non signature packet found")
        sig = parse_signature(body)
        if sig.issuer_key_id is None:
            raise StructuralError("signature doesn't have an issuer")
        keys = keys_by_id_usage(keyring, sig.issuer_key_id, KEY_FLAG_SIGN)
        if keys:
            break
    err = None
    for key in keys:
        try:
            verify_signature(key.public_key, signed, sig)
            return key.entity
        except PGPError as ex:
            err = ex
    raise err


# ---- bftkv layer -------------------------------------------------------------------------------

def signature_verify(keyring: List[Entity], tbs: bytes, sig_data: bytes) -> Optional[str]:
    """PGPSignature.Verify, crypto_pgp.go:319-330.  None = nil error."""
    r = Reader(sig_data or b"")
    err: Optional[str] = ERR_INVALID_SIGNATURE
    while r.remaining() > 0:
        try:
            check_detached_signature(keyring, tbs, r)
            err = None
        except PGPError:
            return ERR_INVALID_SIGNATURE
    return err


def signature_verify_with_certificate(cert_data: bytes, tbs: bytes, sig_data: bytes) -> Optional[str]:
    """PGPSignature.VerifyWithCertificate, crypto_pgp.go:332-344: keyring = the one entity of
    `cert` (Issuer() takes the first entity of sig.Cert, :396-405)."""
    ents = read_entities(cert_data)
    return signature_verify(ents[:1], tbs, sig_data)


def signers(keyring: List[Entity], sig_data: bytes) -> List[int]:
    """PGPSignature.Signers, crypto_pgp.go:373-390: issuer ids of every parseable signature
    packet whose issuer is a PRIMARY key id in the keyring (getCertById).  Returns key ids."""
    out = []
    r = Reader(sig_data or b"")
    while True:
        try:
            pk = read_packet(r)
        except PGPError:
            break
        if pk is None:
            break
        tag, body = pk
        if tag not in KNOWN_TAGS:
            continue
        if tag != 2:
            continue            # other packet kinds parse fine in x/crypto and are ignored by the switch
        try:
            sig = parse_signature(body)
        except PGPError:
            break               # r.Next() error ends the loop
        if sig.version != 4:
            continue            # *SignatureV3 is not matched by the type switch
        if sig.issuer_key_id is None:
            raise RuntimeError("nil pointer dereference (reference panics here)")
        if any(e.primary_key.key_id == sig.issuer_key_id for e in keyring):
            out.append(sig.issuer_key_id)
    return out


def collective_verify(keyring: List[Entity], tbs: bytes, ss_data: bytes, quorum) -> Tuple[Optional[str], bool]:
    """PGPCollectiveSignature.Verify, crypto_pgp.go:485-500.  `quorum` has
    is_sufficient(list_of_nodes_with_.id).  Returns (error, completed)."""
    from .wotqs_oracle import Node
    r = Reader(ss_data or b"")
    verified = []
    while r.remaining() > 0:
        try:
            ent = check_detached_signature(keyring, tbs, r)
        except PGPError:
            continue
        verified.append(Node(ent.primary_key.key_id))
        if quorum.is_sufficient(verified):
            return None, True
    return ERR_INSUFFICIENT, False


def collective_combine(keyring: List[Entity], ss_type: int, ss_data: bytes, s_type: int, s_data: bytes, quorum):
    """PGPCollectiveSignature.Combine, crypto_pgp.go:506-515.  Returns (ok, new_type, new_data)."""
    from .wotqs_oracle import Node
    if ss_type == 0:
        ss_type = s_type
    elif ss_type != s_type:
        return False, ss_type, ss_data
    ss_data = (ss_data or b"") + (s_data or b"")
    ids = signers(keyring, ss_data)
    return quorum.is_sufficient([Node(i) for i in ids]), ss_type, ss_data


# ---- transport messages: PGPMessage.Decrypt's signature half --------------------------------------------------
# crypto/pgp/crypto_pgp.go:453-471 -> openpgp.ReadMessage -> readSignedMessage + signatureCheckReader
# (golang.org/x/crypto/openpgp/read.go @53104e6ec876, restated from the published module; its source is not in the
# reference tree).  The input here is what the SymmetricallyEncrypted packet decrypts to — the host keeps the RSA
# private-key operation and the AES-CFB/MDC layer — i.e. the packet stream
#     [compressed]  one-pass signature (tag 4)   literal data (tag 11)   signature (tag 2)

ERR_DECRYPTION_FAILED = "crypto: decryption failed"                              # ReadMessage returned an error
ERR_TRANSPORT_SECURITY = "crypto: invalid transport security data"                # !(IsEncrypted && IsSigned)
ERR_MESSAGE_BODY = "message body / nonce error"                                   # ioutil.ReadAll or base64 error, returned as is
ERR_MESSAGE_UNSUPPORTED = "unsupported message form (compressed data)"            # NOT restated: the reference would inflate and go on


@dataclass
class MessageResult:
    err: Optional[str] = None
    plain: Optional[bytes] = None
    nonce: Optional[bytes] = None
    signed_by_key_id: int = 0
    signer_known: bool = False          # md.SignedBy != nil


_B64 = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"


def go_base64_std_decode(src: bytes) -> Optional[bytes]:
    """encoding/base64 StdEncoding.DecodeString (Go 1.13): CR and LF are skipped anywhere, every other byte outside the
    alphabet is an error, padding is mandatory ('xx==' / 'xxx=' close the input: only CR / LF may follow), trailing bits
    are not checked.  None = CorruptInputError."""
    out = bytearray()
    chars = [c for c in src if c not in (0x0D, 0x0A)]
    i, n = 0, len(chars)
    while i < n:
        quad = chars[i:i + 4]
        i += 4
        vals = []
        for j, c in enumerate(quad):
            if c == 0x3D:                                  # '='
                if j < 2:
                    return None
                if j == 2 and (len(quad) < 4 or quad[3] != 0x3D):
                    return None
                if i < n:                                  # data after the padding
                    return None
                break
            k = _B64.find(bytes([c]))
            if k < 0:
                return None
            vals.append(k)
        else:
            if len(quad) < 4:
                return None                                # incomplete quantum without padding
        if len(vals) == 4:
            out += bytes([(vals[0] << 2) | (vals[1] >> 4), ((vals[1] & 15) << 4) | (vals[2] >> 2), ((vals[2] & 3) << 6) | vals[3]])
        elif len(vals) == 3:
            out += bytes([(vals[0] << 2) | (vals[1] >> 4), ((vals[1] & 15) << 4) | (vals[2] >> 2)])
        elif len(vals) == 2:
            out += bytes([(vals[0] << 2) | (vals[1] >> 4)])
        else:
            return None
    return bytes(out)


def message_verify(keyring: List[Entity], stream: bytes) -> MessageResult:
    """PGPMessage.Decrypt on the decrypted packet stream.  Mirrors, in order:
      readSignedMessage's FindLiteralData loop  (compressed -> unsupported here; one-pass signature: must be IsLast,
        hashForSignature(p.Hash, p.SigType) must succeed, IsSigned = true, SignedBy = first key of
        KeysByIdUsage(KeyId, KeyFlagSign); the LAST one-pass packet before the literal wins)
      Decrypt's `!(m.IsEncrypted && m.IsSigned)` check
      ioutil.ReadAll(m.UnverifiedBody): with a known signer the packet after the literal data is read at EOF and
        verified with SignedBy.PublicKey.VerifySignature / VerifySignatureV3 over a hash made with the ONE-PASS packet's
        algorithm and signature type; with an unknown signer NOTHING after the literal is looked at and SignatureError
        stays nil (the reference then returns peer = nil, err = nil)
      base64.StdEncoding.DecodeString(FileName)
    """
    res = MessageResult()
    r = Reader(stream or b"")
    ops = None
    lit = None
    try:
        while True:
            at_literal = r.pos < len(r.data) and r.data[r.pos] & 0x80 and ((r.data[r.pos] & 0x3F) if r.data[r.pos] & 0x40 else ((r.data[r.pos] & 0x3F) >> 2)) == 11
            try:
                pk = read_packet(r)
            except PGPError:
                if at_literal:          # the literal body ends early: ioutil.ReadAll's error, after the IsSigned check
                    res.err = ERR_MESSAGE_BODY if ops is not None else ERR_TRANSPORT_SECURITY
                    return res
                raise
            if pk is None:
                raise StructuralError("EOF before literal data")          # packets.Next() -> io.EOF -> ReadMessage fails
            tag, body = pk
            if tag not in KNOWN_TAGS:
                continue
            if tag == 8:
                res.err = ERR_MESSAGE_UNSUPPORTED
                return res
            if tag == 4:
                if len(body) < 13:
                    raise StructuralError("short one-pass signature")
                if body[0] != 3:
                    raise UnsupportedError("one-pass-signature packet version")
                if body[2] not in HASH_BY_ID:
                    raise UnsupportedError("hash function")
                if not body[12]:
                    raise UnsupportedError("nested signatures")
                if body[2] == 3:
                    raise UnsupportedError("hash not available")           # crypto.RIPEMD160 is not linked into bftkv
                if body[1] not in (0x00, 0x01):
                    raise UnsupportedError("unsupported signature type")
                ops = {"sig_type": body[1], "hash_id": body[2], "pk_algo": body[3], "key_id": int.from_bytes(body[4:12], "big")}
            elif tag == 2:
                parse_signature(body)                                       # parsed by packet.Read, ignored by the switch
            elif tag == 11:
                if len(body) < 2 or len(body) < 2 + body[1] + 4:
                    raise StructuralError("short literal data")
                nl = body[1]
                lit = {"binary": body[0] == ord("b"), "name": body[2:2 + nl], "body": body[6 + nl:]}
                break
    except PGPError:
        res.err = ERR_DECRYPTION_FAILED
        return res
    if ops is None:
        res.err = ERR_TRANSPORT_SECURITY
        return res
    res.signed_by_key_id = ops["key_id"]
    keys = keys_by_id_usage(keyring, ops["key_id"], KEY_FLAG_SIGN)
    res.signer_known = bool(keys)
    sig_err = None
    if keys:
        try:
            while True:                                                     # packets.Next(): unknown types are skipped
                pk = read_packet(r)
                if pk is None:
                    raise StructuralError("EOF")
                if pk[0] in KNOWN_TAGS:
                    break
            if pk[0] != 2:
                raise StructuralError("LiteralData not followed by signature")
            sig = parse_signature(pk[1])
            # the hash object was made from the one-pass packet: its algorithm digests, its type canonicalises
            import copy
            fake = copy.copy(sig)
            fake.sig_type = ops["sig_type"]
            if ops["hash_id"] != sig.hash_id:
                if hashlib.new(HASH_BY_ID[ops["hash_id"]]).digest_size != hashlib.new(HASH_BY_ID[sig.hash_id]).digest_size:
                    raise SignatureError("digest made with the one-pass hash does not fit the signature's hash")
                raise UnsupportedError("oracle: same-length digest under another algorithm's DigestInfo not restated")
            verify_signature(keys[0].public_key, lit["body"], fake)
        except PGPError as ex:
            sig_err = ex
    nonce = go_base64_std_decode(lit["name"])
    if nonce is None:
        res.err = ERR_MESSAGE_BODY
        return res
    res.plain, res.nonce = lit["body"], nonce
    res.err = ERR_INVALID_SIGNATURE if sig_err is not None else None
    return res


# ---- Client.Read from raw answers (protocol/client.go:250-268 over transport.Multicast's per-response work) -------------
ST_OK, ST_INVALID, ST_OTHER, ST_NONCE, ST_UNVERIFIED = 0, 1, 3, 7, 8


def read_response_status(keyring: List[Entity], msg: bytes, nonce: bytes, pre: int = 0):
    """One answer as transport.Multicast + Client.processResponse see it -> (status class, t, value bytes).
    status: ST_OK / ST_UNVERIFIED (good answers: err == nil all the way), ST_INVALID (m.SignatureError), ST_NONCE
    (ErrTransportNonceMismatch, transport.go:121-124), ST_OTHER (every other error: ReadMessage, missing one-pass packet,
    body / base64 errors, packet.Parse, or `pre`: the transport failed earlier)."""
    from . import packet_oracle
    if pre:
        return ST_OTHER, 0, b""
    r = message_verify(keyring, msg)
    if r.err == ERR_INVALID_SIGNATURE:
        return ST_INVALID, 0, b""
    if r.err is not None:
        return ST_OTHER, 0, b""
    if r.nonce != nonce:
        return ST_NONCE, 0, b""
    t, value = 0, b""
    if r.plain:                                                    # processResponse: `if res.Data != nil && len(res.Data) > 0`
        try:
            _, value, t, _, _, _ = packet_oracle.parse(r.plain)
        except (EOFError, ValueError):
            return ST_OTHER, 0, b""
        value = value or b""
    return (ST_OK if r.signer_known else ST_UNVERIFIED), t, value
