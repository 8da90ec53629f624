"""Oracle restatement of bftkv's wire packet <x, v, t, sig, ss, auth>.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
Follows packet/packet.go of the reference:
  Serialize :35-60, Parse :62-115, WriteChunk/ReadChunk :117-139,
  seek2tbs :141-154, TBS :156-168, TBSS :170-190,
  writeSignature :192-213, readSignature :215-235.
"""
import struct
from dataclasses import dataclass
from typing import Optional

SIGNATURE_TYPE_NIL = 0   # packet.go:13
SIGNATURE_TYPE_PGP = 1   # packet.go:14


@dataclass
class SignaturePacket:    # packet.go:25-31
    type: int = 0
    version: int = 0
    completed: bool = False
    data: bytes = b""
    cert: bytes = b""


def write_chunk(chunk: Optional[bytes]) -> bytes:          # packet.go:117-124
    chunk = chunk or b""
    return struct.pack(">Q", len(chunk)) + chunk


def write_signature(sig: Optional[SignaturePacket]) -> bytes:   # packet.go:192-213
    if sig is None:
        sig = SignaturePacket()
    return (bytes([sig.type & 0xFF]) + struct.pack(">I", sig.version)
            + (b"\x01" if sig.completed else b"\x00")
            + write_chunk(sig.data) + write_chunk(sig.cert))


def serialize(*args) -> bytes:                              # packet.go:35-60
    out = b""
    for i, arg in enumerate(args):
        if i in (0, 1, 5):
            out += write_chunk(arg)
        elif i == 2:
            out += struct.pack(">Q", arg)
        elif i in (3, 4):
            out += write_signature(arg)
    return out


class _Reader:
    def __init__(self, b: bytes):
        self.b, self.pos = b, 0

    def read(self, n: int) -> bytes:
        if self.pos >= len(self.b) and n > 0:
            raise EOFError
        if self.pos + n > len(self.b):
            raise ValueError("unexpected EOF")
        r = self.b[self.pos:self.pos + n]
        self.pos += n
        return r


def _read_chunk(r: _Reader) -> Optional[bytes]:             # packet.go:126-139
    (l,) = struct.unpack(">Q", r.read(8))
    if l == 0:
        return None
    return r.read(l)


def _read_signature(r: _Reader) -> Optional[SignaturePacket]:   # packet.go:215-235
    sig = SignaturePacket()
    sig.type = r.read(1)[0]
    (sig.version,) = struct.unpack(">I", r.read(4))
    sig.completed = r.read(1)[0] != 0
    sig.data = _read_chunk(r) or b""
    sig.cert = _read_chunk(r) or b""
    if sig.type == SIGNATURE_TYPE_NIL:
        return None
    return sig


def parse(pkt: bytes):                                      # packet.go:62-115
    """Returns (variable, value, t, sig, ss, auth); trailing fields may be None/0
    when the packet ends early (io.EOF at a field boundary is not an error)."""
    r = _Reader(pkt)
    variable = _read_chunk(r)          # EOF here IS an error in the reference
    value, t, sig, ss, auth = None, 0, None, None, None
    try:
        value = _read_chunk(r)
        (t,) = struct.unpack(">Q", r.read(8))
        sig = _read_signature(r)
        ss = _read_signature(r)
        auth = _read_chunk(r)
    except EOFError:
        pass
    return variable, value, t, sig, ss, auth


def _seek2tbs(pkt: bytes) -> int:                           # packet.go:141-154
    pos = 0
    (l,) = struct.unpack(">q", pkt[pos:pos + 8]); pos += 8 + l
    (l,) = struct.unpack(">q", pkt[pos:pos + 8]); pos += 8 + l
    pos += 8
    return pos


def tbs(pkt: bytes) -> bytes:                               # packet.go:156-168
    off = _seek2tbs(pkt)
    if off > len(pkt):
        raise ValueError("unexpected EOF")
    return pkt[:off]


def tbss(pkt: bytes) -> bytes:                              # packet.go:170-190
    off = _seek2tbs(pkt)
    r = _Reader(pkt)
    r.pos = off
    _read_signature(r)
    return pkt[:r.pos]
