"""ctypes loader for oracle/libbftq_oracle.so (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py)."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libbftq_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError("oracle/libbftq_oracle.so missing: run `make -C oracle`")
        _lib = C.CDLL(_PATH)
        _lib.orc_rsa_verify_batch.restype = C.c_int
        _lib.orc_rsa_verify_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_sha256.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        _lib.orc_sha256_3.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p]
        _lib.orc_tally_batch.argtypes = [C.c_void_p] * 3 + [C.c_uint32] + [C.c_void_p] * 3 + [C.c_uint64, C.c_void_p]
        _lib.orc_read_decide_batch.argtypes = [C.c_void_p] * 3 + [C.c_uint32] + [C.c_void_p] * 5 + [C.c_uint64] + [C.c_void_p] * 3
    return _lib


def rsa_verify_batch(moduli, exps, key_idx, sig, digest, hash_alg=8, strict_range=False, threads=1):
    n_be = np.frombuffer(b"".join(int(n).to_bytes(256, "big") for n in moduli), np.uint8).copy()
    exps = np.ascontiguousarray(np.asarray(exps, np.uint32))
    key_idx = np.ascontiguousarray(key_idx, np.uint32)
    sig = np.ascontiguousarray(sig, np.uint8)
    digest = np.ascontiguousarray(digest, np.uint8)
    n = key_idx.shape[0]
    out = np.empty(n, np.uint8)
    rc = lib().orc_rsa_verify_batch(n_be.ctypes.data, exps.ctypes.data, len(moduli), key_idx.ctypes.data, sig.ctypes.data,
                                    digest.ctypes.data, hash_alg, n, int(strict_range), threads, out.ctypes.data)
    if rc:
        raise ValueError("oracle: bad key or hash id")
    return out


def sha256(m: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().orc_sha256(m, len(m), out)
    return out.raw


def tally_batch(qcs, op_off, signer_id, status):
    """qcs: list of (f, min, threshold, suff, [member ids])."""
    params = np.array([v for q in qcs for v in q[:4]], np.int32)
    moff = np.zeros(len(qcs) + 1, np.uint32)
    for i, q in enumerate(qcs):
        moff[i + 1] = moff[i] + len(q[4])
    members = np.array([m for q in qcs for m in q[4]], np.uint64)
    op_off = np.ascontiguousarray(op_off, np.uint32)
    signer_id = np.ascontiguousarray(signer_id, np.uint64)
    status = np.ascontiguousarray(status, np.uint8)
    n_ops = op_off.shape[0] - 1
    out = np.empty(n_ops, np.uint8)
    lib().orc_tally_batch(params.ctypes.data, moff.ctypes.data, members.ctypes.data if len(members) else None, len(qcs),
                          op_off.ctypes.data, signer_id.ctypes.data, status.ctypes.data, n_ops, out.ctypes.data)
    return out


def _qc_arrays(qcs):
    params = np.array([v for q in qcs for v in q[:4]], np.int32)
    moff = np.zeros(len(qcs) + 1, np.uint32)
    for i, q in enumerate(qcs):
        moff[i + 1] = moff[i] + len(q[4])
    members = np.array([m for q in qcs for m in q[4]], np.uint64)
    return params, moff, members


def read_decide_batch(qcs, op_off, signer_id, status, ts, value_id):
    """Client.Read's decision per op, responders in arrival order (oracle/c orc_read_decide_batch).
    Returns (decision u8, winner u32, decided_at u32)."""
    params, moff, members = _qc_arrays(qcs)
    op_off = np.ascontiguousarray(op_off, np.uint32)
    signer_id = np.ascontiguousarray(signer_id, np.uint64)
    status = np.ascontiguousarray(status, np.uint8)
    ts = np.ascontiguousarray(ts, np.uint64)
    value_id = np.ascontiguousarray(value_id, np.uint32)
    n_ops = op_off.shape[0] - 1
    dec, win, at = np.empty(n_ops, np.uint8), np.empty(n_ops, np.uint32), np.empty(n_ops, np.uint32)
    lib().orc_read_decide_batch(params.ctypes.data, moff.ctypes.data, members.ctypes.data if len(members) else None, len(qcs),
                                op_off.ctypes.data, signer_id.ctypes.data, status.ctypes.data, ts.ctypes.data, value_id.ctypes.data,
                                n_ops, dec.ctypes.data, win.ctypes.data, at.ctypes.data)
    return dec, win, at


# ---- libcrypto stand-in for the Go CPU path (oracle/c/libcrypto_baseline.c) --------------------
_LC_PATH = os.path.join(_HERE, "libbftq_libcrypto.so")
_lc = None


def libcrypto_available():
    return os.path.exists(_LC_PATH)


def libcrypto_rsa_verify_batch(moduli, exps, key_idx, sig, digest, threads=1):
    """RSA-2048 / SHA-256 PKCS#1 v1.5 batch verify through OpenSSL EVP_PKEY_verify on `threads` pthreads."""
    global _lc
    if _lc is None:
        if not libcrypto_available():
            raise ImportError("oracle/libbftq_libcrypto.so missing: run `make -C oracle` (needs OpenSSL headers)")
        _lc = C.CDLL(_LC_PATH)
        _lc.lcb_rsa_verify_batch.restype = C.c_int
        _lc.lcb_rsa_verify_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    n_be = np.frombuffer(b"".join(int(n).to_bytes(256, "big") for n in moduli), np.uint8).copy()
    exps = np.ascontiguousarray(np.asarray(exps, np.uint32))
    key_idx = np.ascontiguousarray(key_idx, np.uint32)
    sig = np.ascontiguousarray(sig, np.uint8)
    digest = np.ascontiguousarray(digest, np.uint8)
    n = key_idx.shape[0]
    out = np.empty(n, np.uint8)
    rc = _lc.lcb_rsa_verify_batch(n_be.ctypes.data, exps.ctypes.data, len(moduli), key_idx.ctypes.data, sig.ctypes.data, digest.ctypes.data,
                                  n, threads, out.ctypes.data)
    if rc:
        raise ValueError("libcrypto baseline: bad key")
    return out
