"""Thin host-side handle on a libbftq engine (one per GPU).  numpy arrays / raw pointers in,
numpy arrays out; all arithmetic happens in the CUDA kernels behind the C ABI."""
import ctypes as C
import numpy as np
from . import _lib

HASH_SHA256 = 8
F_STRICT_RANGE = 1
ST_OK, ST_BAD_SIGNATURE, ST_HASH_TAG, ST_MALFORMED, ST_UNKNOWN_SIGNER, ST_UNSUPPORTED, ST_MISSING = range(7)
DIGEST_LEN = {1: 16, 2: 20, 3: 20, 8: 32, 9: 48, 10: 64, 11: 28}


def _ptr(a):
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):           # torch tensor (host pinned or device)
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Engine:
    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.bftq_init(device, C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if self._h:
            self._lib.bftq_shutdown(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return self._lib.bftq_device_sm_count(self._h)

    @property
    def key_count(self):
        return self._lib.bftq_key_count(self._h)

    def register_rsa_keys(self, moduli, exps) -> int:
        """moduli: iterable of Python ints (or (K,256) uint8 big-endian array); returns first index."""
        if not isinstance(moduli, np.ndarray):
            moduli = np.frombuffer(b"".join(int(n).to_bytes(256, "big") for n in moduli), dtype=np.uint8)
        moduli = np.ascontiguousarray(moduli, dtype=np.uint8).reshape(-1, 256)
        exps = np.ascontiguousarray(np.asarray(exps, dtype=np.uint32))
        assert exps.shape[0] == moduli.shape[0]
        first = C.c_uint32()
        _lib.check(self._lib.bftq_register_rsa_keys(self._h, _ptr(moduli), _ptr(exps), moduli.shape[0], C.byref(first)))
        return first.value

    def rsa_verify_batch(self, key_idx, sig_be, digest, hash_alg=HASH_SHA256, flags=0, out=None):
        """Host buffers (numpy or pinned torch tensors).  Returns uint8 status per item."""
        n = int(key_idx.shape[0])
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        _lib.check(self._lib.bftq_rsa_verify_batch(self._h, _ptr(key_idx), _ptr(sig_be), _ptr(digest),
                                                   hash_alg, n, flags, _ptr(out)))
        return out

    def rsa_verify_batch_dev(self, d_key_idx, d_sig, d_digest, n, d_status, hash_alg=HASH_SHA256, flags=0, stream=0):
        _lib.check(self._lib.bftq_rsa_verify_batch_dev(self._h, _ptr(d_key_idx), _ptr(d_sig), _ptr(d_digest),
                                                       hash_alg, n, flags, _ptr(d_status), C.c_void_p(stream)))

    def stats(self):
        s = _lib.Stats()
        _lib.check(self._lib.bftq_stats(self._h, C.byref(s)))
        return {"items": s.items, "launches": s.launches, "h2d_bytes": s.h2d_bytes, "d2h_bytes": s.d2h_bytes}

    def measure_int_peak(self) -> float:
        v = C.c_double()
        _lib.check(self._lib.bftq_measure_int_peak(self._h, C.byref(v)))
        return v.value
