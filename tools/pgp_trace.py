"""One traced packet-level batch call (BFTQ_TRACE): per-chunk host timeline on stderr."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bftkv_b200 import Engine, _lib, workload
from bftkv_b200.crypto_gpu import Keyring, _blob

N = 65536
w = workload.make_pgp_verify_batch(N)
eng = Engine(0)
kr = Keyring(eng)
kr.register(w["keyring"])
lib = _lib.load()
tb, to = _blob(w["tbs"])
sb, so = _blob(w["sigs"])
p = lambda a: C.c_void_p(a.ctypes.data)
err = np.zeros(N, np.int32)
call = lambda: _lib.check(lib.bftq_signature_verify_batch(kr._h, p(tb), p(to), p(sb), p(so), N, p(err)))
for _ in range(5):
    call()
for _ in range(3):
    t0 = time.perf_counter()
    call()
    print("untraced call %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
import threading
os.environ.pop("BFTQ_HOST_THREADS", None)
errs = [np.zeros(N, np.int32) for _ in range(2)]


def caller(c, n):
    for _ in range(n):
        _lib.check(lib.bftq_signature_verify_batch(kr._h, p(tb), p(to), p(sb), p(so), N, p(errs[c])))


def both(n):
    ths = [threading.Thread(target=caller, args=(c, n)) for c in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]


both(5)
t0 = time.perf_counter()
both(10)
print("2 callers untraced: %.3f ms per batch" % ((time.perf_counter() - t0) / 20 * 1e3), file=sys.stderr)
os.environ["BFTQ_TRACE"] = "1"
t0 = time.perf_counter()
both(2)
print("2 callers traced: %.3f ms per batch" % ((time.perf_counter() - t0) / 4 * 1e3), file=sys.stderr)
