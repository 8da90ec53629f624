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
os.environ["BFTQ_TRACE"] = "1"
t0 = time.perf_counter()
call()
print("traced call %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
del os.environ["BFTQ_TRACE"]
# the flat path for comparison, pageable and pinned
import torch
from bftkv_b200 import workload as wl
f = wl.make_verify_batch(N, 16)
eng.register_rsa_keys([k["n"] for k in f["keys"]], [k["e"] for k in f["keys"]])
idx = (f["key_idx"] + 16).astype(np.uint32)
idx[f["expect"] == 4] = 99999
for pin in (False, True):
    a = [torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(f["sig"]), torch.from_numpy(f["digest"]), torch.empty(N, dtype=torch.uint8)]
    if pin:
        a = [x.pin_memory() for x in a]
    for _ in range(3):
        eng.rsa_verify_batch(a[0], a[1], a[2], out=a[3])
    t0 = time.perf_counter()
    for _ in range(10):
        eng.rsa_verify_batch(a[0], a[1], a[2], out=a[3])
    dt = (time.perf_counter() - t0) / 10
    assert np.array_equal(a[3].numpy(), f["expect"])
    print("flat call, pinned=%s: %.3f ms  (%.1f M/s)" % (pin, dt * 1e3, N / dt / 1e6), file=sys.stderr)
