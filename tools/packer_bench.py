"""Host-side packer throughput (no GPU needed): bftq_signature_plan_measure over the OpenPGP-packet form
of BASELINE configs[1].  python tools/packer_bench.py [n_items] [threads...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bftkv_b200 import _lib, workload
from bftkv_b200.crypto_gpu import Keyring, _blob


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    threads = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8, 16]
    t0 = time.time()
    w = workload.make_pgp_verify_batch(n)
    print("generated %d items in %.1f s" % (n, time.time() - t0), file=sys.stderr)
    lib = _lib.load()
    kr = Keyring(None)
    kr.register(w["keyring"])
    tb, to = _blob(w["tbs"])
    sb, so = _blob(w["sigs"])
    p = lambda a: C.c_void_p(a.ctypes.data)
    for th in threads:
        best = 1e9
        for _ in range(5):
            nt, sec = C.c_uint64(), C.c_double()
            _lib.check(lib.bftq_signature_plan_measure(kr._h, p(tb), p(to), p(sb), p(so), n, th, C.byref(nt), C.byref(sec)))
            best = min(best, sec.value)
        print("threads=%2d  tuples=%d  %.3f ms  %.2f M items/s" % (th, nt.value, best * 1e3, n / best / 1e6))


if __name__ == "__main__":
    main()
