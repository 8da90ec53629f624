"""Packet-level end-to-end sweep on one GPU: bftq_signature_verify_batch over the OpenPGP form of BASELINE
configs[1] under different chunk sizes / worker-thread counts / concurrent callers.
python tools/pgp_e2e_experiment.py > gpurun_out/pgp_e2e_experiment.json"""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from bftkv_b200 import Engine, _lib, workload
from bftkv_b200.crypto_gpu import Keyring, _blob

N = 65536


def main():
    w = workload.make_pgp_verify_batch(N)
    eng = Engine(0)
    kr = Keyring(eng)
    kr.register(w["keyring"])
    lib = _lib.load()
    tb, to = _blob(w["tbs"])
    sb, so = _blob(w["sigs"])
    p = lambda a: C.c_void_p(a.ctypes.data)
    rows = []

    def run(chunk, threads, callers, steps=24):
        if chunk:
            os.environ["BFTQ_PLAN_CHUNK"] = str(chunk)
        else:
            os.environ.pop("BFTQ_PLAN_CHUNK", None)
        if threads:
            os.environ["BFTQ_HOST_THREADS"] = str(threads)
        else:
            os.environ.pop("BFTQ_HOST_THREADS", None)
        errs = [np.zeros(N, np.int32) for _ in range(callers)]

        def caller(c, n):
            for _ in range(n):
                _lib.check(lib.bftq_signature_verify_batch(kr._h, p(tb), p(to), p(sb), p(so), N, p(errs[c])))
        ths = [threading.Thread(target=caller, args=(c, 3)) for c in range(callers)]      # warm up with the same concurrency
        [t.start() for t in ths]
        [t.join() for t in ths]
        share = [steps // callers + (1 if c < steps % callers else 0) for c in range(callers)]
        t0 = time.perf_counter()
        ths = [threading.Thread(target=caller, args=(c, share[c])) for c in range(callers)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        dt = time.perf_counter() - t0
        for c in range(callers):
            assert np.array_equal(errs[c] == 0, w["expect_ok"])
        rows.append({"gpu_parse": os.environ.get("BFTQ_GPU_PARSE", "1"), "chunk": chunk, "threads": threads, "callers": callers, "verifies_per_sec": N * steps / dt, "ms_per_batch": dt / steps * 1e3})
        print(rows[-1], file=sys.stderr)

    for gp, ft, callers in (("1", 8, 2), ("1", 2, 2), ("1", 8, 1), ("1", 2, 1), ("1", 4, 4), ("0", 0, 2), ("0", 0, 4), ("0", 0, 1)):
        os.environ["BFTQ_GPU_PARSE"] = gp
        if ft:
            os.environ["BFTQ_FAST_THREADS"] = str(ft)
        else:
            os.environ.pop("BFTQ_FAST_THREADS", None)
        run(0, 0, callers)
        rows[-1]["fast_threads"] = ft
        rows[-1]["cuda_device_max_connections"] = os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "default (8)")
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
