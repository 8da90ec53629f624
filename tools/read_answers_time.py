"""Times bftq_read_responses_batch alone: python tools/read_answers_time.py [ops] [ss_signers] — raw 16-replica answers from
page-locked blobs, prints answers/s and the share parsed on the GPU."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bftkv_b200 import Engine, workload  # noqa: E402
from bftkv_b200.crypto_gpu import Keyring, _blob, read_responses_batch  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
SS = int(sys.argv[2]) if len(sys.argv) > 2 else 11
ra = workload.make_read_answers(M, 16, ss_signers=SS, mix=workload.HARD_MIX)
eng = Engine(0)
kr = Keyring(eng)
kr.register(ra["keyring"])
blob, off = _blob(ra["msgs"])
pin = (eng.host_copy(blob), eng.host_copy(off))
qcs = [(5, 16, 6, 11, ra["ids"])]
for _ in range(2):
    got = read_responses_batch(kr, qcs, ra["op_off"], ra["peer_ids"], None, ra["nonces"], pre_status=ra["pre_status"], blobs=pin)
assert np.array_equal(got["status"] != 0, ra["expect_status"] != 0)
s0 = eng.stats()
t0 = time.perf_counter()
n = 5
for _ in range(n):
    read_responses_batch(kr, qcs, ra["op_off"], ra["peer_ids"], None, ra["nonces"], pre_status=ra["pre_status"], blobs=pin)
dt = (time.perf_counter() - t0) / n
s1 = eng.stats()
print("answers %d bytes/answer %d: %.2f ms per call = %.2f M answers/s, %.1f GB/s H2D; gpu-parsed %d host %d" % (
    M * 16, int(off[-1]) // (M * 16), dt * 1e3, M * 16 / dt / 1e6, (s1["h2d_bytes"] - s0["h2d_bytes"]) / n / dt / 1e9,
    (s1["msg_gpu_items"] - s0["msg_gpu_items"]) // n, (s1["msg_host_items"] - s0["msg_host_items"]) // n))
