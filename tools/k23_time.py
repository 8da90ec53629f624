"""K2 (read_tally_kernel, 65 536 ops x 16 responders) and K3 (lagrange_combine_kernel<8>, 17 476 combines of 10 shares mod the P-256 order)
alone, device-resident, CUDA events — the two secondary kernels of bench.py's roofline_secondary, small enough to run under ncu:
  ncu --set full -k regex:'read_tally|lagrange_combine' --launch-skip 4 -c 4 python tools/k23_time.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bftkv_b200 import Engine, workload, _lib  # noqa: E402

dev = torch.device("cuda", 0)
eng = Engine(0)
stream = torch.cuda.Stream(device=dev)
R, M = 16, 65536
w = workload.make_verify_batch(4096, 16, corrupt_rate=0.0, unknown_rate=0.0)
eng.register_rsa_keys([k["n"] for k in w["keys"]], [k["e"] for k in w["keys"]])
ro = workload.make_read_ops(w, M, R, seed=4)
quorum = eng.quorum_create([(5, 16, 6, 11, list(range(16)))])
NQ = M * R
dq = {k: torch.from_numpy(v).to(dev) for k, v in [("off", ro["op_off"].astype(np.int32)), ("idx", ro["key_idx"].astype(np.int32)), ("sig", ro["sig"]),
                                                 ("dig", ro["digest"]), ("ts", ro["ts"].astype(np.int64)), ("val", ro["value_id"].astype(np.int32))]}
dq_st = torch.from_numpy(ro["expect_status"]).to(dev)
dq_bits = torch.empty(M, dtype=torch.uint8, device=dev)
dq_win = torch.empty(M, dtype=torch.int32, device=dev)


def k2():
    eng.verify_tally_batch_dev(quorum, dq["off"], dq["idx"], dq["sig"], dq["dig"], M, 0, dq_st, dq_bits, d_ts=dq["ts"], d_value_id=dq["val"], d_winner=dq_win,
                               stream=stream.cuda_stream)


def timed(fn, reps):
    for _ in range(3):
        fn()
    stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    stream.synchronize()
    return a.elapsed_time(b) / reps


k2_ms = timed(k2, 20)
q256 = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
Bc, kc = 262144 // 15, 10
rgen = np.random.default_rng(7)
xs = np.stack([rgen.permutation(15)[:kc] + 1 for _ in range(Bc)]).astype(np.int32)
ysb = rgen.integers(0, 256, (Bc, kc, 32), dtype=np.uint8)
ysb[:, :, 0] &= 0x7F
dl = [torch.from_numpy(x).to(dev) for x in (xs, ysb)]
dl_out, dl_st = torch.empty((Bc, 32), dtype=torch.uint8, device=dev), torch.empty(Bc, dtype=torch.uint8, device=dev)
mb = np.frombuffer(q256.to_bytes(32, "big"), np.uint8).copy()


def k3():
    _lib.check(eng._lib.bftq_lagrange_combine_batch_dev(eng._h, mb.ctypes.data_as(C.c_void_p), 32, kc, C.c_void_p(dl[0].data_ptr()), C.c_void_p(dl[1].data_ptr()), Bc,
                                                        C.c_void_p(dl_out.data_ptr()), C.c_void_p(dl_st.data_ptr()), C.c_void_p(stream.cuda_stream)))


k3_ms = timed(k3, 10)
from oracle import sss_oracle  # noqa: E402  (checker only)
out = dl_out.cpu().numpy()
for j in (0, 1, 12345 % Bc, Bc - 1):
    exp = sss_oracle.calculate_secret([(int(xs[j, i]), int.from_bytes(ysb[j, i].tobytes(), "big")) for i in range(kc)], q256)
    assert int.from_bytes(out[j].tobytes(), "big") == exp
assert not dl_st.cpu().numpy().any()
print({"k2_read_tally_ms": k2_ms, "k2_ops": M, "k3_lagrange_ms": k3_ms, "k3_combines": Bc, "k3_combines_per_s": Bc / k3_ms * 1e3})
