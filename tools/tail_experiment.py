"""How much of K1's time at N = 65536 is the partially filled last wave?  Compares per-item throughput of
(a) back-to-back launches on one stream, (b) launches alternating over two / three streams (the next batch's
blocks fill the tail), (c) a batch that is an exact multiple of the resident capacity."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bftkv_b200 import Engine, workload

dev = torch.device("cuda", 0)
w = workload.make_verify_batch(65536, 16, seed=1, corrupt_seed=2, threads=16)
eng = Engine(0)
eng.register_rsa_keys([k["n"] for k in w["keys"]], [k["e"] for k in w["keys"]])


def run(n_items, n_streams, steps=24, copies=8):
    reps = (n_items + 65535) // 65536
    idx = np.tile(w["key_idx"].astype(np.int32), reps)[:n_items]
    sig = np.tile(w["sig"], (reps, 1))[:n_items]
    dig = np.tile(w["digest"], (reps, 1))[:n_items]
    d = [(torch.from_numpy(idx).to(dev), torch.from_numpy(sig).to(dev), torch.from_numpy(dig).to(dev),
          torch.empty(n_items, dtype=torch.uint8, device=dev)) for _ in range(copies)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    for i in range(4):
        eng.rsa_verify_batch_dev(d[i % copies][0], d[i % copies][1], d[i % copies][2], n_items, d[i % copies][3], stream=streams[i % n_streams].cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for i in range(steps):
        c = i % copies
        eng.rsa_verify_batch_dev(d[c][0], d[c][1], d[c][2], n_items, d[c][3], stream=streams[i % n_streams].cuda_stream)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    exp = np.tile(w["expect"], reps)[:n_items]
    assert np.array_equal(d[0][3].cpu().numpy(), exp)
    return n_items * steps / ms / 1e3

sm = eng.sm_count if hasattr(eng, "sm_count") else 148
cap = 148 * 4 * 32
out = {}
for name, n, ns in [("65536_1stream", 65536, 1), ("65536_2streams", 65536, 2), ("65536_3streams", 65536, 3), ("%d_1stream" % (4 * cap), 4 * cap, 1),
                    ("%d_1stream" % (3 * cap), 3 * cap, 1), ("131072_1stream", 131072, 1), ("262144_1stream", 262144, 1)]:
    out[name] = round(run(n, ns), 3)
    print(name, out[name], "M verifies/s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tail_experiment.json", "w"))
