"""Where does the host-API (e2e) leg lose time against the device-resident one?"""
import sys, os, time, threading, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bftkv_b200 import Engine, workload

dev = torch.device("cuda", 0)
N = 65536
w = workload.make_verify_batch(N, 16, seed=1, corrupt_seed=2, threads=16)
eng = Engine(0)
eng.register_rsa_keys([k["n"] for k in w["keys"]], [k["e"] for k in w["keys"]])
idx, sig, dig = w["key_idx"].astype(np.int32), w["sig"], w["digest"]

# raw PCIe
h = torch.from_numpy(sig).pin_memory(); d = torch.empty_like(h, device=dev)
torch.cuda.synchronize()
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("H2D 16 MiB pinned: %.3f ms  %.1f GB/s" % (dt * 1e3, h.numel() / dt / 1e9), flush=True)


def run(callers, chunk, total_calls=48):
    bufs = []
    for c in range(callers):
        lo = (c * chunk) % N
        sl = slice(lo, lo + chunk) if lo + chunk <= N else slice(0, chunk)
        bufs.append((torch.from_numpy(idx[sl].copy()).pin_memory(), torch.from_numpy(sig[sl].copy()).pin_memory(),
                     torch.from_numpy(dig[sl].copy()).pin_memory(), torch.empty(chunk, dtype=torch.uint8).pin_memory(), sl))
    for b in bufs:
        for _ in range(3): eng.rsa_verify_batch(b[0], b[1], b[2], out=b[3])
    per = total_calls * (N // chunk) // callers
    def work(b):
        for _ in range(per): eng.rsa_verify_batch(b[0], b[1], b[2], out=b[3])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(b,)) for b in bufs]
    [t.start() for t in ths]; [t.join() for t in ths]
    dt = time.perf_counter() - t0
    for b in bufs: assert np.array_equal(b[3].numpy(), w["expect"][b[4]])
    return per * callers * chunk / dt / 1e6

res = {}
for callers, chunk in [(1, 65536), (2, 65536), (3, 65536), (4, 65536), (2, 32768), (4, 32768), (4, 16384), (8, 16384), (8, 8192)]:
    r = run(callers, chunk)
    res["%dx%d" % (callers, chunk)] = round(r, 2)
    print("callers=%d chunk=%d: %.2f M verifies/s" % (callers, chunk, r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/e2e_experiment.json", "w"))
