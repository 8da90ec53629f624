"""Request sharding across GPUs: contiguous operation ranges, so the R tuples of one
operation land on one GPU and the tally needs no exchange.  No collective on the data path; ranks
only meet at a barrier and to reduce the elapsed time (max over ranks)."""
import numpy as np


def op_range(n_ops: int, world: int, rank: int):
    """GPU `rank` gets ops [rank*M/G, (rank+1)*M/G) — integer arithmetic, ranges tile [0, n_ops)."""
    return (n_ops * rank) // world, (n_ops * (rank + 1)) // world


def slice_ops(op_off: np.ndarray, lo: int, hi: int):
    """Offsets of ops [lo, hi) rebased to 0 and the tuple range they cover."""
    t0, t1 = int(op_off[lo]), int(op_off[hi])
    return (op_off[lo:hi + 1] - op_off[lo]).astype(np.uint32), t0, t1


def max_over_ranks(value_ms: float, dist=None, device=None) -> float:
    """Elapsed time of a multi-rank step = the slowest rank's device time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value_ms
    import torch
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
