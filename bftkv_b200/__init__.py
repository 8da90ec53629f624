"""bftkv_b200 — B200-native batched signature-verify + quorum-tally engine for yahoo/bftkv's hot path.
Host-side mirror of the reference interfaces over libbftq.so (include/bftq.h)."""
from .engine import Engine  # noqa: F401
