"""Python emulation of the planned Montgomery SQUARING for bftkv_b200/csrc/rsa_verify_r32.cuh (DESIGN.md §7, K1 next
steps): a^2 first, by 512-bit blocks with symmetry, then the reduction half of the existing loop with the high half
of a^2 fed in at the top.  Checks the block assignment, the scatter table and the feed bookkeeping limb for limb
against big-int arithmetic before any CUDA is written.

Lane L of a 4-lane group owns A_L = a[16L .. 16L+16).  Every lane runs the same three product slots
  slot 0   D_L  = A_L^2                              (inner symmetry: 136 instead of 256 limb products)
  slot 1   X_L  = A_L * A_{(L+1) mod 4}              (256)   -> the four pairs at cyclic distance 1: 01 12 23 30
  slot 2   H_L  = C_L * half_L(A_{(L&1)+2})           (128)   -> pairs 02 and 13, each split between two lanes:
                 C_L = A_{L&1} (own block for L < 2, a copy for L >= 2), half_L = low 8 limbs (L < 2) / high 8 (L >= 2)
= 520 limb products per lane instead of 1024, in lock-step.  The results (all but D_L doubled) land at
  D_L @ 32L      X_L @ 16 (L + (L+1) mod 4)      H_L @ 16 ((L&1) + (L&1) + 2) + 8 (L >= 2)
and are scattered through shared memory into T = a^2 as 8 units of 16 limbs, lane r keeping units r and 4 + r."""
import random

B = 1 << 32
T4, W = 4, 16


def limbs(x, n):
    return [(x >> (32 * i)) & (B - 1) for i in range(n)]


def val(ls):
    return sum(v << (32 * i) for i, v in enumerate(ls))


def square_blocks(a):
    """Per lane: the three slot results as (absolute limb position, value, doubled?)."""
    A = [val(limbs(a, 64)[16 * L:16 * L + 16]) for L in range(T4)]
    out = []
    macs = []
    for L in range(T4):
        res = []
        # slot 0: own square with inner symmetry (count the limb products actually needed)
        al = limbs(A[L], 16)
        d = sum(al[i] * al[i] << (64 * i) for i in range(16)) + 2 * sum(al[i] * al[j] << (32 * (i + j)) for i in range(16) for j in range(i + 1, 16))
        assert d == A[L] * A[L]
        res.append((32 * L, d, False))
        M = (L + 1) % 4
        res.append((16 * (L + M), A[L] * A[M], True))
        c = A[L & 1]
        other = A[(L & 1) + 2]
        half = (other & ((1 << 256) - 1)) if L < 2 else (other >> 256)
        res.append((16 * ((L & 1) + (L & 1) + 2) + (8 if L >= 2 else 0), c * half, True))
        out.append(res)
        macs.append(16 * 17 // 2 + 256 + 128)
    return out, macs


def scatter(results):
    """Shared-memory scatter: unit u (16 limbs) = sum of the pieces of every slot result that overlap it.  Returns the
    8 unit sums as plain integers (each may exceed 2^512: the overflow belongs to the next unit)."""
    units = [0] * 8
    table = [[] for _ in range(8)]                     # the static contribution table the kernel would hold
    for L, res in enumerate(results):
        for slot, (pos, v, dbl) in enumerate(res):
            n = 32 if slot < 2 else 24
            ls = limbs(v, n)
            for u in range(8):
                lo, hi = max(pos, 16 * u), min(pos + n, 16 * u + 16)
                if lo >= hi:
                    continue
                piece = val(ls[lo - pos:hi - pos]) << (32 * (lo - 16 * u))
                units[u] += piece * (2 if dbl else 1)
                table[u].append((L, slot, lo - pos, hi - lo, lo - 16 * u, dbl))
    return units, table


def redc_with_feed(t, n, n0inv):
    """The reduction half of mont_mul: window = low 64 limbs of t, two quotient digits per round, the window moves
    down two limbs per round and limbs 64 + 2 rnd, 65 + 2 rnd of t enter at the top."""
    tl = limbs(t, 130)
    win = val(tl[:64])
    for rnd in range(32):
        q0 = (win & (B - 1)) * n0inv & (B - 1)
        win += q0 * n
        assert win & (B - 1) == 0
        q1 = ((win >> 32) & (B - 1)) * n0inv & (B - 1)
        win += (q1 * n) << 32
        assert win & (B * B - 1) == 0
        win >>= 64
        win += (tl[64 + 2 * rnd] + (tl[65 + 2 * rnd] << 32)) << (32 * 62)
    return win + (val(tl[128:]) << (32 * 64))


if __name__ == "__main__":
    random.seed(2)
    R = 1 << 2048
    worst = 0
    tab0 = None
    for it in range(300):
        n = random.getrandbits(2048) | (1 << 2047) | 1
        a = random.getrandbits(2048)
        if it % 7 == 0:
            a = R - 1
        if it % 11 == 0:
            n = R - 1
        if it % 13 == 0:
            a = 0
        res, macs = square_blocks(a)
        units, table = scatter(res)
        assert sum(u << (512 * i) for i, u in enumerate(units)) == a * a          # the block assignment tiles a^2 exactly once
        assert all(m == 520 for m in macs)
        worst = max(worst, max(len(t) for t in table))
        tab0 = tab0 or table
        assert table == tab0                                                       # the scatter table is static
        n0inv = (-pow(n, -1, B)) % B
        got = redc_with_feed(a * a, n, n0inv)
        want = (a * a + ((a * a * (-pow(n, -1, R))) % R) * n) // R
        assert got == want and got < R + n
    print("squaring emulation ok: 520 limb products per lane, at most %d contributions per unit" % worst)
    for u, t in enumerate(tab0):
        print("unit %d (lane %d %s):" % (u, u % 4, "lo" if u < 4 else "hi"),
              ", ".join("lane%d.slot%d[%d:+%d]->%d%s" % (L, s, o, ln, d, "x2" if dbl else "") for L, s, o, ln, d, dbl in t))
