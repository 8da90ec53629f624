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


# ---- limb-level emulation: the same chains, accumulators and carries the CUDA code will use ---------------------
def chain16(acc, base, x_limbs, m):
    """Chain<16>::run(p = acc + base, carry limbs acc[base+16], acc[base+17], operands x_limbs (8 of them), m)."""
    c = 0
    for k in range(8):
        lo = base + 2 * k
        v = acc[lo] + (acc[lo + 1] << 32) + x_limbs[k] * m + c
        acc[lo] = v & (B - 1); acc[lo + 1] = (v >> 32) & (B - 1); c = v >> 64
    for k in (base + 16, base + 17):
        v = acc[k] + c
        acc[k] = v & (B - 1); c = v >> 32
    assert c == 0


def block_mul(x, rows):
    """acc = x (16 limbs) * rows (NR limbs): per row an even chain (x0, x2, ..) at offset i and an odd chain
    (x1, x3, ..) at offset i + 1, one accumulator of NR + 19 limbs."""
    nr = len(rows)
    acc = [0] * (nr + 19)
    for i, m in enumerate(rows):
        chain16(acc, i, x[0::2], m)
        chain16(acc, i + 1, x[1::2], m)
    assert all(v == 0 for v in acc[nr + 16:])
    return acc[:nr + 16]


SLOT_LEN = (32, 32, 24)
SLOT_BASE = (0, 32, 64)            # word offsets of the three slot results in a lane's shared-memory column


def mont_sqr_limbs(a, n, n0inv, table):
    al = [limbs(a, 64)[16 * L:16 * L + 16] for L in range(T4)]
    nl = [limbs(n, 64)[16 * L:16 * L + 16] for L in range(T4)]
    # slots -> shared memory (one 88-word column per lane)
    sm = []
    for L in range(T4):
        cpy = al[L & 1]
        src = al[L | 2]
        rows2 = src[:8] if L < 2 else src[8:]
        col = block_mul(al[L], al[L]) + block_mul(al[L], al[(L + 1) % 4]) + block_mul(cpy, rows2)
        assert len(col) == 88
        sm.append(col)
    # unit sums: 16 limbs + overflow word, contributions added once or twice, predicated per position
    unit, ov = [None] * 8, [0] * 8
    for u in range(8):
        v = [0] * 16
        o = 0
        for (L, slot, off, ln, dst, dbl) in table[u]:
            for _ in range(2 if dbl else 1):
                c = 0
                for p in range(16):
                    w = sm[L][SLOT_BASE[slot] + off + p - dst] if dst <= p < dst + ln else 0
                    s = v[p] + w + c
                    v[p] = s & (B - 1); c = s >> 32
                o += c
        unit[u], ov[u] = v, o
    # carry normalisation, lo units (lanes 0..3) then hi units, as mont_mul's tail does it
    def normalise(vs, hs, carry_into_first):
        fb = [carry_into_first] + hs[:3]
        g, ones = [0] * 4, [False] * 4
        for r in range(4):
            c = fb[r]
            for p in range(16):
                s = vs[r][p] + c
                vs[r][p] = s & (B - 1); c = s >> 32
            g[r] = c
            ones[r] = all(x == B - 1 for x in vs[r])
        cin_ = [0] * 5
        for r in range(1, 5):
            cin_[r] = g[r - 1] | (1 if ones[r - 1] and cin_[r - 1] else 0)
        for r in range(4):
            c = cin_[r]
            for p in range(16):
                s = vs[r][p] + c
                vs[r][p] = s & (B - 1); c = s >> 32
        return hs[3] + cin_[4]
    top_lo = normalise(unit[:4], ov[:4], 0)
    top_hi = normalise(unit[4:], ov[4:], top_lo)
    assert top_hi == 0
    tval = sum(val(unit[u]) << (512 * u) for u in range(8))
    assert tval == a * a
    # reduction: E/O/Z/cin as in mont_mul, the a x b chains gone, the high units fed in at the top lane
    E = [unit[r] + [0, 0, 0] for r in range(T4)]
    O = [[0] * 17 for _ in range(T4)]
    cin = [0] * T4
    Z = [0] * T4
    def chain(arr, idxpairs, xs, m, c=0):
        for lo, xi in idxpairs:
            v = arr[lo] + (arr[lo + 1] << 32) + xs[xi] * m + c
            arr[lo] = v & (B - 1); arr[lo + 1] = (v >> 32) & (B - 1); c = v >> 64
        return c
    def end(arr, idx, c, n2):
        for k in range(n2):
            v = arr[idx + k] + c; arr[idx + k] = v & (B - 1); c = v >> 32
        assert c == 0
    for owner in range(T4):
        for jj in range(0, W, 2):
            q0 = ((E[0][0] + Z[0] + cin[0]) & (B - 1)) * n0inv & (B - 1)
            for r in range(T4):
                c = chain(E[r], [(k, k) for k in range(0, W, 2)], nl[r], q0, cin[r]); end(E[r], 16, c, 2)
                c = chain(O[r], [(k - 1, k) for k in range(1, W, 2)], nl[r], q0); end(O[r], 16, c, 1)
            s0 = [E[r][0] + Z[r] for r in range(T4)]
            c0 = [x >> 32 for x in s0]; p0 = [x & (B - 1) for x in s0]
            q1 = ((E[0][1] + O[0][0] + c0[0]) & (B - 1)) * n0inv & (B - 1)
            p1 = [0] * T4
            for r in range(T4):
                c = chain(O[r], [(k, k) for k in range(0, W, 2)], nl[r], q1); end(O[r], 16, c, 1)
                c = chain(E[r], [(k + 1, k) for k in range(1, W, 2)], nl[r], q1); end(E[r], 18, c, 1)
                s = E[r][1] + O[r][0] + c0[r]; p1[r] = s & (B - 1); cin[r] = s >> 32
            assert p0[0] == 0 and p1[0] == 0
            f0, f1 = unit[4 + owner][jj], unit[4 + owner][jj + 1]
            for r in range(T4):
                r0 = p0[r + 1] if r < T4 - 1 else f0
                r1 = p1[r + 1] if r < T4 - 1 else f1
                Z[r] = O[r][1]
                E[r] = E[r][2:] + [0, 0]; O[r] = O[r][2:] + [0, 0]
                v = E[r][14] + (E[r][15] << 32) + (E[r][16] << 64) + (E[r][17] << 96) + r0 + (r1 << 32)
                E[r][14] = v & (B - 1); E[r][15] = (v >> 32) & (B - 1); E[r][16] = (v >> 64) & (B - 1); E[r][17] = (v >> 96) & (B - 1)
    tot = 0
    for r in range(T4):
        loc = cin[r] + Z[r] + sum(E[r][k] << (32 * k) for k in range(19)) + sum(O[r][k] << (32 * (k + 1)) for k in range(17))
        assert E[r][17] == 0 and E[r][18] == 0 and O[r][15] == 0 and O[r][16] == 0 and E[r][16] <= 4, (E[r][16:], O[r][15:])
        tot += loc << (32 * W * r)
    return tot


def run_limb_level():
    random.seed(3)
    R = 1 << 2048
    _, table = scatter(square_blocks(random.getrandbits(2048))[0])
    for it in range(200):
        n = random.getrandbits(2048) | (1 << 2047) | 1
        a = random.getrandbits(2048)
        if it % 7 == 0:
            a = R - 1
        if it % 11 == 0:
            n = R - 1
        if it % 13 == 0:
            a = (1 << 2048) - (1 << 1024) - 1
        if it % 17 == 0:
            a = 0
        n0inv = (-pow(n, -1, B)) % B
        got = mont_sqr_limbs(a, n, n0inv, table)
        assert got == (a * a + ((a * a * (-pow(n, -1, R))) % R) * n) // R and got < R + n, it
    print("limb-level squaring emulation ok")


if __name__ == "__main__":
    run_limb_level()
