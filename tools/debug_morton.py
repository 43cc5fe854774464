"""Bisects the uint64 Morton-decode mismatch found by tests/cpp/int_check: the same gather_bits chain (morton.h:85-101)
through the Python mirror, against numpy, with the inputs evaluated at different points."""
import sys; sys.path.insert(0, '.')
import numpy as np, enoki_b200 as ek

def magic(dim, level, nbits=64):
    maxb = nbits // dim; bs = min(1 << (level - 1), maxb); count = 0; value = 0; mask = 1 << (nbits - 1)
    for i in range(nbits):
        value >>= 1
        if count < maxb and (i // bs) % dim == 0:
            count += 1; value |= mask
    return value

def gather_bits(y, dim=2, xp=None):
    for Level in range(6, 0, -1):
        il = 6 - Level + 1
        m = magic(dim, il); sh = (1 << (il - 1)) * (dim - 1); sh = sh if sh < 64 else 0
        y = y & (np.uint64(m) if xp is np else ek.UInt64(m))
        if sh:
            y = y | (y >> (np.uint64(sh) if xp is np else ek.UInt64(sh)))
    return y

def scatter_bits(x, dim=2, xp=None):
    for Level in range(6, 0, -1):
        m = magic(dim, Level); sh = (1 << (Level - 1)) * (dim - 1); sh = sh if sh < 64 else 0
        if sh:
            x = x | (x << (np.uint64(sh) if xp is np else ek.UInt64(sh)))
        x = x & (np.uint64(m) if xp is np else ek.UInt64(m))
    return x

n = 100003
i = np.arange(n, dtype=np.uint64)
a = (i * np.uint64(2654435761) + np.uint64(12345)) & np.uint64(0xffffffff)
b = ((i ^ (i >> np.uint64(3))) * np.uint64(40503) + np.uint64(7)) & np.uint64(0xffffffff)
m2 = scatter_bits(a, xp=np) | (scatter_bits(b, xp=np) << np.uint64(1))
assert (gather_bits(m2, xp=np) == a).all() and (gather_bits(m2 >> np.uint64(1), xp=np) == b).all()

def report(tag, got, want):
    got = got.numpy()
    bad = int((got != want).sum())
    print(f"{tag:34s} mismatches={bad}", "" if not bad else f" first idx {np.flatnonzero(got != want)[:3]} got {got[got != want][:2]} want {want[got != want][:2]}")
    return bad

M = ek.UInt64.copy(m2)
report("V1 gather_bits(evaluated m2)", gather_bits(M), a)
report("V1b gather_bits(m2 >> 1)", gather_bits(M >> ek.UInt64(1)), b)
A = ek.UInt64.copy(a); B = ek.UInt64.copy(b)
E = scatter_bits(A) | (scatter_bits(B) << ek.UInt64(1))
report("V2 encode only", E, m2)
E = scatter_bits(A) | (scatter_bits(B) << ek.UInt64(1))
dx = gather_bits(E); dy = gather_bits(E >> ek.UInt64(1))
bad = report("V3 fused encode+decode x", dx, a)
bad += report("V3 fused encode+decode y", dy, b)
E = scatter_bits(A) | (scatter_bits(B) << ek.UInt64(1))
dx = gather_bits(E); dy = gather_bits(E >> ek.UInt64(1))
r = dx + dy * ek.UInt64(3)
del dx, dy, E
print(ek.debug_plan()) if "--plan" in sys.argv else None
report("V4 fused, combined dx + 3 dy", r, a + b * np.uint64(3))
I = ek.UInt64.arange(n)
X = I * ek.UInt64(2654435761) + ek.UInt64(12345)
Y = (I ^ (I >> ek.UInt64(3))) * ek.UInt64(40503) + ek.UInt64(7)
Am = X & ek.UInt64(0xffffffff); Bm = Y & ek.UInt64(0xffffffff)
E = scatter_bits(Am) | (scatter_bits(Bm) << ek.UInt64(1))
r = gather_bits(E) + gather_bits(E >> ek.UInt64(1)) * ek.UInt64(3)
del I, X, Y, Am, Bm, E
report("V5 everything fused from arange", r, a + b * np.uint64(3))
