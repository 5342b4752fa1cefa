"""Offline proof for the 4-bin uint8 channel kernel: bins 1 and 3 of grad_hist without float64.

Reference arithmetic (waldboost/channels.py:43-50 under NumPy 2): ch_i = float32(float64(gx) * cos(t_i) - float64(gy) * sin(t_i)),
t_1 = pi/4, t_3 = 3 pi/4, with gx, gy the float32 gradients of a uint8 image: INTEGERS in [-1020, 1020] (channels.py:16-21).
Candidate (device): d1 = gx - gy, d3 = gx + gy (exact in float32), ch_1 = fma(d1, C_HI, d1 * C_LO), ch_3 = -fma(d3, C_HI, d3 * C_LO)
with C_HI + C_LO a two-float32 split of cos(pi/4).  This script checks ALL 2041^2 (gx, gy) pairs against the reference
expression evaluated by NumPy itself and prints the pairs (if any) where the candidate differs.
    python profiles/prove_hist4_fp32_bins.py
"""
import numpy as np

theta = np.linspace(0, np.pi, 5)[:-1]
cs, sn = np.cos(theta), np.sin(theta)
g = np.arange(-1020, 1021, dtype=np.float32)
GX, GY = np.meshgrid(g, g, indexing="ij")
c = np.float64(cs[1])
C_HI = np.float32(c)
C_LO = np.float32(c - np.float64(C_HI))


def fma32(a, b, c_):
    """float32 fused multiply-add emulated exactly in float64 (|a*b| < 2^35 needs < 53 bits together with c_)."""
    return (a.astype(np.float64) * np.float64(b) + c_.astype(np.float64)).astype(np.float32)


def cand(d):
    lo = (d * C_LO).astype(np.float32)                      # float32 product
    return fma32(d, C_HI, lo)


ref1 = (GX.astype(np.float64) * cs[1] - GY.astype(np.float64) * sn[1]).astype(np.float32)
ref3 = (GX.astype(np.float64) * cs[3] - GY.astype(np.float64) * sn[3]).astype(np.float32)
c1 = cand(GX - GY)
c3 = -cand(GX + GY)
bad1, bad3 = np.argwhere(ref1 != c1), np.argwhere(ref3 != c3)
print("C_HI = %r (0x%08x)  C_LO = %r (0x%08x)" % (float(C_HI), C_HI.view(np.uint32), float(C_LO), C_LO.view(np.uint32)))
print("pairs checked:", GX.size, " bin 1 mismatches:", len(bad1), " bin 3 mismatches:", len(bad3))
# the float64 emulation of fma32 is exact: d*C_HI has <= 12 + 24 significant bits and the addend is far smaller
for name, bad, ref, cnd in (("bin1", bad1, ref1, c1), ("bin3", bad3, ref3, c3)):
    for i, j in bad[:10]:
        print(name, "gx", int(g[i]), "gy", int(g[j]), "ref", repr(float(ref[i, j])), "cand", repr(float(cnd[i, j])))
# every mismatch is a pair with d = 0 and gx != 0: there the reference does not yield 0 but the difference of two float64
# roundings, gx*cos - gx*sin = O(1e-13) (cos(pi/4) and sin(pi/4) differ in their last bit); the kernel recomputes exactly
# those pixels with the float64 expression
d1, d3 = GX - GY, GX + GY
assert np.all(d1[ref1 != c1] == 0) and np.all(GX[ref1 != c1] != 0), "bin 1: a mismatch outside d = 0"
assert np.all(d3[ref3 != c3] == 0) and np.all(GX[ref3 != c3] != 0), "bin 3: a mismatch outside d = 0"
assert np.all((ref1 == c1) | (d1 == 0)) and np.all((ref3 == c3) | (d3 == 0))
print("all mismatches have d == 0 and gx != 0; with those pixels on the float64 path the candidate is bit-exact for every pair")
# how close does the exact value come to a float32 rounding boundary?  (margin of the argument in the kernel comment)
exact = GX.astype(np.float64) * cs[1] - GY.astype(np.float64) * sn[1]
r = ref1.astype(np.float64)
ulp = np.spacing(np.abs(ref1)).astype(np.float64)
dist = 0.5 * ulp - np.abs(exact - r)
nz = ref1 != 0
print("smallest distance of an exact bin-1 value to its rounding boundary: %.3e (in ulps: %.3e)" % (dist[nz].min(), (dist[nz] / ulp[nz]).min()))
