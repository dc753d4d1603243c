"""Developer tool: near-minimax (Chebyshev-fit) polynomial coefficients of csrc/fast_math.cuh, computed with mpmath at 60 digits.

  exp(r)  = 1 + r + r^2 Q(r),            |r| <= ln2 / 2          Q = (exp(r) - 1 - r) / r^2,          degree 9
  log(m)  = 2 f + 2 f s G(s),  f = (m - 1) / (m + 1), s = f^2,   G = (atanh(sqrt s) / sqrt s - 1) / s, degree 7,  m in [sqrt(1/2), sqrt 2]
"""
import mpmath as mp

mp.mp.dps = 60


def Q(r):
    r = mp.mpf(r)
    if abs(r) < mp.mpf("1e-20"):
        return mp.mpf(1) / 2 + r / 6
    return (mp.e ** r - 1 - r) / (r * r)


def G(s):
    s = mp.mpf(s)
    if s < mp.mpf("1e-30"):
        return mp.mpf(1) / 3 + s / 5
    t = mp.sqrt(s)
    return (mp.atanh(t) / t - 1) / s


def fit(f, a, b, n):
    c, err = mp.chebyfit(f, [a, b], n, error=True)   # highest power first
    return [float(x) for x in c[::-1]], float(err)


if __name__ == "__main__":
    half = mp.log(2) / 2 * mp.mpf("1.0005")
    q, eq = fit(Q, -half, half, 10)
    fmax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
    g, eg = fit(G, 0, fmax * fmax * mp.mpf("1.001"), 8)
    print("// exp: Q, lowest power first; max fit error %.2e (times r^2 <= 0.12)" % eq)
    print("FM_EXP_Q = {" + ", ".join("%.17g" % v for v in q) + "};")
    print("// log: G, lowest power first; max fit error %.2e (times 2 f s <= 0.01)" % eg)
    print("FM_LOG_G = {" + ", ".join("%.17g" % v for v in g) + "};")
    ln2 = mp.log(2)
    # ln2_hi: ln2 with the low 21 bits of the mantissa cleared (e * ln2_hi exact for |e| < 2^11)
    import struct
    bits = struct.unpack("<Q", struct.pack("<d", float(ln2)))[0] & ~((1 << 21) - 1)
    hi = struct.unpack("<d", struct.pack("<Q", bits))[0]
    lo = float(ln2 - mp.mpf(hi))
    print("LN2_HI = %.17g; LN2_LO = %.17g; INV_LN2 = %.17g;" % (hi, lo, float(1 / ln2)))
