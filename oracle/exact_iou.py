"""Exact rational convex-quad IoU (fractions.Fraction).  TEST INFRASTRUCTURE ONLY.

Used to pin the double-precision clip of pp_oracle.c / the Boost stand-in: the inputs are
doubles (exactly representable as Fractions), every intersection point and area is computed
without rounding, and only the final quotient is converted to float.
"""
from fractions import Fraction


def _area2(poly):
    n = len(poly)
    s = Fraction(0)
    for i in range(n):
        j = (i + 1) % n
        s += poly[i][0] * poly[j][1] - poly[j][0] * poly[i][1]
    return s  # twice the signed (CCW-positive) area


def exact_iou(a_ring_ccw, g_ring_cw):
    a = [(Fraction(float(x)), Fraction(float(y))) for x, y in a_ring_ccw]
    g = [(Fraction(float(x)), Fraction(float(y))) for x, y in g_ring_cw][::-1]
    subj = list(a)
    m = len(g)
    for e in range(m):
        c0, c1 = g[e], g[(e + 1) % m]
        ex, ey = c1[0] - c0[0], c1[1] - c0[1]
        nxt = []
        n = len(subj)
        for i in range(n):
            cur, prv = subj[i], subj[i - 1]
            dc = ex * (cur[1] - c0[1]) - ey * (cur[0] - c0[0])
            dp = ex * (prv[1] - c0[1]) - ey * (prv[0] - c0[0])
            cin, pin = dc >= 0, dp >= 0
            if cin != pin:
                t = dp / (dp - dc)
                nxt.append((prv[0] + t * (cur[0] - prv[0]), prv[1] + t * (cur[1] - prv[1])))
            if cin:
                nxt.append(cur)
        subj = nxt
        if not subj:
            break
    if len(subj) < 3:
        return 0.0
    inter = _area2(subj)
    if inter <= 0:
        return 0.0
    union = _area2(a) + _area2(g) - inter
    return float(inter / union)
