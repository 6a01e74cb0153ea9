"""Generates the committed golden fixtures from the reference's own compiled
code (oracle/_ref/libcassref.so, built from /root/reference by oracle/Makefile).
Run in the builder container:  python tests/golden/make_golden.py
Fixtures are small (<1 MB) .npz files: inputs are regenerated from seeds by
tests/cases.py, outputs are stored."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O      # noqa: E402
from tests import cases              # noqa: E402


def pack(out):
    d = {}
    for z in range(len(out[0])):
        d["rcv%d" % z] = out[0][z]; d["dnr%d" % z] = out[1][z]; d["typ%d" % z] = out[2][z]
        d["cf%d" % z] = out[3][z]; d["ext%d" % z] = out[4][z]
    d["orphan"] = out[5]
    return d


def golden_cases():
    """name -> (donor tuples, receptor xyz, kwargs)"""
    g = {}
    for v in "abc":
        A, B = cases.c1(v, n=17, m=11)
        g["c1%s_BfromA_o2" % v] = ([A], B[3:], dict(order=2, nature=1, penalty=1, extrap=1))
        g["c1%s_AfromB_o2" % v] = ([B], A[3:], dict(order=2, nature=1, penalty=1, extrap=1))
    Bg, Oc = cases.c2(nb=20, nc=(33, 9, 17))
    rng = np.random.default_rng(7)
    P = rng.uniform(-3.2, 3.2, (3, 3000)); P[2] = rng.uniform(-0.1, 2.1, 3000)
    for order in (2, 3, 5):
        n = 600 if order == 5 else 3000
        g["c2_multi_o%d" % order] = ([Bg, Oc], (P[0][:n], P[1][:n], P[2][:n]),
                                     dict(order=order, nature=0, penalty=1, extrap=1))
        g["c2_OfromBg_o%d" % order] = ([Bg], (Oc[3][::7][:n], Oc[4][::7][:n], Oc[5][::7][:n]),
                                       dict(order=order, nature=1, penalty=1, extrap=1))
    # order 4 (type 44) and regular Cartesian donors (interpDataType=0, InterpCart); `types` = per-zone interpDataType
    n = 3000
    g["c2_multi_o4"] = ([Bg, Oc], (P[0][:n], P[1][:n], P[2][:n]), dict(order=4, nature=0, penalty=1, extrap=1))
    g["c2_OfromBg_o4"] = ([Bg], (Oc[3][::7][:n], Oc[4][::7][:n], Oc[5][::7][:n]),
                          dict(order=4, nature=1, penalty=1, extrap=1))
    for order in (2, 3, 4):
        g["c2_multi_cart_o%d" % order] = ([Bg, Oc], (P[0][:n], P[1][:n], P[2][:n]),
                                          dict(order=order, nature=1, penalty=1, extrap=1, types=[0, 1]))
    return g


def zones_of(dz, kw):
    kw = dict(kw)
    types = kw.pop("types", None) or [1] * len(dz)
    return [O.RefZone(*z, interpDataType=t) for z, t in zip(dz, types)], kw


def main():
    for name, (dz, pts, kw) in golden_cases().items():
        if os.path.exists(os.path.join(HERE, name + ".npz")) and "--all" not in sys.argv:
            continue                       # committed fixtures are not rewritten
        zones, kw = zones_of(dz, kw)
        out = O.set_interp_data(pts[0], pts[1], pts[2], zones, **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack(out))
        print(name, [a.size for a in out[0]], out[5].size)


if __name__ == "__main__":
    main()
