#!/usr/bin/env python
"""Generates tests/golden/jerkcar.npz from the reference's own fixtures.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

Inputs  (reference repo): examples/jerkcar/{uvec,yacchist,yposhist}.csv  -- the driver's inputs
Outputs (reference repo): examples/jerkcar/{vanilla,information,sqrt}.csv -- what the reference's
        Vanilla / Information / SquareRoot filters printed (examples/jerkcar/main.go:133-161),
        12 columns = (value, +2 sigma, -2 sigma) x (position, velocity, acceleration, bias), "%f".
The arrays are stored verbatim (parsed floats); nothing is computed here.
"""
import os

import numpy as np

REF = "/root/reference/examples/jerkcar"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jerkcar.npz")


def table(name):
    rows = []
    with open(os.path.join(REF, name)) as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith("#") or line.startswith("position"):
                continue
            rows.append([float(v) for v in line.split(",")])
    return np.array(rows)


def main():
    u = np.array([float(l.split(",")[0]) for l in open(os.path.join(REF, "uvec.csv")) if l.strip()])
    yacc = np.array([float(v) for v in open(os.path.join(REF, "yacchist.csv")).readline().split(",")])
    ypos = np.array([float(v) for v in open(os.path.join(REF, "yposhist.csv")).readline().split(",")])
    out = {"uvec": u, "yacc": yacc, "ypos": ypos}
    for name in ("vanilla", "information", "sqrt"):
        out[name] = table(name + ".csv")
        print(name, out[name].shape)
    print("u", u.shape, "yacc", yacc.shape, "ypos", ypos.shape, "nan in ypos:", int(np.isnan(ypos).sum()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
