#!/usr/bin/env python
"""tools/derive_corner_series.py -- derives the series that ecg_moment_interior_kernel (csrc/ecg.cu, corner_series2)
evaluates instead of the 8-corner stencil sum of an interior voxel.

    g(q) = sum_{d in {+-1}^3} d.(q + d) / |q + d|^3,          q = lead - voxel (bordered voxel units)

With f = 1/|q|:  d.(q+d)/|q+d|^3 = -(d.grad f)(q+d), so  g = -d/ds [ sum_d f(q + s d) ]_{s=1}  and
sum_d f(q + s d) = 8 cosh(s dx) cosh(s dy) cosh(s dz) f = 8 sum_n s^(2n) T_2n f,
T_2n = sum_{i+j+k=n} dx^(2i) dy^(2j) dz^(2k) / ((2i)! (2j)! (2k)!).   f is harmonic: T_2 f = 0.
T_2n f is a cubic harmonic: a polynomial in e2 = xi_z xi_y + xi_y xi_x + xi_x xi_z and e3 = xi_z xi_y xi_x
(xi_i = q_i^2 / |q|^2, so xi_z + xi_y + xi_x = 1) times |q|^-(2n+1).  Prints the terms  -16 n T_2n f  of g.
Needs sympy; takes about a minute.  Usage: python tools/derive_corner_series.py [max_n=4]"""
import sys

import sympy as sp
from sympy import factorial as fac
from sympy.polys.polyfuncs import symmetrize


def main():
    max_n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    x, y, z = sp.symbols("x y z", real=True)
    X, Y, Z = sp.symbols("X Y Z")
    e2, e3 = sp.symbols("e2 e3")
    r = sp.sqrt(x * x + y * y + z * z)
    f = 1 / r
    for n in range(2, max_n + 1):
        t = 0
        for i in range(n + 1):
            for j in range(n + 1 - i):
                k = n - i - j
                t += sp.diff(f, x, 2 * i, y, 2 * j, z, 2 * k) / (fac(2 * i) * fac(2 * j) * fac(2 * k))
        num = sp.Poly(sp.expand(sp.simplify(t * r ** (4 * n + 1))), x, y, z)   # homogeneous of degree 2n, even in x, y, z
        expr = 0
        for (a, b, c), co in num.terms():
            assert a % 2 == 0 and b % 2 == 0 and c % 2 == 0
            expr += co * X ** (a // 2) * Y ** (b // 2) * Z ** (c // 2)
        sym, rem, defs = symmetrize(sp.expand(expr), [X, Y, Z], formal=True)
        assert rem == 0
        s1, s2, s3 = (d[0] for d in defs)
        term = sp.expand(-16 * n * sym.subs(s1, 1).subs({s2: e2, s3: e3}))
        print("|q|^-%d * ( %s )" % (2 * n + 1, term))


if __name__ == "__main__":
    main()
