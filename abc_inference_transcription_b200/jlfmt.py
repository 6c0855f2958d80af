"""Julia-compatible text formatting for the reference's tab-separated files.

`DelimitedFiles.writedlm` prints every Float64 with `print(io, x)`: the shortest digit string that
round-trips (Ryu), fixed notation for 1e-5 < |x| < 1e6 and `d.ddde±x` otherwise, always with a decimal
point; NaN, Inf, -Inf spelled like that.  Python's repr() yields the same shortest digits, so only the
layout has to be rebuilt.  Used for the S2 / E3 / A1 file layouts (abc_simulation.jl:47-61, 89-95,
compute_errors.jl:66-68, accepted_particles.jl:23-30).
"""
import math
from decimal import Decimal

import numpy as np


def jl_float(x):
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:       # shortest form has no trailing zeros
        digits.pop()
        exp += 1
    nd = len(digits)
    e10 = exp + nd - 1                                # x = d.ddd * 10^e10
    ds = "".join(map(str, digits))
    s = "-" if sign else ""
    if -5 < e10 < 6 or (e10 == -5 and False):
        if e10 >= 0:
            if nd <= e10 + 1:
                return s + ds + "0" * (e10 + 1 - nd) + ".0"
            return s + ds[:e10 + 1] + "." + ds[e10 + 1:]
        return s + "0." + "0" * (-e10 - 1) + ds
    mant = ds[0] + "." + (ds[1:] if nd > 1 else "0")
    return f"{s}{mant}e{e10}"


def writedlm_rows(fh, a):
    """writedlm(io, A): one line per row, tab separated"""
    a = np.asarray(a)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if np.issubdtype(a.dtype, np.integer):
        for row in a:
            fh.write("\t".join(str(int(v)) for v in row) + "\n")
    else:
        for row in a:
            fh.write("\t".join(jl_float(v) for v in row) + "\n")


def readdlm(path):
    """readdlm(path) for the numeric files of the pipeline -> 2-D float64 (Julia spellings of NaN/Inf accepted)"""
    rows = []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line == "":
                continue
            rows.append([float(tok.replace("Inf", "inf").replace("NaN", "nan")) for tok in line.split("\t")])
    return np.array(rows, dtype=np.float64).reshape(len(rows), -1)
