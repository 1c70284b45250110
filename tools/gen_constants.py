"""Derive the double / double-double constants used by the math headers with mpmath
(50 digits) and print them as C hex-float literals.  Development tool; its output is
pasted into oracle/ossg_oracle_math.h and csrc/ossg_math.cuh."""
import mpmath as mp
mp.mp.prec = 400

def d(x):            # round to nearest double
    return float(mp.mpf(x))

def dd(x):
    hi = d(x); lo = d(mp.mpf(x) - mp.mpf(hi)); return hi, lo

def trunc_bits(x, bits):   # truncate x>0 to `bits` significant bits
    x = mp.mpf(x); e = mp.floor(mp.log(x, 2))
    s = mp.mpf(2) ** (e - bits + 1)
    return float(mp.floor(x / s) * s)

def show(name, v):
    print(f"{name:12s} = {v!r:28} /* {float.hex(v)} */")

pi = mp.pi
hi, lo = dd(pi / 180); show("DEG2RAD_HI", hi); show("DEG2RAD_LO", lo)
show("RAD2DEG", d(180 / mp.mpf(d(pi))))   # Julia: 180/π with π::Float64
show("M180_PI", -180 / float(mp.mpf(d(pi))))
show("P360_PI", 360 / float(mp.mpf(d(pi))))
show("PI_180", float(d(pi)) / 180)
hi, lo = dd(pi / 2); show("PIO2_HI", hi); show("PIO2_LO", lo)
hi, lo = dd(pi / 4); show("PIO4_HI", hi)
show("INVPIO2", d(2 / pi))
p1 = trunc_bits(pi / 2, 33); show("PIO2_1", p1)
r1 = pi / 2 - mp.mpf(p1); show("PIO2_1T", d(r1))
p2 = trunc_bits(r1, 33); show("PIO2_2", p2)
r2 = r1 - mp.mpf(p2); show("PIO2_2T", d(r2))
p3 = trunc_bits(r2, 33); show("PIO2_3", p3)
r3 = r2 - mp.mpf(p3); show("PIO2_3T", d(r3))
hi, lo = dd(mp.log(2)); show("LN2_HI", hi); show("LN2_LO", lo)
show("INVLN2", d(1 / mp.log(2)))
for k in range(2, 12):
    hi, lo = dd(1 / mp.factorial(k)); show(f"INVFACT{k}_HI", hi); show(f"INVFACT{k}_LO", lo)
for i, a in enumerate([mp.atan(mp.mpf(1) / 2), mp.atan(1), mp.atan(mp.mpf(3) / 2), pi / 2]):
    hi, lo = dd(a); show(f"ATANHI{i}", hi); show(f"ATANLO{i}", lo)
