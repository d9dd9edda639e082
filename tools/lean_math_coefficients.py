# derive the constants of the lean FP64 sincos / atan2 (mpmath, 200 digits)
import mpmath as mp, struct
mp.mp.dps = 120
def d(x):  # nearest double
    return float(mp.mpf(x))
def hexd(x):
    return struct.pack('>d', x).hex()
def trunc_bits(x, bits):
    m, e = mp.frexp(mp.mpf(x))
    q = mp.floor(m * 2**bits) / 2**bits
    return float(mp.ldexp(q, e))
pio2 = mp.pi/2
hi = d(pio2)
mid = trunc_bits(pio2 - hi, 42)
lo = d(pio2 - hi - mid)
print("pio2", repr(hi), hexd(hi), repr(mid), hexd(mid), repr(lo), hexd(lo))
print("2/pi", repr(d(2/mp.pi)), hexd(d(2/mp.pi)))

def cheb_fit(f, a, b, n):
    # interpolate at Chebyshev nodes, return monomial coefficients in z (high precision)
    xs = [ (a+b)/2 + (b-a)/2*mp.cos(mp.pi*(2*k+1)/(2*n)) for k in range(n)]
    A = mp.matrix(n, n); y = mp.matrix(n, 1)
    for i, x in enumerate(xs):
        for j in range(n): A[i, j] = x**j
        y[i] = f(x)
    c = mp.lu_solve(A, y)
    return [c[i] for i in range(n)]
def maxerr(f, c, a, b, rel=None, m=4001):
    worst = 0
    for k in range(m):
        x = a + (b-a)*mp.mpf(k)/(m-1)
        p = sum(ck*x**j for j, ck in enumerate(c))
        e = abs(p - f(x))
        worst = max(worst, e)
    return worst

# atan(t)/t as a function of z = t^2 on [0, tan(pi/8)^2]
L = mp.tan(mp.pi/8)**2
def fa(z):
    if z == 0: return mp.mpf(1)
    t = mp.sqrt(z); return mp.atan(t)/t
for n in (12, 13, 14):
    c = cheb_fit(fa, mp.mpf(0), L*mp.mpf('1.0001'), n)
    cd = [d(x) for x in c]
    print("atan n", n, "err(exact coef)", mp.nstr(maxerr(fa, c, 0, L), 5), "err(double coef)", mp.nstr(maxerr(fa, [mp.mpf(x) for x in cd], 0, L), 5))
    if n == 13: atan_c = cd
print("atan coefficients (z^0..):")
for x in atan_c: print("   %s, // %s" % (repr(x), hexd(x)))

# sin(r)/r and cos(r) in z = r^2 on [0, (pi/4)^2]
Z = (mp.pi/4)**2 * mp.mpf('1.001')
def fs(z):
    if z == 0: return mp.mpf(1)
    r = mp.sqrt(z); return mp.sin(r)/r
def fc(z):
    return mp.cos(mp.sqrt(z))
for n in (7, 8):
    c = cheb_fit(fs, mp.mpf(0), Z, n); cd=[d(x) for x in c]
    print("sin n", n, mp.nstr(maxerr(fs, c, 0, Z),5), mp.nstr(maxerr(fs,[mp.mpf(x) for x in cd],0,Z),5))
    if n == 7: sin_c = cd
    c = cheb_fit(fc, mp.mpf(0), Z, n+1); cd=[d(x) for x in c]
    print("cos n", n+1, mp.nstr(maxerr(fc, c, 0, Z),5), mp.nstr(maxerr(fc,[mp.mpf(x) for x in cd],0,Z),5))
    if n == 7: cos_c = cd
print("sin:"); [print("   %s, // %s" % (repr(x), hexd(x))) for x in sin_c]
print("cos:"); [print("   %s, // %s" % (repr(x), hexd(x))) for x in cos_c]
