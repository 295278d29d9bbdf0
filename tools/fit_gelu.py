import numpy as np
from scipy.special import erf
from scipy.optimize import least_squares
R2 = 36.0
x = np.linspace(-6, 6, 24001)
gel = 0.5*x*(1+erf(x/np.sqrt(2)))
def model(c, x, dt=np.float64):
    x = x.astype(dt)
    t = np.minimum(x*x, dt(R2))
    p = np.full_like(x, c[-1], dtype=dt)
    for ci in c[-2::-1]:
        p = (p*t + dt(ci)).astype(dt)
    w = (x*p).astype(dt)           # = -2*log2e*u
    e = np.exp2(w).astype(dt)
    return (x/(dt(1)+e)).astype(dt)
for deg in (3,4,5):
    c0 = np.concatenate([[0.7978845608, 0.0356774], np.zeros(deg-1)]) * (-2*np.log2(np.e))
    w = np.ones_like(x); best=None
    for it in range(80):
        r = least_squares(lambda c: w*(model(c,x)-gel), c0, method='lm', xtol=1e-15, ftol=1e-15)
        c0 = r.x
        err = np.abs(model(c0,x)-gel)
        if best is None or err.max()<best[0]: best=(err.max(), c0.copy())
        w = w*(1+ 2*err/err.max()); w/=w.mean()
    c = best[1]
    print(deg+1, "coeffs; max abs err", best[0]); print("   ", ", ".join(f"{v:.10e}f" for v in c))
    # all fp16 values, float32 arithmetic
    h = np.arange(0, 0x7c00, dtype=np.uint16).view(np.float16).astype(np.float64)
    h = np.concatenate([h, -h])
    g = 0.5*h*(1+erf(h/np.sqrt(2)))
    m32 = model(c.astype(np.float32), h, np.float32).astype(np.float64)
    e = np.abs(m32-g)
    ulp = np.maximum(np.abs(g), 2.0**-14) * 2.0**-11
    print("    fp16 grid: max abs", e.max(), "at", h[e.argmax()], " max err in fp16 half-ulps", (e/ulp).max(), "at", h[(e/ulp).argmax()])
    # monotonic check of p(t)>0
    t = np.linspace(0,36,1000); p = np.polyval(c[::-1], t); print("    p(t) range", p.min(), p.max())
