"""Counter-based synthetic config-2 curves, bit-identical on the host (numpy) and on the GPU (torch).

SURVEY 8(d) asks for bench inputs keyed by (seed, curve, point) so that the CPU arm and the GPU arm fit
THE SAME curves.  Every value is a pure function of its counter: a splitmix64-style integer hash (64-bit
wrap-around multiplies, xor-shifts), converted to a 53-bit uniform, then only IEEE +, -, *, / and floor in a
fixed order -- no libm call (sin/log differ between glibc and CUDA by an ulp), and each operation is a
separate array operation on both sides, so nothing is contracted into an FMA.  The signal is a sine-like wave
built from the classical parabolic approximation refined once (max error ~1e-3, far below the noise
sigma = 0.1, so the fits see the same problem as with a true sine); the noise is the sum of four uniforms
scaled to unit variance.

    x[b, i] = (i + 0.3 (2 u - 1)) * (1/(m - 1)),  x[b, 0] = 0, x[b, m-1] = 1   (jittered grid, sorted)
    y[b, i] = A_b * wave(f_b * x + p_b) + sigma * noise,   A in [0.5, 2), f in [1, 4), p in [0, 1)

`c2_curves(np, ...)` returns curve-major [nb][m] arrays; `c2_curves(torch, ..., device=dev)` the same values
as point-major [m][nb] tensors (the layout the fit kernels read).
"""

_M64 = (1 << 64) - 1


def _s64(v):
    """two's-complement int64 value of the unsigned 64-bit constant v"""
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_K1 = _s64(0x9E3779B97F4A7C15)
_K2 = _s64(0xBF58476D1CE4E5B9)
_K3 = _s64(0x94D049BB133111EB)
_KS = _s64(0xD6E8FEB86659FD93)   # seed multiplier
_KC = _s64(0xA0761D6478BD642F)   # curve multiplier
_KP = _s64(0xE7037ED1A0B428DB)   # point multiplier
_KT = _s64(0x8EBC6AF09C88C6E3)   # stream multiplier


def _lsr(z, k):
    return (z >> k) & ((1 << (64 - k)) - 1)


def _hash_uniform(xp, ctr):
    """ctr: int64 array -> float64 uniforms in [0, 1) (53 bits)"""
    z = ctr * _K1
    z = (z ^ _lsr(z, 30)) * _K2
    z = (z ^ _lsr(z, 27)) * _K3
    z = z ^ _lsr(z, 31)
    u = _lsr(z, 11)
    if xp.__name__ == "torch":
        return u.to(xp.float64) * (2.0 ** -53)
    return u.astype(xp.float64) * (2.0 ** -53)


def _wave(xp, v):
    """sine-like 1-periodic wave of the phase v (in turns): parabola 16 r (1/2 - r) on the first half period,
    mirrored on the second, refined once by P (0.775 + 0.225 |P|); only +, -, *, floor, abs, where."""
    r = v - xp.floor(v)
    half = r >= 0.5
    q = xp.where(half, r - 0.5, r)
    p = (q * 16.0) * (0.5 - q)
    p = p * (p * 0.225 + 0.775)
    return xp.where(half, -p, p)


def c2_curves(xp, b0, nb, m, sigma=0.1, seed=1234, device=None):
    """curves b0 .. b0+nb-1 of the config-2 workload.  numpy: (x, y) curve-major [nb][m];
    torch: point-major [m][nb] on `device`."""
    is_torch = xp.__name__ == "torch"
    if is_torch:
        b = xp.arange(b0, b0 + nb, dtype=xp.int64, device=device)[None, :]
        i = xp.arange(m, dtype=xp.int64, device=device)[:, None]
        f64 = lambda a: a.to(xp.float64)
    else:
        with xp.errstate(over="ignore"):
            return _c2_numpy(xp, b0, nb, m, sigma, seed)
    base = b * _KC + _s64(seed * _KS)
    return _c2_body(xp, base, i, f64(i), m, sigma, is_torch)


def _c2_numpy(xp, b0, nb, m, sigma, seed):
    b = xp.arange(b0, b0 + nb, dtype=xp.int64)[:, None]
    i = xp.arange(m, dtype=xp.int64)[None, :]
    base = b * xp.int64(_KC) + xp.int64(_s64(seed * _KS))
    return _c2_body(xp, base, i, i.astype(xp.float64), m, sigma, False)


def _c2_body(xp, base, i, fi, m, sigma, is_torch):
    k = (lambda v: v) if is_torch else (lambda v: xp.int64(v))
    pt = base + i * k(_KP)

    def stream(ctr, s):
        return _hash_uniform(xp, ctr + k(_s64(s * _KT)))

    u = stream(pt, 1)
    # multiply by the reciprocal instead of dividing: torch's CUDA division by a Python scalar is itself
    # implemented as a multiplication by 1/scalar, numpy's is a true division -- a product is the same everywhere
    x = (fi + (u * 2.0 - 1.0) * 0.3) * (1.0 / float(m - 1))
    first, last = (i == 0), (i == m - 1)
    x = xp.where(first, xp.zeros_like(x), x)
    x = xp.where(last, xp.ones_like(x), x)
    amp = stream(base, 101) * 1.5 + 0.5
    frq = stream(base, 102) * 3.0 + 1.0
    ph = stream(base, 103)
    y = amp * _wave(xp, frq * x + ph)
    noise = ((stream(pt, 2) + stream(pt, 3)) + (stream(pt, 4) + stream(pt, 5)) - 2.0) * 1.7320508075688772
    y = y + noise * sigma
    if is_torch:
        return x.contiguous(), y.contiguous()
    return xp.ascontiguousarray(x), xp.ascontiguousarray(y)
