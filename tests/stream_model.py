"""numpy model of the FAST-math Jacobi formulation and of the row-streaming temporal block
used by opticalflowhs_b200/csrc/hs_stream.cu.  Test infrastructure: it documents the kernel's
bookkeeping (stage lags, first-row replicate, virtual bottom row, chunk warm-up, column halo)
and lets the CPU suite check that streaming == direct sweeps bit-for-bit and that the
formulation stays within the north-star tolerance of the oracle.

fma(a,b,c) is emulated through float64 (exact product, one extra rounding at worst)."""
import numpy as np

F = np.float32


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)


def normalise(Ex, Ey, Et, rho):
    """coefficients a,b,c = (Ex,Ey,Et) / sqrt(rho + Ex^2 + Ey^2)  (k_deriv epilogue)."""
    den = fma(Ey, Ey, fma(Ex, Ex, F(rho)))
    r = (1.0 / np.sqrt(den.astype(np.float64))).astype(F)
    return (Ex * r).astype(F), (Ey * r).astype(F), (Et * r).astype(F)


def hsum(c):
    """h = W + E with clamp-to-edge columns."""
    l = np.concatenate([c[..., :1], c[..., :-1]], axis=-1)
    r = np.concatenate([c[..., 1:], c[..., -1:]], axis=-1)
    return (l + r).astype(F)


def row_terms(c, stencil8):
    h = hsum(c)
    if stencil8:
        return fma(F(2), c, h), h     # G, h  (M = 2h folded into the fma below)
    return c, h


def combine(p, G, k):
    return (k * (p + G).astype(F)).astype(F)


def p_of(g_prev, h, stencil8):
    return fma(F(2), h, g_prev) if stencil8 else (g_prev + h).astype(F)


def update(ub, vb, a, b, c):
    t = fma(a, ub, fma(b, vb, c))
    return fma(-a, t, ub), fma(-b, t, vb)


def sweep_direct(u, v, a, b, c, stencil8=True):
    """One Jacobi iteration, whole frame, FAST formulation (what k_jacobi1<FAST> computes)."""
    k = F(1.0 / 12) if stencil8 else F(0.25)
    out = []
    for x in (u, v):
        G, h = row_terms(x, stencil8)
        Gup = np.concatenate([G[:1], G[:-1]], axis=0)
        Gdn = np.concatenate([G[1:], G[-1:]], axis=0)
        out.append(combine(p_of(Gup, h, stencil8), Gdn, k))
    return update(out[0], out[1], a, b, c)


def stream_block(u, v, a, b, c, T, R0, R1, stencil8=True, lim=None, emax=None):
    """T fused iterations for output rows [R0,R1) by row streaming, exactly like one warp of
    hs_stream.cu (all columns at once; the column halo is modelled separately in
    stream_block_strips).

    EPS criterion (the TRACK instantiation): every stage keeps the input it received one tick earlier -- the old value
    of the row it puts out now.  Stages S >= lim pass that through instead of iterating (tail blocks and the replay
    launch: the block then advances exactly `lim` sweeps); stages S < lim reduce max |new - old| over the owned rows
    [R0, R1) into emax[S]."""
    H = u.shape[0]
    k = F(1.0 / 12) if stencil8 else F(0.25)
    rs = max(R0 - T, 0)
    lim = T if lim is None else lim
    p = [[None, None] for _ in range(T)]
    g = [[None, None] for _ in range(T)]
    prev = [None] * T
    out_u = np.full_like(u, np.nan)
    out_v = np.full_like(v, np.nan)

    def track(S, rho, un, vn):                  # time step S+1 of row rho against its time step S
        if emax is not None and R0 <= rho < R1:
            d = max(float(np.abs(un - prev[S][0]).max()), float(np.abs(vn - prev[S][1]).max()))
            emax[S] = max(emax[S], d)

    def recv(S, rho, cu, cv):
        if S == T:
            if R0 <= rho < R1:
                out_u[rho], out_v[rho] = cu, cv
            return
        if S >= lim:                            # pass-through stage: out = previous input (row rho - 1, unchanged)
            old, prev[S] = prev[S], (cu, cv)
            if rho != rs:
                recv(S + 1, rho - 1, old[0], old[1])
            return
        bars = []
        for f, cx in enumerate((cu, cv)):
            G, h = row_terms(cx, stencil8)
            if rho == rs:                       # first row this stage sees: replicate upwards
                p[S][f] = p_of(G, h, stencil8)
                g[S][f] = G
                bars.append(None)
            else:
                bars.append(combine(p[S][f], G, k))
                p[S][f] = p_of(g[S][f], h, stencil8)
                g[S][f] = G
        if bars[0] is None:
            prev[S] = (cu, cv)
            return
        un, vn = update(bars[0], bars[1], a[rho - 1], b[rho - 1], c[rho - 1])
        track(S, rho - 1, un, vn)
        prev[S] = (cu, cv)
        recv(S + 1, rho - 1, un, vn)

    def virt(S):                                # row H == row H-1 (clamp at the bottom edge)
        if S >= lim:
            recv(S + 1, H - 1, prev[S][0], prev[S][1])
            return
        bars = [combine(p[S][f], g[S][f], k) for f in range(2)]
        un, vn = update(bars[0], bars[1], a[H - 1], b[H - 1], c[H - 1])
        track(S, H - 1, un, vn)
        recv(S + 1, H - 1, un, vn)

    for r in range(rs, R1 - 1 + T + 1):
        if r <= H - 1:
            recv(0, r, u[r], v[r])
        elif r - H < T:
            virt(r - H)
    return out_u, out_v


def stream_block_strips(u, v, a, b, c, T, chunk, valid_w, halo, stencil8=True):
    """Whole-frame T-block out of independent (column strip x row chunk) units."""
    H, W = u.shape
    ou, ov = np.empty_like(u), np.empty_like(v)
    for x0 in range(0, W, valid_w):
        lo, hi = max(x0 - halo, 0), min(x0 + valid_w + halo, W)
        sl = slice(lo, hi)
        for R0 in range(0, H, chunk):
            R1 = min(R0 + chunk, H)
            su, sv = stream_block(u[:, sl], v[:, sl], a[:, sl], b[:, sl], c[:, sl], T, R0, R1, stencil8)
            w = min(valid_w, W - x0)
            ou[R0:R1, x0:x0 + w] = su[R0:R1, x0 - lo:x0 - lo + w]
            ov[R0:R1, x0:x0 + w] = sv[R0:R1, x0 - lo:x0 - lo + w]
    return ou, ov
