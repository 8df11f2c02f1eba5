"""Quick GPU bring-up script (not a test): prints where the CUDA path and the oracle diverge."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
import opticalflowhs_b200 as P

def bits(a): return np.ascontiguousarray(a, np.float32).view(np.uint32)
fr = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "frames.npz")))
g1, g2 = fr["bunny_1"], fr["bunny_2"]
e = P.HSFlow(0)
e.load_pair(g1, g2)
d = e.read_derivatives(); o = O.derivatives(g1.astype(np.float32), g2.astype(np.float32))
print("deriv bit-exact:", [(bits(x) == bits(y)).all() for x, y in zip(d, o)], flush=True)
e.set_math(P.MATH_EXACT).set_params(15.0, 10, 0, True)
e.load_pair(g1, g2).compute(); u, v = e.read_uv(); uo, vo = O.run_cl(g1, g2, 15.0, 10, True)
print("exact bit-exact:", (bits(u) == bits(uo)).all(), (bits(v) == bits(vo)).all(), np.abs(u-uo).max(), flush=True)
e.set_math(P.MATH_FAST).set_kernel(1).set_params(15.0, 10, 0, True, 1)
e.load_pair(g1, g2).compute(); uf, vf = e.read_uv()
print("fast T=1 single-sweep vs oracle:", np.abs(uf-uo).max(), np.abs(vf-vo).max(), flush=True)
for T in (1, 2, 4, 8):
    e.set_kernel(2).set_params(15.0, 10, 0, True, T)
    t0 = time.time()
    e.load_pair(g1, g2).compute(); us, vs = e.read_uv()
    nb = (bits(us) != bits(uf)).sum(); 
    print(f"stream T={T}: mismatching px u {nb} v {(bits(vs)!=bits(vf)).sum()} maxdiff {np.abs(us-uf).max():.3g}  ({time.time()-t0:.2f}s)", flush=True)
    if nb:
        ys, xs = np.nonzero(bits(us) != bits(uf))
        print("   rows", ys.min(), ys.max(), "cols", xs.min(), xs.max(), "first", list(zip(ys[:8], xs[:8])))
e.close()
