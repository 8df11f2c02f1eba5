"""Run one stream-kernel configuration (for ncu): T, CHUNK, WPC, NP, N from the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opticalflowhs_b200 as P
T, CH, WPC = int(os.environ.get("T", 4)), int(os.environ.get("CHUNK", 0)), int(os.environ.get("WPC", 0))
NP, N = int(os.environ.get("NP", 32)), int(os.environ.get("N", 24))
W, H = int(os.environ.get("W", 3840)), int(os.environ.get("H", 2160))
e = P.HSFlow(0)
e.configure(W, H, NP).synth_frames(0, 0, 1234)
e.set_kernel(2).set_tuning(CH, WPC, 0).set_params(15.0, N, 0, True, T)
for _ in range(2):
    e.prepare(); e.iterate(N); e.sync()
ms = e.last_ms(2)
print(f"T={T} chunk={CH} wpc={WPC}: {ms:.3f} ms  {NP*W*H*N/ms/1e3:.0f} Mpx-it/s")
e.close()
