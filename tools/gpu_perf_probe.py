"""GPU perf probe (not a test): sweeps temporal block / chunk rows on synthetic frames."""
import sys, os, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opticalflowhs_b200 as P

W, H, NP = int(os.environ.get("W", 3840)), int(os.environ.get("H", 2160)), int(os.environ.get("NP", 16))
N = int(os.environ.get("N", 48))
configs = []
for T in (1, 2, 3, 4, 5, 6, 8):
    for chunk in (0, 64, 128, 270):
        configs.append((2, T, 0, chunk))
configs = [(1, 1, 0, 0), (1, 1, 0, 32), (1, 1, 0, 256)] + configs
sel = os.environ.get("SEL")
e = P.HSFlow(0)
e.configure(W, H, NP).synth_frames(0, 0, 1234)
print(f"# {W}x{H} x {NP} pairs, N={N}")
for kern, T, wpc, chunk in configs:
    if sel and f"T{T}" not in sel.split(","):
        continue
    try:
        e.set_kernel(kern).set_tuning(chunk, wpc, 0).set_params(15.0, N, 0, True, T)
        e.prepare(); e.iterate(N); e.sync()        # warm-up
        best = 1e9
        for rep in range(3):
            e.prepare(); e.iterate(N); e.sync()
            best = min(best, e.last_ms(2))
        mpx = NP * W * H * N / best / 1e3
        print(f"kern={kern} T={T} wpc={wpc} chunk={chunk:4d}  {best:8.3f} ms  {mpx:10.0f} Mpx-it/s  eff {mpx*28/1e3:8.1f} GB/s", flush=True)
    except Exception as ex:
        print(f"kern={kern} T={T} wpc={wpc} chunk={chunk}: {ex}", flush=True)
e.close()
