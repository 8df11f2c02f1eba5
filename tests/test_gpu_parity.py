"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libhsflow.so (ctypes binding opticalflowhs_b200.HSFlow) and is checked against the CPU oracle.

Bars (BASELINE.json north_star):
  * integer / exactly representable work (gray conversion, derivatives, synthetic frames, dot
    masks): bit-exact;
  * EXACT math iteration: bit-exact with the oracle (Kernels.cl op order, no contraction);
  * FAST math iteration (the throughput path): max |du|,|dv| <= 1e-3 px and mean end-point-error
    difference <= 1e-4 px after N iterations;
  * temporally blocked kernel vs single-sweep kernel, batched vs single, strips vs whole: bit-identical.
"""
import numpy as np
import pytest

from conftest import iou

pytestmark = pytest.mark.gpu

TOL_MAX = 1e-3      # px, north_star
TOL_EPE = 1e-4      # px, north_star


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def P():
    import opticalflowhs_b200 as pkg
    return pkg


@pytest.fixture()
def eng(P):
    e = P.HSFlow(0)
    yield e
    e.close()


def epe_diff(u, v, uo, vo):
    return abs(float(np.mean(np.hypot(u, v))) - float(np.mean(np.hypot(uo, vo))))


def rand_frames(rng, h, w):
    return rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (h, w), dtype=np.uint8)


ODD_SHAPES = [(1, 1), (1, 5), (7, 1), (5, 7), (17, 33), (9, 130), (66, 257), (40, 116), (23, 240), (31, 124)]

# ---- derivatives: bit-exact -------------------------------------------------------------------

def test_derivatives_gray_bit_exact(eng, oracle, frames):
    for name in ("city", "bunny"):
        g1, g2 = frames[f"{name}_1"], frames[f"{name}_2"]
        eng.load_pair(g1, g2)
        d = eng.read_derivatives()
        o = oracle.derivatives(g1.astype(np.float32), g2.astype(np.float32))
        for x, y in zip(d, o):
            assert (bits(x) == bits(y)).all()


def test_derivatives_bgr_fused_gray_bit_exact(eng, oracle, frames):
    b1, b2 = frames["bunny_1_bgr"], frames["bunny_2_bgr"]
    eng.load_pair(b1, b2)
    d = eng.read_derivatives()
    o = oracle.derivatives(oracle.bgr2gray(b1).astype(np.float32), oracle.bgr2gray(b2).astype(np.float32))
    for x, y in zip(d, o):
        assert (bits(x) == bits(y)).all()
    rng = np.random.default_rng(3)
    for (h, w) in [(3, 5), (9, 85), (4, 128), (6, 43)]:
        b1 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        b2 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        eng.load_pair(b1, b2)
        d = eng.read_derivatives()
        o = oracle.derivatives(oracle.bgr2gray(b1).astype(np.float32), oracle.bgr2gray(b2).astype(np.float32))
        for x, y in zip(d, o):
            assert (bits(x) == bits(y)).all(), (h, w)


def test_derivatives_odd_shapes_and_f32(eng, oracle):
    rng = np.random.default_rng(5)
    for (h, w) in ODD_SHAPES:
        g1, g2 = rand_frames(rng, h, w)
        eng.load_pair(g1, g2)
        for x, y in zip(eng.read_derivatives(), oracle.derivatives(g1.astype(np.float32), g2.astype(np.float32))):
            assert (bits(x) == bits(y)).all(), (h, w)
        f1 = rng.standard_normal((h, w)).astype(np.float32) * 50
        f2 = rng.standard_normal((h, w)).astype(np.float32) * 50
        eng.load_pair(f1, f2)
        for x, y in zip(eng.read_derivatives(), oracle.derivatives(f1, f2)):
            assert (bits(x) == bits(y)).all(), (h, w)


# ---- EXACT math: bit-exact with the oracle ------------------------------------------------------

@pytest.mark.parametrize("update_v", [False, True])
def test_exact_city_100_iterations_bitwise(eng, P, oracle, frames, update_v):
    g1, g2 = frames["city_1"], frames["city_2"]
    eng.set_math(P.MATH_EXACT).set_params(15.0, 100, P.STENCIL_CL8, update_v)
    eng.load_pair(g1, g2).compute()
    u, v = eng.read_uv()
    uo, vo = oracle.run_cl(g1, g2, 15.0, 100, update_v)
    assert (bits(u) == bits(uo)).all() and (bits(v) == bits(vo)).all()


def test_exact_odd_shapes_and_warm_start_bitwise(eng, P, oracle):
    rng = np.random.default_rng(11)
    eng.set_math(P.MATH_EXACT)
    for (h, w) in ODD_SHAPES:
        g1, g2 = rand_frames(rng, h, w)
        d = oracle.derivatives(g1.astype(np.float32), g2.astype(np.float32))
        u0 = rng.standard_normal((h, w)).astype(np.float32)
        v0 = rng.standard_normal((h, w)).astype(np.float32)
        for upd in (True, False):
            eng.set_params(3.0, 9, P.STENCIL_CL8, upd).set_warm_start(True)
            eng.load_pair(g1, g2).write_uv(u0, v0)
            eng.compute()
            u, v = eng.read_uv()
            uo, vo = oracle.jacobi(*d, 3.0, 9, upd, u0, v0)
            assert (bits(u) == bits(uo)).all() and (bits(v) == bits(vo)).all(), (h, w, upd)
    eng.set_warm_start(False)


def test_exact_cv4_stencil_bitwise(eng, P, oracle, frames):
    g1, g2 = frames["bunny_1"], frames["bunny_2"]
    d = oracle.derivatives(g1.astype(np.float32), g2.astype(np.float32))
    eng.set_math(P.MATH_EXACT).set_params(15.0, 20, P.STENCIL_CV4, True).set_lambda(0.1)
    eng.load_pair(g1, g2).compute()
    u, v = eng.read_uv()
    uo, vo = oracle.jacobi_general(*d, np.float32(0.25), np.float32(0.0), np.float32(1.0) / np.float32(0.1), 20, True)
    assert (bits(u) == bits(uo)).all() and (bits(v) == bits(vo)).all()


# ---- golden masks through the GPU path ------------------------------------------------------------

@pytest.mark.parametrize("name,n,dots", [("city", 10, 677), ("bunny", 10, 1373), ("bunny", 2, 924)])
@pytest.mark.parametrize("math", ["exact", "fast"])
def test_golden_masks_literal_mode(eng, P, frames, masks, name, n, dots, math):
    eng.set_math(P.MATH_EXACT if math == "exact" else P.MATH_FAST).set_params(15.0, n, P.STENCIL_CL8, False)
    eng.load_pair(frames[f"{name}_1"], frames[f"{name}_2"]).compute()
    m, cnt = eng.dot_mask(0, 4, 0.5)
    g = masks[f"{name}_cl_a15_n{n}"]
    assert cnt == dots and iou(m, g) == 1.0
    u, v = eng.read_uv()
    assert not v.any()


# ---- the shipped pictures through the GPU path, pixel for pixel ---------------------------------------------------

@pytest.mark.parametrize("key,name,n", [("city_cl_a15_n10", "city", 10), ("bunny_cl_a15_n10", "bunny", 10), ("bunny_cl_a15_n2", "bunny", 2)])
def test_shipped_cl_pictures_from_gpu_fields(eng, P, oracle, frames, pictures, key, name, n):
    """The fields the GPU computes in LITERAL mode, drawn like cpp:758-770 and saved like cvSaveImage, ARE the shipped
    *_cl_out.jpg: every pixel in EXACT math; FAST math (streaming kernel) may move an end point that sits on an integer."""
    pytest.importorskip("cv2")
    g1, g2 = frames[f"{name}_1"], frames[f"{name}_2"]
    eng.set_math(P.MATH_EXACT).set_params(15.0, n, P.STENCIL_CL8, False)
    eng.load_pair(g1, g2).compute()
    u, v = eng.read_uv()
    assert (oracle.jpeg_roundtrip(oracle.render_flow(u, v, 0.5, 1.0)) == pictures[key]).all()
    eng.set_math(P.MATH_FAST).set_params(15.0, n, P.STENCIL_CL8, False, 0)
    eng.load_pair(g1, g2).compute()
    uf, vf = eng.read_uv()
    assert (oracle.render_flow(uf, vf, 0.5, 1.0) != oracle.render_flow(u, v, 0.5, 1.0)).any(axis=2).sum() <= 60


@pytest.mark.parametrize("name", ["city", "bunny"])
def test_shipped_cv_pictures_from_gpu_fields(eng, P, oracle, frames, pictures, name):
    """OpenCV-mode path (blur, Sobel estimator, 4-neighbour stencil, lambda = 0.1, 10 iterations, eps 1e-6): the restated
    cvCalcOpticalFlowHS is pinned pixel-exactly by *_cv_out.jpg.  EXACT math follows its rounding sequence: bit-identical
    fields, so the GPU's picture is the shipped one pixel for pixel.  FAST math is within 1e-3 px, so its picture may
    differ in a few line end points out of the ~1 000 - 2 000 drawn."""
    pytest.importorskip("cv2")
    g1, g2 = frames[f"{name}_1"], frames[f"{name}_2"]
    uo, vo, _ = oracle.run_cv(g1, g2, 0.1, 10, eps=1e-6)
    ref = oracle.render_flow(uo, vo, 1.0, 0.5)
    assert (oracle.jpeg_roundtrip(ref) == pictures[f"{name}_cv_l0.1_n10"]).all()
    for math in (P.MATH_EXACT, P.MATH_FAST):
        eng.set_math(math).set_deriv(P.DERIV_CV).set_params(0.0, 10, P.STENCIL_CV4, True, 0).set_lambda(0.1).set_epsilon(1e-6)
        eng.load_pair(g1, g2).compute()
        u, v = eng.read_uv()
        assert np.abs(u - uo).max() <= TOL_MAX and np.abs(v - vo).max() <= TOL_MAX
        if math == P.MATH_EXACT:                    # the rounding sequence of cvCalcOpticalFlowHS: bit-identical, so the picture is too
            assert (bits(u) == bits(uo)).all() and (bits(v) == bits(vo)).all()
            assert (oracle.jpeg_roundtrip(oracle.render_flow(u, v, 1.0, 0.5)) == pictures[f"{name}_cv_l0.1_n10"]).all()
        else:
            differing = (oracle.render_flow(u, v, 1.0, 0.5) != ref).any(axis=2).sum()
            assert differing <= 120, (math, differing)
    eng.set_deriv(P.DERIV_CL).set_epsilon(0.0)


# ---- FAST math: north-star tolerance ------------------------------------------------------------

@pytest.mark.parametrize("name", ["city", "bunny"])
@pytest.mark.parametrize("T", [1, 2, 4, 8])
def test_fast_within_tolerance_of_oracle(eng, P, oracle, frames, name, T):
    g1, g2 = frames[f"{name}_1"], frames[f"{name}_2"]
    uo, vo = oracle.run_cl(g1, g2, 15.0, 100, True)
    eng.set_math(P.MATH_FAST).set_params(15.0, 100, P.STENCIL_CL8, True, T)
    eng.load_pair(g1, g2).compute()
    u, v = eng.read_uv()
    assert np.abs(u - uo).max() <= TOL_MAX and np.abs(v - vo).max() <= TOL_MAX
    assert epe_diff(u, v, uo, vo) <= TOL_EPE
    # the formulation is in fact far tighter than the contract
    assert np.abs(u - uo).max() <= 2e-5 and np.abs(v - vo).max() <= 2e-5


def test_fast_fields_match_reference_generated_fixture(eng, P, fields, frames):
    eng.set_math(P.MATH_FAST).set_params(15.0, 100, P.STENCIL_CL8, True, 4)
    eng.load_pair(frames["city_1"], frames["city_2"]).compute()
    u, v = eng.read_uv()
    assert np.abs(u[::4, ::4] - fields["city_n100_full_u"]).max() <= TOL_MAX
    assert np.abs(v[::4, ::4] - fields["city_n100_full_v"]).max() <= TOL_MAX


# ---- temporally blocked kernel == single-sweep kernel, bit for bit ----------------------------------

def run_fast(eng, P, g1, g2, n, T, kernel, stencil=0, chunk=0, wpc=0, alpha=15.0):
    eng.set_math(P.MATH_FAST).set_kernel(kernel).set_tuning(chunk, wpc, 0)
    eng.set_params(alpha, n, stencil, True, T)
    eng.load_pair(g1, g2).compute()
    out = eng.read_uv()
    eng.set_kernel(0).set_tuning(0, 0, 0)
    return out


@pytest.mark.parametrize("T", [1, 2, 3, 4, 5, 6, 7, 8])
def test_stream_kernel_bit_identical_to_single_sweep(eng, P, T):
    rng = np.random.default_rng(100 + T)
    shapes = [(1, 1), (3, 9), (2, 130), (5, 116), (19, 113), (40, 124), (37, 250), (70, 131), (9, 500)]
    for (h, w) in shapes:
        g1, g2 = rand_frames(rng, h, w)
        n = 2 * T + (1 if T > 1 else 0)          # full blocks plus a remainder block
        ref = run_fast(eng, P, g1, g2, n, 1, 1)
        for chunk, wpc in ((0, 0), (7, 3), (1, 1)):
            out = run_fast(eng, P, g1, g2, n, T, 2, chunk=chunk, wpc=wpc)
            assert (bits(out[0]) == bits(ref[0])).all() and (bits(out[1]) == bits(ref[1])).all(), (h, w, T, chunk, wpc)


@pytest.mark.parametrize("T", [1, 4, 6])
def test_literal_mode_on_the_stream_kernel_bit_identical_to_single_sweep(eng, P, oracle, T):
    """update_v = 0 (the shipped kernel never writes v, Kernels.cl:87-89) runs on the temporally blocked kernel with a
    zeroed b plane while v == 0; it must equal the single-sweep LITERAL kernel bit for bit and stay within the
    tolerance of the oracle; a v field written by the caller sends it back to the single-sweep kernel."""
    rng = np.random.default_rng(300 + T)
    for (h, w) in [(1, 1), (5, 116), (40, 124), (37, 250), (70, 131)]:
        g1, g2 = rand_frames(rng, h, w)
        n = 2 * T + 1
        eng.set_math(P.MATH_FAST).set_kernel(1).set_params(15.0, n, P.STENCIL_CL8, False, 1)
        eng.load_pair(g1, g2).compute()
        ref = eng.read_uv()
        eng.set_kernel(0).set_params(15.0, n, P.STENCIL_CL8, False, T)
        assert eng.temporal_block == T
        eng.load_pair(g1, g2).compute()
        out = eng.read_uv()
        assert (bits(out[0]) == bits(ref[0])).all() and not out[1].any(), (h, w, T)
        uo, vo = oracle.run_cl(g1, g2, 15.0, n, False)
        assert np.abs(out[0] - uo).max() <= TOL_MAX and not vo.any()
        ex, ey, et = eng.read_derivatives()             # still the raw derivatives, Ey included
        assert np.abs(ey).max() > 0 or h == 1
    # a caller-supplied v field: v is no longer zero, the iteration has to keep it constant (single-sweep kernel)
    g1, g2 = rand_frames(rng, 40, 124)
    v0 = rng.standard_normal((40, 124)).astype(np.float32)
    d = oracle.derivatives(g1.astype(np.float32), g2.astype(np.float32))
    uo, vo = oracle.jacobi(*d, 15.0, 9, update_v=False, u0=np.zeros_like(v0), v0=v0)
    eng.set_kernel(0).set_params(15.0, 9, P.STENCIL_CL8, False, T)
    eng.load_pair(g1, g2).prepare()
    eng.write_uv(np.zeros_like(v0), v0)
    assert eng.temporal_block == 1
    eng.iterate(9)
    u, v = eng.read_uv()
    assert (bits(v) == bits(v0)).all() and np.abs(u - uo).max() <= TOL_MAX
    eng.set_kernel(0)


@pytest.mark.parametrize("T", [2, 4, 7])
def test_stream_kernel_cv4_stencil_bit_identical(eng, P, T):
    rng = np.random.default_rng(7 + T)
    for (h, w) in [(11, 37), (33, 260)]:
        g1, g2 = rand_frames(rng, h, w)
        ref = run_fast(eng, P, g1, g2, 3 * T, 1, 1, stencil=1)
        out = run_fast(eng, P, g1, g2, 3 * T, T, 2, stencil=1, chunk=5)
        assert (bits(out[0]) == bits(ref[0])).all() and (bits(out[1]) == bits(ref[1])).all()


def test_stream_kernel_real_frames_bit_identical(eng, P, frames):
    g1, g2 = frames["city_1"], frames["city_2"]
    ref = run_fast(eng, P, g1, g2, 100, 1, 1)
    for T in (4, 6, 8):
        out = run_fast(eng, P, g1, g2, 100, T, 0)
        assert (bits(out[0]) == bits(ref[0])).all() and (bits(out[1]) == bits(ref[1])).all(), T


# ---- batches ----------------------------------------------------------------------------------------

def test_batch_with_sub_batches_equals_single_pairs(P, oracle):
    W, H, n = 200, 96, 7
    single = []
    with P.HSFlow(0) as e:
        e.set_params(15.0, 21, P.STENCIL_CL8, True, 4)
        for k in range(n):
            f1, f2 = oracle.synth_pair(W, H, seed=1234 + k)
            e.load_pair(f1, f2).compute()
            single.append(e.read_uv())
    for T, kern in ((4, 0), (1, 1)):
        with P.HSFlow(0) as e:
            e.set_tuning(0, 0, 3)                     # sub-batches of 3 -> 3 + 3 + 1
            e.set_params(15.0, 21, P.STENCIL_CL8, True, T).set_kernel(kern)
            e.configure(W, H, n)
            for k in range(n):
                e.set_frames(*oracle.synth_pair(W, H, seed=1234 + k), pair=k)
            e.compute()
            for k in range(n):
                u, v = e.read_uv(k)
                assert (bits(u) == bits(single[k][0])).all() and (bits(v) == bits(single[k][1])).all(), (T, k)


def test_device_synthetic_frames_equal_oracle_generator(P, oracle):
    W, H, n = 333, 77, 3
    with P.HSFlow(0) as e:
        e.configure(W, H, n).synth_frames(0, 0, 1234)
        e.set_params(15.0, 0)
        for k in range(n):
            f1, f2 = oracle.synth_pair(W, H, seed=1234 + k)
            o = oracle.derivatives(f1.astype(np.float32), f2.astype(np.float32))
            for x, y in zip(e.read_derivatives(k), o):
                assert (bits(x) == bits(y)).all()
        # a strip of a taller frame
        e.configure(W, 20, 1).synth_frames(200, 150, 99)
        f1, f2 = oracle.synth_pair(W, 200, seed=99, row0=150, rows=20)
        for x, y in zip(e.read_derivatives(0), oracle.derivatives(f1.astype(np.float32), f2.astype(np.float32))):
            assert (bits(x) == bits(y)).all()


@pytest.mark.parametrize("W,H,n,math_exact,sub", [(256, 120, 11, False, 0), (232, 64, 41, False, 4), (250, 60, 41, False, 4),
                                                  (128, 48, 37, True, 4), (232, 64, 70, False, 16)])
def test_pipelined_host_batch_equals_per_pair_compute(P, oracle, W, H, n, math_exact, sub):
    """hsflow_run_batch_host against loading every pair separately: one sub-batch; ramped sub-batch sizes with planar
    staging (W % 4 == 0, streaming kernel); the 2-D copy fallback for W % 4 != 0 and for the single-sweep kernel."""
    frames = np.empty((n, 2, H, W), np.uint8)
    for k in range(n):
        frames[k, 0], frames[k, 1] = oracle.synth_pair(W, H, seed=50 + k)
    u = np.empty((n, H, W), np.float32)
    v = np.empty((n, H, W), np.float32)
    math = P.MATH_EXACT if math_exact else P.MATH_FAST
    with P.HSFlow(0) as e:
        e.set_math(math).set_params(15.0, 12, P.STENCIL_CL8, True, 4).set_tuning(sub_batch=sub)
        e.run_batch_host(frames, u, v)
        u[:] = 0; v[:] = 0
        e.run_batch_host(frames, u, v)                  # slots and staging are reused
        with pytest.raises(P.HSFlowError):              # the call leaves no current field on the device
            e.read_uv()
        with pytest.raises(ValueError):                 # raw pointers go to C: dtype and layout are checked
            e.run_batch_host(frames, u.astype(np.float64), v)
        with pytest.raises(ValueError):
            e.run_batch_host(frames[:, :, :, ::2], u, v)
    with P.HSFlow(0) as e:
        e.set_math(math).set_params(15.0, 12, P.STENCIL_CL8, True, 4)
        for k in range(n):
            e.load_pair(frames[k, 0], frames[k, 1]).compute()
            us, vs = e.read_uv()
            assert (bits(u[k]) == bits(us)).all() and (bits(v[k]) == bits(vs)).all(), k


@pytest.mark.parametrize("sequence", [False, True])
def test_pipelined_bgr_frames_and_sampled_fields(P, oracle, sequence):
    """hsflow_run_pipeline_host with interleaved BGR frames (gray conversion of cpp:727-728 fused into the derivative
    kernel) and with the consumer-shaped read-back (stride-4 samples, cpp:762-767), against per-pair computes on the
    gray frames the oracle's fixed-point BGR2GRAY produces; hsflow_sample_uv against read_uv."""
    W, H, n, N = 236, 70, 23, 14
    rng = np.random.default_rng(8)
    nf = n + 1 if sequence else 2 * n
    bgr = rng.integers(0, 256, (nf, H, W, 3), dtype=np.uint8)
    gray = np.stack([oracle.bgr2gray(f) for f in bgr])
    pair = (lambda k: (k, k + 1)) if sequence else (lambda k: (2 * k, 2 * k + 1))
    want = []
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, 4)
        for k in range(n):
            a, b = pair(k)
            e.load_pair(gray[a], gray[b]).compute()
            want.append(e.read_uv())
            us, vs = e.sample_uv(0, 4)
            assert (bits(us) == bits(want[-1][0][::4, ::4])).all() and (bits(vs) == bits(want[-1][1][::4, ::4])).all()
        us, vs = e.sample_uv(0, 5)
        assert us.shape == (14, 48) and (bits(us) == bits(want[-1][0][::5, ::5])).all()
    frames = bgr if sequence else bgr.reshape(n, 2, H, W, 3)
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, 4).set_tuning(sub_batch=4)
        u, v = np.empty((n, H, W), np.float32), np.empty((n, H, W), np.float32)
        e.run_pipeline_host(frames, u, v, sequence=sequence)
        for k in range(n):
            assert (bits(u[k]) == bits(want[k][0])).all() and (bits(v[k]) == bits(want[k][1])).all(), k
        for step in (4, 3):
            gh, gw = -(-H // step), -(-W // step)
            us, vs = np.empty((n, gh, gw), np.float32), np.empty((n, gh, gw), np.float32)
            e.run_pipeline_host(frames, us, vs, sequence=sequence, sample_step=step)
            for k in range(n):
                assert (bits(us[k]) == bits(want[k][0][::step, ::step])).all(), (step, k)
                assert (bits(vs[k]) == bits(want[k][1][::step, ::step])).all(), (step, k)


@pytest.mark.parametrize("sequence", [False, True])
def test_pipeline_over_several_handles_equals_one_handle(P, oracle, sequence):
    """hsflow_run_pipeline_host_multi: pair sharding inside one process (contiguous blocks of pairs, one host thread per
    handle -- here three handles, spread over the GPUs that exist) against the single-handle call; ragged last block."""
    import torch
    W, H, n, N = 232, 64, 14, 11
    nf = n + 1 if sequence else 2 * n
    frames = np.stack([oracle.synth_pair(W, H, seed=300 + k)[k & 1] for k in range(nf)])
    if not sequence:
        frames = frames.reshape(n, 2, H, W)
    u1, v1 = np.empty((n, H, W), np.float32), np.empty((n, H, W), np.float32)
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, 4)
        e.run_pipeline_host(frames, u1, v1, sequence=sequence)
    ndev = torch.cuda.device_count()
    engs = [P.HSFlow(k % ndev) for k in range(3)]
    try:
        for e in engs:
            e.set_params(15.0, N, P.STENCIL_CL8, True, 4).set_tuning(sub_batch=2)
        for step in (0, 4):
            shape = (n, -(-H // step), -(-W // step)) if step else (n, H, W)
            u, v = np.empty(shape, np.float32), np.empty(shape, np.float32)
            P.HSFlow.run_pipeline_host_multi(engs, frames, u, v, sequence=sequence, sample_step=step)
            ru, rv = (u1[:, ::step, ::step], v1[:, ::step, ::step]) if step else (u1, v1)
            assert (bits(u) == bits(ru)).all() and (bits(v) == bits(rv)).all(), step
        with pytest.raises(P.HSFlowError):              # an error inside one block comes out with its handle's message
            engs[1].set_strip(False, True)
            P.HSFlow.run_pipeline_host_multi(engs, frames, u, v, sequence=sequence, sample_step=4)
    finally:
        for e in engs:
            e.close()


def test_compute_range_on_pair_slots_and_frames_written_on_the_device(P, oracle):
    """hsflow_compute_range (two halves of the pair slots computed independently -- what the JPEG ingest uses to overlap
    decode and compute) and the device-side frame entries hsflow_map_frames / hsflow_set_frames_*_dev, against per-pair
    computes; torch only moves bytes into the mapped planes."""
    import torch
    from opticalflowhs_b200.sharding import DeviceView
    W, H, n, N = 200, 64, 6, 13
    pairs = [oracle.synth_pair(W, H, seed=90 + k) for k in range(n)]
    want = []
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, 4)
        for f1, f2 in pairs:
            e.load_pair(f1, f2).compute()
            want.append(e.read_uv())
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, 4).set_tuning(sub_batch=3).configure(W, H, n)
        assert e.sub_batch == 3
        d1, d2, rp, pp = e.map_frames(bgr=False)
        for k, (f1, f2) in enumerate(pairs):             # write the frames into the mapped planes on the device
            for base, f in ((d1, f1), (d2, f2)):
                plane = torch.as_tensor(DeviceView(base + k * pp, (H, rp), "|u1"), device="cuda")
                plane[:, :W].copy_(torch.from_numpy(f).cuda())
        torch.cuda.synchronize()
        e.compute_range(3, 3).compute_range(0, 3).sync()
        for k in range(n):
            u, v = e.read_uv(k)
            assert (bits(u) == bits(want[k][0])).all() and (bits(v) == bits(want[k][1])).all(), k
        with pytest.raises(P.HSFlowError):
            e.compute_range(0, 4)                        # more than sub_batch pairs
        with pytest.raises(P.HSFlowError):
            e.compute_range(4, 3)                        # outside the configured pairs
        # device-to-device entry with BGR frames
        bgr = np.stack([pairs[0][0]] * 3, axis=2).copy(), np.stack([pairs[0][1]] * 3, axis=2).copy()
        t1, t2 = torch.from_numpy(bgr[0]).cuda(), torch.from_numpy(bgr[1]).cuda()
        e.configure(W, H, 1).set_frames_bgr_dev(t1.data_ptr(), t2.data_ptr(), 3 * W).compute()
        u, v = e.read_uv()
        assert (bits(u) == bits(want[0][0])).all() and (bits(v) == bits(want[0][1])).all()


# ---- strips: host-mediated halo exchange on one GPU ---------------------------------------------------

def test_frame_sequence_pipeline_and_push_frame_equal_per_pair_compute(P, oracle):
    """hsflow_run_sequence_host (pair k = frames k, k+1; one upload per frame) and the streaming hsflow_push_frame_gray8
    loop (cpp:800-842 with the cpp:834 frame hand-over as a pointer swap) against loading every pair separately."""
    from opticalflowhs_b200.hsflow import pinned_empty
    W, H, n = 232, 100, 37                               # 38 frames -> 37 pairs: several sub-batches, a ragged last one
    frames = pinned_empty((n + 1, H, W), np.uint8)
    for k in range(n + 1):
        frames[k] = oracle.synth_pair(W, H, seed=500 + k)[k & 1]
    uo, vo = pinned_empty((n, H, W), np.float32), pinned_empty((n, H, W), np.float32)
    want = []
    with P.HSFlow(0) as e:
        e.set_params(15.0, 17, P.STENCIL_CL8, True, 4)
        for k in range(n):
            e.load_pair(frames[k], frames[k + 1]).compute()
            want.append(e.read_uv())
    with P.HSFlow(0) as e:
        e.set_params(15.0, 17, P.STENCIL_CL8, True, 4).set_tuning(sub_batch=4)
        e.run_sequence_host(frames, uo, vo)
        for k in range(n):
            assert (bits(uo[k]) == bits(want[k][0])).all() and (bits(vo[k]) == bits(want[k][1])).all(), k
        e.run_sequence_host(frames[:2], uo[:1], vo[:1])  # shortest sequence, re-configures the handle
        assert (bits(uo[0]) == bits(want[0][0])).all()
    with P.HSFlow(0) as e:
        e.set_params(15.0, 17, P.STENCIL_CL8, True, 4).configure(W, H, 1)
        e.push_frame(frames[0]).compute()
        u, v = e.read_uv()
        assert not u.any() and not v.any()              # a single frame: both planes equal, zero flow
        for k in range(6):
            e.push_frame(frames[k + 1]).compute()
            u, v = e.read_uv()
            assert (bits(u) == bits(want[k][0])).all() and (bits(v) == bits(want[k][1])).all(), k


@pytest.mark.parametrize("T,ghost", [(1, 1), (4, 4), (4, 8), (3, 6)])
def test_row_strips_with_halo_exchange_equal_whole_frame(P, oracle, T, ghost):
    W, H, N, nstrips = 180, 90, 2 * ghost + 3, 3
    f1, f2 = oracle.synth_pair(W, H, seed=5)
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        e.load_pair(f1, f2).compute()
        whole = e.read_uv()
    bounds = [H * k // nstrips for k in range(nstrips + 1)]
    engs, ext = [], []
    for k in range(nstrips):
        lo, hi = bounds[k], bounds[k + 1]
        a, b = max(lo - ghost, 0), min(hi + ghost + 1, H)     # +1 frame row for the j+1 derivative tap
        e = P.HSFlow(0)
        e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        e.configure(W, b - a, 1).set_strip(k == 0, k == nstrips - 1)
        e.set_frames(f1[a:b], f2[a:b]).prepare()
        engs.append(e)
        ext.append((a, b, lo, hi))
    done = 0
    while done < N:
        step = min(ghost, N - done)
        for e in engs:
            e.iterate(step)
        done += step
        cur = [e.read_uv() for e in engs]
        gu, gv = np.zeros((H, W), np.float32), np.zeros((H, W), np.float32)
        for (a, b, lo, hi), (u, v) in zip(ext, cur):
            gu[lo:hi], gv[lo:hi] = u[lo - a:hi - a], v[lo - a:hi - a]
        for e, (a, b, lo, hi) in zip(engs, ext):           # refresh every strip's ghost rows
            e.write_uv(gu[a:b], gv[a:b]).halo_refreshed()
    assert (bits(gu) == bits(whole[0])).all() and (bits(gv) == bits(whole[1])).all()
    for e in engs:
        e.close()


@pytest.mark.parametrize("T,ghost,N,rows,chunk", [(1, 1, 7, 96, 0), (4, 4, 22, 200, 0), (6, 6, 60, 400, 0), (6, 6, 45, 400, 64),
                                                  (0, 6, 31, 300, 0), (8, 8, 19, 120, 0), (4, 6, 40, 130, 0)])
def test_peer_strips_three_handles_on_one_gpu_equal_whole_frame(P, T, ghost, N, rows, chunk):
    """The fused compute + halo-exchange path (k_jacobi_stream<.., PEER = true>: seam rows stored into the neighbours'
    buffers, epoch words, cuStreamWaitValue32 between launches) with three strips living in ONE process on ONE GPU,
    connected through the same-process branch of hsflow_strip_connect.  Tall strips have many row chunks, so the seam
    chunks run first and the epoch is published while the interior still runs; chunk = 64 keeps every unit counted.
    Two consecutive runs on the same connection; bit-identical to the whole frame."""
    from opticalflowhs_b200.sharding import LocalStripSolver
    W, H, world = 1500, rows * 3 + 9, 3
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        e.configure(W, H, 1).synth_frames(0, 0, 4321).compute()
        whole = e.read_uv()
    engs = [P.HSFlow(0) for _ in range(world)]
    try:
        for e in engs:
            e.set_params(15.0, N, P.STENCIL_CL8, True, T).set_tuning(chunk_rows=chunk)
        s = LocalStripSolver(engs, W, H, ghost)
        s.load_synth(4321)
        for rep in range(2):
            s.run(N, slice_blocks=3).sync()
            u, v = s.gather_uv()
            assert (bits(u) == bits(whole[0])).all() and (bits(v) == bits(whole[1])).all(), rep
        s.close()
    finally:
        for e in engs:
            e.close()


# ---- OpenCV-mode path -------------------------------------------------------------------------------------

@pytest.mark.parametrize("name,floor", [("city", 0.99), ("bunny", 0.98)])
def test_cv_mode_matches_restated_opencv_and_golden_masks(eng, P, oracle, frames, masks, name, floor):
    g1, g2 = frames[f"{name}_1"], frames[f"{name}_2"]
    uo, vo, it = oracle.run_cv(g1, g2, 0.1, 10, eps=0)
    for math in (P.MATH_EXACT, P.MATH_FAST):
        eng.set_math(math).set_deriv(P.DERIV_CV).set_params(15.0, 10, P.STENCIL_CV4, True, 4).set_lambda(0.1)
        eng.load_pair(g1, g2).compute()
        u, v = eng.read_uv()
        assert np.abs(u - uo).max() <= TOL_MAX and np.abs(v - vo).max() <= TOL_MAX
        m, _ = eng.dot_mask(0, 4, 1.0)
        assert iou(m, masks[f"{name}_cv_l0.1_n10"]) >= floor
    eng.set_deriv(P.DERIV_CL)


# ---- EPS termination: cvTermCriteria(CV_TERMCRIT_ITER | CV_TERMCRIT_EPS, it, eps), cv.cpp:29 ------------------

def _weights(P, stencil):
    if stencil == P.STENCIL_CL8:
        return np.float32(1.0 / 6), np.float32(1.0 / 12), np.float32(15.0) * np.float32(15.0)
    return np.float32(0.25), np.float32(0.0), np.float32(1.0) / np.float32(0.1)


@pytest.mark.parametrize("stencil_name", ["cl8", "cv4"])
def test_epsilon_stop_sweep_count_and_field_bitwise(eng, P, oracle, stencil_name):
    stencil = P.STENCIL_CL8 if stencil_name == "cl8" else P.STENCIL_CV4
    f1, f2 = oracle.synth_pair(160, 120, seed=3)
    d = oracle.derivatives(f1.astype(np.float32), f2.astype(np.float32))
    we, wd, rho = _weights(P, stencil)
    seen = set()
    for eps, max_iter in [(5e-2, 300), (5e-2, 301), (4e-3, 300), (4e-3, 301), (1e-12, 40), (0.0, 17)]:
        uo, vo, it = oracle.jacobi_general_eps(*d, we, wd, rho, max_iter, eps, True)
        eng.set_math(P.MATH_EXACT).set_params(15.0, max_iter, stencil, True)
        if stencil == P.STENCIL_CV4:
            eng.set_lambda(0.1)
        eng.set_epsilon(eps).load_pair(f1, f2).compute()
        u, v = eng.read_uv()
        assert eng.iterations_done() == it, (eps, max_iter, eng.iterations_done(), it)
        assert (bits(u) == bits(uo)).all() and (bits(v) == bits(vo)).all(), (eps, max_iter)
        seen.add((it < max_iter, (max_iter - it) & 1))
    assert {(True, 0), (True, 1), (False, 0)} <= seen      # stopped early in either buffer, and ran to the cap
    eng.set_epsilon(0.0)


def test_epsilon_stop_per_pair_in_batches_and_split_iterate(P, oracle):
    W, H, n, max_iter, eps = 144, 80, 5, 260, 6e-3
    frames, want = [], []
    we, wd, rho = _weights(P, P.STENCIL_CL8)
    for k in range(n):
        f1, f2 = oracle.synth_pair(W, H, seed=40 + 7 * k)
        if k == 3:
            f2 = f1.copy()                              # no motion: converges after the first sweep
        frames.append((f1, f2))
        d = oracle.derivatives(f1.astype(np.float32), f2.astype(np.float32))
        want.append(oracle.jacobi_general_eps(*d, we, wd, rho, max_iter, eps, True))
    assert len({w[2] for w in want}) >= 3 and want[3][2] == 1
    for sub in (0, 2):                                  # one sub-batch (prepare + iterate) / sub-batches of 2
        with P.HSFlow(0) as e:
            e.set_tuning(0, 0, sub).set_math(P.MATH_EXACT).set_params(15.0, max_iter, P.STENCIL_CL8, True).set_epsilon(eps)
            e.configure(W, H, n)
            for k, (f1, f2) in enumerate(frames):
                e.set_frames(f1, f2, pair=k)
            if sub == 0:                                # the split form: two iterate calls continue one session
                e.prepare(); e.iterate(101); e.iterate(max_iter - 101); e.sync()
            else:
                e.compute()
            for k in range(n):
                u, v = e.read_uv(k)
                assert (bits(u) == bits(want[k][0])).all() and (bits(v) == bits(want[k][1])).all(), (sub, k)
                assert e.iterations_done(k) == want[k][2], (sub, k, e.iterations_done(k), want[k][2])


def test_epsilon_stop_fast_math_and_strip_mode_refusal(eng, P, oracle):
    f1, f2 = oracle.synth_pair(200, 96, seed=11)
    d = oracle.derivatives(f1.astype(np.float32), f2.astype(np.float32))
    we, wd, rho = _weights(P, P.STENCIL_CL8)
    uo, vo, it = oracle.jacobi_general_eps(*d, we, wd, rho, 400, 3e-3, True)
    eng.set_math(P.MATH_FAST).set_params(15.0, 400, P.STENCIL_CL8, True, 6).set_epsilon(3e-3)
    assert eng.temporal_block == 4                      # FAST math: blocks of 4 tracked sweeps + replay
    eng.load_pair(f1, f2).compute()
    u, v = eng.read_uv()
    assert abs(eng.iterations_done() - it) <= 2 and it < 400
    assert np.abs(u - uo).max() <= TOL_MAX and np.abs(v - vo).max() <= TOL_MAX
    eng.set_strip(False, True)
    with pytest.raises(P.HSFlowError):
        eng.prepare()
    eng.set_strip(True, True).set_epsilon(0.0)


@pytest.mark.parametrize("stencil_name,W,H", [("cl8", 200, 96), ("cv4", 200, 96), ("cl8", 333, 70), ("cl8", 130, 9)])
def test_epsilon_stop_on_blocked_kernel_equals_per_sweep_check_bitwise(P, oracle, stencil_name, W, H):
    """cvTermCriteria(ITER | EPS) (OpticalFlowOpenCV.cpp:29) on the temporally blocked kernel: blocks of up to 4 sweeps
    tracking every sweep's max-norm, and a replay of the block for the pair that met the criterion inside it, against
    the single-sweep kernel that checks after every sweep (itself bit-identical to the oracle's sweep rule in EXACT
    math): same sweep count, same field, for stops at every position inside a block, in either ping-pong buffer, for
    the iteration cap, for tail blocks (max_iter % 4 != 0) and for a LITERAL-mode run."""
    stencil = P.STENCIL_CL8 if stencil_name == "cl8" else P.STENCIL_CV4
    f1, f2 = oracle.synth_pair(W, H, seed=23)
    positions = set()
    cases = [(e, m, True) for e in (8e-2, 5e-2, 3e-2, 2e-2, 1e-2, 6e-3, 4e-3, 2.5e-3) for m in (300,)]
    cases += [(4e-3, 37, True), (1e-12, 41, True), (1e-12, 40, True), (0.5, 300, True), (1e9, 300, True), (6e-3, 300, False), (3e-2, 2, True)]
    for eps, max_iter, upd in cases:
        res = []
        for kernel in (1, 0):                           # 1: single-sweep kernel only; 0: auto = blocked TRACK kernel
            with P.HSFlow(0) as e:
                e.set_kernel(kernel).set_params(15.0, max_iter, stencil, upd, 0).set_epsilon(eps)
                if stencil == P.STENCIL_CV4:
                    e.set_lambda(0.1)
                assert e.temporal_block == (1 if kernel == 1 else 4)
                l0 = e.kernel_launches
                e.load_pair(f1, f2).compute()
                res.append((e.iterations_done(), e.read_uv(), e.kernel_launches - l0))
        (it1, (u1, v1), l1), (it4, (u4, v4), l4) = res
        assert it1 == it4, (eps, max_iter, upd, it1, it4)
        assert (bits(u1) == bits(u4)).all() and (bits(v1) == bits(v4)).all(), (eps, max_iter, upd, it1)
        positions.add((it1 - 1) % 4 if it1 < max_iter else -1)
    assert -1 in positions and len(positions - {-1}) >= 3, positions    # stops at (nearly) every position inside a block, and the cap


def test_epsilon_stop_on_blocked_kernel_per_pair_batches_split_iterate_and_pipeline(P, oracle):
    W, H, n, max_iter, eps = 144, 80, 6, 260, 6e-3
    frames = np.empty((n, 2, H, W), np.uint8)
    for k in range(n):
        frames[k, 0], frames[k, 1] = oracle.synth_pair(W, H, seed=40 + 7 * k)
    frames[3, 1] = frames[3, 0]                          # no motion: converges after the first sweep
    want = []
    with P.HSFlow(0) as e:
        e.set_kernel(1).set_params(15.0, max_iter, P.STENCIL_CL8, True).set_epsilon(eps)
        for k in range(n):
            e.load_pair(frames[k, 0], frames[k, 1]).compute()
            want.append((e.read_uv(), e.iterations_done()))
    assert len({w[1] for w in want}) >= 3 and want[3][1] == 1
    for sub in (0, 2):                                  # one sub-batch (prepare + iterate) / sub-batches of 2
        with P.HSFlow(0) as e:
            e.set_tuning(0, 0, sub).set_params(15.0, max_iter, P.STENCIL_CL8, True).set_epsilon(eps)
            e.configure(W, H, n)
            for k in range(n):
                e.set_frames(frames[k, 0], frames[k, 1], pair=k)
            if sub == 0:                                # the split form: iterate calls continue one session, odd split points
                e.prepare(); e.iterate(101); e.iterate(2); e.iterate(max_iter - 103); e.sync()
            else:
                e.compute()
            for k in range(n):
                u, v = e.read_uv(k)
                assert e.iterations_done(k) == want[k][1], (sub, k, e.iterations_done(k), want[k][1])
                assert (bits(u) == bits(want[k][0][0])).all() and (bits(v) == bits(want[k][0][1])).all(), (sub, k)
    with P.HSFlow(0) as e:                              # through the host pipeline (ragged sub-batches of 4)
        e.set_tuning(0, 0, 4).set_params(15.0, max_iter, P.STENCIL_CL8, True).set_epsilon(eps)
        u, v = np.empty((n, H, W), np.float32), np.empty((n, H, W), np.float32)
        e.run_batch_host(frames, u, v)
        for k in range(n):
            assert (bits(u[k]) == bits(want[k][0][0])).all() and (bits(v[k]) == bits(want[k][0][1])).all(), k


# ---- full-size frames: window check through the domain of dependence -----------------------------------------

@pytest.mark.parametrize("W,H,N,T", [(3840, 2160, 100, 4), (3840, 2160, 100, 8), (1920, 1080, 200, 6)])
def test_full_size_frame_windows_against_oracle_crops(P, oracle, W, H, N, T):
    K = 48
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, T)
        e.configure(W, H, 1).synth_frames(0, 0, 1234).compute()
        u, v = e.read_uv()
    assert np.isfinite(u).all() and np.isfinite(v).all()
    f1, f2 = oracle.synth_pair(W, H, seed=1234)
    # windows at the corners, on strip seams (multiples of the valid strip width) and in the middle
    vw = 128 - 2 * ((T + 3) // 4 * 4)
    spots = [(0, 0), (H - K, W - K), (0, W - K), (H - K, 0), (H // 2, 7 * vw - K // 2), (H // 3, W // 2)]
    for (y, x) in spots:
        y0, y1 = max(y - N, 0), min(y + K + N, H)
        x0, x1 = max(x - N, 0), min(x + K + N, W)
        uo, vo = oracle.run_cl(f1[y0:y1, x0:x1], f2[y0:y1, x0:x1], 15.0, N, True)
        # the derivative tap j+1/i+1 and N sweeps stay inside the crop except at true image edges
        yy, xx = y - y0, x - x0
        du = np.abs(u[y:y + K, x:x + K] - uo[yy:yy + K, xx:xx + K]).max()
        dv = np.abs(v[y:y + K, x:x + K] - vo[yy:yy + K, xx:xx + K]).max()
        assert du <= TOL_MAX and dv <= TOL_MAX, (y, x, du, dv)


def test_bench_geometry_batch_windows_against_oracle_crops(P, oracle):
    """The timed configuration of bench.py (BASELINE.json configs[3]): 4K pairs, 100 iterations, FULL mode, automatic
    temporal block (7), MANY pairs in ONE launch (pair coordinate z up to 47, several waves of work units).  Windows
    of the first, a middle and the last pair -- at corners, on strip seams (multiples of 112 columns) and on row-chunk
    seams -- against the oracle on their domains of dependence."""
    W, H, N, K, pairs, seed0 = 3840, 2160, 100, 40, 48, 1234
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, 0)
        e.configure(W, H, pairs).synth_frames(0, 0, seed0)
        assert e.sub_batch == pairs and e.temporal_block == 7       # throughput regime: the deep block
        l0 = e.kernel_launches
        e.compute()
        assert e.kernel_launches - l0 == 1 + (N + 6) // 7          # one derivative launch + 15 blocks (10 x 7 + 5 x 6) over all pairs
        fields = {z: e.read_uv(z) for z in (0, pairs // 2 + 1, pairs - 1)}
    spots = [(0, 0), (H - K, W - K), (H // 2 - K // 2, 9 * 112 - K // 2), (3 * 24 - K // 2, 20 * 112 - K // 2), (1111, 1777)]
    for z, (u, v) in fields.items():
        assert np.isfinite(u).all() and np.isfinite(v).all()
        for (y, x) in spots:
            du, dv = oracle.window_error(u[y:y + K, x:x + K], v[y:y + K, x:x + K], W, H, N, seed0 + z, y, x)
            assert du <= TOL_MAX and dv <= TOL_MAX, (z, y, x, du, dv)


def test_baseline_config1_bunny_cv_path_as_written(P, oracle, frames):
    """BASELINE.json configs[0] exactly as the reference runs it (main.cpp:4, 8, 20; OpticalFlowOpenCV.cpp:27-29): the bunny
    pair, lambda = 0.1, 100 iterations, cvTermCriteria(ITER | EPS, 100, 1e-6), both 3x3 blurs.  The comparator is the
    restated cvCalcOpticalFlowHS (cv210.dll's source is not in the reference tree; the restatement is pinned pixel-exactly
    by the shipped *_cv_out.jpg, tests/test_oracle.py): EXACT math bit-identical in field and sweep count, FAST math within
    1e-3 px."""
    g1, g2 = frames["bunny_1"], frames["bunny_2"]
    uo, vo, it = oracle.run_cv(g1, g2, 0.1, 100, eps=1e-6)
    for math in (P.MATH_EXACT, P.MATH_FAST):
        with P.HSFlow(0) as e:
            e.set_math(math).set_deriv(P.DERIV_CV).set_params(0.0, 100, P.STENCIL_CV4, True, 0).set_lambda(0.1).set_epsilon(1e-6)
            e.load_pair(g1, g2).compute()
            u, v = e.read_uv()
            done = e.iterations_done(0)
        assert np.abs(u - uo).max() <= TOL_MAX and np.abs(v - vo).max() <= TOL_MAX, math
        assert epe_diff(u, v, uo, vo) <= TOL_EPE
        if math == P.MATH_EXACT:
            assert done == it
            assert (bits(u) == bits(uo)).all() and (bits(v) == bits(vo)).all()
        else:
            assert abs(done - it) <= 1 or done == it == 100


def test_16k_frame_500_iterations_windows_against_oracle_crops(P, oracle):
    """BASELINE.json configs[4] at full size on one GPU (auto temporal block): windows on strip and chunk seams, at the
    image corners and in the middle, each checked against the oracle on its domain of dependence."""
    W = H = 16384
    N, K = 500, 40
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, 0)
        e.configure(W, H, 1).synth_frames(0, 0, 1234).compute()
        assert e.temporal_block == 7
        u, v = e.read_uv()
    assert np.isfinite(u[::7, ::5]).all() and np.isfinite(v[::7, ::5]).all()
    spots = [(0, 0), (H - K, W - K), (2049 - K // 2, 73 * 112 - K // 2), (H // 2 + 3, W // 3)]
    for (y, x) in spots:
        y0, y1 = max(y - N, 0), min(y + K + N, H)
        x0, x1 = max(x - N, 0), min(x + K + N, W)
        f1, f2 = oracle.synth_pair(W, H, seed=1234, row0=y0, rows=y1 - y0)
        uo, vo = oracle.run_cl(np.ascontiguousarray(f1[:, x0:x1]), np.ascontiguousarray(f2[:, x0:x1]), 15.0, N, True)
        yy, xx = y - y0, x - x0
        du = np.abs(u[y:y + K, x:x + K] - uo[yy:yy + K, xx:xx + K]).max()
        dv = np.abs(v[y:y + K, x:x + K] - vo[yy:yy + K, xx:xx + K]).max()
        assert du <= TOL_MAX and dv <= TOL_MAX, (y, x, du, dv)


# ---- CUDA graph replay of repeated computes (launch-bound small frames) ------------------------------------------

@pytest.mark.parametrize("mode", ["full", "literal", "eps", "exact"])
def test_graph_replay_equals_eager_compute(P, oracle, mode):
    """hsflow_compute on a small frame: first call eager, second captured into a CUDA graph, later ones replayed.  Every
    call must give the eager result bit for bit -- also after the frame planes were rewritten in place, after the
    camera-loop pointer swap (another graph), and with the EPS criterion deciding on the device inside the graph."""
    W, H, N = 600, 480, 100
    seq = [oracle.synth_pair(W, H, seed=60 + k)[0] for k in range(5)]

    def make(graph_mode):
        e = P.HSFlow(0)
        e.set_graph(graph_mode).set_params(15.0, N, P.STENCIL_CL8, mode != "literal", 0)
        if mode == "eps":
            e.set_epsilon(2e-2)
        if mode == "exact":
            e.set_math(P.MATH_EXACT).set_params(15.0, 12, P.STENCIL_CL8, True, 0)
        return e

    ref, eng = make(1), make(0)
    try:
        for rep in range(4):                            # same planes, frames rewritten in place from the third call on
            a, b = (seq[0], seq[1]) if rep < 2 else (seq[rep], seq[rep - 1])
            for e in (ref, eng):
                if rep == 0:
                    e.load_pair(a, b)
                elif rep >= 2:
                    e.set_frames(a, b)
                e.compute()
            (u0, v0), (u1, v1) = ref.read_uv(), eng.read_uv()
            assert (bits(u0) == bits(u1)).all() and (bits(v0) == bits(v1)).all(), rep
            assert ref.iterations_done() == eng.iterations_done()
        l0, l1 = ref.kernel_launches, eng.kernel_launches
        ref.compute(); eng.compute()
        assert ref.kernel_launches - l0 == eng.kernel_launches - l1      # a replay counts the launches it contains
        for k in range(1, 5):                           # camera loop: the plane pointers swap with every frame
            for e in (ref, eng):
                e.push_frame(seq[k]).compute()
            (u0, v0), (u1, v1) = ref.read_uv(), eng.read_uv()
            assert (bits(u0) == bits(u1)).all() and (bits(v0) == bits(v1)).all(), k
    finally:
        ref.close(); eng.close()


# ---- error behaviour of the boundary ------------------------------------------------------------------------------

def test_error_codes_and_messages(eng, P):
    with pytest.raises(P.HSFlowError):
        eng.compute()                                   # nothing loaded
    with pytest.raises(P.HSFlowError):
        eng.set_params(15.0, -1)
    with pytest.raises(P.HSFlowError):
        eng.set_params(15.0, 10, 5)
    with pytest.raises(P.HSFlowError):
        eng.configure(0, 10, 1)
    eng.configure(16, 16, 2)
    with pytest.raises(P.HSFlowError):
        eng.set_frames(np.zeros((16, 16), np.uint8), np.zeros((16, 16), np.uint8), pair=2)
    with pytest.raises(P.HSFlowError):
        P.HSFlow(999)
    # zero iterations: u = v = 0 (cpp:331-332)
    eng.set_params(15.0, 0)
    eng.load_pair(np.full((8, 8), 3, np.uint8), np.full((8, 8), 9, np.uint8)).compute()
    u, v = eng.read_uv()
    assert not u.any() and not v.any()
    assert eng.kernel_launches > 0
