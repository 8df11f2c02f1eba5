"""End-to-end through the reference's UNCHANGED main.cpp linked against the drop-in classes
(opticalflowhs_b200/bin/OpticalFlowHS, built by opticalflowhs_b200.build_host where the
reference tree exists) and through the image I/O of libhsflow_host.so.  GPU box only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, iou

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "opticalflowhs_b200", "bin", "OpticalFlowHS")
HOSTLIB = os.path.join(ROOT, "opticalflowhs_b200", "libhsflow_host.so")


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def read_ppm(path):
    with open(path, "rb") as f:
        data = f.read()
    assert data[:2] == b"P6"
    parts = data.split(b"\n", 3)
    w, h = map(int, parts[1].split())
    return np.frombuffer(parts[3], np.uint8, w * h * 3).reshape(h, w, 3)


def dots_from_picture(img):
    mx = img.max(axis=2)
    h, w = mx.shape
    m = np.zeros(((h + 3) // 4, (w + 3) // 4), bool)
    for i in range(0, h, 4):
        for j in range(0, w, 4):
            m[i // 4, j // 4] = (mx[max(i - 1, 0):i + 2, max(j - 1, 0):j + 2] > 120).sum() >= 5
    return m


def run_main(args, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([EXE] + args, cwd=cwd, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=120, env=e)


@pytest.mark.skipif(not os.path.exists(EXE), reason="reference main.cpp binary not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name,n", [("city", 10), ("bunny", 10), ("bunny", 2)])
def test_unchanged_main_cl_disk_reproduces_shipped_pictures(tmp_path, frames, masks, name, n):
    write_pgm(tmp_path / "a.pgm", frames[f"{name}_1"])
    write_pgm(tmp_path / "b.pgm", frames[f"{name}_2"])
    # argv layout of main:91-100: -cl -hd in1 in2 out alpha iterations group-size device
    r = run_main(["-cl", "-hd", "a.pgm", "b.pgm", "out.ppm", "15", str(n), "1", "GPU"], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OpenCL dysk!" in r.stdout and "Avg time:" in r.stdout
    m = dots_from_picture(read_ppm(tmp_path / "out.ppm"))
    assert iou(m, masks[f"{name}_cl_a15_n{n}"]) == 1.0     # the class defaults to the shipped (LITERAL) kernel semantics


@pytest.mark.skipif(not os.path.exists(EXE), reason="reference main.cpp binary not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name,n", [("city", 10), ("bunny", 10), ("bunny", 2)])
def test_unchanged_main_writes_the_shipped_picture_pixel_for_pixel(tmp_path, oracle, frames, pictures, name, n):
    """End to end through the reference's unchanged main.cpp: the picture our class writes, saved as JPEG at cvSaveImage's
    default quality and loaded again, IS the shipped *_cl_out.jpg -- every pixel (threshold decisions, circles, line end
    points trunc(x + u) of all grid points).  Bit-exact arithmetic (HSFLOW_EXACT=1) must hit it exactly; the default FAST
    arithmetic (fields within 4e-6 px) is allowed a handful of end points that sit on an integer boundary."""
    pytest.importorskip("cv2")
    write_pgm(tmp_path / "a.pgm", frames[f"{name}_1"])
    write_pgm(tmp_path / "b.pgm", frames[f"{name}_2"])
    gold = pictures[f"{name}_cl_a15_n{n}"]
    for env, budget in (({"HSFLOW_EXACT": "1"}, 0), ({}, 60)):
        r = run_main(["-cl", "-hd", "a.pgm", "b.pgm", "out.ppm", "15", str(n), "1", "GPU"], tmp_path, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        bgr = np.ascontiguousarray(read_ppm(tmp_path / "out.ppm")[..., ::-1])
        if budget == 0:
            assert (oracle.jpeg_roundtrip(bgr) == gold).all()
        else:
            u, v = oracle.run_cl(frames[f"{name}_1"], frames[f"{name}_2"], 15.0, n, False)
            assert (bgr != oracle.render_flow(u, v, 0.5, 1.0)).any(axis=2).sum() <= budget


@pytest.mark.skipif(not os.path.exists(EXE), reason="reference main.cpp binary not built")
@pytest.mark.parametrize("name", ["city", "bunny"])
def test_unchanged_main_cv_picture_against_the_shipped_one(tmp_path, oracle, frames, pictures, name):
    """-cv -hd (OpticalFlowOpenCV::runFromImg, cv.cpp:7-52) with lambda = 0.1, 10 iterations: the restated
    cvCalcOpticalFlowHS reproduces *_cv_out.jpg exactly (tests/test_oracle.py).  With HSFLOW_EXACT=1 the GPU follows its
    rounding sequence and the picture main writes is the shipped one, pixel for pixel; the default FAST arithmetic may
    put a few of the ~1 000 - 2 000 line end points on the other side of an integer."""
    pytest.importorskip("cv2")
    write_pgm(tmp_path / "a.pgm", frames[f"{name}_1"])
    write_pgm(tmp_path / "b.pgm", frames[f"{name}_2"])
    u, v, _ = oracle.run_cv(frames[f"{name}_1"], frames[f"{name}_2"], 0.1, 10, eps=1e-6)
    ref = oracle.render_flow(u, v, 1.0, 0.5)
    gold = pictures[f"{name}_cv_l0.1_n10"]
    assert (oracle.jpeg_roundtrip(ref) == gold).all()
    for env in ({"HSFLOW_EXACT": "1"}, {}):
        r = run_main(["-cv", "-hd", "a.pgm", "b.pgm", "out.ppm", ".1", "10"], tmp_path, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        bgr = np.ascontiguousarray(read_ppm(tmp_path / "out.ppm")[..., ::-1])
        if env:                                             # rounding sequence of cvCalcOpticalFlowHS: the shipped picture, every pixel
            assert (oracle.jpeg_roundtrip(bgr) == gold).all()
        else:
            differing = (bgr != ref).any(axis=2).sum()
            assert differing <= 120, differing              # out of 288 000 / 101 760 pixels


@pytest.mark.skipif(not os.path.exists(EXE), reason="reference main.cpp binary not built")
def test_unchanged_main_cv_disk_and_error_paths(tmp_path, frames, masks):
    write_pgm(tmp_path / "a.pgm", frames["city_1"])
    write_pgm(tmp_path / "b.pgm", frames["city_2"])
    r = run_main(["-cv", "-hd", "a.pgm", "b.pgm", "out.ppm", ".1", "10"], tmp_path)     # main:127-138
    assert r.returncode == 0 and "OpenCV dysk!" in r.stdout, r.stdout + r.stderr
    m = dots_from_picture(read_ppm(tmp_path / "out.ppm"))
    assert iou(m, masks["city_cv_l0.1_n10"]) >= 0.99
    r = run_main(["-cl", "-hd", "missing.pgm", "b.pgm", "o.ppm", "15", "10", "1", "GPU"], tmp_path)
    assert "Input image error." in r.stdout                                             # cpp:722-725
    r = run_main(["-cl", "-hd", "a.pgm"], tmp_path)
    assert "Bledna lista argumentow!" in r.stdout                                       # main:40-43
    r = run_main(["-cl", "-cam", "15", "10", "1", "GPU"], tmp_path)                    # no capture device
    assert "capture is NULL" in r.stderr


@pytest.mark.skipif(not os.path.exists(EXE), reason="reference main.cpp binary not built")
def test_frame_sequence_replaces_camera_loop(tmp_path, oracle):
    for k in range(4):
        f1, _ = oracle.synth_pair(160, 120, seed=3, row0=0)
        write_pgm(tmp_path / f"f{k:04d}.pgm", np.roll(f1, 2 * k, axis=1))
    r = run_main(["-cl", "-cam", "15", "20", "1", "GPU"], tmp_path,
                 env={"HSFLOW_FRAMES": "f%04d.pgm", "HSFLOW_FRAMES_OUT": "o%04d.ppm", "HSFLOW_UPDATE_V": "1"})
    assert r.returncode == 0 and "Avg time:" in r.stdout, r.stdout + r.stderr
    assert all((tmp_path / f"o{k:04d}.ppm").exists() for k in (1, 2, 3))


def test_image_io_pnm_exact_and_jpeg_roundtrip(tmp_path, oracle):
    L = C.CDLL(HOSTLIB)
    L.hsimg_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint8))]
    L.hsimg_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.hsimg_last_error.restype = C.c_char_p

    def read(path):
        w, h, ch, p = C.c_int(), C.c_int(), C.c_int(), C.POINTER(C.c_uint8)()
        rc = L.hsimg_read(str(path).encode(), C.byref(w), C.byref(h), C.byref(ch), C.byref(p))
        assert rc == 0, L.hsimg_last_error()
        a = np.ctypeslib.as_array(p, shape=(h.value, w.value, ch.value)).copy()
        L.hsimg_free(p)
        return a

    f1, _ = oracle.synth_pair(320, 200, seed=11)
    bgr = np.stack([f1, np.roll(f1, 5, 0), np.roll(f1, 9, 1)], axis=2).copy()
    assert L.hsimg_write(str(tmp_path / "g.pgm").encode(), f1.ctypes.data, 320, 200, 1) == 0
    assert (read(tmp_path / "g.pgm")[..., 0] == f1).all()
    assert L.hsimg_write(str(tmp_path / "c.ppm").encode(), bgr.ctypes.data, 320, 200, 3) == 0
    assert (read(tmp_path / "c.ppm") == bgr).all()
    smooth = np.kron(f1[::8, ::8], np.ones((8, 8), np.uint8))[:200, :320]
    img = np.stack([smooth] * 3, axis=2).copy()
    assert L.hsimg_write(str(tmp_path / "s.jpg").encode(), img.ctypes.data, 320, 200, 3) == 0, L.hsimg_last_error()
    back = read(tmp_path / "s.jpg")                       # nvJPEG encode -> nvJPEG decode
    assert back.shape == img.shape
    assert np.abs(back.astype(int) - img).mean() < 3.0


# ---- f2 ingest: JPEG -> device frames without a host bounce (include/hsflow_ingest.h) -------------------------------

def _jpeg_frames(oracle, n, W=320, H=200):
    """Smooth colour frames (so that decoders agree to about one grey level) as JPEG bitstreams + the cv2 decode."""
    cv2 = pytest.importorskip("cv2")
    base, _ = oracle.synth_pair(W, H, seed=21)
    base = cv2.GaussianBlur(base, (0, 0), 2.0)
    streams, decoded = [], []
    for k in range(n):
        g = np.roll(base, (k, 2 * k), axis=(0, 1))
        bgr = np.stack([g, np.roll(g, 3, 1), 255 - g], axis=2).copy()
        ok, enc = cv2.imencode(".jpg", bgr, [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])
        assert ok
        streams.append(enc.tobytes())
        decoded.append(cv2.imdecode(enc, cv2.IMREAD_COLOR))
    return streams, decoded


def test_jpeg_pair_decoded_on_gpu_into_frame_planes(tmp_path, oracle):
    """hsingest_load_pair_jpeg / _files: nvJPEG writes BGR straight into the handle's planes, k_deriv<BGR8> does the gray
    conversion.  Against (1) the oracle on the pixels the same decoder delivers to the host (tight) and (2) the oracle on
    cv2.imdecode's pixels (another decoder: loose)."""
    import opticalflowhs_b200 as P
    from opticalflowhs_b200 import ingest
    L = C.CDLL(HOSTLIB)
    L.hsimg_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint8))]
    streams, cvdec = _jpeg_frames(oracle, 2)
    assert ingest.jpeg_info(streams[0]) == (320, 200, 3)
    paths = []
    for k, s in enumerate(streams):
        paths.append(str(tmp_path / f"f{k}.jpg"))
        with open(paths[-1], "wb") as f:
            f.write(s)
    host = []
    for p in paths:                                       # the same decoder, delivered to the host by hsimg_read
        w, h, ch, ptr = C.c_int(), C.c_int(), C.c_int(), C.POINTER(C.c_uint8)()
        assert L.hsimg_read(p.encode(), C.byref(w), C.byref(h), C.byref(ch), C.byref(ptr)) == 0
        host.append(np.ctypeslib.as_array(ptr, shape=(h.value, w.value, ch.value)).copy())
        L.hsimg_free(ptr)
    assert np.abs(host[0].astype(int) - cvdec[0]).mean() < 1.0        # nvJPEG vs libjpeg: about a grey level
    N = 30
    uo, vo = oracle.run_cl(oracle.bgr2gray(host[0]), oracle.bgr2gray(host[1]), 15.0, N, True)
    uc, vc = oracle.run_cl(oracle.bgr2gray(cvdec[0]), oracle.bgr2gray(cvdec[1]), 15.0, N, True)
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, 4)
        for load in (lambda: ingest.load_pair_jpeg(e, streams[0], streams[1]), lambda: ingest.load_pair_files(e, paths[0], paths[1])):
            load()
            e.compute()
            u, v = e.read_uv()
            assert np.abs(u - uo).max() <= 1e-3 and np.abs(v - vo).max() <= 1e-3
            assert np.abs(u - uc).mean() < 0.05 and np.abs(v - vc).mean() < 0.05
        # OpenCV-mode path on BGR planes (k_bgr2gray + blur + Sobel estimator)
        e.set_deriv(P.DERIV_CV).set_params(0.0, N, P.STENCIL_CV4, True, 4).set_lambda(0.1)
        ingest.load_pair_jpeg(e, streams[0], streams[1])
        e.compute()
        u, v = e.read_uv()
        uo, vo, _ = oracle.run_cv(oracle.bgr2gray(host[0]), oracle.bgr2gray(host[1]), 0.1, N, eps=0)
        assert np.abs(u - uo).max() <= 1e-3 and np.abs(v - vo).max() <= 1e-3
        # camera-loop step: swap + decode of the new frame only
        e.set_deriv(P.DERIV_CL).set_params(15.0, N, P.STENCIL_CL8, True, 4)
        ingest.load_pair_files(e, paths[0], paths[0])
        ingest.push_frame_file(e, paths[1])
        e.compute()
        u, v = e.read_uv()
        uo, vo = oracle.run_cl(oracle.bgr2gray(host[0]), oracle.bgr2gray(host[1]), 15.0, N, True)
        assert np.abs(u - uo).max() <= 1e-3 and np.abs(v - vo).max() <= 1e-3


def test_jpeg_video_batch_decode_overlapped_with_compute(oracle):
    """hsingest_run_jpeg_batch: batched nvJPEG decode of the next chunk of pairs overlapping the compute of the current
    one, pairs and consecutive-frame (sequence) layouts, full and sampled fields.  Checked against the oracle on cv2's
    decode (other decoder: loose) and for consistency between the layouts (same decoder: exact)."""
    import opticalflowhs_b200 as P
    from opticalflowhs_b200 import ingest
    n, N = 9, 20
    streams, cvdec = _jpeg_frames(oracle, n, 256, 160)
    gray = [oracle.bgr2gray(f) for f in cvdec]
    with P.HSFlow(0) as e:
        e.set_params(15.0, N, P.STENCIL_CL8, True, 4)
        us, vs, st = ingest.run_jpeg_batch(e, streams, sequence=True)
        assert us.shape == (n - 1, 160, 256) and st["images"] == n          # every frame decoded exactly once
        for k in range(n - 1):
            uo, vo = oracle.run_cl(gray[k], gray[k + 1], 15.0, N, True)
            assert np.abs(us[k] - uo).mean() < 0.05 and np.abs(vs[k] - vo).mean() < 0.05, k
        pairs = [s for k in range(n - 1) for s in (streams[k], streams[k + 1])]
        up, vp, st = ingest.run_jpeg_batch(e, pairs, sequence=False)
        assert st["images"] == 2 * (n - 1)
        assert (up.view(np.uint32) == us.view(np.uint32)).all() and (vp.view(np.uint32) == vs.view(np.uint32)).all()
        u4, v4, _ = ingest.run_jpeg_batch(e, streams, sequence=True, sample_step=4)
        assert (u4.view(np.uint32) == us[:, ::4, ::4].view(np.uint32)).all() and (v4.view(np.uint32) == vs[:, ::4, ::4].view(np.uint32)).all()
        with pytest.raises(P.HSFlowError):
            ingest.run_jpeg_batch(e, streams[:1], sequence=True)


@pytest.mark.skipif(not os.path.exists(EXE), reason="reference main.cpp binary not built")
def test_unchanged_main_cv_camera_loop_reads_frame_files(tmp_path, oracle):
    """OpticalFlowOpenCV::runFromCamera (cv.cpp:56-131) through the unchanged main: frames from HSFLOW_FRAMES, each paired
    with the previous one, with and without use_previous."""
    for k in range(4):
        f1, _ = oracle.synth_pair(160, 120, seed=3, row0=0)
        write_pgm(tmp_path / f"f{k:04d}.pgm", np.roll(f1, 2 * k, axis=1))
    for warm in ("0", "1"):
        r = run_main(["-cv", "-cam", "0.1", "20"], tmp_path,
                     env={"HSFLOW_FRAMES": "f%04d.pgm", "HSFLOW_FRAMES_OUT": f"c{warm}_%04d.ppm", "HSFLOW_USE_PREVIOUS": warm})
        assert r.returncode == 0 and "Avg time:" in r.stdout, r.stdout + r.stderr
        assert all((tmp_path / f"c{warm}_{k:04d}.ppm").exists() for k in (1, 2, 3))
    r = run_main(["-cv", "-cam", "0.1", "20"], tmp_path)              # no frame source: the reference's error path
    assert "capture is NULL" in r.stdout + r.stderr
