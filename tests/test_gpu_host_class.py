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
