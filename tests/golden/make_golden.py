#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box, so the outputs are committed).

  frames.npz  : the four shipped input frames decoded (cv2.imread) and converted to gray with
                cv2.cvtColor(BGR2GRAY) (== OpenCV 2.1 fixed point, HSOpticalFlowOpenCL.cpp:727-728),
                plus the decoded BGR of the bunny pair for the BGR->gray ingest path.
  masks.npz   : stride-4 dot masks recovered from the shipped OUTPUT pictures
                (OpticalFlowHS/{city,bunny}_cl_out.jpg, Release/bunny_cl_out.jpg, *_cv_out.jpg):
                a dot is present at (i, j) iff >= 5 of the 3x3 pixels centred there have
                max(B,G,R) > 120 (the blue filled circle of cpp:766 / cv.cpp:42).
  pictures.npz: the five shipped OUTPUT pictures decoded (cv2.imread), BGR uint8.  They are exact known-answer vectors
                for BOTH paths: drawing the oracle's fields like cpp:758-770 / cv.cpp:32-46 (cv2.circle + cv2.line),
                JPEG-encoding at quality 95 (the cvSaveImage default) and decoding again reproduces every pixel of
                *_cl_out.jpg (alpha = 15, N = 10 / 2, LITERAL) and of *_cv_out.jpg (lambda = 0.1, N = 10, eps 1e-6).
  fields.npz  : u/v produced HERE by the reference's own Kernels.cl compiled for the host
                (oracle/_ref/libclref.so): city pair, alpha=15, N=100, LITERAL and FULL modes,
                sampled on the stride-4 drawing grid; and bunny N=10.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle as O  # noqa: E402

REF = "/root/reference"


def dots_from_picture(path):
    out = cv2.imread(path, 1)
    mx = out.max(axis=2)
    h, w = mx.shape
    m = np.zeros(((h + 3) // 4, (w + 3) // 4), bool)
    for i in range(0, h, 4):
        for j in range(0, w, 4):
            ys = slice(max(i - 1, 0), min(i + 2, h))
            xs = slice(max(j - 1, 0), min(j + 2, w))
            m[i // 4, j // 4] = (mx[ys, xs] > 120).sum() >= 5
    return m


def main():
    frames = {}
    for name in ("city", "bunny"):
        for k in (1, 2):
            bgr = cv2.imread(f"{REF}/OpticalFlowHS/{name}_{k}.jpg", 1)
            frames[f"{name}_{k}"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
            if name == "bunny":
                frames[f"{name}_{k}_bgr"] = bgr
    np.savez_compressed(os.path.join(HERE, "frames.npz"), **frames)

    masks = {
        "city_cl_a15_n10": dots_from_picture(f"{REF}/OpticalFlowHS/city_cl_out.jpg"),
        "bunny_cl_a15_n10": dots_from_picture(f"{REF}/OpticalFlowHS/bunny_cl_out.jpg"),
        "bunny_cl_a15_n2": dots_from_picture(f"{REF}/Release/bunny_cl_out.jpg"),
        "city_cv_l0.1_n10": dots_from_picture(f"{REF}/OpticalFlowHS/city_cv_out.jpg"),
        "bunny_cv_l0.1_n10": dots_from_picture(f"{REF}/OpticalFlowHS/bunny_cv_out.jpg"),
    }
    np.savez_compressed(os.path.join(HERE, "masks.npz"), **masks)

    pictures = {
        "city_cl_a15_n10": cv2.imread(f"{REF}/OpticalFlowHS/city_cl_out.jpg", 1),
        "bunny_cl_a15_n10": cv2.imread(f"{REF}/OpticalFlowHS/bunny_cl_out.jpg", 1),
        "bunny_cl_a15_n2": cv2.imread(f"{REF}/Release/bunny_cl_out.jpg", 1),
        "city_cv_l0.1_n10": cv2.imread(f"{REF}/OpticalFlowHS/city_cv_out.jpg", 1),
        "bunny_cv_l0.1_n10": cv2.imread(f"{REF}/OpticalFlowHS/bunny_cv_out.jpg", 1),
    }
    np.savez_compressed(os.path.join(HERE, "pictures.npz"), **pictures)

    fields = {}
    for name, n in (("city", 100), ("bunny", 10)):
        g1, g2 = frames[f"{name}_1"], frames[f"{name}_2"]
        for mode, upd in (("literal", False), ("full", True)):
            u, v = O.ref_run(g1, g2, 15.0, n, update_v=upd)
            fields[f"{name}_n{n}_{mode}_u"] = u[::4, ::4].copy()
            fields[f"{name}_n{n}_{mode}_v"] = v[::4, ::4].copy()
    np.savez_compressed(os.path.join(HERE, "fields.npz"), **fields)
    for f in ("frames.npz", "masks.npz", "pictures.npz", "fields.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
