#!/usr/bin/env python
"""Generates tests/golden/ref_alignment_pair.npz from the reference's own test data (/root/reference/test_data, used by
src/opt/test/test_alignment.cc:50-84 "TestPairAlignment" through test_alignment_util.cc:134-300). /root/reference does not exist on
the GPU box, so the decoded arrays travel as a fixture:
  a_gray, b_gray   cv::imread(IMREAD_GRAYSCALE) of images/{a,b}_image.png       (what Problem::LoadImages reads, image.cc:48)
  a_bgr            cv::imread of images/a_image.png                             (point colours, test_alignment_util.cc:168-192)
  a_depth          cv::imread(IMREAD_UNCHANGED) of images/a_depth.png, uint16   (:149-151)
  plus the numbers of small_offset.txt / identical_images.txt.
Run in the build container:  python tests/golden/make_alignment_fixture.py"""
import os

import cv2
import numpy as np

SRC = "/root/reference/test_data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_alignment_pair.npz")


def read_info(name):
    tok = open(os.path.join(SRC, name)).read().split()
    assert tok[0] == "calibration"
    calib = [float(v) for v in tok[1:8]]
    i = tok.index("a_t_b")
    a_t_b = np.array([float(v) for v in tok[i + 1:i + 13]]).reshape(3, 4)
    depth = float(tok[tok.index("average_scene_depth") + 1])
    paths = {k: tok[tok.index(k) + 1] for k in ("a_image", "a_depth", "b_image", "b_depth")}
    return calib, a_t_b, depth, paths


def main():
    out = {}
    for key, name in (("small_offset", "small_offset.txt"), ("identical", "identical_images.txt")):
        calib, a_t_b, depth, paths = read_info(name)
        out[key + "_calibration"] = np.array(calib)          # width height fx fy cx cy depth_factor
        out[key + "_a_t_b"] = a_t_b
        out[key + "_average_scene_depth"] = np.array(depth)
        out[key + "_a_gray"] = cv2.imread(os.path.join(SRC, paths["a_image"]), cv2.IMREAD_GRAYSCALE)
        out[key + "_b_gray"] = cv2.imread(os.path.join(SRC, paths["b_image"]), cv2.IMREAD_GRAYSCALE)
        out[key + "_a_bgr"] = cv2.imread(os.path.join(SRC, paths["a_image"]))
        out[key + "_a_depth"] = cv2.imread(os.path.join(SRC, paths["a_depth"]), cv2.IMREAD_UNCHANGED)
        assert out[key + "_a_depth"].dtype == np.uint16
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes;", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
