"""Regenerates tests/golden/pyrlk_scene.npz — the input of the reference's own LK integration
test (tests/pyrlk.cc:14-50): two 100x100 u8 images, a 5-px white square at (50,50) / (52,52)
(draw::square), blurred with cv::GaussianBlur(9x9, sigmaX=3, sigmaY=5, BORDER_REPLICATE).
Needs python cv2 (present in the build container, not needed at test time)."""
import os

import cv2
import numpy as np


def square(img, center, width, fill):
    # vpp/draw/square.hh: pixels p with |p - center|_inf <= width / 2 (integer division)
    h = width // 2
    r, c = center
    img[r - h:r + h + 1, c - h:c + h + 1] = fill


i1 = np.zeros((100, 100), np.uint8)
i2 = np.zeros((100, 100), np.uint8)
square(i1, (50, 50), 5, 255)
square(i2, (52, 52), 5, 255)
b1 = cv2.GaussianBlur(i1, (9, 9), sigmaX=3, sigmaY=5, borderType=cv2.BORDER_REPLICATE)
b2 = cv2.GaussianBlur(i2, (9, 9), sigmaX=3, sigmaY=5, borderType=cv2.BORDER_REPLICATE)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "pyrlk_scene.npz"), i1=b1, i2=b2)
print("wrote pyrlk_scene.npz", b1.max(), b2.max())
b1.tofile(os.path.join(os.path.dirname(__file__), "pyrlk_i1_100x100.u8"))  # raw copies for the C++ test (tests/cpp/algo_tests.cu)
b2.tofile(os.path.join(os.path.dirname(__file__), "pyrlk_i2_100x100.u8"))
