"""Generates tests/golden/area_resize.npz with OpenCV itself (run in the build container: `python -m oracle.make_golden_resize`).

Each case is a small random uint8 video [T, 3, H, W] and the result of the reference loader's scaling step on it,
    cv2.resize(frame.transpose(1, 2, 0), (int(W * f), int(H * f)), None, None, None, cv2.INTER_AREA)      (io/dataset.py:1473-1482),
for the scale factors of the reference's experiments (full_comparison.py:107-110, 124-125: GEM GoPro 0.1 after the row crop,
GEM front 0.3, DR(eye)VE GoPro 0.4, DR(eye)VE front 1/3) on reduced frame sizes that take the same code paths (integral /
non-integral scale per axis), plus 2 x 2 and odd sizes.  The oracle (oracle/area_resize.py) and the CUDA kernel are tested
against these files; the oracle is additionally compared with cv2 at the full frame sizes when cv2 is importable.
"""
import os

import cv2
import numpy as np

CASES = [  # (name, T, H, W, factor)
    ("gem_gopro_0.1", 2, 87, 1000, 0.1),         # x: integral scale 10, y: 87 -> 8 (10.875): general path, like 864 x 3840 -> 86 x 384
    ("gem_gopro_0.1_integral", 1, 90, 640, 0.1),  # both integral (10 x 10 box sums)
    ("gem_front_0.3", 2, 108, 110, 0.3),          # 1080 x 1088 -> 324 x 326 at a tenth of the size: scales 3.375 / 3.333
    ("dreyeve_gopro_0.4", 2, 108, 192, 0.4),      # scale 2.5 on both axes
    ("dreyeve_front_third", 2, 72, 96, 1 / 3.0),  # integral 3 x 3: cvRound(sum * (1.f / 9))
    ("half", 2, 60, 80, 0.5),                     # 2 x 2: (sum + 2) >> 2
    ("half_odd", 1, 61, 83, 0.5),                 # 61 -> 30, 83 -> 41: non-integral
    ("odd_0.37", 1, 97, 131, 0.37),
    ("near_one", 1, 50, 70, 0.9),
]


def main():
    rng = np.random.default_rng(20240607)
    out = {}
    for name, T, H, W, f in CASES:
        x = rng.integers(0, 256, size=(T, 3, H, W), dtype=np.uint8)
        size = (int(W * f), int(H * f))
        y = np.stack([cv2.resize(fr.transpose(1, 2, 0), size, None, None, None, cv2.INTER_AREA).transpose(2, 0, 1) for fr in x])
        out[name + "/x"], out[name + "/y"], out[name + "/factor"] = x, y, np.float64(f)
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "area_resize.npz")
    np.savez_compressed(path, cv2_version=np.array(cv2.__version__), **out)
    print("wrote", os.path.normpath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
