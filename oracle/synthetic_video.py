"""Deterministic synthetic video for the tracker-loop tests (TEST INFRASTRUCTURE ONLY): a textured rectangle
drifting over a textured background, uint8 BGR frames like cv2.imread returns (tools/test.py:325)."""
import numpy as np


def make_frames(n=6, h=240, w=320, seed=0):
    rng = np.random.RandomState(seed)
    bg = (rng.rand(h // 8 + 1, w // 8 + 1, 3) * 255).astype(np.uint8)
    bg = np.kron(bg, np.ones((8, 8, 1), np.uint8))[:h, :w]
    obj = (rng.rand(8, 6, 3) * 255).astype(np.uint8)
    obj = np.kron(obj, np.ones((8, 8, 1), np.uint8))          # 64 x 48 object
    frames, boxes = [], []
    x, y = 120.0, 80.0
    for i in range(n):
        f = bg.copy()
        xi, yi = int(round(x)), int(round(y))
        f[yi:yi + obj.shape[0], xi:xi + obj.shape[1]] = obj
        frames.append(f)
        boxes.append((xi, yi, obj.shape[1], obj.shape[0]))
        x += 3.0
        y += 2.0
    return frames, boxes
