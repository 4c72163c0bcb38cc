"""Plain-numpy restatement of cv2.resize(src, dsize) with INTER_LINEAR on 8-bit images — OpenCV's fixed-point
scheme (modules/imgproc/src/resize.cpp, third-party dependency of tools/test.py:105; the reference pins
opencv_python==3.4.3.18, this image has 4.13 — the 8-bit linear path is unchanged).  TEST INFRASTRUCTURE ONLY: it
documents the arithmetic the CUDA crop kernel reproduces and is itself checked bit-for-bit against cv2 in
tests/test_crop.py."""
import numpy as np


def resize_linear_u8(src, dsize):
    """OpenCV's 8-bit INTER_LINEAR resize in fixed point (resize.cpp: HResizeLinear + VResizeLinear, 11-bit coefficients)."""
    H, W, C = src.shape
    dw, dh = dsize
    def coeffs(dst_n, src_n, clamp_frac=True):
        scale = 1.0 / (dst_n / src_n)
        ofs = np.zeros(dst_n, np.int64); a = np.zeros((dst_n, 2), np.int64)
        for d in range(dst_n):
            f = np.float32((d + 0.5) * scale - 0.5)     # float fx
            s = int(np.floor(f)); f = np.float32(f - s)
            if clamp_frac:
                if s < 0: f = np.float32(0); s = 0
                if s >= src_n - 1: f = np.float32(0); s = src_n - 1
            ofs[d] = s
            a[d, 0] = int(np.rint(np.float32(1.0 - f) * np.float32(2048)))   # saturate_cast<short> = cvRound
            a[d, 1] = int(np.rint(f * np.float32(2048)))
        return ofs, a
    xo, xa = coeffs(dw, W); yo, ya = coeffs(dh, H, clamp_frac=False)   # rows are clipped instead (resize.cpp)
    s = src.astype(np.int64)
    x1 = np.minimum(xo + 1, W - 1)
    hrow = s[:, xo, :] * xa[:, 0][None, :, None] + s[:, x1, :] * xa[:, 1][None, :, None]   # [H, dw, C] ints
    y0 = np.clip(yo, 0, H - 1); y1 = np.clip(yo + 1, 0, H - 1)
    S0 = hrow[y0]; S1 = hrow[y1]
    b0 = ya[:, 0][:, None, None]; b1 = ya[:, 1][:, None, None]
    out = ((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2)
    return np.clip(out, 0, 255).astype(np.uint8)
