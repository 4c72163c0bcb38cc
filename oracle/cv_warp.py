"""Plain-numpy restatement of cv2.warpAffine(src f32, M, dsize, INTER_LINEAR, BORDER_CONSTANT, borderValue) as used by
crop_back() in siamese_track (tools/test.py:263-275) — OpenCV imgwarp.cpp: forward map inverted in double, fixed-point
source coordinates (AB_BITS 10, 1/32-pixel sub-positions), float bilinear table.  TEST INFRASTRUCTURE ONLY; checked
bit for bit against cv2 in tests/test_crop.py."""
import numpy as np


def warp_affine_f32(src, M, dsize, border=-1.0):
    """cv2.warpAffine(src f32 1-channel, M, dsize, INTER_LINEAR, BORDER_CONSTANT, borderValue) — imgwarp.cpp"""
    W, H = dsize
    M = np.asarray(M, dtype=np.float64).copy().reshape(6)
    D = M[0]*M[4] - M[1]*M[3]
    D = 1.0/D if D != 0 else 0.0
    A11, A22 = M[4]*D, M[0]*D
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22
    b1 = -M[0]*M[2] - M[1]*M[5]; b2 = -M[3]*M[2] - M[4]*M[5]
    M[2] = b1; M[5] = b2
    AB_BITS, INTER_BITS = 10, 5
    AB_SCALE = 1 << AB_BITS
    round_delta = AB_SCALE // 32 // 2
    x = np.arange(W)
    adelta = np.rint(M[0]*x*AB_SCALE).astype(np.int64)
    bdelta = np.rint(M[3]*x*AB_SCALE).astype(np.int64)
    t = np.arange(32, dtype=np.float32) / np.float32(32)
    tab1 = np.stack([np.float32(1) - t, t], 1)           # [32][2]
    sh, sw = src.shape
    out = np.empty((H, W), np.float32)
    bv = np.float32(border)
    for y in range(H):
        X0 = int(np.rint((M[1]*y + M[2])*AB_SCALE)) + round_delta
        Y0 = int(np.rint((M[4]*y + M[5])*AB_SCALE)) + round_delta
        X = (X0 + adelta) >> (AB_BITS - INTER_BITS)
        Y = (Y0 + bdelta) >> (AB_BITS - INTER_BITS)
        sx = np.clip(X >> INTER_BITS, -32768, 32767); sy = np.clip(Y >> INTER_BITS, -32768, 32767)
        fx = X & 31; fy = Y & 31
        w00 = tab1[fy, 0]*tab1[fx, 0]; w01 = tab1[fy, 0]*tab1[fx, 1]; w10 = tab1[fy, 1]*tab1[fx, 0]; w11 = tab1[fy, 1]*tab1[fx, 1]
        def px(yy, xx):
            ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
            return np.where(ok, src[np.clip(yy, 0, sh-1), np.clip(xx, 0, sw-1)], bv)
        v = px(sy, sx)*w00 + px(sy, sx+1)*w01 + px(sy+1, sx)*w10 + px(sy+1, sx+1)*w11
        out[y] = v.astype(np.float32)
    return out
