"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the ROI front-end that precedes the regressor:
crop the 2D-detector box out of a uint8 frame, resize to the network input, BGR->RGB, normalise, NCHW float32.

Reference call sites: `Regressor.crop` + `IEModel._preprocess` (torchdet3d/utils/ie_wrappers.py:155-158,18-21:
`frame[y0:y1, x0:x1]` then `cv.resize(img, (w, h))`, i.e. OpenCV INTER_LINEAR on uint8), BGR->RGB
(`ConvertColor`, utils/transforms.py:10-17), normalisation constants `configs/default_config.py:9-10` applied as
albumentations.Normalize does ((img - mean*255) * (1 / (std*255)) in float32), `transpose(2, 0, 1)`.

`reference(...)` calls OpenCV itself (cv2 is in the image); `resize_bilinear_u8(...)` restates OpenCV's fixed-point
uint8 bilinear rule (11-bit coefficients, the `(((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2` vertical pass of
imgproc/resize.cpp) so the CUDA kernel can be bit-exact; tests pin the restatement to cv2 bit for bit."""
import numpy as np

MEAN = (0.5931, 0.4690, 0.4229)      # configs/default_config.py:9
STD = (0.2471, 0.2214, 0.2157)       # configs/default_config.py:10
COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def norm_constants(mean=MEAN, std=STD):
    """(mean*255, 1/(std*255)) as float32, the two arrays albumentations.Normalize applies."""
    m = (np.asarray(mean, dtype=np.float32) * np.float32(255.0)).astype(np.float32)
    inv = (np.float32(1.0) / (np.asarray(std, dtype=np.float32) * np.float32(255.0))).astype(np.float32)
    return m, inv


def _axis_tables(src, dst):
    """OpenCV's per-axis source index and 11-bit weights for INTER_LINEAR: (i0[dst], i1[dst], w0[dst], w1[dst])."""
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def resize_bilinear_u8(img, out_h, out_w):
    """cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR) for uint8 HxWxC, restated."""
    h, w = img.shape[:2]
    sx, fx = _axis_tables(w, out_w)
    lo, hi = sx < 0, sx >= w - 1
    fx = np.where(lo | hi, np.float32(0), fx)
    sx = np.where(lo, 0, np.where(hi, w - 1, sx))
    sx1 = np.minimum(sx + 1, w - 1)
    a1 = np.rint(fx * np.float32(COEF_SCALE)).astype(np.int64)
    a0 = np.rint((np.float32(1.0) - fx) * np.float32(COEF_SCALE)).astype(np.int64)
    sy, fy = _axis_tables(h, out_h)
    b1 = np.rint(fy * np.float32(COEF_SCALE)).astype(np.int64)
    b0 = np.rint((np.float32(1.0) - fy) * np.float32(COEF_SCALE)).astype(np.int64)
    y0 = np.clip(sy, 0, h - 1)
    y1 = np.clip(sy + 1, 0, h - 1)
    src = img.astype(np.int64)
    rows = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]          # horizontal pass, scaled by 2^11
    r0, r1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def crop_resize_normalize(frame, boxes, out_h=224, out_w=224, mean=MEAN, std=STD, bgr=True):
    """Restated pipeline: frame uint8 [H,W,3], boxes int [N,4] (x0,y0,x1,y1) -> float32 [N,3,out_h,out_w]."""
    m, inv = norm_constants(mean, std)
    out = np.empty((len(boxes), 3, out_h, out_w), dtype=np.float32)
    for n, (x0, y0, x1, y1) in enumerate(np.asarray(boxes).tolist()):
        r = resize_bilinear_u8(frame[y0:y1, x0:x1], out_h, out_w)
        if bgr:
            r = r[:, :, ::-1]
        v = (r.astype(np.float32) - m) * inv
        out[n] = v.transpose(2, 0, 1)
    return out


def reference(frame, boxes, out_h=224, out_w=224, mean=MEAN, std=STD, bgr=True):
    """The same through OpenCV itself (ie_wrappers.py:155-158,18-21 + transforms.py:16-17 + Normalize)."""
    import cv2
    m, inv = norm_constants(mean, std)
    out = np.empty((len(boxes), 3, out_h, out_w), dtype=np.float32)
    for n, (x0, y0, x1, y1) in enumerate(np.asarray(boxes).tolist()):
        r = cv2.resize(np.ascontiguousarray(frame[y0:y1, x0:x1]), (out_w, out_h))
        if bgr:
            r = cv2.cvtColor(r, cv2.COLOR_BGR2RGB)
        out[n] = ((r.astype(np.float32) - m) * inv).transpose(2, 0, 1)
    return out
