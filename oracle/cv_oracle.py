"""TEST INFRASTRUCTURE ONLY — integer-exact numpy restatements of the OpenCV arithmetic the
anonymisation path relies on.  Only tests/, smoke() and bench.py's CPU-baseline legs import this.

The arithmetic lives in a third-party dependency whose source is not under /root/reference:
opencv-python (pin 4.9.0.80, reference requirements.txt:109; 4.13.0 in this image).  Call sites in
the reference: cv2.resize INTER_LINEAR at detect_face/utils/utils.py:15, cv2.resize INTER_NEAREST at
combine_detect.py:157-158.  cv2.GaussianBlur has NO call site in the reference; BASELINE.json's
north_star defines it as the oracle for the Gaussian mode.  Parity is "unpinned by the reference"
(it has no tests); it is pinned here by tests/test_oracle_golden.py (test_cv_live_against_this_image), which checks every function below
bit-for-bit against the cv2 build in this image, and by tests/golden/cv_golden.npz.
"""
import math
import numpy as np

# ----------------------------------------------------------------------------------------------
# Gaussian taps: cv::getGaussianKernel (float64) then 8.8 fixed point with error diffusion
# (OpenCV smooth: getGaussianKernelFixedPoint_ED); SURVEY §8c.
# ----------------------------------------------------------------------------------------------
_SMALL = {
    1: [1.0],
    3: [0.25, 0.5, 0.25],
    5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
    7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
}


def gaussian_kernel_f64(ksize, sigma):
    if ksize % 2 == 1 and ksize <= 7 and sigma <= 0:
        return np.array(_SMALL[ksize], np.float64)
    sigma_x = sigma if sigma > 0 else ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8
    scale2x = -0.5 / (sigma_x * sigma_x)
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(scale2x * x * x)
    return k / k.sum()


def gaussian_taps_q8(ksize, sigma):
    """int taps summing to 256 (half-to-even rounding with error diffusion from the outside in)."""
    g = gaussian_kernel_f64(ksize, sigma)
    q = np.zeros(ksize, np.int64)
    err = 0.0
    half = ksize // 2
    s = 0
    for i in range(half):
        adj = g[i] * 256.0 + err
        qi = int(np.rint(adj))          # rint = half-to-even, as cvRound
        err = adj - qi
        q[i] = q[ksize - 1 - i] = qi
        s += 2 * qi
    q[half] = 256 - s
    return q


def _reflect101(p, n):
    if n == 1:
        return 0
    while p < 0 or p >= n:
        p = -p if p < 0 else 2 * (n - 1) - p
    return p


def _reflect_index(n, r):
    return np.array([_reflect101(i, n) for i in range(-r, n + r)], np.int64)


def gaussian_blur_u8(roi, ksize, sigma=0.0):
    """== cv2.GaussianBlur(roi, (ksize,ksize), sigma) for uint8 HxWxC with BORDER_REFLECT_101
    applied inside the ROI (a numpy view carries no parent-matrix information)."""
    roi = np.asarray(roi)
    assert roi.dtype == np.uint8
    if roi.ndim == 2:
        return gaussian_blur_u8(roi[:, :, None], ksize, sigma)[:, :, 0]
    h, w, _ = roi.shape
    q = gaussian_taps_q8(ksize, sigma)
    r = ksize // 2
    src = roi.astype(np.int64)
    xi = _reflect_index(w, r)
    hbuf = np.zeros_like(src)
    for t in range(ksize):
        hbuf += q[t] * src[:, xi[t:t + w], :]
    hbuf = np.minimum(hbuf, 65535)                  # u16 saturate (never reached: taps sum to 256)
    yi = _reflect_index(h, r)
    v = np.zeros_like(src)
    for t in range(ksize):
        v += q[t] * hbuf[yi[t:t + h], :, :]
    return ((v + 32768) >> 16).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# INTER_NEAREST (cv2.resize): src = min(floor(dst * (1.0 / (dst_size/src_size))), src_size-1), float64
# ----------------------------------------------------------------------------------------------
def nearest_map(dst_n, src_n):
    inv = 1.0 / (float(dst_n) / float(src_n))
    return np.minimum(np.floor(np.arange(dst_n, dtype=np.float64) * inv).astype(np.int64), src_n - 1)


def mosaic_box(img, x1, y1, x2, y2, mosaic_level=8):
    """reference combine_detect.py:138-161 — returns a new array."""
    img = img.copy()
    h, w = img.shape[:2]
    x1, y1, x2, y2 = max(0, x1), max(0, y1), min(w, x2), min(h, y2)
    if x2 <= x1 or y2 <= y1:
        return img
    bw, bh = x2 - x1, y2 - y1
    sw, sh = max(1, bw // mosaic_level), max(1, bh // mosaic_level)
    xs = nearest_map(sw, bw)[nearest_map(bw, sw)]    # down map composed with up map
    ys = nearest_map(sh, bh)[nearest_map(bh, sh)]
    area = img[y1:y2, x1:x2].copy()
    img[y1:y2, x1:x2] = area[ys][:, xs]
    return img


def gauss_box(img, x1, y1, x2, y2, ksize=31, sigma=0.0):
    """the a16 idiom of SURVEY §8a: img[y1:y2,x1:x2] = GaussianBlur(img[y1:y2,x1:x2]) with the same
    clip rules as the mosaic (combine_detect.py:145-150).  Returns a new array."""
    img = img.copy()
    h, w = img.shape[:2]
    x1, y1, x2, y2 = max(0, x1), max(0, y1), min(w, x2), min(h, y2)
    if x2 <= x1 or y2 <= y1:
        return img
    img[y1:y2, x1:x2] = gaussian_blur_u8(img[y1:y2, x1:x2], ksize, sigma)
    return img


def composite(img, boxes, mode="mosaic", mosaic_level=8, ksize=31, sigma=0.0):
    """sequential application, each box on the result of the previous one
    (combine_detect.py:242-249); boxes are int()-truncated toward zero first (:243)."""
    out = img.copy()
    for b in boxes:
        x1, y1, x2, y2 = (int(v) for v in b)
        if mode == "mosaic":
            out = mosaic_box(out, x1, y1, x2, y2, mosaic_level)
        else:
            out = gauss_box(out, x1, y1, x2, y2, ksize, sigma)
    return out


# ----------------------------------------------------------------------------------------------
# INTER_LINEAR uint8 (cv2.resize, down-scale / identity) and the reference letterbox
# (detect_face/utils/utils.py:8-18)
# ----------------------------------------------------------------------------------------------
def _linear_coeffs(dst_n, src_n, clamp=True):
    """cv2 resize.cpp coefficient tables.  Along x (clamp=True) a tap that starts left of the row or on its last pixel is
    replaced by that pixel with weight 1; along y (clamp=False) the weights are KEPT and the two row indices are clipped
    instead — when up-scaling, the first / last output rows therefore blend the border row with itself, and the two
    separately truncated products can come out one below the single full-weight product."""
    scale = np.float64(src_n) / np.float64(dst_n)
    idx = np.zeros(dst_n, np.int64)
    a0 = np.zeros(dst_n, np.int64)
    a1 = np.zeros(dst_n, np.int64)
    for d in range(dst_n):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(math.floor(fx))
        fx = np.float32(fx - sx)
        if clamp:
            if sx < 0:
                fx, sx = np.float32(0), 0
            if sx >= src_n - 1:
                fx, sx = np.float32(0), src_n - 1
        idx[d] = sx
        # cv: saturate_cast<short>(v * INTER_RESIZE_COEF_SCALE) with cvRound (half-to-even)
        a0[d] = int(np.rint(np.float32(np.float32(1.0) - fx) * np.float32(2048)))
        a1[d] = int(np.rint(np.float32(fx) * np.float32(2048)))
    return idx, a0, a1


def resize_linear_u8(img, dst_w, dst_h):
    """== cv2.resize(img, (dst_w,dst_h)) (INTER_LINEAR) for uint8 HxWx3, down- and up-scaling.  An exact 2x reduction in
    both directions is what OpenCV turns into its fast INTER_AREA path ((a+b+c+d+2)>>2), which the bilinear arithmetic
    below reproduces bit for bit (both taps 1024: no truncation loss)."""
    h, w = img.shape[:2]
    if (dst_w, dst_h) == (w, h):
        return img.copy()
    xi, xa0, xa1 = _linear_coeffs(dst_w, w)
    yi, ya0, ya1 = _linear_coeffs(dst_h, h, clamp=False)
    s = img.astype(np.int64)
    x1 = np.minimum(xi + 1, w - 1)
    rows = s[:, xi, :] * xa0[None, :, None] + s[:, x1, :] * xa1[None, :, None]     # (h,dst_w,3), x2048
    y0 = np.clip(yi, 0, h - 1)
    y1 = np.clip(yi + 1, 0, h - 1)
    r0, r1 = rows[y0], rows[y1]
    b0, b1 = ya0[:, None, None], ya1[:, None, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox(img, W, H):
    """detect_face/utils/utils.py:8-18 — returns uint8 canvas (the reference makes it float64 = 128,
    values are identical integers)."""
    ih, iw = img.shape[:2]
    scale = min(W / iw, H / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    r = resize_linear_u8(img, nw, nh)
    canvas = np.full((H, W, 3), 128, np.uint8)
    canvas[(H - nh) // 2:nh + (H - nh) // 2, (W - nw) // 2:nw + (W - nw) // 2] = r
    return canvas
