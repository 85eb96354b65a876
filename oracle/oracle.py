"""ctypes binding of the CPU oracle (oracle/pano_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])


def build(force=False):
    """Compile the oracle shared libraries with the committed Makefile."""
    if force or not os.path.exists(os.path.join(_HERE, "libpano_oracle.so")):
        subprocess.check_call(["make", "-C", _HERE, "-j3"], stdout=subprocess.DEVNULL)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """variant: '' (serial -O2), 'omp' (OpenMP matcher), 'O0' (the reference's default flags)."""

    def __init__(self, variant=""):
        build()
        name = "libpano_oracle%s.so" % ("_" + variant if variant else "")
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build(force=True)
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_detect.restype = C.c_int
        L.orc_match.restype = C.c_int
        L.orc_ransac.restype = C.c_int
        L.orc_find_homography4.restype = C.c_int
        L.orc_canvas_geometry.restype = C.c_int
        L.orc_compose.restype = C.c_int
        L.orc_stitch_pair.restype = C.c_int
        L.orc_invert33.restype = C.c_int
        L.orc_num_threads.restype = C.c_int

    # ---- stage functions -------------------------------------------------------------
    @staticmethod
    def _img(img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        assert img.ndim == 3 and img.shape[2] == 3
        return img

    def gray(self, img):
        img = self._img(img)
        h, w, _ = img.shape
        out = np.empty((h, w), np.uint8)
        self.lib.orc_gray(_p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]), _p(out, C.c_uint8))
        return out

    def gaussian_kernel(self, ksize=5, sigma=1.0):
        out = np.empty((ksize, ksize), np.float64)
        self.lib.orc_gaussian_kernel(ksize, C.c_double(sigma), _p(out, C.c_double))
        return out

    def convolve(self, plane, kern):
        plane = np.ascontiguousarray(plane, np.float64)
        kern = np.ascontiguousarray(kern, np.float64)
        h, w = plane.shape
        out = np.empty_like(plane)
        self.lib.orc_convolve(_p(plane, C.c_double), w, h, _p(kern, C.c_double), kern.shape[0],
                              _p(out, C.c_double))
        return out

    def harris_response(self, img, k=0.04):
        img = self._img(img)
        h, w, _ = img.shape
        out = np.empty((h, w), np.float64)
        self.lib.orc_harris_response(_p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]),
                                     C.c_double(k), _p(out, C.c_double))
        return out

    def detect(self, img, k=0.04, thresh=1e6, nbhd=3):
        img = self._img(img)
        h, w, _ = img.shape
        cap = w * h
        n = self.lib.orc_detect(_p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]), C.c_double(k),
                                C.c_double(thresh), nbhd, None, 0)
        xy = np.empty((max(n, 1), 2), np.int32)
        n2 = self.lib.orc_detect(_p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]), C.c_double(k),
                                 C.c_double(thresh), nbhd, _p(xy, C.c_int32), n)
        assert n2 == n and n <= cap
        return xy[:n]

    def match(self, kq, kt, imq, imt, patch=5, max_ssd=1e8, offset=0):
        imq, imt = self._img(imq), self._img(imt)
        kq = np.ascontiguousarray(kq, np.int32).reshape(-1, 2)
        kt = np.ascontiguousarray(kt, np.int32).reshape(-1, 2)
        out = np.empty(max(len(kq), 1), MATCH_DTYPE)
        n = self.lib.orc_match(_p(kq, C.c_int32), len(kq), _p(kt, C.c_int32), len(kt),
                               _p(imq, C.c_uint8), imq.shape[1], imq.shape[0], C.c_size_t(imq.strides[0]),
                               _p(imt, C.c_uint8), imt.shape[1], imt.shape[0], C.c_size_t(imt.strides[0]),
                               patch, C.c_double(max_ssd), offset, out.ctypes.data_as(C.c_void_p), len(out))
        return out[:n]

    def match_knn(self, kq, kt, imq, imt, patch=5, descriptor=0, ratio=0.75):
        """checker of the engine's opt-in pano_match_knn (2-NN + Lowe's ratio; NOT a reference function).
        Returns (matches, runner-up distances)."""
        imq, imt = self._img(imq), self._img(imt)
        kq = np.ascontiguousarray(kq, np.int32).reshape(-1, 2)
        kt = np.ascontiguousarray(kt, np.int32).reshape(-1, 2)
        out = np.empty(max(len(kq), 1), MATCH_DTYPE)
        second = np.empty(max(len(kq), 1), np.float32)
        self.lib.orc_match_knn.restype = C.c_int
        n = self.lib.orc_match_knn(_p(kq, C.c_int32), len(kq), _p(kt, C.c_int32), len(kt),
                                   _p(imq, C.c_uint8), imq.shape[1], imq.shape[0], C.c_size_t(imq.strides[0]),
                                   _p(imt, C.c_uint8), imt.shape[1], imt.shape[0], C.c_size_t(imt.strides[0]),
                                   patch, descriptor, C.c_double(ratio), out.ctypes.data_as(C.c_void_p),
                                   _p(second, C.c_float), len(out))
        return out[:n], second[:n]

    def knn_binary_descriptor(self, img, x, y):
        img = self._img(img)
        bits = np.zeros(8, np.uint32)
        self.lib.orc_knn_binary_descriptor(_p(img, C.c_uint8), C.c_size_t(img.strides[0]), int(x), int(y), _p(bits, C.c_uint32))
        return bits

    def find_homography4(self, src, dst):
        src = np.ascontiguousarray(src, np.float32).reshape(4, 2)
        dst = np.ascontiguousarray(dst, np.float32).reshape(4, 2)
        H = np.empty((3, 3), np.float64)
        ok = self.lib.orc_find_homography4(_p(src, C.c_float), _p(dst, C.c_float), _p(H, C.c_double))
        return H if ok else None

    def ransac(self, kp1, kp2, matches, iters=1000, nsamples=4, thr=3.0, seed=12345, div_mode=0):
        kp1 = np.ascontiguousarray(kp1, np.int32).reshape(-1, 2)
        kp2 = np.ascontiguousarray(kp2, np.int32).reshape(-1, 2)
        matches = np.ascontiguousarray(matches, MATCH_DTYPE)
        m = len(matches)
        H = np.zeros((3, 3), np.float64)
        best = C.c_int(0)
        best_iter = C.c_int(-1)
        draws = C.c_uint64(0)
        samples = np.full((iters, 4), -1, np.int32)
        counts = np.full(iters, -2, np.int32)
        mask = np.zeros(max(m, 1), np.uint8)
        ok = self.lib.orc_ransac(_p(kp1, C.c_int32), _p(kp2, C.c_int32), matches.ctypes.data_as(C.c_void_p),
                                 m, iters, nsamples, C.c_double(thr), C.c_uint32(seed), div_mode,
                                 _p(H, C.c_double), C.byref(best), _p(samples, C.c_int32),
                                 _p(counts, C.c_int32), _p(mask, C.c_uint8), C.byref(draws),
                                 C.byref(best_iter))
        return dict(ok=bool(ok), H=H if ok else None, best_count=best.value, best_iter=best_iter.value,
                    samples=samples, counts=counts, inlier_mask=mask[:m].astype(bool), draws=draws.value)

    def mt19937(self, seed, n):
        out = np.empty(n, np.uint32)
        self.lib.orc_mt19937(C.c_uint32(seed), n, _p(out, C.c_uint32))
        return out

    def shuffle_iota(self, seed, n, reps=1, skip=0):
        out = np.empty((reps, 4), np.int32)
        self.lib.orc_shuffle_iota(C.c_uint32(seed), C.c_uint64(skip), n, reps, _p(out, C.c_int32))
        return out

    def perspective_transform(self, pts, H):
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        H = np.ascontiguousarray(H, np.float64)
        out = np.empty_like(pts)
        self.lib.orc_perspective_transform(_p(pts, C.c_float), len(pts), _p(H, C.c_double), _p(out, C.c_float))
        return out

    def invert33(self, M):
        M = np.ascontiguousarray(M, np.float64)
        out = np.empty((3, 3), np.float64)
        ok = self.lib.orc_invert33(_p(M, C.c_double), _p(out, C.c_double))
        return out if ok else None

    def canvas_geometry(self, wl, hl, wr, hr, H):
        H = np.ascontiguousarray(H, np.float64)
        geom = np.zeros(4, np.int32)
        TH = np.empty((3, 3), np.float64)
        ok = self.lib.orc_canvas_geometry(wl, hl, wr, hr, _p(H, C.c_double), _p(geom, C.c_int32),
                                          _p(TH, C.c_double))
        return bool(ok), tuple(int(v) for v in geom), TH

    def warp_perspective(self, src, M, dsize):
        src = self._img(src)
        M = np.ascontiguousarray(M, np.float64)
        dw, dh = dsize
        dst = np.empty((dh, dw, 3), np.uint8)
        self.lib.orc_warp_perspective(_p(src, C.c_uint8), src.shape[1], src.shape[0],
                                      C.c_size_t(src.strides[0]), _p(M, C.c_double), _p(dst, C.c_uint8),
                                      dw, dh, C.c_size_t(dst.strides[0]))
        return dst

    def compose(self, left, right, H):
        left, right = self._img(left), self._img(right)
        H = np.ascontiguousarray(H, np.float64)
        ok, (cw, ch, ox, oy), _ = self.canvas_geometry(left.shape[1], left.shape[0], right.shape[1],
                                                       right.shape[0], H)
        if not ok:
            return None
        canvas = np.empty((ch, cw, 3), np.uint8)
        geom = np.zeros(4, np.int32)
        r = self.lib.orc_compose(_p(left, C.c_uint8), left.shape[1], left.shape[0], C.c_size_t(left.strides[0]),
                                 _p(right, C.c_uint8), right.shape[1], right.shape[0], C.c_size_t(right.strides[0]),
                                 _p(H, C.c_double), _p(canvas, C.c_uint8), C.c_size_t(canvas.nbytes),
                                 _p(geom, C.c_int32))
        assert r == 1
        return canvas

    def stitch_pair(self, left, right, seed=12345, max_canvas_px=None):
        """ref: src/serial/main.cpp:311-391.  Returns dict(status, canvas, H, stats, times_ms)."""
        left, right = self._img(left), self._img(right)
        cap_px = max_canvas_px or 6 * (left.shape[0] * left.shape[1] + right.shape[0] * right.shape[1])
        buf = np.empty(cap_px * 3, np.uint8)
        geom = np.zeros(4, np.int32)
        H = np.zeros((3, 3), np.float64)
        stats = np.zeros(4, np.int32)
        times = np.zeros(4, np.float64)
        st = self.lib.orc_stitch_pair(_p(left, C.c_uint8), left.shape[1], left.shape[0], C.c_size_t(left.strides[0]),
                                      _p(right, C.c_uint8), right.shape[1], right.shape[0], C.c_size_t(right.strides[0]),
                                      C.c_uint32(seed), _p(buf, C.c_uint8), C.c_size_t(buf.nbytes),
                                      _p(geom, C.c_int32), _p(H, C.c_double), _p(stats, C.c_int32),
                                      _p(times, C.c_double))
        canvas = None
        if st == 1:
            cw, ch = int(geom[0]), int(geom[1])
            canvas = buf[: cw * ch * 3].reshape(ch, cw, 3).copy()
        return dict(status=st, canvas=canvas, H=H, geom=tuple(int(v) for v in geom),
                    stats=dict(kl=int(stats[0]), kr=int(stats[1]), m=int(stats[2]), best=int(stats[3])),
                    times_ms=dict(detect=times[0], match=times[1], ransac=times[2], warp=times[3]))

    def stitch_fold(self, images, seed=12345):
        """ref: src/serial/main.cpp:395-414 stitchAllImages (left fold; failed step skipped)."""
        pano = self._img(images[0])
        log = []
        for im in images[1:]:
            r = self.stitch_pair(pano, im, seed=seed)
            log.append(r)
            if r["status"] == 1:
                pano = r["canvas"]
        return pano, log

    # ---- chain mode (the engine's shardable multi-image mode) ---------------
    def mul33(self, A, B):
        A = np.ascontiguousarray(A, np.float64); B = np.ascontiguousarray(B, np.float64)
        out = np.empty((3, 3), np.float64)
        self.lib.orc_mul33(_p(A, C.c_double), _p(B, C.c_double), _p(out, C.c_double))
        return out

    def pair_homography(self, left, right, seed=12345):
        """steps 1-3 of stitchTwoImages (ref: src/serial/main.cpp:316-332): H right -> left or None"""
        kl, kr = self.detect(left), self.detect(right)
        m = self.match(kr, kl, right, left)
        if len(m) == 0:
            return None
        r = self.ransac(kr, kl, m, seed=seed)
        return r["H"] if r["ok"] else None

    def chain_geometry(self, sizes, Hs):
        f = np.float32
        minX, minY, maxX, maxY = f(0), f(0), f(sizes[0][0]), f(sizes[0][1])
        for (w, h), H in list(zip(sizes, Hs))[1:]:
            pts = self.perspective_transform(np.float32([[0, 0], [w, 0], [w, h], [0, h]]), H)
            for x, y in pts:
                minX, minY, maxX, maxY = min(minX, f(x)), min(minY, f(y)), max(maxX, f(x)), max(maxY, f(y))
        T = np.array([[1, 0, float(-minX)], [0, 1, float(-minY)], [0, 0, 1.0]])
        cw, ch = int(np.ceil(f(maxX - minX))), int(np.ceil(f(maxY - minY)))
        return (cw, ch, int(-minX), int(-minY)), T

    def stitch_chain(self, images, seed=12345):
        """H(i <- i+1) per adjacent pair, composed into image 0's frame, image 0 placed at its integer
        offset, every other image warped by T*H and overlaid with the non-black rule."""
        images = [self._img(i) for i in images]
        pair_H = [self.pair_homography(images[i], images[i + 1], seed) for i in range(len(images) - 1)]
        Hs = [np.eye(3)]
        for H in pair_H:
            if H is None:
                break
            Hs.append(self.mul33(Hs[-1], H))
        sizes = [(im.shape[1], im.shape[0]) for im in images]
        (cw, ch, x0, y0), T = self.chain_geometry(sizes, Hs)
        canvas = np.zeros((ch, cw, 3), np.uint8)
        canvas[y0:y0 + images[0].shape[0], x0:x0 + images[0].shape[1]] = images[0]
        for im, H in list(zip(images, Hs))[1:]:
            w = self.warp_perspective(im, self.mul33(T, H), (cw, ch))
            nz = w.any(axis=2)
            canvas[nz] = w[nz]
        return canvas, pair_H

    def num_threads(self):
        return self.lib.orc_num_threads()
