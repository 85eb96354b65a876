"""ctypes binding of oracle/_ref: the UNMODIFIED reference sources compiled against oracle/cvshim.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
--impl reference legs).  The product package never imports this module.

oracle/_ref is built by `make -C oracle ref` where /root/reference exists (this container); on
the GPU box the prebuilt files that travelled with the snapshot are used.  `available()` says
whether a variant can be loaded.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = os.environ.get("PANO_REFERENCE_ROOT", "/root/reference")

MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])

VARIANTS = {"": "libpano_ref.so", "O0": "libpano_ref_O0.so", "omp": "libpano_ref_omp.so",
            "omp_O0": "libpano_ref_omp_O0.so"}


def build():
    """(Re)build oracle/_ref when the reference sources are present; otherwise keep what is there."""
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "serial", "main.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-j4", "ref", "REF=" + REFERENCE_ROOT],
                              stdout=subprocess.DEVNULL)


def available(variant=""):
    return os.path.exists(os.path.join(_REF_DIR, VARIANTS[variant]))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_STAGE_RE = re.compile(r"^(Harris Corner Detection|Harris Corner Matching|RANSAC Homography Estimation|"
                       r"Image Stitching|Total Stitching Process)(?: \(OpenMP\))?: ([0-9.]+) ms", re.M)


def parse_stage_lines(log):
    """The reference's own Timer lines (ref: src/serial/main.cpp:183,242,302,389,412) -> {stage: [ms...]}"""
    out = {}
    for name, ms in _STAGE_RE.findall(log):
        out.setdefault(name, []).append(float(ms))
    return out


class Reference:
    """variant: '' serial -O2, 'O0' serial at the reference's own CMake flags, 'omp' / 'omp_O0' the
    reference's OpenMP pipeline (src/openmp/main.cpp; different NMS tie rule, sampling and match
    order than serial: a timing baseline, not a parity target)."""

    def __init__(self, variant=""):
        if not available(variant):
            build()
        path = os.path.join(_REF_DIR, VARIANTS[variant])
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/%s is missing and %s is not present to build it"
                                    % (VARIANTS[variant], REFERENCE_ROOT))
        self.variant = variant
        self.lib = L = C.CDLL(path)
        for f in ("ref_detect", "ref_match", "ref_ransac", "ref_stitch_pair", "ref_stitch_all",
                  "ref_take_log", "ref_num_threads", "ref_is_openmp"):
            getattr(L, f).restype = C.c_int

    @staticmethod
    def _img(img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        assert img.ndim == 3 and img.shape[2] == 3
        return img

    def num_threads(self):
        return self.lib.ref_num_threads()

    def take_log(self):
        buf = C.create_string_buffer(1 << 16)
        self.lib.ref_take_log(buf, len(buf))
        return buf.value.decode()

    def gaussian_kernel(self, ksize=5, sigma=1.0):
        out = np.empty((ksize, ksize), np.float64)
        self.lib.ref_gaussian_kernel(ksize, C.c_double(sigma), _p(out, C.c_double))
        return out

    def convolve(self, plane, kern):
        plane = np.ascontiguousarray(plane, np.float64)
        kern = np.ascontiguousarray(kern, np.float64)
        out = np.empty_like(plane)
        self.lib.ref_convolve(_p(plane, C.c_double), plane.shape[1], plane.shape[0], _p(kern, C.c_double),
                              kern.shape[0], _p(out, C.c_double))
        return out

    def detect(self, img, k=0.04, thresh=1e6, nbhd=3):
        img = self._img(img)
        h, w, _ = img.shape
        xy = np.empty((w * h, 2), np.int32)
        n = self.lib.ref_detect(_p(img, C.c_uint8), w, h, C.c_size_t(img.strides[0]), C.c_double(k),
                                C.c_double(thresh), nbhd, _p(xy, C.c_int32), len(xy))
        return xy[:n].copy()

    def match(self, kq, kt, imq, imt, patch=5, max_ssd=1e8, offset=0):
        imq, imt = self._img(imq), self._img(imt)
        kq = np.ascontiguousarray(kq, np.int32).reshape(-1, 2)
        kt = np.ascontiguousarray(kt, np.int32).reshape(-1, 2)
        out = np.empty(max(len(kq), 1), MATCH_DTYPE)
        n = self.lib.ref_match(_p(kq, C.c_int32), len(kq), _p(kt, C.c_int32), len(kt),
                               _p(imq, C.c_uint8), imq.shape[1], imq.shape[0], C.c_size_t(imq.strides[0]),
                               _p(imt, C.c_uint8), imt.shape[1], imt.shape[0], C.c_size_t(imt.strides[0]),
                               patch, C.c_double(max_ssd), offset, out.ctypes.data_as(C.c_void_p), len(out))
        return out[:n]

    def ransac(self, kp1, kp2, matches, iters=1000, nsamples=4, thr=3.0, seed=12345):
        kp1 = np.ascontiguousarray(kp1, np.int32).reshape(-1, 2)
        kp2 = np.ascontiguousarray(kp2, np.int32).reshape(-1, 2)
        matches = np.ascontiguousarray(matches, MATCH_DTYPE)
        H = np.zeros((3, 3), np.float64)
        ok = self.lib.ref_ransac(_p(kp1, C.c_int32), len(kp1), _p(kp2, C.c_int32), len(kp2),
                                 matches.ctypes.data_as(C.c_void_p), len(matches), iters, nsamples,
                                 C.c_double(thr), C.c_uint(seed), _p(H, C.c_double))
        return H if ok else None

    def stitch_pair(self, left, right, seed=12345):
        """ref: stitchTwoImages.  dict(status, canvas, times_ms from the reference's own Timer lines)."""
        left, right = self._img(left), self._img(right)
        cap = 6 * 3 * (left.shape[0] * left.shape[1] + right.shape[0] * right.shape[1])
        buf = np.empty(cap, np.uint8)
        wh = np.zeros(2, np.int32)
        self.take_log()
        st = self.lib.ref_stitch_pair(_p(left, C.c_uint8), left.shape[1], left.shape[0], C.c_size_t(left.strides[0]),
                                      _p(right, C.c_uint8), right.shape[1], right.shape[0], C.c_size_t(right.strides[0]),
                                      C.c_uint(seed), _p(buf, C.c_uint8), C.c_size_t(cap), _p(wh, C.c_int32))
        log = self.take_log()
        canvas = buf[: int(wh[0]) * int(wh[1]) * 3].reshape(int(wh[1]), int(wh[0]), 3).copy() if st == 1 else None
        return dict(status=st, canvas=canvas, log=log, times_ms=parse_stage_lines(log))

    def stitch_all(self, images, seed=12345):
        """ref: stitchAllImages (left fold; a failed step keeps the previous panorama)."""
        images = [self._img(i) for i in images]
        n = len(images)
        ptrs = (C.POINTER(C.c_uint8) * n)(*[_p(i, C.c_uint8) for i in images])
        ws = (C.c_int * n)(*[i.shape[1] for i in images])
        hs = (C.c_int * n)(*[i.shape[0] for i in images])
        cap = 8 * 3 * sum(i.shape[0] * i.shape[1] for i in images)
        buf = np.empty(cap, np.uint8)
        wh = np.zeros(2, np.int32)
        self.take_log()
        st = self.lib.ref_stitch_all(ptrs, ws, hs, n, C.c_uint(seed), _p(buf, C.c_uint8), C.c_size_t(cap),
                                     _p(wh, C.c_int32))
        log = self.take_log()
        canvas = buf[: int(wh[0]) * int(wh[1]) * 3].reshape(int(wh[1]), int(wh[0]), 3).copy() if st == 1 else None
        return dict(status=st, canvas=canvas, log=log, times_ms=parse_stage_lines(log))
