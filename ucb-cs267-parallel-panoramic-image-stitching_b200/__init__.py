"""pano_b200 — host-side mirror of the reference's GPU stage interface over the C ABI.

The product is ``libpano_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI in
``include/pano_b200.h``).  This module is the thin Python binding used by tests and bench.py;
it mirrors the names and argument meaning of the reference's stage entry points:

    gpuHarrisCornerDetectorDetect   ref: src/gpu/harris_detector.cuh:5-9
    gpuHarrisMatchKeyPoints         ref: src/gpu/harris_matcher.cuh:5-9
    GpuRansacHomographyCalculator   ref: src/gpu/ransac.cuh:8-36
    convolveCUDA                    ref: src/gpu/convolution.cuh:5
    stitchTwoImages / stitchAllImages   ref: src/gpu/main.cpp:322-449

with the SERIAL path's semantics (ref: src/serial/main.cpp).  There is no CPU fallback: if the
library is missing or no sm_100 GPU is present, construction raises.

Images are ``numpy`` uint8 arrays (H, W, 3) BGR on the host, or CUDA ``torch`` uint8 tensors of
the same shape (then nothing bulk crosses PCIe).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpano_b200.so")

PANO_OK = 0
PANO_ERR_INVALID, PANO_ERR_CUDA, PANO_ERR_NO_MATCHES, PANO_ERR_TOO_FEW_MATCHES = 1, 2, 3, 4
PANO_ERR_NO_HOMOGRAPHY, PANO_ERR_ROI, PANO_ERR_CAPACITY, PANO_ERR_UNSUPPORTED, PANO_ERR_NO_DEVICE = 5, 6, 7, 8, 9
PANO_ERR_BUSY = 10
STATUS_NAMES = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "NO_MATCHES", 4: "TOO_FEW_MATCHES", 5: "NO_HOMOGRAPHY",
                6: "ROI", 7: "CAPACITY", 8: "UNSUPPORTED", 9: "NO_DEVICE", 10: "BUSY"}
PROFILE_CLASSES = ["harris_fused", "scan_scatter", "descriptor_gather", "match_tc", "emit", "replay", "dlt", "score",
                   "warp"]
MEM_HOST, MEM_DEVICE = 0, 1

MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])

# every symbol include/pano_b200.h declares
EXPORTED_SYMBOLS = [
    "pano_default_harris_opts", "pano_default_ransac_opts", "pano_create", "pano_destroy", "pano_set_seed",
    "pano_last_error", "pano_version", "pano_kernel_launches", "pano_set_matcher", "pano_detect",
    "pano_harris_response", "pano_convolve_f64", "pano_match", "pano_ransac", "pano_canvas_geometry",
    "pano_warp_overlay", "pano_warp_perspective", "pano_stitch_pair", "pano_get_canvas", "pano_canvas_device",
    "pano_stitch_fold", "pano_stitch_batch", "pano_stream", "pano_pair_homography", "pano_mul33",
    "pano_chain_geometry", "pano_warp_accumulate", "pano_set_stream", "pano_set_replay_mode",
    "pano_set_fold_mode", "pano_stitch_pair_async", "pano_pair_query", "pano_pair_wait", "pano_set_profile", "pano_get_profile",
    "pano_default_knn_opts", "pano_match_knn",
    "pano_detect_async", "pano_match_async", "pano_ransac_async", "pano_warp_overlay_async", "pano_stitch_fold_async",
    "pano_stitch_batch_async", "pano_set_match_mode", "pano_replay_work_estimate",
]


KNN_PATCH_SSD, KNN_BINARY = 0, 1


class ReplayWork(C.Structure):
    """pano_replay_work (measurement aid: what the chunked shuffle replay plans for one RANSAC run; include/pano_b200.h)"""
    _fields_ = [("chunk_iterations", C.c_int32), ("chunks", C.c_int32), ("steps", C.c_uint32),
                ("candidates_per_chunk", C.c_uint32), ("diagonals_per_chunk", C.c_uint32), ("reserved", C.c_uint32),
                ("rejections_mean", C.c_double), ("rejections_sigma", C.c_double), ("cells", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class KnnOptions(C.Structure):
    """pano_knn_opts (opt-in 2-NN / Lowe-ratio matcher; include/pano_b200.h)"""
    _fields_ = [("patchSize_", C.c_int), ("descriptor_", C.c_int), ("ratio_", C.c_double)]

    def __init__(self, patchSize_=5, descriptor_=0, ratio_=0.75):
        super().__init__(patchSize_, descriptor_, ratio_)


class HarrisCornerOptions(C.Structure):
    """ref: src/serial/main.cpp:28-34"""
    _fields_ = [("k_", C.c_double), ("nmsThresh_", C.c_double), ("nmsNeighborhood_", C.c_int),
                ("patchSize_", C.c_int), ("maxSSDThresh_", C.c_double)]

    def __init__(self, k_=0.04, nmsThresh_=1e6, nmsNeighborhood_=3, patchSize_=5, maxSSDThresh_=1e8):
        super().__init__(k_, nmsThresh_, nmsNeighborhood_, patchSize_, maxSSDThresh_)


class RansacOptions(C.Structure):
    """ref: src/serial/main.cpp:36-40, src/gpu/ransac.cuh:10-14"""
    _fields_ = [("numIterations_", C.c_int), ("numSamples_", C.c_int), ("distanceThreshold_", C.c_double)]

    def __init__(self, numIterations_=1000, numSamples_=4, distanceThreshold_=3.0):
        super().__init__(numIterations_, numSamples_, distanceThreshold_)


class CanvasInfo(C.Structure):
    _fields_ = [("canvas_w", C.c_int), ("canvas_h", C.c_int), ("left_x", C.c_int), ("left_y", C.c_int),
                ("TH", C.c_double * 9)]


class PairResult(C.Structure):
    _fields_ = [("status", C.c_int), ("n_kp_left", C.c_int), ("n_kp_right", C.c_int), ("n_matches", C.c_int),
                ("best_inliers", C.c_int), ("best_iteration", C.c_int), ("H", C.c_double * 9),
                ("canvas", CanvasInfo), ("ms_detect", C.c_float), ("ms_match", C.c_float),
                ("ms_ransac", C.c_float), ("ms_warp", C.c_float), ("ms_total", C.c_float)]

    def as_dict(self):
        return dict(status=self.status, status_name=STATUS_NAMES.get(self.status, str(self.status)),
                    kl=self.n_kp_left, kr=self.n_kp_right, m=self.n_matches, best=self.best_inliers,
                    best_iter=self.best_iteration, H=np.array(self.H[:], np.float64).reshape(3, 3),
                    canvas=(self.canvas.canvas_w, self.canvas.canvas_h, self.canvas.left_x, self.canvas.left_y),
                    ms=dict(detect=self.ms_detect, match=self.ms_match, ransac=self.ms_ransac,
                            warp=self.ms_warp, total=self.ms_total))


# numpy view of an array of PairResult (same layout as the C struct): vectorised access to a batch's results
PAIR_DTYPE = np.dtype([("status", "<i4"), ("n_kp_left", "<i4"), ("n_kp_right", "<i4"), ("n_matches", "<i4"),
                       ("best_inliers", "<i4"), ("best_iteration", "<i4"), ("H", "<f8", (9,)),
                       ("canvas", [("canvas_w", "<i4"), ("canvas_h", "<i4"), ("left_x", "<i4"), ("left_y", "<i4"),
                                   ("TH", "<f8", (9,))]),
                       ("ms_detect", "<f4"), ("ms_match", "<f4"), ("ms_ransac", "<f4"), ("ms_warp", "<f4"),
                       ("ms_total", "<f4")], align=True)
assert PAIR_DTYPE.itemsize == C.sizeof(PairResult), (PAIR_DTYPE.itemsize, C.sizeof(PairResult))


def results_array(results):
    """structured numpy view (PAIR_DTYPE) of a ctypes PairResult array"""
    return np.frombuffer(results, dtype=PAIR_DTYPE)


class PanoError(RuntimeError):
    def __init__(self, status, msg=""):
        super().__init__("pano_b200: %s (%d) %s" % (STATUS_NAMES.get(status, "?"), status, msg))
        self.status = status


def build(verbose=False):
    """Compile libpano_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE, "-j8"], stdout=out)
    return LIB_PATH


def load_library():
    """Loads the C-ABI library.  Raises if it has not been built: there is no fallback."""
    if not os.path.exists(LIB_PATH):
        raise PanoError(PANO_ERR_NO_DEVICE, "libpano_b200.so is not built (run __graft_entry__.build())")
    lib = C.CDLL(LIB_PATH)
    lib.pano_last_error.restype = C.c_char_p
    lib.pano_version.restype = C.c_char_p
    lib.pano_kernel_launches.restype = C.c_uint64
    return lib


def _is_torch_cuda(x):
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


class _Img:
    """pointer / geometry of an image argument (numpy host array or CUDA torch tensor)"""

    def __init__(self, img):
        if _is_torch_cuda(img):
            assert img.dim() == 3 and img.shape[2] == 3 and str(img.dtype) == "torch.uint8"
            assert img.stride(2) == 1 and img.stride(1) == 3
            self.keep = img
            self.ptr = C.c_void_p(img.data_ptr())
            self.h, self.w = int(img.shape[0]), int(img.shape[1])
            self.stride = int(img.stride(0))
            self.mem = MEM_DEVICE
        else:
            a = np.asarray(img)
            assert a.ndim == 3 and a.shape[2] == 3 and a.dtype == np.uint8
            if a.strides[2] != 1 or a.strides[1] != 3:
                a = np.ascontiguousarray(a)
            self.keep = a
            self.ptr = C.c_void_p(a.ctypes.data)
            self.h, self.w = a.shape[0], a.shape[1]
            self.stride = a.strides[0]
            self.mem = MEM_HOST


class Engine:
    """One context = one device + one stream (ref: single-threaded caller)."""

    def __init__(self, device=0, seed=12345):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        st = self.lib.pano_create(int(device), C.c_uint32(seed), C.byref(self.ctx))
        if st != PANO_OK:
            self.ctx = None
            raise PanoError(st, "pano_create failed (an sm_100 GPU is required; there is no CPU path)")
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pano_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, allow=()):
        if st != PANO_OK and st not in allow:
            raise PanoError(st, (self.lib.pano_last_error(self.ctx) or b"").decode())
        return st

    def set_seed(self, seed):
        self._check(self.lib.pano_set_seed(self.ctx, C.c_uint32(seed)))

    def set_matcher(self, which):
        """0 = tensor-core matcher (product), 1 = SIMT cross-check kernel"""
        self._check(self.lib.pano_set_matcher(self.ctx, int(which)))

    def set_match_mode(self, mode, ratio=0.75, descriptor=KNN_PATCH_SSD):
        """matcher of the fused calls: 0 = the reference's (default), 1 = 2-NN + Lowe's ratio test (opt-in; results differ
        from the reference's by design)"""
        self._check(self.lib.pano_set_match_mode(self.ctx, int(mode), C.c_double(ratio), int(descriptor)))

    def set_fold_mode(self, mode):
        """0 = the reference's fold (re-detect on the panorama), 1 = incremental (carry the keypoints; opt-in)"""
        self._check(self.lib.pano_set_fold_mode(self.ctx, int(mode)))

    def set_replay_mode(self, mode):
        """0 = chunked speculative shuffle replay (lowest latency), 1 = resident one-CTA replay (least work)"""
        self._check(self.lib.pano_set_replay_mode(self.ctx, int(mode)))

    def kernel_launches(self):
        return int(self.lib.pano_kernel_launches(self.ctx))

    # ---- stage entry points (reference names) ----------------------------------------
    def gpuHarrisCornerDetectorDetect(self, image, k=0.04, nmsThresh=1e6, nmsNeighborhood=3):
        im = _Img(image)
        assert im.mem == MEM_HOST, "stage API with device images: use detect_device"
        opts = HarrisCornerOptions(k, nmsThresh, nmsNeighborhood)
        count = C.c_int(0)
        self._check(self.lib.pano_detect(self.ctx, im.ptr, im.w, im.h, C.c_size_t(im.stride), im.mem, C.byref(opts),
                                         None, 0, C.byref(count)))
        xy = np.empty((max(count.value, 1), 2), np.int32)
        self._check(self.lib.pano_detect(self.ctx, im.ptr, im.w, im.h, C.c_size_t(im.stride), im.mem, C.byref(opts),
                                         xy.ctypes.data_as(C.c_void_p), count.value, C.byref(count)))
        return xy[:count.value]

    def harrisResponse(self, image, k=0.04):
        im = _Img(image)
        assert im.mem == MEM_HOST
        out = np.empty((im.h, im.w), np.float64)
        self._check(self.lib.pano_harris_response(self.ctx, im.ptr, im.w, im.h, C.c_size_t(im.stride), im.mem,
                                                  C.c_double(k), out.ctypes.data_as(C.c_void_p)))
        return out

    def convolveCUDA(self, input64f, kernel):
        a = np.ascontiguousarray(input64f, np.float64)
        kern = np.ascontiguousarray(kernel, np.float64)
        out = np.empty_like(a)
        self._check(self.lib.pano_convolve_f64(self.ctx, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0],
                                               kern.ctypes.data_as(C.c_void_p), kern.shape[0], MEM_HOST,
                                               out.ctypes.data_as(C.c_void_p)))
        return out

    def gpuHarrisMatchKeyPoints(self, keypointsL, keypointsR, image1, image2, patchSize=5, maxSSDThresh=1e8,
                                offset=0):
        """(query keypoints, train keypoints, query image, train image) as in the reference."""
        kq = np.ascontiguousarray(keypointsL, np.int32).reshape(-1, 2)
        kt = np.ascontiguousarray(keypointsR, np.int32).reshape(-1, 2)
        iq, it = _Img(image1), _Img(image2)
        assert iq.mem == MEM_HOST and it.mem == MEM_HOST
        opts = HarrisCornerOptions(patchSize_=patchSize, maxSSDThresh_=maxSSDThresh)
        out = np.empty(max(len(kq), 1), MATCH_DTYPE)
        count = C.c_int(0)
        self._check(self.lib.pano_match(self.ctx, kq.ctypes.data_as(C.c_void_p), len(kq),
                                        kt.ctypes.data_as(C.c_void_p), len(kt), iq.ptr, iq.w, iq.h,
                                        C.c_size_t(iq.stride), it.ptr, it.w, it.h, C.c_size_t(it.stride), MEM_HOST,
                                        C.byref(opts), int(offset), out.ctypes.data_as(C.c_void_p), len(out),
                                        C.byref(count)))
        return out[:count.value]

    def matchKnn(self, keypointsQ, keypointsT, imageQ, imageT, patchSize=5, descriptor=KNN_PATCH_SSD, ratio=0.75):
        """Opt-in 2-nearest-neighbour matching with Lowe's ratio test (north star item (c); NOT the reference's
        matcher, which is gpuHarrisMatchKeyPoints).  Returns (matches, runner-up distances)."""
        kq = np.ascontiguousarray(keypointsQ, np.int32).reshape(-1, 2)
        kt = np.ascontiguousarray(keypointsT, np.int32).reshape(-1, 2)
        iq, it = _Img(imageQ), _Img(imageT)
        assert iq.mem == MEM_HOST and it.mem == MEM_HOST
        opts = KnnOptions(patchSize_=patchSize, descriptor_=descriptor, ratio_=ratio)
        out = np.empty(max(len(kq), 1), MATCH_DTYPE)
        second = np.empty(max(len(kq), 1), np.float32)
        count = C.c_int(0)
        self._check(self.lib.pano_match_knn(self.ctx, kq.ctypes.data_as(C.c_void_p), len(kq),
                                            kt.ctypes.data_as(C.c_void_p), len(kt), iq.ptr, iq.w, iq.h,
                                            C.c_size_t(iq.stride), it.ptr, it.w, it.h, C.c_size_t(it.stride), MEM_HOST,
                                            C.byref(opts), out.ctypes.data_as(C.c_void_p),
                                            second.ctypes.data_as(C.c_void_p), len(out), C.byref(count)))
        return out[:count.value], second[:count.value]

    def computeHomography(self, keypoints1, keypoints2, matches, options=None, details=False):
        """GpuRansacHomographyCalculator::computeHomography.  Returns H (3x3) or None (the
        reference's empty Mat); with details=True a dict with samples/counts/inlier mask."""
        o = options or RansacOptions()
        k1 = np.ascontiguousarray(keypoints1, np.int32).reshape(-1, 2)
        k2 = np.ascontiguousarray(keypoints2, np.int32).reshape(-1, 2)
        m = np.ascontiguousarray(matches, MATCH_DTYPE)
        H = np.zeros((3, 3), np.float64)
        best, best_iter = C.c_int(0), C.c_int(-1)
        it = max(o.numIterations_, 1)
        samples = np.full((it, 4), -1, np.int32)
        counts = np.full(it, -2, np.int32)
        mask = np.zeros(max(len(m), 1), np.uint8)
        st = self.lib.pano_ransac(self.ctx, k1.ctypes.data_as(C.c_void_p), len(k1), k2.ctypes.data_as(C.c_void_p),
                                  len(k2), m.ctypes.data_as(C.c_void_p), len(m), MEM_HOST, C.byref(o),
                                  H.ctypes.data_as(C.c_void_p), C.byref(best), C.byref(best_iter),
                                  samples.ctypes.data_as(C.c_void_p) if details else None,
                                  counts.ctypes.data_as(C.c_void_p) if details else None,
                                  mask.ctypes.data_as(C.c_void_p) if details else None)
        self._check(st, allow=(PANO_ERR_TOO_FEW_MATCHES, PANO_ERR_NO_HOMOGRAPHY))
        Hret = H if st == PANO_OK else None
        if not details:
            return Hret
        return dict(ok=st == PANO_OK, status=st, H=Hret, best_count=best.value, best_iter=best_iter.value,
                    samples=samples, counts=counts, inlier_mask=mask[:len(m)].astype(bool))

    def canvasGeometry(self, wl, hl, wr, hr, H):
        H = np.ascontiguousarray(H, np.float64)
        info = CanvasInfo()
        st = self.lib.pano_canvas_geometry(wl, hl, wr, hr, H.ctypes.data_as(C.c_void_p), C.byref(info))
        return st == PANO_OK, (info.canvas_w, info.canvas_h, info.left_x, info.left_y), \
            np.array(info.TH[:], np.float64).reshape(3, 3)

    def warpPerspective(self, src, M, dsize):
        im = _Img(src)
        assert im.mem == MEM_HOST
        M = np.ascontiguousarray(M, np.float64)
        dw, dh = dsize
        dst = np.empty((dh, dw, 3), np.uint8)
        self._check(self.lib.pano_warp_perspective(self.ctx, im.ptr, im.w, im.h, C.c_size_t(im.stride), MEM_HOST,
                                                   M.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p),
                                                   dw, dh, C.c_size_t(dst.strides[0])))
        return dst

    def warpOverlay(self, left, right, H):
        """warp + left copy + overlay of stitchTwoImages for a given H; None if the ROI fails."""
        L, R = _Img(left), _Img(right)
        assert L.mem == MEM_HOST and R.mem == MEM_HOST
        H = np.ascontiguousarray(H, np.float64)
        ok, (cw, ch, _, _), _ = self.canvasGeometry(L.w, L.h, R.w, R.h, H)
        if not ok:
            return None
        canvas = np.empty((ch, cw, 3), np.uint8)
        info = CanvasInfo()
        self._check(self.lib.pano_warp_overlay(self.ctx, L.ptr, L.w, L.h, C.c_size_t(L.stride), R.ptr, R.w, R.h,
                                               C.c_size_t(R.stride), MEM_HOST, H.ctypes.data_as(C.c_void_p),
                                               canvas.ctypes.data_as(C.c_void_p), C.c_size_t(canvas.strides[0]),
                                               C.c_size_t(canvas.nbytes), C.byref(info)))
        return canvas

    # ---- fused pipeline ----------------------------------------------------------------
    def stitchTwoImages(self, leftImage, rightImage, harrisOpts=None, ransacOpts=None, fetch=True):
        """ref: stitchTwoImages.  Returns (canvas or None, result dict).  With CUDA tensors as
        inputs and fetch=False nothing but the small result record crosses PCIe."""
        ho, ro = harrisOpts or HarrisCornerOptions(), ransacOpts or RansacOptions()
        L, R = _Img(leftImage), _Img(rightImage)
        assert L.mem == R.mem
        res = PairResult()
        st = self.lib.pano_stitch_pair(self.ctx, L.ptr, L.w, L.h, C.c_size_t(L.stride), R.ptr, R.w, R.h,
                                       C.c_size_t(R.stride), L.mem, C.byref(ho), C.byref(ro), C.byref(res))
        self._check(st, allow=(PANO_ERR_NO_MATCHES, PANO_ERR_TOO_FEW_MATCHES, PANO_ERR_NO_HOMOGRAPHY, PANO_ERR_ROI))
        d = res.as_dict()
        canvas = None
        if st == PANO_OK and fetch:
            canvas = self.getCanvas(device=(L.mem == MEM_DEVICE))
        return canvas, d

    def stitchTwoImagesAsync(self, leftImage, rightImage, harrisOpts=None, ransacOpts=None, stream_ptr=None):
        """pano_stitch_pair_async: returns a handle; .done() polls, .result() waits and returns the result dict.
        The inputs are kept alive by the handle; the canvas is fetched with getCanvas afterwards."""
        ho, ro = harrisOpts or HarrisCornerOptions(), ransacOpts or RansacOptions()
        L, R = _Img(leftImage), _Img(rightImage)
        assert L.mem == R.mem
        res = PairResult()
        st = self.lib.pano_stitch_pair_async(self.ctx, L.ptr, L.w, L.h, C.c_size_t(L.stride), R.ptr, R.w, R.h,
                                             C.c_size_t(R.stride), L.mem, C.byref(ho), C.byref(ro),
                                             C.c_void_p(stream_ptr) if stream_ptr else None, C.byref(res))
        self._check(st)
        eng = self

        class Handle:
            keep = (L, R, ho, ro, res)
            status = None

            def done(self):
                if self.status is None:
                    s = eng.lib.pano_pair_query(eng.ctx)
                    if s == PANO_ERR_BUSY:
                        return False
                    self.status = s
                return True

            def result(self):
                if self.status is None:
                    self.status = eng.lib.pano_pair_wait(eng.ctx)
                eng._check(self.status, allow=(PANO_ERR_NO_MATCHES, PANO_ERR_TOO_FEW_MATCHES, PANO_ERR_NO_HOMOGRAPHY,
                                               PANO_ERR_ROI))
                return res.as_dict()
        return Handle()

    # ---- asynchronous forms of the stage calls (pano_*_async; completion through pano_pair_query / pano_pair_wait) ----
    def _async_handle(self, keep, finish, allow=()):
        eng = self

        class Handle:
            status = None

            def done(self):
                if self.status is None:
                    s = eng.lib.pano_pair_query(eng.ctx)
                    if s == PANO_ERR_BUSY:
                        return False
                    self.status = s
                return True

            def result(self):
                if self.status is None:
                    self.status = eng.lib.pano_pair_wait(eng.ctx)
                eng._check(self.status, allow=allow)
                return finish(self.status)
        h = Handle()
        h.keep = keep
        return h

    def gpuHarrisCornerDetectorDetectAsync(self, image, k=0.04, nmsThresh=1e6, nmsNeighborhood=3, cap=1 << 18, stream_ptr=None):
        """pano_detect_async: handle whose .result() is the keypoint array (cap = output capacity in keypoints)"""
        im = _Img(image)
        assert im.mem == MEM_HOST
        opts = HarrisCornerOptions(k, nmsThresh, nmsNeighborhood)
        xy = np.empty((cap, 2), np.int32)
        count = C.c_int(0)
        self._check(self.lib.pano_detect_async(self.ctx, im.ptr, im.w, im.h, C.c_size_t(im.stride), im.mem, C.byref(opts),
                                               xy.ctypes.data_as(C.c_void_p), cap, C.byref(count),
                                               C.c_void_p(stream_ptr) if stream_ptr else None))
        return self._async_handle((im, opts, xy, count), lambda st: xy[:count.value])

    def gpuHarrisMatchKeyPointsAsync(self, keypointsL, keypointsR, image1, image2, patchSize=5, maxSSDThresh=1e8, offset=0,
                                     stream_ptr=None):
        kq = np.ascontiguousarray(keypointsL, np.int32).reshape(-1, 2)
        kt = np.ascontiguousarray(keypointsR, np.int32).reshape(-1, 2)
        iq, it = _Img(image1), _Img(image2)
        assert iq.mem == MEM_HOST and it.mem == MEM_HOST
        opts = HarrisCornerOptions(patchSize_=patchSize, maxSSDThresh_=maxSSDThresh)
        out = np.empty(max(len(kq), 1), MATCH_DTYPE)
        count = C.c_int(0)
        self._check(self.lib.pano_match_async(self.ctx, kq.ctypes.data_as(C.c_void_p), len(kq), kt.ctypes.data_as(C.c_void_p),
                                              len(kt), iq.ptr, iq.w, iq.h, C.c_size_t(iq.stride), it.ptr, it.w, it.h,
                                              C.c_size_t(it.stride), MEM_HOST, C.byref(opts), int(offset),
                                              out.ctypes.data_as(C.c_void_p), len(out), C.byref(count),
                                              C.c_void_p(stream_ptr) if stream_ptr else None))
        return self._async_handle((kq, kt, iq, it, opts, out, count), lambda st: out[:count.value])

    def computeHomographyAsync(self, keypoints1, keypoints2, matches, options=None, stream_ptr=None):
        """pano_ransac_async: .result() is (H or None, best inlier count, best iteration)"""
        o = options or RansacOptions()
        k1 = np.ascontiguousarray(keypoints1, np.int32).reshape(-1, 2)
        k2 = np.ascontiguousarray(keypoints2, np.int32).reshape(-1, 2)
        m = np.ascontiguousarray(matches, MATCH_DTYPE)
        H = np.zeros((3, 3), np.float64)
        best, best_iter = C.c_int(0), C.c_int(-1)
        self._check(self.lib.pano_ransac_async(self.ctx, k1.ctypes.data_as(C.c_void_p), len(k1), k2.ctypes.data_as(C.c_void_p),
                                               len(k2), m.ctypes.data_as(C.c_void_p), len(m), MEM_HOST, C.byref(o),
                                               H.ctypes.data_as(C.c_void_p), C.byref(best), C.byref(best_iter), None, None, None,
                                               C.c_void_p(stream_ptr) if stream_ptr else None))
        return self._async_handle((k1, k2, m, o, H, best, best_iter),
                                  lambda st: (H if st == PANO_OK else None, best.value, best_iter.value),
                                  allow=(PANO_ERR_TOO_FEW_MATCHES, PANO_ERR_NO_HOMOGRAPHY))

    def set_profile(self, on=True):
        self._check(self.lib.pano_set_profile(self.ctx, 1 if on else 0))

    def get_profile(self):
        """{kernel class: (total ms, launches)} since set_profile(True)"""
        n = len(PROFILE_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_int * n)()
        self.lib.pano_get_profile(self.ctx, ms, cnt, n)
        return {PROFILE_CLASSES[i]: (ms[i], cnt[i]) for i in range(n)}

    def getCanvas(self, device=False, out=None):
        w, h = C.c_int(0), C.c_int(0)
        self._check(self.lib.pano_get_canvas(self.ctx, None, 0, 0, MEM_HOST, C.byref(w), C.byref(h)))
        if device:
            import torch
            t = out if out is not None else torch.empty((h.value, w.value, 3), dtype=torch.uint8,
                                                        device="cuda:%d" % self.device)
            self._check(self.lib.pano_get_canvas(self.ctx, C.c_void_p(t.data_ptr()), C.c_size_t(t.stride(0)),
                                                 C.c_size_t(t.numel()), MEM_DEVICE, C.byref(w), C.byref(h)))
            return t
        a = out if out is not None else np.empty((h.value, w.value, 3), np.uint8)
        self._check(self.lib.pano_get_canvas(self.ctx, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.strides[0]),
                                             C.c_size_t(a.nbytes), MEM_HOST, C.byref(w), C.byref(h)))
        return a

    def stitchAllImages(self, images, harrisOpts=None, ransacOpts=None, fetch=True):
        """ref: stitchAllImages (left fold).  Returns (panorama, [per-step result dicts])."""
        ho, ro = harrisOpts or HarrisCornerOptions(), ransacOpts or RansacOptions()
        ims = [_Img(i) for i in images]
        n = len(ims)
        assert n >= 1 and all(i.mem == ims[0].mem for i in ims)
        ptrs = (C.c_void_p * n)(*[i.ptr for i in ims])
        ws = (C.c_int * n)(*[i.w for i in ims])
        hs = (C.c_int * n)(*[i.h for i in ims])
        ss = (C.c_size_t * n)(*[i.stride for i in ims])
        results = (PairResult * max(n - 1, 1))()
        self._check(self.lib.pano_stitch_fold(self.ctx, ptrs, ws, hs, ss, n, ims[0].mem, C.byref(ho), C.byref(ro),
                                              results))
        pano = self.getCanvas(device=(ims[0].mem == MEM_DEVICE)) if fetch else None
        return pano, [results[i].as_dict() for i in range(n - 1)]


    def makeBatch(self, lefts, rights, canvases_out=None):
        """Pointer tables of a batch, built once and reusable across stitchBatch calls (a caller that stitches the
        same buffers repeatedly - or bench.py's timed region - does not pay the per-image Python conversions again)."""
        Ls, Rs = [_Img(i) for i in lefts], [_Img(i) for i in rights]
        n = len(Ls)
        assert n == len(Rs) and n > 0
        L0, R0 = Ls[0], Rs[0]
        assert all((i.w, i.h, i.stride, i.mem) == (L0.w, L0.h, L0.stride, L0.mem) for i in Ls)
        assert all((i.w, i.h, i.stride, i.mem) == (R0.w, R0.h, R0.stride, L0.mem) for i in Rs)
        b = type("Batch", (), {})()
        b.keep = (Ls, Rs, canvases_out)
        b.n, b.L0, b.R0 = n, L0, R0
        b.lp = (C.c_void_p * n)(*[i.ptr for i in Ls])
        b.rp = (C.c_void_p * n)(*[i.ptr for i in Rs])
        b.results = (PairResult * n)()
        b.cp, b.cap = None, 0
        if canvases_out is not None:
            assert len(canvases_out) == n
            ptrs, cap = [], 0
            for c in canvases_out:
                if _is_torch_cuda(c) or hasattr(c, "data_ptr"):
                    ptrs.append(C.c_void_p(c.data_ptr())); cap = int(c.numel())
                else:
                    ptrs.append(C.c_void_p(c.ctypes.data)); cap = int(c.nbytes)
            b.cp, b.cap = (C.c_void_p * n)(*ptrs), cap
        return b

    def stitchBatch(self, lefts=None, rights=None, harrisOpts=None, ransacOpts=None, canvases_out=None, batch=None,
                    raw=False):
        """n independent pairs of one geometry (throughput mode).  lefts/rights: lists of host
        arrays or CUDA tensors; canvases_out: optional list of preallocated flat uint8 buffers
        (same memory kind) receiving the tightly packed canvases; or batch= a makeBatch() object.  Returns
        ([result dicts], device ms for the whole batch); with raw=True the ctypes PairResult array instead of dicts."""
        ho, ro = harrisOpts or HarrisCornerOptions(), ransacOpts or RansacOptions()
        b = batch if batch is not None else self.makeBatch(lefts, rights, canvases_out)
        ms = C.c_float(0)
        self._check(self.lib.pano_stitch_batch(self.ctx, b.n, b.lp, b.rp, b.L0.w, b.L0.h, C.c_size_t(b.L0.stride), b.R0.w,
                                               b.R0.h, C.c_size_t(b.R0.stride), b.L0.mem, C.byref(ho), C.byref(ro),
                                               b.results, b.cp, C.c_size_t(b.cap), C.byref(ms)))
        if raw:
            return b.results, ms.value
        return [b.results[i].as_dict() for i in range(b.n)], ms.value

    # ---- chain mode --------------------------------------------------
    def pairHomography(self, leftImage, rightImage, harrisOpts=None, ransacOpts=None):
        """detect + match + RANSAC only: result dict with 'H' (right -> left) and 'status'"""
        ho, ro = harrisOpts or HarrisCornerOptions(), ransacOpts or RansacOptions()
        L, R = _Img(leftImage), _Img(rightImage)
        assert L.mem == R.mem
        res = PairResult()
        st = self.lib.pano_pair_homography(self.ctx, L.ptr, L.w, L.h, C.c_size_t(L.stride), R.ptr, R.w, R.h,
                                           C.c_size_t(R.stride), L.mem, C.byref(ho), C.byref(ro), C.byref(res))
        self._check(st, allow=(PANO_ERR_NO_MATCHES, PANO_ERR_TOO_FEW_MATCHES, PANO_ERR_NO_HOMOGRAPHY))
        return res.as_dict()

    def mul33(self, A, B):
        A = np.ascontiguousarray(A, np.float64); B = np.ascontiguousarray(B, np.float64)
        out = np.empty((3, 3), np.float64)
        self.lib.pano_mul33(A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def composeChain(self, pair_H):
        """pair_H[i] = H(image i <- image i+1) or None for a failed pair.  Returns the list of
        H(image 0 <- image i) for the longest prefix of the chain that is connected."""
        Hs = [np.eye(3)]
        for H in pair_H:
            if H is None:
                break
            Hs.append(self.mul33(Hs[-1], H))
        return Hs

    def replayWork(self, n_matches, iterations=1000, target_candidates=0.0):
        """operation count of the chunked shuffle replay for this match count (host-side planning only)"""
        w = ReplayWork()
        st = self.lib.pano_replay_work_estimate(int(n_matches), int(iterations), C.c_double(target_candidates), C.byref(w))
        if st != PANO_OK:
            raise PanoError(st, "pano_replay_work_estimate")
        return w.as_dict()

    def chainGeometry(self, sizes, Hs):
        """sizes: [(w, h)], Hs: H(0 <- i).  Returns ok, (cw, ch, x0, y0), T"""
        n = len(Hs)
        ws = (C.c_int * n)(*[s[0] for s in sizes[:n]]); hs = (C.c_int * n)(*[s[1] for s in sizes[:n]])
        flat = np.ascontiguousarray(np.stack(Hs), np.float64)
        info = CanvasInfo()
        st = self.lib.pano_chain_geometry(n, ws, hs, flat.ctypes.data_as(C.c_void_p), C.byref(info))
        return st == PANO_OK, (info.canvas_w, info.canvas_h, info.left_x, info.left_y), \
            np.array(info.TH[:], np.float64).reshape(3, 3)

    def renderChainBand(self, images, Hs, geom, T, y0, band_h, out=None):
        """rows [y0, y0 + band_h) of the chain canvas: image 0 placed at its integer offset, then every image
        i >= 1 warped by T * H(0 <- i), non-black pixels overwriting.  The band stays on the device while the images
        are accumulated into it and crosses PCIe once, into `out` (e.g. this rank's rows of a shared pinned
        canvas) or into a new host array."""
        import torch
        cw, ch, x0, yy0 = geom
        dev = torch.device("cuda", self.device)
        band = torch.zeros((band_h, cw, 3), dtype=torch.uint8, device=dev)
        dimgs = [im if _is_torch_cuda(im) else torch.from_numpy(np.ascontiguousarray(im)).to(dev) for im in images[:len(Hs)]]
        torch.cuda.current_stream(dev).synchronize()
        for i, H in enumerate(Hs):
            M = np.array([[1, 0, x0], [0, 1, yy0], [0, 0, 1.0]]) if i == 0 else self.mul33(T, H)
            im = _Img(dimgs[i])
            M = np.ascontiguousarray(M, np.float64)
            self._check(self.lib.pano_warp_accumulate(self.ctx, im.ptr, im.w, im.h, C.c_size_t(im.stride), MEM_DEVICE,
                                                      M.ctypes.data_as(C.c_void_p), C.c_void_p(band.data_ptr()),
                                                      cw, ch, int(y0), int(band_h), C.c_size_t(band.stride(0))))
        if out is None:
            return band.cpu().numpy()
        torch.from_numpy(out).copy_(band)
        return out

    def stitchChain(self, images, harrisOpts=None, ransacOpts=None):
        """single-GPU chain mode: returns (panorama or None, [pair result dicts])"""
        res = [self.pairHomography(images[i], images[i + 1], harrisOpts, ransacOpts) for i in range(len(images) - 1)]
        Hs = self.composeChain([r["H"] if r["status"] == 0 else None for r in res])
        sizes = [(np.asarray(im).shape[1], np.asarray(im).shape[0]) for im in images]
        ok, geom, T = self.chainGeometry(sizes, Hs)
        if not ok:
            return None, res
        return self.renderChainBand(images, Hs, geom, T, 0, geom[1]), res

    def set_stream(self, stream_ptr):
        """enqueue on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); None resets"""
        self._check(self.lib.pano_set_stream(self.ctx, C.c_void_p(stream_ptr) if stream_ptr else None))

    def stream_ptr(self):
        self.lib.pano_stream.restype = C.c_void_p
        return self.lib.pano_stream(self.ctx)


# reference-style free functions bound to a default engine ------------------------------
_default = None


def default_engine():
    global _default
    if _default is None:
        _default = Engine()
    return _default


class GpuRansacHomographyCalculator:
    """ref: src/gpu/ransac.cuh:8-36"""
    Options = RansacOptions

    def __init__(self, options=None, engine=None):
        self.options_ = options or RansacOptions()
        self.engine = engine or default_engine()

    def computeHomography(self, keypoints1, keypoints2, matches):
        return self.engine.computeHomography(keypoints1, keypoints2, matches, self.options_)
