"""Host-side mirror of the reference's ``ORB_SLAM2::ORBextractor`` (include/ORBextractor.h:45-111)
over the C ABI.  Same constructor arguments, same call semantics, same getters; the work runs in
the CUDA kernels of ``csrc/`` -- there is no CPU implementation behind this class.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import KEYPOINT_DTYPE, OrbParams, check, lib, ptr


class ORBextractor:
    """``ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)`` -- ORBextractor.h:51.

    ``max_size`` = (width, height) upper bound of the images, ``max_batch`` = images per batched
    call; both only size the device buffers.
    """

    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST,
                 max_size=(1241, 376), max_batch=1, device=0):
        self._h = C.c_void_p()
        self.nfeatures, self.scaleFactor, self.nlevels = int(nfeatures), float(scaleFactor), int(nlevels)
        self.iniThFAST, self.minThFAST = int(iniThFAST), int(minThFAST)
        self.max_batch = int(max_batch)
        prm = OrbParams(self.nfeatures, self.scaleFactor, self.nlevels, self.iniThFAST, self.minThFAST)
        check(lib().obs_extractor_create(C.byref(prm), int(max_size[0]), int(max_size[1]), self.max_batch,
                                         int(device), C.byref(self._h)))
        n = self.nlevels
        self._scale, self._inv, self._s2, self._is2 = (np.empty(n, np.float32) for _ in range(4))
        self._fpl = np.empty(n, np.int32)
        check(lib().obs_extractor_tables(self._h, ptr(self._scale), ptr(self._inv), ptr(self._s2), ptr(self._is2), ptr(self._fpl)))
        self._last_n = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().obs_extractor_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # ---- getters, ORBextractor.h:63-83
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return np.float32(self.scaleFactor)

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv.copy()

    def GetScaleSigmaSquares(self):
        return self._s2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._is2.copy()

    @property
    def mnFeaturesPerLevel(self):
        return self._fpl.copy()

    @property
    def capacity(self):
        return lib().obs_extractor_max_keypoints(self._h)

    # ---- operator(), ORBextractor.cc:1043-1105
    def __call__(self, image, mask=None):
        """Returns (keypoints, descriptors): a KEYPOINT_DTYPE array (cv::KeyPoint layout, level-major
        order) and an N x 32 uint8 array.  The mask is ignored, as in the reference.  An empty image
        gives empty outputs; zero keypoints give a 0 x 32 descriptor array (the reference releases it)."""
        image = np.asarray(image)
        if image.size == 0:
            self._last_n = 0
            return np.empty(0, KEYPOINT_DTYPE), np.empty((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise ValueError("image must be 8-bit single channel (the reference asserts CV_8UC1, ORBextractor.cc:1050)")
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        cap = self.capacity
        kps = np.empty(cap, KEYPOINT_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        check(lib().obs_extract(self._h, ptr(image), image.shape[1], image.shape[0], image.strides[0],
                                ptr(kps), ptr(desc), cap, C.byref(n)))
        self._last_n = 1
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, out=None, copy=True):
        """operator() over equally shaped host images in one call.  `images`: a list of 2-D arrays or one
        (n, h, w) array; page-locked arrays (``_capi.pinned_empty``) are DMA-ed without staging.
        `out` = (keypoints[n, cap], descriptors[n, cap, 32], counts[n]) reuses caller (ideally pinned) buffers;
        with copy=False the per-image results are views into them."""
        if isinstance(images, np.ndarray) and images.ndim == 3:
            assert images.dtype == np.uint8 and images.strides[2] == 1
            nimg, h, w = images.shape
            stride = images.strides[1]
            ptrs = [images.ctypes.data + i * images.strides[0] for i in range(nimg)]
        else:
            images = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
            h, w = images[0].shape
            assert all(im.shape == (h, w) for im in images)
            nimg, stride = len(images), w
            ptrs = [im.ctypes.data for im in images]
        cap = self.capacity
        if out is None:
            out = (np.empty((nimg, cap), KEYPOINT_DTYPE), np.empty((nimg, cap, 32), np.uint8), np.zeros(nimg, np.int32))
        kps, desc, counts = out
        arr = (C.c_void_p * nimg)(*ptrs)
        check(lib().obs_extract_batch(self._h, arr, nimg, w, h, stride, ptr(kps), ptr(desc), kps.shape[1], ptr(counts)))
        self._last_n = nimg
        if copy:
            return [(kps[i, :counts[i]].copy(), desc[i, :counts[i]].copy()) for i in range(nimg)]
        return [(kps[i, :counts[i]], desc[i, :counts[i]]) for i in range(nimg)]

    def submit_batch(self, images, keypoints, descriptors, counts, n=None):
        """obs_extract_batch_submit on page-locked arrays: images [F, H, stride] uint8, keypoints [F, cap] KEYPOINT_DTYPE,
        descriptors [F, cap, 32] uint8, counts [F] int32.  Returns once enqueued; ``wait_batch`` sleeps until the results are there."""
        n = images.shape[0] if n is None else int(n)
        check(lib().obs_extract_batch_submit(self._h, ptr(images), n, images.shape[2], images.shape[1], images.strides[1],
                                             ptr(keypoints), ptr(descriptors), keypoints.shape[1], ptr(counts)))
        self._last_n = n

    def wait_batch(self):
        check(lib().obs_extract_batch_wait(self._h))

    def extract_device(self, d_images, n_images, w, h, stride, image_stride, stream=None):
        """Images already resident in HBM (``d_images`` = device address); results stay on the device."""
        check(lib().obs_extract_batch_device(self._h, C.c_void_p(int(d_images)), int(n_images), int(w), int(h),
                                             int(stride), int(image_stride), C.c_void_p(int(stream) if stream else 0)))
        self._last_n = int(n_images)

    def fetch(self):
        nimg, cap = self._last_n, self.capacity
        kps = np.empty((nimg, cap), KEYPOINT_DTYPE)
        desc = np.empty((nimg, cap, 32), np.uint8)
        counts = np.zeros(nimg, np.int32)
        check(lib().obs_extractor_fetch(self._h, ptr(kps), ptr(desc), cap, ptr(counts)))
        return [(kps[i, :counts[i]].copy(), desc[i, :counts[i]].copy()) for i in range(nimg)]

    def fetch_counts(self):
        counts = np.zeros(self._last_n, np.int32)
        check(lib().obs_extractor_fetch_counts(self._h, ptr(counts)))
        return counts

    def results_device(self):
        p, rb, cap = C.c_void_p(), C.c_size_t(), C.c_int()
        check(lib().obs_extractor_results_device(self._h, C.byref(p), C.byref(rb), C.byref(cap)))
        return p.value, rb.value, cap.value

    # ---- front-end neighbours (device-resident): Tracking.cc:202-263, Frame.cc:883-904
    def gray_from_color(self, d_src, n_images, w, h, channels, rgb_order, src_stride, src_image_stride, d_gray, gray_stride,
                        gray_image_stride, stream=None):
        check(lib().obs_gray_from_color(self._h, C.c_void_p(int(d_src)), int(n_images), int(w), int(h), int(channels), int(bool(rgb_order)),
                                        int(src_stride), int(src_image_stride), C.c_void_p(int(d_gray)), int(gray_stride),
                                        int(gray_image_stride), C.c_void_p(int(stream) if stream else 0)))

    def depth_to_float(self, d_src, n_images, w, h, src_stride, src_image_stride, factor, d_dst, dst_stride, dst_image_stride, stream=None):
        check(lib().obs_depth_to_float(self._h, C.c_void_p(int(d_src)), int(n_images), int(w), int(h), int(src_stride), int(src_image_stride),
                                       float(factor), C.c_void_p(int(d_dst)), int(dst_stride), int(dst_image_stride),
                                       C.c_void_p(int(stream) if stream else 0)))

    def stereo_from_rgbd(self, d_depth, depth_stride, depth_image_stride, mbf, stream=None):
        """Frame::ComputeStereoFromRGBD on the last extraction; returns the device addresses of mvuRight and mvDepth."""
        pu, pd = C.c_void_p(), C.c_void_p()
        check(lib().obs_stereo_from_rgbd(self._h, C.c_void_p(int(d_depth)), int(depth_stride), int(depth_image_stride), float(mbf),
                                         C.c_void_p(int(stream) if stream else 0), C.byref(pu), C.byref(pd)))
        return pu.value, pd.value

    STAGES = ("pyramid", "fast", "quadtree", "blur", "describe")

    def set_profiling(self, on):
        check(lib().obs_extractor_set_profiling(self._h, int(bool(on))))

    def stage_ms(self):
        """Summed per-stage device time (ms) over the calls recorded since set_profiling(True)."""
        ms = np.zeros(len(self.STAGES), np.float32)
        st, nc, ns = C.c_float(), C.c_int(), C.c_int()
        check(lib().obs_extractor_stage_ms(self._h, ptr(ms), C.byref(st), C.byref(nc), C.byref(ns)))
        out = {k: float(v) for k, v in zip(self.STAGES, ms)}
        out["stereo"] = float(st.value)
        return out, nc.value, ns.value

    # ---- mvImagePyramid, ORBextractor.h:85 (downloaded on demand; border-less)
    def level(self, level, image_index=0, blurred=False):
        w, h = C.c_int(), C.c_int()
        check(lib().obs_extractor_get_level(self._h, image_index, level, int(blurred), None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        check(lib().obs_extractor_get_level(self._h, image_index, level, int(blurred), ptr(out), out.strides[0], C.byref(w), C.byref(h)))
        return out

    @property
    def mvImagePyramid(self):
        return [self.level(l) for l in range(self.nlevels)]

    # ---- stage outputs for parity tests
    def _keys(self, fn, level, image_index):
        n = C.c_int()
        check(fn(self._h, image_index, level, None, 0, C.byref(n)))
        out = np.empty((max(n.value, 1), 3), np.int32)
        check(fn(self._h, image_index, level, ptr(out), n.value, C.byref(n)))
        return out[:n.value]

    def candidates(self, level, image_index=0):
        return self._keys(lib().obs_extractor_get_candidates, level, image_index)

    def selected(self, level, image_index=0):
        return self._keys(lib().obs_extractor_get_selected, level, image_index)


def ComputeStereoMatches(left, right, mbf, minD, maxD, out=None):
    """Frame::ComputeStereoMatches (src/Frame.cc:706-880) on the last extraction of two extractors.
    Returns (mvuRight, mvDepth) per image of the batch: float32 arrays over the left keypoints
    (views into `out` = (uRight[n, cap], depth[n, cap]) when given)."""
    cap = left.capacity
    nimg = left._last_n
    if out is None:
        ur = np.empty((nimg, cap), np.float32)
        dp = np.empty((nimg, cap), np.float32)
    else:
        ur, dp = out
    check(lib().obs_stereo_match(left._h, right._h, float(mbf), float(minD), float(maxD), ptr(ur), ptr(dp), ur.shape[1]))
    counts = left.fetch_counts()
    if out is None:
        return [(ur[i, :counts[i]].copy(), dp[i, :counts[i]].copy()) for i in range(nimg)]
    return [(ur[i, :counts[i]], dp[i, :counts[i]]) for i in range(nimg)]


def stereo_match_device(left, right, mbf, minD, maxD, stream=None):
    pu, pd = C.c_void_p(), C.c_void_p()
    check(lib().obs_stereo_match_device(left._h, right._h, float(mbf), float(minD), float(maxD),
                                        C.c_void_p(int(stream) if stream else 0), C.byref(pu), C.byref(pd)))
    return pu.value, pd.value


class StereoFrames:
    """The stereo ``Frame`` constructor's device work for batches of frames (src/Frame.cc:78-90: ExtractORB on both eyes, then
    ComputeStereoMatches) through ``obs_stereo_frames_submit`` / ``obs_stereo_frames_wait``: one extractor pair, page-locked
    input and output buffers, one C call per batch.  ``submit`` returns as soon as the batch is enqueued; several instances
    driven round-robin from one thread overlap each other's transfers and kernels.

    ``left`` / ``right``: page-locked [F, H, stride] uint8 arrays to fill (``self.left[i, :, :W] = image``)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, size, frames, device=0):
        from ._capi import StereoIO, pinned_empty
        W, H = int(size[0]), int(size[1])
        self.W, self.H, self.F = W, H, int(frames)
        self.eL = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_size=(W, H), max_batch=frames, device=device)
        self.eR = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_size=(W, H), max_batch=frames, device=device)
        cap = self.cap = self.eL.capacity
        self.stride = W
        self.left = pinned_empty((frames, H, W), np.uint8)
        self.right = pinned_empty((frames, H, W), np.uint8)
        mk = lambda: (pinned_empty((frames, cap), KEYPOINT_DTYPE), pinned_empty((frames, cap, 32), np.uint8), pinned_empty((frames,), np.int32))
        self.kpL, self.descL, self.nL = mk()
        self.kpR, self.descR, self.nR = mk()
        self.uRight = pinned_empty((frames, cap), np.float32)
        self.depth = pinned_empty((frames, cap), np.float32)
        a = _capi.addr
        self._io = StereoIO(a(self.left), a(self.right), a(self.kpL), a(self.descL), a(self.nL), a(self.kpR), a(self.descR), a(self.nR),
                            a(self.uRight), a(self.depth))
        self.h2d_bytes = 2 * frames * H * W
        self.d2h_bytes = 2 * frames * (cap * 60 + 4) + 2 * frames * cap * 4

    def submit(self, mbf, minD, maxD, frames=None):
        n = self.F if frames is None else int(frames)
        check(lib().obs_stereo_frames_submit(self.eL._h, self.eR._h, C.byref(self._io), n, self.W, self.H, self.stride, self.cap,
                                             float(mbf), float(minD), float(maxD)))
        self.eL._last_n = self.eR._last_n = n

    def wait(self):
        check(lib().obs_stereo_frames_wait(self.eL._h, self.eR._h))

    def __call__(self, mbf, minD, maxD, frames=None):
        self.submit(mbf, minD, maxD, frames)
        self.wait()
        return self.results(self.F if frames is None else int(frames))

    def results(self, n=None):
        """[(keysL, descL, keysR, descR, uRight, depth)] per frame (views of the page-locked buffers)."""
        n = self.F if n is None else n
        out = []
        for i in range(n):
            a, b = int(self.nL[i]), int(self.nR[i])
            out.append((self.kpL[i, :a], self.descL[i, :a], self.kpR[i, :b], self.descR[i, :b], self.uRight[i, :a], self.depth[i, :a]))
        return out
