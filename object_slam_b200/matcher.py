"""Host-side mirror of the reference's ``ORB_SLAM2::ORBmatcher`` (include/ORBmatcher.h:37-102) over the
C ABI, for the searches Tracking runs on every frame.  The reference walks Frame / MapPoint objects;
here the same quantities are flat arrays named after the members they come from.  Every array
argument may be a numpy array (host memory) or an ``int`` device address (e.g. ``tensor.data_ptr()``);
the work runs in the CUDA kernels of ``csrc/matcher.cu`` -- there is no CPU implementation behind this.
"""
import ctypes as C

import numpy as np

from ._capi import (KEYPOINT_DTYPE, BowSide, FrameParams, FrameView, KeyFramePointsView, LastFrameView, MapPointView, addr, check, lib, ptr)

FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48       # include/Frame.h:41-42


def _f32(a):
    return a if isinstance(a, int) or a is None else np.ascontiguousarray(a, np.float32)


def _i32(a):
    return a if isinstance(a, int) or a is None else np.ascontiguousarray(a, np.int32)


def _u8(a):
    return a if isinstance(a, int) or a is None else np.ascontiguousarray(a, np.uint8)


class FrameSet:
    """A batch of frames as the matchers see them (mvKeysUn, mDescriptors, mvuRight, image bounds, camera,
    the 64x48 grid built on the device).  ``bounds`` = (mnMinX, mnMaxX, mnMinY, mnMaxY); ``camera`` =
    (fx, fy, cx, cy, mbf, mb)."""

    def __init__(self, matcher, scale_factors, bounds, camera=(1, 1, 0, 0, 0, 0), max_frames=1, max_keypoints=2048):
        self.matcher = matcher
        prm = FrameParams()
        prm.min_x, prm.max_x, prm.min_y, prm.max_y = (float(b) for b in bounds)
        prm.fx, prm.fy, prm.cx, prm.cy, prm.mbf, prm.mb = (float(c) for c in camera)
        sf = np.asarray(scale_factors, np.float32)
        prm.nlevels = len(sf)
        for i, v in enumerate(sf):
            prm.scale_factors[i] = float(v)
        self.scale_factors = sf
        self._h = C.c_void_p()
        check(lib().obs_frame_set_create(matcher._h, C.byref(prm), int(max_frames), int(max_keypoints), C.byref(self._h)))
        self.max_frames = int(max_frames)
        self.cap = (int(max_keypoints) + 31) & ~31
        self.counts = []

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().obs_frame_set_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def upload(self, frames):
        """frames: list of (keys_un, descriptors, u_right or None) host arrays."""
        views = (FrameView * len(frames))()
        keep = []
        for v, (k, d, ur) in zip(views, frames):
            k = np.ascontiguousarray(k, KEYPOINT_DTYPE)
            d = _u8(d)
            ur = _f32(ur)
            keep.append((k, d, ur))
            v.n = len(k)
            v.keys_un, v.descriptors, v.u_right = addr(k), addr(d), addr(ur)
        check(lib().obs_frame_set_upload(self._h, views, len(frames)))
        self.counts = [len(k) for k, _, _ in keep]
        return self

    def from_extractor(self, extractor, d_u_right=None):
        check(lib().obs_frame_set_from_extractor(self._h, extractor._h, ptr(d_u_right)))
        self.counts = None
        return self

    def __len__(self):
        return lib().obs_frame_set_count(self._h)

    def grid(self, frame=0):
        start = np.empty(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, np.int32)
        idx = np.empty(self.cap, np.int32)
        check(lib().obs_frame_set_grid(self._h, frame, ptr(start), ptr(idx), self.cap))
        return start, idx[:start[-1]].copy()


class ORBmatcher:
    """``ORBmatcher(nnratio=0.6, checkOri=True)`` -- ORBmatcher.h:41."""

    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30      # ORBmatcher.cc:37-39

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self._h = C.c_void_p()
        check(lib().obs_matcher_create(int(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().obs_matcher_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    @property
    def stream(self):
        return lib().obs_matcher_stream(self._h)

    def sync(self):
        check(lib().obs_matcher_sync(self._h))

    def last_rounds(self):
        out = np.zeros(4096, np.int32)
        n = lib().obs_matcher_last_rounds(self._h, ptr(out), len(out))
        return out[:max(n, 0)].copy()

    def frame_set(self, *args, **kw):
        return FrameSet(self, *args, **kw)

    # ---- ORBmatcher.cc:1647-1663
    def DescriptorDistance(self, a, b):
        a = _u8(a).reshape(-1, 32)
        b = _u8(b).reshape(-1, 32)
        out = np.empty(len(a), np.int32)
        check(lib().obs_descriptor_distance(self._h, ptr(a), ptr(b), len(a), ptr(out)))
        return out

    # ---- ORBmatcher.cc:1601-1642
    def ComputeThreeMaxima(self, bin_sizes):
        h = _i32(bin_sizes)
        h2 = h.reshape(-1, h.shape[-1])
        ind = np.empty((len(h2), 3), np.int32)
        check(lib().obs_compute_three_maxima(self._h, ptr(h2), len(h2), h2.shape[1], ptr(ind)))
        return ind if h.ndim > 1 else tuple(int(i) for i in ind[0])

    def _outs(self, frames, kp_match, n_matches):
        B = len(frames)
        if kp_match is None:
            kp_match = np.empty((B, frames.cap), np.int32)
        if n_matches is None:
            n_matches = np.empty(B, np.int32)
        return kp_match, n_matches

    # ---- ORBmatcher.cc:45-129
    def SearchByProjection(self, frames, in_view, proj_x, proj_y, proj_xr, scale_level, view_cos, descriptors,
                           observations, th=1.0, n_points=None, per_frame=False, kp_observations=None,
                           kp_match=None, n_matches=None):
        """SearchByProjection(Frame&, const vector<MapPoint*>&, th) for every frame of ``frames``.  Returns
        (n_matches[B], kp_match[B, cap]): kp_match[b, k] = index of the map point assigned to keypoint k
        (-1 = none)."""
        v = MapPointView()
        arrs = [_u8(in_view), _f32(proj_x), _f32(proj_y), _f32(proj_xr), _i32(scale_level), _f32(view_cos),
                _u8(descriptors), _i32(observations)]
        if n_points is None:
            n_points = arrs[0].shape[-1] if not isinstance(arrs[0], int) else None
        v.n, v.per_frame = int(n_points), int(bool(per_frame))
        (v.in_view, v.proj_x, v.proj_y, v.proj_xr, v.scale_level, v.view_cos, v.descriptors, v.observations) = (addr(a) for a in arrs)
        kp_match, n_matches = self._outs(frames, kp_match, n_matches)
        check(lib().obs_search_by_projection(self._h, frames._h, C.byref(v), float(th), self.mfNNratio,
                                             ptr(_i32(kp_observations)), ptr(kp_match), ptr(n_matches)))
        return n_matches, kp_match

    # ---- ORBmatcher.cc:1328-1470
    def SearchByProjectionLast(self, current, has_point, world_pos, octave, angle, descriptors, observations,
                               tcw_last, tcw_current, th, mono, n_points=None, per_frame=False,
                               kp_observations=None, kp_match=None, n_matches=None):
        """SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono).  kp_match: -1 untouched,
        -2 reset to NULL by the rotation check, else the last-frame keypoint whose map point was assigned."""
        v = LastFrameView()
        arrs = [_u8(has_point), _f32(world_pos), _i32(octave), _f32(angle), _u8(descriptors), _i32(observations),
                _f32(tcw_last), _f32(tcw_current)]
        if n_points is None:
            n_points = arrs[0].shape[-1]
        v.n, v.per_frame = int(n_points), int(bool(per_frame))
        (v.has_point, v.world_pos, v.octave, v.angle, v.descriptors, v.observations, v.tcw_last, v.tcw_current) = (addr(a) for a in arrs)
        kp_match, n_matches = self._outs(current, kp_match, n_matches)
        check(lib().obs_search_by_projection_last(self._h, current._h, C.byref(v), float(th), int(bool(mono)),
                                                  int(self.mbCheckOrientation), ptr(_i32(kp_observations)),
                                                  ptr(kp_match), ptr(n_matches)))
        return n_matches, kp_match

    def _kf_view(self, pts, tcw, n_points, per_frame):
        v = KeyFramePointsView()
        arrs = dict(valid=_u8(pts["valid"]), world_pos=_f32(pts["world_pos"]), min_distance=_f32(pts["min_distance"]),
                    max_distance=_f32(pts["max_distance"]), max_distance_raw=_f32(pts["max_distance_raw"]),
                    normal=_f32(pts.get("normal")), angle=_f32(pts.get("angle")), descriptors=_u8(pts["descriptors"]), tcw=_f32(tcw))
        if n_points is None:
            n_points = arrs["valid"].shape[-1]
        v.n, v.per_frame = int(n_points), int(bool(per_frame))
        for k, a in arrs.items():
            setattr(v, k, addr(a))
        return v, arrs

    # ---- ORBmatcher.cc:1472-1599
    def SearchByProjectionKeyFrame(self, current, pts, tcw, th, orb_dist, n_points=None, per_frame=False, kp_taken=None,
                                   kp_match=None, n_matches=None):
        """SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist).  pts: dict(valid, world_pos,
        min_distance, max_distance, max_distance_raw, angle, descriptors)."""
        v, keep = self._kf_view(pts, tcw, n_points, per_frame)
        kp_match, n_matches = self._outs(current, kp_match, n_matches)
        check(lib().obs_search_by_projection_keyframe(self._h, current._h, C.byref(v), float(th), int(orb_dist), int(self.mbCheckOrientation),
                                                      ptr(_i32(kp_taken)), ptr(kp_match), ptr(n_matches)))
        return n_matches, kp_match

    # ---- ORBmatcher.cc:290-403
    def SearchByProjectionSim3(self, keyframes, pts, tcw, th, n_points=None, per_frame=False, kp_taken=None, kp_match=None, n_matches=None):
        """SearchByProjection(KeyFrame* pKF, Scw, vpPoints, vpMatched, th) after the decomposition of Scw; pts as above with
        `normal` instead of `angle`."""
        v, keep = self._kf_view(pts, tcw, n_points, per_frame)
        kp_match, n_matches = self._outs(keyframes, kp_match, n_matches)
        check(lib().obs_search_by_projection_sim3(self._h, keyframes._h, C.byref(v), int(th), ptr(_i32(kp_taken)), ptr(kp_match), ptr(n_matches)))
        return n_matches, kp_match

    # ---- ORBmatcher.cc:825-966 / :974-1100 (search half)
    def FuseSearch(self, keyframes, pts, tcw, th, camera_centre=None, sim3=False, n_points=None, per_frame=False):
        """Projection + best-keypoint search of ORBmatcher::Fuse for every keyframe of the set; pts as for SearchByProjectionSim3
        (valid = pMP && !isBad() && !IsInKeyFrame(pKF)).  camera_centre: [B, 3] GetCameraCenter() (keyframe variant).  Returns
        (best_idx, best_dist), each [B, n_points]; the caller accepts best_dist <= TH_LOW and applies Replace / AddObservation."""
        v, keep = self._kf_view(pts, tcw, n_points, per_frame)
        B = len(keyframes)
        best_idx = np.empty((B, v.n), np.int32); best_dist = np.empty((B, v.n), np.int32)
        check(lib().obs_fuse_search(self._h, keyframes._h, C.byref(v), ptr(_f32(camera_centre)), float(th), int(bool(sim3)),
                                    ptr(best_idx), ptr(best_dist)))
        return best_idx, best_dist

    # ---- ORBmatcher.cc:1102-1326
    def SearchBySim3(self, kf1, kf2, pts1, pts2, t1w, t2w, t21, t12, th, per_frame=False):
        """SearchBySim3 for every keyframe pair (kf1[i], kf2[i]); pts1 / pts2: dict(valid, world_pos, min_distance, max_distance,
        max_distance_raw, descriptors) indexed by the keyframes' keypoints; t21 = [sR21 | t21], t12 = [sR12 | t12].
        Returns (n_found [B], match12 [B, n1])."""
        v1, keep1 = self._kf_view(pts1, t1w, None, per_frame)
        v2, keep2 = self._kf_view(pts2, t2w, None, per_frame)
        B = len(kf1)
        a21, a12 = _f32(t21), _f32(t12)
        m12 = np.empty((B, v1.n), np.int32); nf = np.empty(B, np.int32)
        check(lib().obs_search_by_sim3(self._h, kf1._h, kf2._h, C.byref(v1), C.byref(v2), ptr(a21), ptr(a12), float(th), ptr(m12), ptr(nf)))
        return nf, m12

    # ---- ORBmatcher.cc:405-520
    def SearchForInitialization(self, f1, f2, prev_matched, window_size=10, matches12=None, n_matches=None):
        """Returns (n_matches[B], vnMatches12[B, cap1]); ``prev_matched`` ([B, cap1, 2] float32) is updated in place."""
        B = len(f1)
        if matches12 is None:
            matches12 = np.empty((B, f1.cap), np.int32)
        if n_matches is None:
            n_matches = np.empty(B, np.int32)
        if not isinstance(prev_matched, int):
            assert prev_matched.dtype == np.float32 and prev_matched.flags.c_contiguous
        check(lib().obs_search_for_initialization(self._h, f1._h, f2._h, ptr(prev_matched), ptr(matches12), int(window_size),
                                                  self.mfNNratio, int(self.mbCheckOrientation), ptr(n_matches)))
        return n_matches, matches12

    KNN2_AUTO, KNN2_POPC, KNN2_TENSOR = 0, 1, 2

    def set_knn2_engine(self, engine):
        """Which kernel knn2 runs: KNN2_TENSOR (tcgen05 int8 contraction), KNN2_POPC, or KNN2_AUTO (by size)."""
        check(lib().obs_matcher_set_knn2_engine(self._h, int(engine)))

    # ---- brute force best / second best + ratio (candidate loop of SearchByBoW, ORBmatcher.cc:200-229)
    def knn2(self, descriptors, pairs, n_keyframes=None, n_desc=None, th_low=None, best_idx=None, best_dist=None,
             second_dist=None, want_dists=True):
        d = _u8(descriptors)
        if n_keyframes is None:
            n_keyframes, n_desc = d.shape[0], d.shape[1]
        pr = _i32(pairs)
        n_pairs = len(pr) if not isinstance(pr, int) else None
        if best_idx is None:
            best_idx = np.empty((n_pairs, n_desc), np.int32)
            if want_dists:
                best_dist = np.empty((n_pairs, n_desc), np.int32)
                second_dist = np.empty((n_pairs, n_desc), np.int32)
        elif n_pairs is None:
            raise ValueError("device pair lists need explicit outputs")
        check(lib().obs_hamming_knn2(self._h, ptr(d), int(n_keyframes), int(n_desc), ptr(pr), int(n_pairs),
                                     int(self.TH_LOW if th_low is None else th_low), self.mfNNratio,
                                     ptr(best_idx), ptr(best_dist), ptr(second_dist)))
        return best_idx, best_dist, second_dist

    # ---- DBoW2-gated searches.  A "side" is a list of dicts, one per keyframe / frame, with the keys n, descriptors,
    # keys_un, valid (or None), u_right (or None), node_id, node_start, node_idx (the DBoW2::FeatureVector as a CSR,
    # node ids ascending); pairs are side1[i] x side2[i].
    @staticmethod
    def _pack_side(frames, use_valid=True):
        B = len(frames)
        cap = max(max(int(f["n"]) for f in frames), 1)
        node_cap = max(max(len(f["node_id"]) for f in frames), 1)
        pk = dict(n=np.zeros(B, np.int32), descriptors=np.zeros((B, cap, 32), np.uint8), keys_un=np.zeros((B, cap), KEYPOINT_DTYPE),
                  n_nodes=np.zeros(B, np.int32), node_id=np.zeros((B, node_cap), np.uint32),
                  node_start=np.zeros((B, node_cap + 1), np.int32), node_idx=np.zeros((B, cap), np.int32))
        has_valid = use_valid and any(f.get("valid") is not None for f in frames)
        has_ur = any(f.get("u_right") is not None for f in frames)
        pk["valid"] = np.ones((B, cap), np.uint8) if has_valid else None
        pk["u_right"] = np.full((B, cap), -1.0, np.float32) if has_ur else None
        for b, f in enumerate(frames):
            n, k = int(f["n"]), len(f["node_id"])
            pk["n"][b] = n; pk["n_nodes"][b] = k
            pk["descriptors"][b, :n] = f["descriptors"][:n]
            pk["keys_un"][b, :n] = f["keys_un"][:n]
            pk["node_id"][b, :k] = f["node_id"]
            pk["node_start"][b, :k + 1] = f["node_start"]
            pk["node_start"][b, k + 1:] = f["node_start"][-1] if k else 0
            pk["node_idx"][b, :n] = f["node_idx"][:n]
            if has_valid and f.get("valid") is not None:
                pk["valid"][b, :n] = f["valid"][:n]
            if has_ur and f.get("u_right") is not None:
                pk["u_right"][b, :n] = f["u_right"][:n]
        side = BowSide(cap, node_cap, *[ptr(pk[k]) for k in ("n", "descriptors", "keys_un", "valid", "u_right", "n_nodes",
                                                               "node_id", "node_start", "node_idx")])
        return side, pk

    def SearchByBoW(self, side1, side2, keyframe_pair=False, th_low=None):
        """ORBmatcher::SearchByBoW: keyframe_pair=False -> (KeyFrame*, Frame&) src/ORBmatcher.cc:159-288 (side2's valid
        is ignored, read match21: frame keypoint -> keyframe keypoint); keyframe_pair=True -> (KeyFrame*, KeyFrame*)
        :522-655 (read match12).  Returns (n_matches [B], match12 [B, cap1], match21 [B, cap2])."""
        s1, k1 = self._pack_side(side1)
        s2, k2 = self._pack_side(side2, use_valid=keyframe_pair)
        B = len(side1)
        m12 = np.empty((B, s1.cap), np.int32); m21 = np.empty((B, s2.cap), np.int32); nm = np.empty(B, np.int32)
        check(lib().obs_search_by_bow(self._h, C.byref(s1), C.byref(s2), B, int(self.TH_LOW if th_low is None else th_low),
                                      int(keyframe_pair), self.mfNNratio, int(self.mbCheckOrientation), ptr(m12), ptr(m21), ptr(nm)))
        return nm, m12, m21

    def SearchForTriangulation(self, side1, side2, f12, epipole, level_sigma2, scale_factors, bOnlyStereo=False):
        """ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:657-823.  valid = keypoint has no map point yet.
        Returns (n_matches [B], match12 [B, cap1]); vMatchedPairs of pair b = [(i, match12[b, i]) for ascending i with a match]."""
        s1, k1 = self._pack_side(side1)
        s2, k2 = self._pack_side(side2)
        B = len(side1)
        f = np.ascontiguousarray(f12, np.float32).reshape(B, 9)
        ep = np.ascontiguousarray(epipole, np.float32).reshape(B, 2)
        s2l = np.ascontiguousarray(level_sigma2, np.float32); sfl = np.ascontiguousarray(scale_factors, np.float32)
        m12 = np.empty((B, s1.cap), np.int32); nm = np.empty(B, np.int32)
        check(lib().obs_search_for_triangulation(self._h, C.byref(s1), C.byref(s2), B, ptr(f), ptr(ep), ptr(s2l), ptr(sfl), len(sfl),
                                                 int(bOnlyStereo), int(self.mbCheckOrientation), ptr(m12), ptr(nm)))
        return nm, m12

    def ComputeDistinctiveDescriptors(self, descriptors, start):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:345-410) for a batch of map points: descriptors of point p
        are rows start[p]..start[p+1]; returns the index (inside the point's list) of the chosen descriptor, -1 if empty."""
        d = _u8(descriptors); st = _i32(start)
        n_points = len(st) - 1
        best = np.empty(max(n_points, 1), np.int32)
        check(lib().obs_distinctive_descriptors(self._h, ptr(d), ptr(st), n_points, ptr(best)))
        return best[:n_points]

    # ---- object layer: Frame::BuildObject2DsRGBD / BuildObject2DsStereo, src/Frame.cc:240-385 (keypoint-to-mask assignment)
    def AssignKeypointsToMasks(self, keys_un, depth, masks, th_depth, min_keypoints=5):
        """masks: [n_masks, h, w] uint8 (255 = object).  Returns (mask_of_kp [n], mvObjectKpIndices [n, 2], object_of_mask [n_masks], N_O)."""
        k = np.ascontiguousarray(keys_un); d = _f32(depth); mk = _u8(masks)
        n, (nm, h, w) = len(k), mk.shape
        mo = np.empty(n, np.int32); ok = np.empty((n, 2), np.int32); om = np.empty(nm, np.int32); no = np.zeros(1, np.int32)
        check(lib().obs_assign_keypoints_to_masks(self._h, ptr(k), ptr(d), n, ptr(mk), nm, w, h, w, w * h, float(th_depth), int(min_keypoints),
                                                  ptr(mo), ptr(ok), ptr(om), ptr(no)))
        return mo, ok, om, int(no[0])

    # ---- Frame::ExtractHSVHistogramsFromMask, src/Frame.cc:388-414
    def ExtractHSVHistogramsFromMasks(self, im_bgr, masks):
        """im_bgr: [h, w, 3] uint8; masks: [n_masks, h, w] uint8.  Returns [n_masks, 94] float32 (V | S | H bins, L1-normalised)."""
        im = _u8(im_bgr); mk = _u8(masks)
        nm, h, w = mk.shape
        hist = np.empty((nm, 94), np.float32)
        check(lib().obs_hsv_histograms(self._h, ptr(im), 3 * w, ptr(mk), nm, w, h, w, w * h, ptr(hist)))
        return hist

    # ---- Frame::UndistortKeyPoints, src/Frame.cc:644-674
    def UndistortKeyPoints(self, keys, K, dist_coef):
        """keys: KeyPoint array (mvKeys); K = (fx, fy, cx, cy); dist_coef = (k1, k2, p1, p2[, k3]).  Returns mvKeysUn."""
        k = np.ascontiguousarray(keys); d = np.ascontiguousarray(dist_coef, np.float32)
        out = np.empty_like(k)
        check(lib().obs_undistort_keypoints(self._h, ptr(k), len(k), *[float(v) for v in K], ptr(d), len(d), ptr(out)))
        return out

    def UndistortPoints(self, pts, K, dist_coef):
        p = np.ascontiguousarray(pts, np.float32).reshape(-1, 2); d = np.ascontiguousarray(dist_coef, np.float32)
        out = np.empty_like(p)
        check(lib().obs_undistort_points(self._h, ptr(p), len(p), *[float(v) for v in K], ptr(d), len(d), ptr(out)))
        return out

    # ---- Object2D::Object2D, src/ObjectTypes.cc:23
    def DistanceTransform(self, masks):
        """cv::distanceTransform(~mask, DIST_L2, DIST_MASK_PRECISE) per mask; masks [n_masks, h, w] uint8 -> [n_masks, h, w] float32."""
        mk = _u8(masks)
        nm, h, w = mk.shape
        out = np.empty((nm, h, w), np.float32)
        check(lib().obs_distance_transform(self._h, ptr(mk), nm, w, h, w, w * h, ptr(out)))
        return out
