"""Seeded synthetic inputs of the benchmark shapes.  numpy only.

There is no dataset in the repository or on the GPU box, so every test and bench line uses
these generators; ``seed`` is the frame index unless stated otherwise.
"""
import numpy as np

KITTI_SHAPE = (376, 1241)   # rows, cols  (Examples/Stereo/KITTI00-02.yaml)
TUM_SHAPE = (480, 640)      # rows, cols  (Examples/RGB-D/TUM1.yaml)
KITTI_FX = 718.856
KITTI_BF = 386.1448
TUM_BF = 40.0


def _gauss_blur_f32(img, sigma):
    r = max(1, int(3 * sigma + 0.5))
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2).astype(np.float32)
    k /= k.sum()
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i, kv in enumerate(k):
        out += kv * p[:, i:i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(img)
    for i, kv in enumerate(k):
        out2 += kv * p[i:i + img.shape[0], :]
    return out2


def blocky_canvas(shape, seed):
    """Noise-free float canvas: grey background with w*h/400 random grey rectangles, blurred."""
    h, w = shape
    rng = np.random.default_rng(seed)
    img = np.full((h, w), 128.0, np.float32)
    n = (w * h) // 400
    xs = rng.integers(0, w, n)
    ys = rng.integers(0, h, n)
    ws = rng.integers(4, 60, n)
    hs = rng.integers(4, 60, n)
    gs = rng.integers(0, 256, n)
    for x, y, ww, hh, g in zip(xs, ys, ws, hs, gs):
        img[y:y + hh, x:x + ww] = g
    return _gauss_blur_f32(img, 0.8)


def _finish(canvas, rng):
    out = canvas + rng.normal(0.0, 2.0, canvas.shape).astype(np.float32)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def blocky_image(shape, seed):
    """The default generator: ~2.6k level-0 FAST candidates at KITTI shape."""
    return _finish(blocky_canvas(shape, seed), np.random.default_rng(seed + 1_000_003))


def noise_image(shape, seed):
    """White-noise stress case (~42k level-0 candidates at KITTI shape)."""
    return np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)


def flat_image(shape, value=128):
    return np.full(shape, value, np.uint8)


def stereo_pair(shape, seed):
    """Left = blocky image; right = the same canvas shifted by a per-row-band disparity in
    [5, 80] px plus independent noise."""
    h, w = shape
    canvas = blocky_canvas(shape, seed)
    rng = np.random.default_rng(seed + 2_000_003)
    left = _finish(canvas, rng)
    band = 47
    nb = (h + band - 1) // band
    disp = rng.integers(5, 81, nb)
    right_c = np.empty_like(canvas)
    for b in range(nb):
        d = int(disp[b])
        rows = slice(b * band, min(h, (b + 1) * band))
        # a point at x in the left eye appears at x - d in the right eye
        right_c[rows, :w - d] = canvas[rows, d:]
        right_c[rows, w - d:] = canvas[rows, w - 1:w]
    right = _finish(right_c, rng)
    return left, right


def random_descriptors(n, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, 32), dtype=np.uint8)


def flip_bits(desc, nflip, rng):
    """Copy of ``desc`` (n x 32 u8) with nflip[i] random bits flipped in row i."""
    out = desc.copy()
    for i in range(out.shape[0]):
        k = int(nflip[i])
        if k:
            pos = rng.choice(256, size=k, replace=False)
            np.bitwise_xor.at(out[i], pos >> 3, (1 << (pos & 7)).astype(np.uint8))
    return out


# ----------------------------------------------------------------------------- matcher inputs
KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
     ("octave", "<i4"), ("class_id", "<i4")])


def scale_factors(nlevels=8, scale_factor=1.2):
    """mvScaleFactor as the ORBextractor constructor builds it (float * double -> float, ORBextractor.cc:415-420)."""
    s = np.ones(nlevels, np.float32)
    for i in range(1, nlevels):
        s[i] = np.float32(np.float64(s[i - 1]) * np.float64(np.float32(scale_factor)))
    return s


def synthetic_frame(shape, n, seed, nlevels=8, stereo_fraction=0.7, bf=TUM_BF):
    """A frame as the matchers see it without running the extractor: level-major keypoints with the
    extractor's per-level share, random descriptors, mvuRight from a synthetic depth (Frame.cc:898-902)
    for `stereo_fraction` of the keypoints (-1 elsewhere).  Returns (keys, descriptors, u_right)."""
    h, w = shape
    rng = np.random.default_rng(seed)
    sf = scale_factors(nlevels)
    share = (1 / 1.2) ** np.arange(nlevels)
    per = np.floor(n * share / share.sum()).astype(int)
    per[-1] += n - per.sum()
    keys = np.zeros(n, KEYPOINT_DTYPE)
    pos = 0
    for l, c in enumerate(per):
        lw, lh = w / sf[l], h / sf[l]
        x = rng.integers(19, max(20, int(lw) - 19), c).astype(np.float32)
        y = rng.integers(19, max(20, int(lh) - 19), c).astype(np.float32)
        keys["x"][pos:pos + c] = x * sf[l] if l else x
        keys["y"][pos:pos + c] = y * sf[l] if l else y
        keys["octave"][pos:pos + c] = l
        keys["size"][pos:pos + c] = np.float32(int(31 * sf[l]))
        pos += c
    keys["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    keys["response"] = rng.integers(7, 120, n).astype(np.float32)
    keys["class_id"] = -1
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    z = rng.uniform(0.5, 8.0, n).astype(np.float32)
    ur = (keys["x"] - np.float32(bf) / z).astype(np.float32)
    ur[rng.random(n) >= stereo_fraction] = -1.0
    return keys, desc, ur


def map_points_for_frame(keys, desc, shape, n_points, seed, nlevels=8, bf=TUM_BF, anchored=0.5):
    """cfg3-style local map projected into a frame (the fields Frame::isInFrustum leaves on each MapPoint).
    A fraction `anchored` of the points sits near the keypoint its descriptor was derived from (sigma 2 px,
    predicted level = octave or octave+1); the rest is uniform over the image with a uniform level.
    Descriptors: a frame descriptor with k ~ U{0..60} flipped bits.  Returns a dict of arrays."""
    h, w = shape
    rng = np.random.default_rng(seed)
    n = len(keys)
    src = rng.integers(0, max(n, 1), n_points)
    anch = rng.random(n_points) < anchored
    u = rng.uniform(0, w, n_points).astype(np.float32)
    v = rng.uniform(0, h, n_points).astype(np.float32)
    level = rng.integers(0, nlevels, n_points).astype(np.int32)
    if n:
        u[anch] = (keys["x"][src[anch]] + rng.normal(0, 2, anch.sum())).astype(np.float32)
        v[anch] = (keys["y"][src[anch]] + rng.normal(0, 2, anch.sum())).astype(np.float32)
        level[anch] = np.minimum(keys["octave"][src[anch]] + rng.integers(0, 2, anch.sum()), nlevels - 1)
        d = flip_bits(desc[src], rng.integers(0, 61, n_points), rng)
    else:
        d = rng.integers(0, 256, (n_points, 32), dtype=np.uint8)
    z = rng.uniform(0.5, 8.0, n_points).astype(np.float32)
    return dict(in_view=(rng.random(n_points) < 0.95).astype(np.uint8), proj_x=u, proj_y=v,
                proj_xr=(u - np.float32(bf) / z).astype(np.float32), scale_level=level,
                view_cos=rng.uniform(0.5, 1.0, n_points).astype(np.float32), descriptors=d,
                observations=rng.integers(0, 4, n_points).astype(np.int32))


def camera_for(shape, bf=TUM_BF):
    """(fx, fy, cx, cy, mbf, mb) of a pinhole camera filling `shape` (TUM1.yaml-like for 640x480)."""
    h, w = shape
    fx = np.float32(0.82 * w)
    return (float(fx), float(fx), w / 2 - 0.5, h / 2 - 0.5, float(bf), float(np.float32(bf) / fx))


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    return (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
            np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))


def motion_pair(shape, n, seed, nlevels=8, bf=TUM_BF, forward=0.0):
    """A last frame with map points and the current frame that re-observes them after a small motion
    (`forward` metres along the optical axis on top of a random few-centimetre motion).  Returns
    (last, current): last = dict(has_point, world_pos, octave, angle, descriptors, observations, tcw_last,
    tcw_current); current = (keys, descriptors, u_right)."""
    h, w = shape
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, mbf, mb = camera_for(shape, bf)
    keysL, descL, _ = synthetic_frame(shape, n, seed + 17, nlevels, bf=bf)
    z = rng.uniform(1.0, 10.0, n)
    Pc = np.stack([(keysL["x"] - cx) / fx * z, (keysL["y"] - cy) / fy * z, z], 1)
    Rl = _rot(*rng.normal(0, 0.02, 3)); tl = rng.normal(0, 0.05, 3)
    Pw = (Pc - tl) @ Rl                     # world = Rl^T (Pc - tl)
    Rc = _rot(*rng.normal(0, 0.01, 3)) @ Rl
    tc = tl + rng.normal(0, 0.03, 3) + np.array([0, 0, -forward])
    Tl = np.hstack([Rl, tl[:, None]]).astype(np.float32)
    Tc = np.hstack([Rc, tc[:, None]]).astype(np.float32)
    Pcur = Pw @ Rc.T + tc
    keysC = keysL.copy()
    keysC["x"] = (fx * Pcur[:, 0] / Pcur[:, 2] + cx + rng.normal(0, 0.7, n)).astype(np.float32)
    keysC["y"] = (fy * Pcur[:, 1] / Pcur[:, 2] + cy + rng.normal(0, 0.7, n)).astype(np.float32)
    rot_noise = rng.normal(0, 4, n)
    wild = rng.random(n) < 0.1
    rot_noise[wild] = rng.uniform(0, 360, wild.sum())
    keysC["angle"] = np.mod(keysL["angle"] + rot_noise, 360).astype(np.float32)
    descC = flip_bits(descL, rng.integers(0, 50, n), rng)
    urC = (keysC["x"] - np.float32(mbf) / Pcur[:, 2]).astype(np.float32)
    urC[rng.random(n) < 0.3] = -1.0
    perm = rng.permutation(n)               # the current frame lists its keypoints in another order
    # keep it level-major like a real frame
    perm = perm[np.argsort(keysC["octave"][perm], kind="stable")]
    last = dict(has_point=(rng.random(n) < 0.8).astype(np.uint8), world_pos=Pw.astype(np.float32),
                octave=keysL["octave"].astype(np.int32), angle=keysL["angle"].copy(),
                descriptors=flip_bits(descL, rng.integers(0, 10, n), rng), observations=rng.integers(0, 3, n).astype(np.int32),
                tcw_last=Tl, tcw_current=Tc)
    return last, (keysC[perm], descC[perm], urC[perm])


def init_pair(shape, n, seed, nlevels=8):
    """Two monocular frames for SearchForInitialization: the second re-observes the first after a small
    image-plane flow.  Returns ((keys1, desc1, None), (keys2, desc2, None), prev_matched[n, 2])."""
    rng = np.random.default_rng(seed)
    k1, d1, _ = synthetic_frame(shape, n, seed + 31, nlevels)
    k2 = k1.copy()
    flow = rng.normal(0, 6, (n, 2))
    k2["x"] = (k1["x"] + flow[:, 0]).astype(np.float32)
    k2["y"] = (k1["y"] + flow[:, 1]).astype(np.float32)
    rot_noise = rng.normal(0, 4, n)
    wild = rng.random(n) < 0.1
    rot_noise[wild] = rng.uniform(0, 360, wild.sum())
    k2["angle"] = np.mod(k1["angle"] + rot_noise, 360).astype(np.float32)
    d2 = flip_bits(d1, rng.integers(0, 70, n), rng)
    # duplicate some descriptors so that several frame-1 keypoints compete for one frame-2 keypoint
    dup = rng.integers(0, n, n // 10)
    d1 = d1.copy()
    d1[dup] = flip_bits(d1[(dup + 1) % n], rng.integers(0, 8, len(dup)), rng)
    perm = rng.permutation(n)
    perm = perm[np.argsort(k2["octave"][perm], kind="stable")]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    return (k1, d1, None), (k2[perm], d2[perm], None), prev


def keyframe_descriptors(n_keyframes, n_desc, seed, planted=0.3):
    """cfg5: uniform random descriptors; a fraction `planted` of keyframe i+1's descriptors are noisy copies
    (0..40 flipped bits) of keyframe i's."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, (n_keyframes, n_desc, 32), dtype=np.uint8)
    m = int(n_desc * planted)
    for k in range(1, n_keyframes):
        src = rng.choice(n_desc, m, replace=False)
        dst = rng.choice(n_desc, m, replace=False)
        d[k, dst] = flip_bits(d[k - 1, src], rng.integers(0, 41, m), rng)
    return d


def keyframe_points(last, seed, camera_centre=(0.0, 0.0, 0.0)):
    """Map-point fields for the relocalisation / loop-closing searches, on top of a `motion_pair` last frame: scale
    invariance distances around the true distance, normals (mean viewing directions) mostly along the ray, 10 % invalid points."""
    rng = np.random.default_rng(seed)
    pos = last["world_pos"]
    n = len(pos)
    d = np.linalg.norm(pos - np.asarray(camera_centre, np.float32), axis=1).astype(np.float32)
    # mfMaxDistance such that MapPoint::PredictScale lands on the keypoint's octave (70 %) or next to it
    delta = rng.choice([0, 0, 0, 0, 0, 0, 0, -1, 1, 2], n)
    # the reference keeps two members per point, mfMinDistance / mfMaxDistance; the invariance limits are 0.8f / 1.2f times those and
    # PredictScale reads mfMaxDistance (MapPoint.cc:475-519) -- the generated fields follow that relation exactly
    raw = (d * np.power(1.2, last["octave"] - 0.5 + delta)).astype(np.float32)
    far = rng.random(n) < 0.05
    raw[far] = (d[far] * 0.4).astype(np.float32)                          # outside the invariance region
    raw_min = (raw / np.float32(1.2 ** 7)).astype(np.float32)
    mind = (np.float32(0.8) * raw_min).astype(np.float32)
    maxd = (np.float32(1.2) * raw).astype(np.float32)
    nrm = pos / np.linalg.norm(pos, axis=1, keepdims=True) + rng.normal(0, 0.3, (n, 3))      # mean viewing direction (camera -> point)
    flip = rng.random(n) < 0.1
    nrm[flip] *= -1                                                        # seen from behind: fails the 60-degree test
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return dict(valid=(rng.random(n) < 0.9).astype(np.uint8), world_pos=pos, min_distance=mind, max_distance=maxd,
                min_distance_raw=raw_min, max_distance_raw=raw, normal=nrm.astype(np.float32), angle=last["angle"],
                descriptors=last["descriptors"])


def feature_vector(node_of_keypoint):
    """DBoW2::FeatureVector of a frame as a CSR: (node_id ascending, node_start, node_idx); inside a node the keypoint
    indices ascend (DBoW2 appends features in index order)."""
    node_of_keypoint = np.asarray(node_of_keypoint, np.int64)
    order = np.argsort(node_of_keypoint, kind="stable")
    ids, counts = np.unique(node_of_keypoint, return_counts=True)
    start = np.zeros(len(ids) + 1, np.int32)
    start[1:] = np.cumsum(counts)
    return ids.astype(np.uint32), start, order.astype(np.int32)


def bow_pair(shape, n, seed, n_nodes=100, nlevels=8, copies=0.6):
    """Two keyframes as the DBoW2-gated matchers read them (SearchByBoW, SearchForTriangulation): keyframe 2 holds
    noisy copies of `copies` of keyframe 1's keypoints (descriptor with 0..45 flipped bits, same row up to a few px
    -- the epipolar geometry of a pure x translation, F12 = [0 0 0; 0 0 -1; 0 1 0] --, common in-plane rotation with
    outliers, mostly the same vocabulary node).  Returns (side1, side2, extra); a side is a dict with n, descriptors,
    keys_un, valid, u_right, node_id, node_start, node_idx."""
    rng = np.random.default_rng(seed)
    k1, d1, ur1 = synthetic_frame(shape, n, seed)
    k2, d2, ur2 = synthetic_frame(shape, n, seed + 7919)
    node1 = rng.integers(0, n_nodes, n) * 37 + 11
    node2 = rng.integers(0, n_nodes, n) * 37 + 11
    m = int(copies * n)
    src = rng.choice(n, m, replace=False)
    dst = rng.choice(n, m, replace=False)
    d2[dst] = flip_bits(d1[src], rng.integers(0, 46, m), rng)
    k2["x"][dst] = np.clip(k1["x"][src] - rng.uniform(2, 60, m).astype(np.float32), 1, shape[1] - 2)
    k2["y"][dst] = k1["y"][src] + rng.choice([0, 0, 0.5, -1, 1.5, -3, 6, -12], m).astype(np.float32) * scale_factors(nlevels)[k1["octave"][src]]
    k2["octave"][dst] = np.clip(k1["octave"][src] + rng.choice([0, 0, 0, 1, -1], m), 0, nlevels - 1)
    rot = np.where(rng.random(m) < 0.85, 12.0 + rng.normal(0, 3, m), rng.uniform(0, 360, m))
    k2["angle"][dst] = np.mod(k1["angle"][src] - rot, 360).astype(np.float32)
    same = rng.random(m) < 0.9
    node2[dst[same]] = node1[src[same]]
    sides = []
    for k, d, ur, node in ((k1, d1, ur1, node1), (k2, d2, ur2, node2)):
        ids, start, idx = feature_vector(node)
        sides.append(dict(n=n, descriptors=d, keys_un=k, valid=(rng.random(n) < 0.7).astype(np.uint8), u_right=ur,
                          node_id=ids, node_start=start, node_idx=idx))
    sf = scale_factors(nlevels)
    extra = dict(f12=np.array([0, 0, 0, 0, 0, -1, 0, 1, 0], np.float32), epipole=(np.float32(shape[1] * 0.7), np.float32(shape[0] * 0.4)),
                 scale_factors=sf, level_sigma2=(sf * sf).astype(np.float32))
    return sides[0], sides[1], extra


def observation_descriptors(n_points, seed, max_obs=24):
    """Descriptor lists of n_points map points for MapPoint::ComputeDistinctiveDescriptors: 1..max_obs observations each,
    noisy copies of one prototype with a few outliers.  Returns (descriptors [total, 32], start [n_points + 1])."""
    rng = np.random.default_rng(seed)
    cnt = rng.integers(1, max_obs + 1, n_points)
    cnt[rng.random(n_points) < 0.05] = 0
    start = np.zeros(n_points + 1, np.int32)
    start[1:] = np.cumsum(cnt)
    desc = np.zeros((int(start[-1]), 32), np.uint8)
    for p in range(n_points):
        c = int(cnt[p])
        if c == 0:
            continue
        proto = rng.integers(0, 256, (1, 32), dtype=np.uint8)
        block = flip_bits(np.repeat(proto, c, 0), rng.integers(0, 60, c), rng)
        out = rng.random(c) < 0.1
        block[out] = rng.integers(0, 256, (int(out.sum()), 32), dtype=np.uint8)
        desc[start[p]:start[p + 1]] = block
    return desc, start


def sim3_pair(shape, n, seed, nlevels=8, bf=TUM_BF):
    """Two keyframes that see the same n physical points from nearby poses, every keypoint carrying its own map point, and the
    similarity between the cameras, as ORBmatcher::SearchBySim3 reads them.  Returns (kf1, kf2, pts1, pts2, poses):
    kf = (keys_un, descriptors, u_right); pts = dict(valid, world_pos, min_distance, max_distance, max_distance_raw, descriptors)
    indexed by the keyframe's keypoints; poses = dict(t1w, t2w, t21, t12) with t21 = [sR21 | t21], t12 = [sR12 | t12]."""
    h, w = shape
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, mbf, mb = camera_for(shape, bf)
    keys1, desc1, ur1 = synthetic_frame(shape, n, seed + 31, nlevels, bf=bf)
    z = rng.uniform(1.0, 10.0, n)
    Pc1 = np.stack([(keys1["x"] - cx) / fx * z, (keys1["y"] - cy) / fy * z, z], 1)
    R1 = _rot(*rng.normal(0, 0.02, 3)); t1 = rng.normal(0, 0.05, 3)
    Pw = (Pc1 - t1) @ R1
    R2 = _rot(*rng.normal(0, 0.015, 3)) @ R1
    t2 = t1 + rng.normal(0, 0.04, 3)
    Pc2 = Pw @ R2.T + t2
    keys2 = keys1.copy()
    keys2["x"] = (fx * Pc2[:, 0] / Pc2[:, 2] + cx + rng.normal(0, 0.8, n)).astype(np.float32)
    keys2["y"] = (fy * Pc2[:, 1] / Pc2[:, 2] + cy + rng.normal(0, 0.8, n)).astype(np.float32)
    desc2 = flip_bits(desc1, rng.integers(0, 60, n), rng)
    ur2 = (keys2["x"] - np.float32(mbf) / Pc2[:, 2]).astype(np.float32)
    perm = rng.permutation(n)
    perm = perm[np.argsort(keys2["octave"][perm], kind="stable")]
    keys2, desc2, ur2, Pw2 = keys2[perm], desc2[perm], ur2[perm], Pw[perm] + rng.normal(0, 0.01, (n, 3))
    s12 = 1.0 + rng.normal(0, 0.01)
    R12 = R1 @ R2.T
    t12 = t1 - R12 @ t2 + rng.normal(0, 0.005, 3)
    sR12 = s12 * R12
    sR21 = (1.0 / s12) * R12.T
    t21 = -sR21 @ t12
    f32 = lambda R, t: np.hstack([R, np.asarray(t)[:, None]]).astype(np.float32)

    def points(P, octave, desc, Rt, tt, sd):
        r = np.random.default_rng(sd)
        d = np.linalg.norm(P @ Rt.T + tt, axis=1).astype(np.float32)          # distance from the camera it is projected into
        delta = r.choice([0, 0, 0, 0, 0, 0, 0, -1, 1, 2], len(P))
        raw = (d * np.power(1.2, octave - 0.5 + delta)).astype(np.float32)
        far = r.random(len(P)) < 0.05
        raw[far] = (d[far] * 0.4).astype(np.float32)          # mfMaxDistance so small that the point falls outside its invariance region
        raw_min = (raw / np.float32(1.2 ** (nlevels - 1))).astype(np.float32)
        mind = (np.float32(0.8) * raw_min).astype(np.float32)  # MapPoint::GetMinDistanceInvariance / GetMaxDistanceInvariance
        maxd = (np.float32(1.2) * raw).astype(np.float32)
        return dict(valid=(r.random(len(P)) < 0.85).astype(np.uint8), world_pos=P.astype(np.float32), min_distance=mind, max_distance=maxd,
                    min_distance_raw=raw_min, max_distance_raw=raw, descriptors=flip_bits(desc, r.integers(0, 12, len(P)), r))

    pts1 = points(Pw, keys1["octave"], desc1, R2, t2, seed + 1)
    pts2 = points(Pw2, keys2["octave"], desc2, R1, t1, seed + 2)
    poses = dict(t1w=f32(R1, t1), t2w=f32(R2, t2), t21=f32(sR21, t21), t12=f32(sR12, t12))
    return (keys1, desc1, ur1), (keys2, desc2, ur2), pts1, pts2, poses


def semantic_masks(shape, n_masks, seed):
    """n_masks instance masks (255 = object) as a detector would return them: filled ellipses and boxes of assorted sizes, some
    overlapping (a keypoint then belongs to the first), some too small to hold a 20 x 20 window."""
    h, w = shape
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    out = np.zeros((n_masks, h, w), np.uint8)
    for m in range(n_masks):
        cx, cy = rng.uniform(0.1 * w, 0.9 * w), rng.uniform(0.1 * h, 0.9 * h)
        rx, ry = rng.uniform(8, 0.3 * w), rng.uniform(8, 0.35 * h)
        if rng.random() < 0.5:
            out[m][((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0] = 255
        else:
            out[m][(abs(xx - cx) <= rx) & (abs(yy - cy) <= ry)] = 255
        if rng.random() < 0.3:                                # a hole
            out[m][(abs(xx - cx) <= rx * 0.2) & (abs(yy - cy) <= ry * 0.2)] = 0
    return out
