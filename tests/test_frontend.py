"""Front-end neighbours of the extractor: colour -> grey, depth scaling,
Frame::ComputeStereoFromRGBD.  CPU: the numpy restatements against cv2; GPU: the kernels against the restatements."""
import numpy as np
import pytest

import oracle
from object_slam_b200 import synth


def test_gray_restatement_matches_cv2_on_all_colours():
    cv2 = pytest.importorskip("cv2")
    r, g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    img = np.stack([r, g, b], -1).reshape(256, 65536, 3)
    assert np.array_equal(oracle.gray_from_color(img, True), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))
    assert np.array_equal(oracle.gray_from_color(img, False), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    img4 = np.concatenate([img[:32], np.full((32, 65536, 1), 200, np.uint8)], -1)
    assert np.array_equal(oracle.gray_from_color(img4, True), cv2.cvtColor(img4, cv2.COLOR_RGBA2GRAY))
    assert np.array_equal(oracle.gray_from_color(img4, False), cv2.cvtColor(img4, cv2.COLOR_BGRA2GRAY))


def _rgbd_case(shape, seed):
    rng = np.random.default_rng(seed)
    gray = synth.blocky_image(shape, seed)
    rgb = np.clip(gray[..., None].astype(np.int32) + rng.integers(-30, 31, (*shape, 3)), 0, 255).astype(np.uint8)
    depth = rng.integers(0, 40000, shape).astype(np.uint16)
    depth[rng.random(shape) < 0.2] = 0                      # holes in the depth map
    return rgb, depth


@pytest.mark.gpu
@pytest.mark.parametrize("shape,channels", [(synth.TUM_SHAPE, 3), (synth.TUM_SHAPE, 4), ((61, 203), 3), ((61, 203), 4)])
def test_gray_and_depth_kernels(gpu, shape, channels):
    import torch
    from object_slam_b200.extractor import ORBextractor
    h, w = shape
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_size=(640, 480), max_batch=3)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (3, h, w, channels), dtype=np.uint8)
    pitch = (w + 15) & ~15
    for rgb_order in (True, False):
        for contiguous in (True, False):                    # aligned rows take the vector path, odd strides the scalar one
            if contiguous:
                d_src = torch.from_numpy(src).to(dev)
                sstride = w * channels
            else:
                d_src = torch.zeros((3, h, w * channels + 5), dtype=torch.uint8, device=dev)
                d_src[:, :, :w * channels] = torch.from_numpy(src.reshape(3, h, w * channels)).to(dev)
                sstride = w * channels + 5
            d_dst = torch.zeros((3, h, pitch), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize()
            ex.gray_from_color(d_src.data_ptr(), 3, w, h, channels, rgb_order, sstride, h * sstride, d_dst.data_ptr(), pitch, h * pitch)
            torch.cuda.synchronize()
            got = d_dst.cpu().numpy()[:, :, :w]
            assert np.array_equal(got, oracle.gray_from_color(src, rgb_order))
            assert not d_dst.cpu().numpy()[:, :, w:].any()
    depth = rng.integers(0, 65536, (3, h, w)).astype(np.uint16)
    d_d = torch.from_numpy(depth.view(np.int16)).to(dev)
    d_f = torch.zeros((3, h, w), dtype=torch.float32, device=dev)
    for factor in (1.0 / 5000.0, 1.0 / 5208.0, 1.0):
        ex.depth_to_float(d_d.data_ptr(), 3, w, h, w * 2, h * w * 2, factor, d_f.data_ptr(), w * 4, h * w * 4)
        torch.cuda.synchronize()
        assert np.array_equal(d_f.cpu().numpy(), oracle.depth_to_float(depth, factor))


@pytest.mark.gpu
def test_rgbd_frame_path_matches_oracle(gpu):
    """RGB + depth in, as Tracking::GrabImageRGBD sees them: grey conversion, extraction, ComputeStereoFromRGBD, frame set with
    mvuRight, SearchByProjection -- all on the device, compared stage by stage with the oracle."""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import matcher_cases as mc
    from object_slam_b200.extractor import ORBextractor
    from object_slam_b200.matcher import ORBmatcher
    shape = synth.TUM_SHAPE
    h, w = shape
    B = 2
    cases = [_rgbd_case(shape, 20 + i) for i in range(B)]
    dev = torch.device("cuda:0")
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_size=(w, h), max_batch=B)
    d_rgb = torch.from_numpy(np.stack([c[0] for c in cases])).to(dev)
    d_dep = torch.from_numpy(np.stack([c[1] for c in cases]).view(np.int16)).to(dev)
    d_gray = torch.empty((B, h, w), dtype=torch.uint8, device=dev)
    d_depf = torch.empty((B, h, w), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    factor, mbf = 1.0 / 5000.0, 40.0
    ex.gray_from_color(d_rgb.data_ptr(), B, w, h, 3, True, w * 3, h * w * 3, d_gray.data_ptr(), w, h * w)
    ex.depth_to_float(d_dep.data_ptr(), B, w, h, w * 2, h * w * 2, factor, d_depf.data_ptr(), w * 4, h * w * 4)
    ex.extract_device(d_gray.data_ptr(), B, w, h, w, h * w)
    pu, pd = ex.stereo_from_rgbd(d_depf.data_ptr(), w * 4, h * w * 4, mbf)
    M = ORBmatcher(0.8, True)
    fs = M.frame_set(ex.GetScaleFactors(), mc.bounds(shape), synth.camera_for(shape), max_frames=B, max_keypoints=ex.capacity)
    fs.from_extractor(ex, pu)
    res = ex.fetch()
    cap = ex.capacity
    torch.cuda.synchronize()
    ur_all = _from_device(pu, (B, cap)); dp_all = _from_device(pd, (B, cap))
    for b in range(B):
        rgb, depth = cases[b]
        gray = oracle.gray_from_color(rgb, True)
        assert np.array_equal(d_gray[b].cpu().numpy(), gray)
        ok, od = oracle.OracleExtractor(1000)(gray)
        k, d = res[b]
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)
        our, odp = oracle.stereo_from_rgbd(ok, oracle.depth_to_float(depth, factor), mbf)
        n = len(k)
        assert np.array_equal(ur_all[b, :n], our) and np.array_equal(dp_all[b, :n], odp)
        assert np.all(ur_all[b, n:] == -1) and (our > 0).sum() > 300
        mp = synth.map_points_for_frame(ok, od, shape, 4000, 50 + b)
        nm, match = M.SearchByProjection(fs, *[mp[key] for key in mc.MP_KEYS], th=3.0)
        on, om = mc.oracle_map((ok, od, our), shape, mp, 3.0, 0.8)
        assert nm[b] == on and np.array_equal(match[b, :n], om)


def _from_device(ptr, shape):
    import torch
    n = int(np.prod(shape))
    out = torch.empty(n, dtype=torch.float32, device="cuda:0")
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert rt.cudaMemcpy(C.c_void_p(out.data_ptr()), C.c_void_p(ptr), n * 4, 3) == 0      # device to device
    return out.cpu().numpy().reshape(shape)
