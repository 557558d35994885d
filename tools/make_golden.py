"""Generates tests/golden/*.npz from the reference itself: the reference's own, unmodified
src/ORBextractor.cc compiled in place (oracle/_ref, bump allocator) run on seeded synthetic
frames.  Needs /root/reference (or a prebuilt oracle/_ref); the fixtures it writes are what
travels.  Stereo fixtures come from the reference's own Frame::ComputeStereoMatches (src/Frame.cc compiled
unmodified into oracle/_ref/libref_matcher.so, oracle/refm.py), fed with the reference extractor's outputs;
the restatement must agree before a fixture is written (its `sad` array is kept as a diagnostic).

    python tools/make_golden.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from object_slam_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [  # name, shape, nfeatures, generator, seed
    ("tum_blocky_s0", synth.TUM_SHAPE, 1000, "blocky_image", 0),
    ("tum_noise_s3", synth.TUM_SHAPE, 1000, "noise_image", 3),
    ("kitti_blocky_s0", synth.KITTI_SHAPE, 2000, "blocky_image", 0),
    ("kitti_noise_s1", synth.KITTI_SHAPE, 2000, "noise_image", 1),
]

def main():
    assert oracle.ref_available() or os.path.exists("/root/reference"), "needs the compiled reference"
    for name, shape, nf, gen, seed in CASES:
        img = getattr(synth, gen)(shape, seed)
        ref = oracle.ReferenceExtractor(nf)
        k, d = ref(img)
        lv = [ref.level(l) for l in range(8)]
        # level checksums keep the fixture small: sum and a position-weighted sum per level
        cks = np.array([[int(p.sum()), int((p.astype(np.int64) * (np.arange(p.size).reshape(p.shape) % 65521)).sum())] for p in lv], np.int64)
        np.savez_compressed(os.path.join(OUT, f"extract_{name}.npz"), keypoints=k, descriptors=d, level_checksums=cks,
                            shape=np.array(shape), nfeatures=nf, generator=gen, seed=seed)
        print(name, len(k))
    # stereo: reference extractor outputs -> the reference's ComputeStereoMatches
    from oracle import refm
    import sys as _sys
    _sys.path.insert(0, os.path.join(os.path.dirname(OUT)))
    from matcher_cases import bounds
    for seed in (0, 5):
        L, R = synth.stereo_pair(synth.KITTI_SHAPE, seed)
        rL, rR = oracle.ReferenceExtractor(2000), oracle.ReferenceExtractor(2000)
        kL, dL = rL(L); kR, dR = rR(R)
        sc = np.empty(8, np.float32); oracle.ref_lib().ref_get_scale_factors(rL._h, sc.ctypes.data_as(oracle.C.c_void_p))
        pl, pr = [rL.level(l) for l in range(8)], [rR.level(l) for l in range(8)]
        rur, rdp, used = refm.stereo_match(kL, dL, kR, dR, pl, pr, sc, synth.KITTI_BF, synth.KITTI_FX, bounds(synth.KITTI_SHAPE))
        assert np.float32(used) == np.float32(synth.KITTI_FX), "mbf / mb must reproduce maxD = fx"
        ur, dp, sad = oracle.stereo_match(kL, dL, kR, dR, pl, pr, sc, (1.0 / sc).astype(np.float32), synth.KITTI_BF, 0.0, synth.KITTI_FX)
        assert np.array_equal(ur, rur) and np.array_equal(dp, rdp), "restatement differs from the compiled reference"
        ur, dp = rur, rdp
        np.savez_compressed(os.path.join(OUT, f"stereo_kitti_s{seed}.npz"), uRight=ur, depth=dp, sad=sad,
                            n_left=len(kL), n_right=len(kR), seed=seed)
        print("stereo", seed, (ur >= 0).sum())

if __name__ == "__main__":
    main()
