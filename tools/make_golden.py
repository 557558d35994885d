"""Generates tests/golden/*.npz from the reference itself: the reference's own, unmodified
src/ORBextractor.cc compiled in place (oracle/_ref, bump allocator) run on seeded synthetic
frames.  Needs /root/reference (or a prebuilt oracle/_ref); the fixtures it writes are what
travels.  Stereo fixtures come from the restated ComputeStereoMatches (the reference's Frame.cc
cannot be compiled here), fed with the reference extractor's outputs.

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
    # stereo: reference extractor outputs -> restated ComputeStereoMatches
    for seed in (0, 5):
        L, R = synth.stereo_pair(synth.KITTI_SHAPE, seed)
        rL, rR = oracle.ReferenceExtractor(2000), oracle.ReferenceExtractor(2000)
        kL, dL = rL(L); kR, dR = rR(R)
        sc = np.empty(8, np.float32); oracle.ref_lib().ref_get_scale_factors(rL._h, sc.ctypes.data_as(oracle.C.c_void_p))
        ur, dp, sad = oracle.stereo_match(kL, dL, kR, dR, [rL.level(l) for l in range(8)], [rR.level(l) for l in range(8)],
                                          sc, (1.0 / sc).astype(np.float32), synth.KITTI_BF, 0.0, synth.KITTI_FX)
        np.savez_compressed(os.path.join(OUT, f"stereo_kitti_s{seed}.npz"), uRight=ur, depth=dp, sad=sad,
                            n_left=len(kL), n_right=len(kR), seed=seed)
        print("stereo", seed, (ur >= 0).sum())

if __name__ == "__main__":
    main()
