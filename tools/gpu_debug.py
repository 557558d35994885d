"""Stage-by-stage comparison of the CUDA path with the oracle (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from object_slam_b200 import synth
from object_slam_b200.extractor import ORBextractor, ComputeStereoMatches

def compare(shape, nf, gen, seed):
    img = gen(shape, seed)
    o = oracle.OracleExtractor(nf)
    ok, od = o(img)
    e = ORBextractor(nf, 1.2, 8, 20, 7, max_size=(shape[1], shape[0]))
    k, d = e(img)
    bad = []
    for l in range(8):
        if not np.array_equal(o.level(l), e.level(l)): bad.append(f"pyr{l}:{(o.level(l)!=e.level(l)).sum()}")
        b = o.level(l, True)
        if b is not None and not np.array_equal(b, e.level(l, blurred=True)): bad.append(f"blur{l}:{(b!=e.level(l,blurred=True)).sum()}")
        oc = o.level_keypoints(l, False)
        c = e.candidates(l)
        occ = np.stack([oc['x'], oc['y'], oc['response']], 1).astype(np.int32)
        if not np.array_equal(occ, c): bad.append(f"cand{l}:{len(occ)}vs{len(c)}")
        os_ = o.level_keypoints(l, True)
        s = e.selected(l)
        oss = np.stack([os_['x'] - 16, os_['y'] - 16, os_['response']], 1).astype(np.int32)
        if not np.array_equal(oss, s):
            same_set = set(map(tuple, oss)) == set(map(tuple, s))
            bad.append(f"sel{l}:{len(oss)}vs{len(s)} sameset={same_set}")
    kp_ok = len(k) == len(ok) and k.tobytes() == ok.tobytes()
    d_ok = d.shape == od.shape and np.array_equal(d, od)
    if not kp_ok and len(k) == len(ok):
        for f in k.dtype.names:
            n = (k[f] != ok[f]).sum()
            if n: bad.append(f"kp.{f}:{n}")
    if not d_ok and d.shape == od.shape: bad.append(f"descbits:{np.unpackbits(d ^ od).sum()}")
    print(shape, nf, gen.__name__, seed, "n", len(ok), len(k), "kp", kp_ok, "desc", d_ok, bad)
    return kp_ok and d_ok and not bad

def stereo(seed):
    shape = synth.KITTI_SHAPE
    L, R = synth.stereo_pair(shape, seed)
    oL, oR = oracle.OracleExtractor(2000), oracle.OracleExtractor(2000)
    kL, dL = oL(L); kR, dR = oR(R)
    t = oL.tables()
    ur, dp, sad = oracle.stereo_match(kL, dL, kR, dR, [oL.level(l) for l in range(8)], [oR.level(l) for l in range(8)],
                                      t['scale'], t['inv_scale'], synth.KITTI_BF, 0, synth.KITTI_FX)
    eL = ORBextractor(2000, 1.2, 8, 20, 7); eR = ORBextractor(2000, 1.2, 8, 20, 7)
    gk, gd = eL(L); eR(R)
    (gur, gdp), = ComputeStereoMatches(eL, eR, synth.KITTI_BF, 0, synth.KITTI_FX)
    ok = len(gur) == len(ur) and np.array_equal(gur, ur) and np.array_equal(gdp, dp)
    print("stereo", seed, "matches", (ur >= 0).sum(), (gur >= 0).sum(), "exact", ok,
          "maxdiff", np.abs(gur - ur).max() if len(gur) == len(ur) else None)
    return ok

if __name__ == "__main__":
    allok = True
    for shape, nf in ((synth.TUM_SHAPE, 1000), (synth.KITTI_SHAPE, 2000)):
        for gen in (synth.blocky_image, synth.noise_image):
            for seed in (0, 1):
                allok &= compare(shape, nf, gen, seed)
    for seed in (0, 1, 2):
        allok &= stereo(seed)
    print("ALL OK" if allok else "MISMATCHES")
