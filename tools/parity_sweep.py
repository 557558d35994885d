#!/usr/bin/env python
"""Randomised parity sweep on the GPU box: random shapes / parameters / generators, extractor (keypoints + descriptors) and stereo
matcher against the oracle, bit for bit.  usage: python tools/parity_sweep.py [n_cases] > gpurun_out/parity_sweep.json"""
import json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from object_slam_b200 import synth
from object_slam_b200.extractor import ORBextractor, ComputeStereoMatches

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(20261017)
cases = []
for i in range(n_cases):
    h = int(rng.integers(90, 520)); w = int(rng.integers(120, 1300))
    nf = int(rng.choice([300, 500, 1000, 1500, 2000, 3000]))
    sf = float(rng.choice([1.2, 1.2, 1.2, 1.1, 1.3, 1.5]))
    nl = int(rng.choice([8, 8, 8, 4, 6, 5]))
    ini = int(rng.choice([20, 20, 20, 12, 30, 40])); mn = int(rng.choice([7, 7, 5, 10]))
    gen = str(rng.choice(["blocky_image", "blocky_image", "noise_image"]))
    cases.append((h, w, nf, sf, nl, ini, mn, gen, int(rng.integers(0, 1 << 30))))

t0 = time.time()
bad, frames, kps, stereo_frames, stereo_bad = [], 0, 0, 0, 0


def oracle_one(c):
    h, w, nf, sf, nl, ini, mn, gen, seed = c
    img = getattr(synth, gen)((h, w), seed)
    o = oracle.OracleExtractor(nf, sf, nl, ini, mn)
    return o(img)


with ThreadPoolExecutor(os.cpu_count() or 4) as pool:
    refs = list(pool.map(oracle_one, cases))
for c, (ok, od) in zip(cases, refs):
    h, w, nf, sf, nl, ini, mn, gen, seed = c
    img = getattr(synth, gen)((h, w), seed)
    try:
        e = ORBextractor(nf, sf, nl, ini, mn, max_size=(w, h))
        k, d = e(img)
    except Exception as ex:                                  # e.g. image too small for the level count: the oracle must agree it is degenerate
        bad.append({"case": c, "error": str(ex)[:120]})
        continue
    frames += 1; kps += len(k)
    if k.tobytes() != ok.tobytes() or not np.array_equal(d, od):
        bad.append({"case": c, "n_gpu": int(len(k)), "n_oracle": int(len(ok))})
# stereo on random shapes (default extractor parameters)
for i in range(max(n_cases // 6, 10)):
    h = int(rng.integers(200, 500)); w = int(rng.integers(400, 1300)); seed = int(rng.integers(0, 1 << 30))
    L, R = synth.stereo_pair((h, w), seed)
    eL = ORBextractor(1500, 1.2, 8, 20, 7, max_size=(w, h)); eR = ORBextractor(1500, 1.2, 8, 20, 7, max_size=(w, h))
    eL(L); eR(R)
    fx = 0.58 * w
    (ur, dp), = ComputeStereoMatches(eL, eR, 0.54 * fx, 0.0, fx)
    oL, oR = oracle.OracleExtractor(1500), oracle.OracleExtractor(1500)
    kL, dL = oL(L); kR, dR = oR(R)
    t = oL.tables()
    our, odp, _ = oracle.stereo_match(kL, dL, kR, dR, [oL.level(l) for l in range(8)], [oR.level(l) for l in range(8)], t["scale"], t["inv_scale"],
                                      0.54 * fx, 0.0, fx)
    stereo_frames += 1
    if not (np.array_equal(ur, our) and np.array_equal(dp, odp)):
        stereo_bad += 1
print(json.dumps({"extractor_cases": n_cases, "extractor_frames_compared": frames, "keypoints_compared": kps, "extractor_mismatches": bad,
                  "stereo_frames_compared": stereo_frames, "stereo_mismatches": stereo_bad, "seconds": round(time.time() - t0, 1)}, indent=1))
