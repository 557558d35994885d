#!/usr/bin/env python
"""Write tests/golden/bow_cases.npz from the REFERENCE ITSELF: SearchByBoW (both overloads), SearchForTriangulation and
ComputeDistinctiveDescriptors of the reference's own, unmodified src/ORBmatcher.cc / MapPoint.cc compiled in place
(oracle/_ref/libref_matcher.so, oracle/refm.py) on seeded synthetic keyframe pairs; the restatement (oracle/match_oracle.cpp) must
agree on every array before the fixture is written.  Needs /root/reference; the fixture travels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle.refm
from test_bow_matchers import golden_cases, GOLDEN
ref = golden_cases(reference=True)
res = golden_cases(stored=ref)
for k in ref:
    assert np.array_equal(ref[k], res[k]), (k, "restatement differs from the compiled reference")
np.savez_compressed(GOLDEN, **ref)
print("wrote", GOLDEN, sorted(ref))
