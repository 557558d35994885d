#!/usr/bin/env python
"""Write tests/golden/bow_cases.npz: outputs of the oracle restatement of SearchByBoW (both overloads),
SearchForTriangulation and ComputeDistinctiveDescriptors on seeded synthetic keyframe pairs (the reference's
ORBmatcher.cc / MapPoint.cc cannot be compiled here: they pull in the un-vendored DBoW2 and g2o)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_bow_matchers import golden_cases, GOLDEN
np.savez_compressed(GOLDEN, **golden_cases())
print("wrote", GOLDEN)
