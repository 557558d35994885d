// Tensor-core brute-force Hamming search (knn2_tc.cu): launcher used by obs_hamming_knn2.
#pragma once
#include "matcher.h"
// bytes of the +-1 int8 expansion of nKeyframes x n descriptors (256 per descriptor)
size_t knn2_tc_expanded_bytes(int nKeyframes, int n);
size_t knn2_tc_scratch_bytes(int nKeyframes);   // bytes of the `used` scratch (keyframe flags + the kernel's work counter)
cudaError_t knn2_tc_peak(double* tops);      // measured int8 tcgen05 throughput in the kernel's MMA shape (TOP/s)
// expands the keyframes of a.desc that a.pairs names into `expanded` (`used`: nKeyframes scratch bytes) and runs the tcgen05
// kernel over a.pairs; same outputs as launch_knn2
cudaError_t launch_knn2_tc(const Knn2Args& a, int nKeyframes, uint8_t* expanded, uint8_t* used, cudaStream_t st);
void knn2_tc_set_cta_pair(bool on);       // tcgen05 kernel with cta_group::2 CTA pairs (default) or one CTA per SM
