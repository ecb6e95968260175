// pcg.cuh -- block-sparse reduced system + preconditioned conjugate gradients
// (used when the reduced system is too large to factor densely).
#pragma once
#include <string>
#include "schur.cuh"

namespace ars {

struct PcgWorkspace {
  long long total_iterations = 0;
  size_t value_count() const { return 0; }
};

inline cudaError_t pcg_init() { return cudaSuccess; }

}  // namespace ars
