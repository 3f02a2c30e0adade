// persistent_inst.cu — instantiations of the render kernel (persistent.cuh).  Compiled once per integrator
// (-DLR_INST_INTEGRATOR=0|1) by build.py so the two translation units build in parallel.
#include "persistent.cuh"

#ifndef LR_INST_INTEGRATOR
#error "compile with -DLR_INST_INTEGRATOR=0|1"
#endif

namespace lr {

#define LR_PASTE2(a, b) a##b
#define LR_PASTE(a, b) LR_PASTE2(a, b)

cudaError_t LR_PASTE(launch_persistent_i, LR_INST_INTEGRATOR)(const DevScene& sc, const DevParams& p, bool count, float* out_sum,
                                                              float* out_sumsq, unsigned long long* counters, unsigned int* next_unit,
                                                              int sm_count, cudaStream_t stream) {
  if (count) return launch_persistent_one<LR_INST_INTEGRATOR, true, true>(sc, p, out_sum, out_sumsq, counters, next_unit, sm_count, stream);
  if (sc.n_nodes > 0) return launch_persistent_one<LR_INST_INTEGRATOR, true, false>(sc, p, out_sum, out_sumsq, counters, next_unit, sm_count, stream);
  return launch_persistent_one<LR_INST_INTEGRATOR, false, false>(sc, p, out_sum, out_sumsq, counters, next_unit, sm_count, stream);
}

}  // namespace lr
