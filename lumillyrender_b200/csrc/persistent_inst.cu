// persistent_inst.cu — instantiations of the render kernel (persistent.cuh).  Compiled once per (integrator, scene has
// a BVH) pair by build.py (-DLR_INST_INTEGRATOR=0|1 -DLR_INST_TREE=0|1; the tree units also get -DLR_DIV_OUT_OF_LINE),
// so the four translation units build in parallel.  The instrumented (COUNT) kernel lives in the tree units.
#include "persistent.cuh"

#if !defined(LR_INST_INTEGRATOR) || !defined(LR_INST_TREE)
#error "compile with -DLR_INST_INTEGRATOR=0|1 -DLR_INST_TREE=0|1"
#endif

namespace lr {

#define LR_PASTE2(a, b, c) a##b##_t##c
#define LR_PASTE(a, b, c) LR_PASTE2(a, b, c)

cudaError_t LR_PASTE(launch_persistent_i, LR_INST_INTEGRATOR, LR_INST_TREE)(const DevScene& sc, const DevParams& p, bool count,
                                                                            float* out_sum, float* out_sumsq, unsigned long long* counters,
                                                                            unsigned int* next_unit, int sm_count, cudaStream_t stream) {
#if LR_INST_TREE
  if (count) return launch_persistent_one<LR_INST_INTEGRATOR, true, true>(sc, p, out_sum, out_sumsq, counters, next_unit, sm_count, stream);
  return launch_persistent_one<LR_INST_INTEGRATOR, true, false>(sc, p, out_sum, out_sumsq, counters, next_unit, sm_count, stream);
#else
  (void)count;
  return launch_persistent_one<LR_INST_INTEGRATOR, false, false>(sc, p, out_sum, out_sumsq, counters, next_unit, sm_count, stream);
#endif
}

}  // namespace lr
