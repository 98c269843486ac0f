// phmm_inst_f32_fast.cu -- instantiates forward_kernel<PolicyF32x2, every Shape of phmm_launch.h, every MODE, EXACT=false> and the work-list (LIST) variants.
#include "phmm_launch.h"
namespace phmm {
void register_f32_fast(KernelTab& tab) { register_all<PolicyF32x2, false, true>(tab); }
}
