// phmm_inst_f32_exact.cu -- instantiates forward_kernel<PolicyF32x2, every Shape of phmm_launch.h, every MODE, EXACT=true>.
#include "phmm_launch.h"
namespace phmm {
void register_f32_exact(KernelTab& tab) { register_all<PolicyF32x2, true, false>(tab); }
}
