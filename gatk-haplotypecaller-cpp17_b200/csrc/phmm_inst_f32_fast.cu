// phmm_inst_f32_fast.cu -- instantiates forward_kernel<PolicyF32x2, every Shape of phmm_launch.h, every MODE, EXACT=false>.
#include "phmm_launch.h"
namespace phmm {
void register_f32_fast(KernelTab& tab) { PHMM_REGISTER_ALL(PolicyF32x2, false); }
}
