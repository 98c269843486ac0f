// phmm_inst_f32_exact.cu -- instantiates forward_kernel<PolicyF32x2, K=1..8, G=32, uniform/general, EXACT=true>.
#include "phmm_launch.h"
namespace phmm {
void register_f32_exact(KernelFn (*tab)[kMaxRowsPerLane]) { PHMM_REGISTER_ALL(PolicyF32x2, true); }
}
