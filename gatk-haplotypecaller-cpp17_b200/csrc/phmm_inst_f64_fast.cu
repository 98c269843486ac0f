// phmm_inst_f64_fast.cu -- instantiates forward_kernel<PolicyF64, K=1..8, G=32, uniform/general, EXACT=false>.
#include "phmm_launch.h"
namespace phmm {
void register_f64_fast(KernelFn (*tab)[kMaxRowsPerLane]) { PHMM_REGISTER_ALL(PolicyF64, false); }
}
