// phmm_inst_f64_exact.cu -- instantiates forward_kernel<PolicyF64, every Shape of phmm_launch.h, every MODE, EXACT=true>.
#include "phmm_launch.h"
namespace phmm {
void register_f64_exact(KernelTab& tab) { register_all<PolicyF64, true, false>(tab); }
}
