// phmm_inst_f64_fast.cu -- instantiates forward_kernel<PolicyF64, every Shape of phmm_launch.h, every MODE, EXACT=false> and the work-list (LIST) variants.
#include "phmm_launch.h"
namespace phmm {
void register_f64_fast(KernelTab& tab) { register_all<PolicyF64, false, true>(tab); }
}
