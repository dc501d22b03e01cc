// alp_k_encode_f64.cu — encode_kernel<double, true> (vector-order layout); one translation unit of libalp_b200.so
#include "alp_k_encode.inc"

namespace alpb200 {
template int launch_encode_impl<double, true>(const double*, uint64_t, const alpb200_rg_state*, const alpb200_column*, void*, void*, bool);
}
