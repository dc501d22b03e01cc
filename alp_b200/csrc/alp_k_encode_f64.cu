// alp_k_encode_f64.cu — encode_kernel<double> (one translation unit of libalp_b200.so)
#include "alp_k_encode.inc"

namespace alpb200 {
template int launch_encode<double>(const double*, uint64_t, const alpb200_rg_state*, const alpb200_column*, void*, void*);
}
