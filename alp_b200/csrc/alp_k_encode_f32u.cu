// alp_k_encode_f32u.cu — encode_kernel<float, false> (completion-order layout, alpb200_encode_unordered_*); one translation unit of libalp_b200.so
#include "alp_k_encode.inc"

namespace alpb200 {
template int launch_encode_impl<float, false>(const float*, uint64_t, const alpb200_rg_state*, const alpb200_column*, void*, void*, bool);
}
