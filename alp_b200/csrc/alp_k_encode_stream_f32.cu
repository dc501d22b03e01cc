// alp_k_encode_stream_f32.cu — encode_stream_kernel<float> (vector-order layout, streaming pipeline); one translation unit of libalp_b200.so
#include "alp_k_encode_stream.inc"

namespace alpb200 {
template int launch_encode_stream<float>(const float*, uint64_t, const alpb200_rg_state*, const alpb200_column*, void*, void*, bool);
}
