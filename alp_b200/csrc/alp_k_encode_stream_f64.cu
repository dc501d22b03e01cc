// alp_k_encode_stream_f64.cu — encode_stream_kernel<double> (vector-order layout, streaming pipeline); one translation unit of libalp_b200.so
#define ALPB200_STREAM_PROFILE_EXPORT 1  // (development profile build: this unit exports the counter reader)
#include "alp_k_encode_stream.inc"

namespace alpb200 {
template int launch_encode_stream<double>(const double*, uint64_t, const alpb200_rg_state*, const alpb200_column*, void*, void*, bool);
}
