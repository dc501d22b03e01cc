// alp_host.h — what the translation units of libalp_b200.so share on the host side: error reporting, the per-device
// bookkeeping and the kernel launchers.  The library is split into one TU per kernel family (decode, encode f64,
// encode f32, init + primitives) so that nvcc compiles them in parallel; nothing here is part of the public ABI.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "alp_b200.h"

namespace alpb200 {

// records the message returned by alpb200_last_error() (thread-local) and returns `code`
int fail(int code, const char* fmt, const char* a = "", const char* b = "");

#define CUDA_TRY(expr)                                                                                    \
	do {                                                                                                  \
		cudaError_t err__ = (expr);                                                                       \
		if (err__ != cudaSuccess) { return ::alpb200::fail(ALPB200_ECUDA, "%s: %s", #expr, cudaGetErrorString(err__)); } \
	} while (0)

#define TRY(expr)                               \
	do {                                        \
		if (int rc__ = (expr)) { return rc__; } \
	} while (0)

constexpr int DEC_WARPS = 8;
constexpr int ENC_MIN_WARPS = 2;  // fewest vectors per encode thread block (EncodeCfg<PT>::WARPS); sizes the workspace

struct DeviceInfo {
	int sms        = 0;
	int smem_optin = 0;
	// scratch words of the decode kernels (work-distribution counter + "block outgrows the stage" flag): one PAIR of slots
	// per launch, handed out round-robin, zeroed on the launch's stream right before the kernel (allocated once per device;
	// nothing on the hot path)
	unsigned long long* counters     = nullptr;
	uint32_t            next_counter = 0;
};
int device_info(DeviceInfo& out);

// ---- launchers (explicitly instantiated for double and float in their TU) ----
template <typename PT>
int launch_decode(const alpb200_column* col, uint64_t first, uint64_t n, PT* d_out, void* stream);
template <typename PT>
int launch_decode_sum(const alpb200_column* col, uint64_t first, uint64_t n, double* d_sum, void* stream, uint32_t flags = 0);
template <typename PT>
int launch_decode_minmax(const alpb200_column* col, uint64_t first, uint64_t n, alpb200_minmax* d_out, void* stream);
template <typename PT>
int launch_decode_filter(const alpb200_column* col, uint64_t first, uint64_t n, uint32_t op, double constant, uint32_t* d_bitmap,
                         uint64_t* d_selected, void* stream);
template <typename PT, bool ORDERED>
int launch_encode_impl(const PT* d_in, uint64_t n, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws, void* stream,
                       bool append);
// the streaming (persistent, warp-specialised) form of the vector-order encoder: same workspace, same bytes (alp_encode_stream.cuh)
template <typename PT>
int launch_encode_stream(const PT* d_in, uint64_t n, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws, void* stream,
                         bool append);
// which kernel serves alpb200_encode_* (vector order): 1 = streaming pipeline (default), 0 = one thread block per 9 / 8 vectors;
// ALPB200_ENCODE_KERNEL=block|stream in the environment overrides it (development A/B, read once)
bool encode_uses_stream();
// ordered = true: blocks in vector order (alpb200_encode_*); false: completion order (alpb200_encode_unordered_*).
// append  = true: the output continues where col->totals says the column ends (meta / d_in / d_states point at the
//                 first vector of this call; packed / exc arrays and their offsets stay those of the whole column).
template <typename PT>
inline int launch_encode(const PT* d_in, uint64_t n, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws, void* stream,
                         bool ordered = true, bool append = false) {
	return ordered ? launch_encode_impl<PT, true>(d_in, n, d_states, col, ws, stream, append)
	               : launch_encode_impl<PT, false>(d_in, n, d_states, col, ws, stream, append);
}
template <typename PT>
int launch_init(const PT* d_in, uint64_t n_values, alpb200_rg_state* d_states, void* ws, void* stream, bool force_rd = false);
// fillers for NULL slots and for the slots behind n_values up to the next multiple of 1024 (alp_prims.cuh)
template <typename PT>
int launch_fill_invalid(PT* d_values, uint64_t n_values, const uint8_t* d_validity, const alpb200_rg_state* d_states, void* stream);

int validate_device(const alpb200_column* col, int value_bytes, uint64_t* h_max_block_bytes, void* stream);

size_t encode_workspace_bytes(uint64_t n_vectors);
size_t init_workspace_bytes(uint64_t n_values);

}  // namespace alpb200
