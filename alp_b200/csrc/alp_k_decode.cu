// alp_k_decode.cu — launchers of the decode kernels and the column validation (one translation unit of libalp_b200.so).
#include <algorithm>

#include "alp_decode.cuh"
#include "alp_host.h"

namespace alpb200 {

template <typename PT>
int launch_decode(const alpb200_column* col, uint64_t first, uint64_t n, PT* d_out, void* stream) {
	if (!col) { return fail(ALPB200_EINVAL, "decode: null argument"); }
	if (first + n > col->n_vectors) { return fail(ALPB200_EINVAL, "decode: vector range outside the column"); }
	if (n == 0) { return ALPB200_OK; }
	if (!d_out || !col->meta || !col->packed) { return fail(ALPB200_EINVAL, "decode: null argument"); }
	if ((reinterpret_cast<uintptr_t>(col->packed) & 127u) != 0) { return fail(ALPB200_EINVAL, "decode: column.packed must be 128-byte aligned"); }
	if ((reinterpret_cast<uintptr_t>(d_out) & 15u) != 0) { return fail(ALPB200_EINVAL, "decode: the output buffer must be 16-byte aligned"); }
	if ((reinterpret_cast<uintptr_t>(col->meta) & 15u) != 0) { return fail(ALPB200_EINVAL, "decode: column.meta must be 16-byte aligned"); }
	DeviceInfo di;
	if (int rc = device_info(di)) { return rc; }
	const uint32_t widest = (sizeof(PT) == 8 ? 66u : 35u) * 128u;
	uint32_t       block  = col->max_block_bytes ? (uint32_t)std::min<uint64_t>(col->max_block_bytes, widest) : widest;
	const uint32_t stage  = ((block + 127u) & ~127u) + STAGE_PAD;
	// narrow blocks: decode into a shared-memory tile and bulk-store it; wide blocks (ALP_RD, bw > 32): direct line stores
	const bool   tile = stage <= 4096 + STAGE_PAD;
	const size_t smem = (size_t)DEC_WARPS * ((tile ? VEC * sizeof(PT) : 0) + 2 * stage) + DEC_WARPS * 2 * sizeof(uint64_t);
	auto         kern = tile ? decode_kernel<PT, DEC_WARPS, true> : decode_kernel<PT, DEC_WARPS, false>;
	CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DEC_WARPS * 32, smem));
	if (per_sm < 1) { return fail(ALPB200_ECUDA, "decode: kernel does not fit on an SM"); }
	const uint64_t want = (n + DEC_WARPS - 1) / DEC_WARPS;
	const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)di.sms * per_sm);
	ColView        view {col->meta, col->packed, col->exc_val, col->exc_pos};
	cudaStream_t        s        = static_cast<cudaStream_t>(stream);
	unsigned long long* counter  = di.counters + di.next_counter;  // [0] work counter, [1] "a block outgrows the stage"
	unsigned long long* oversize = nullptr;
	CUDA_TRY(cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned long long), s));
	if (block < widest) {  // a hint was given: make sure it covers this call's blocks (see hint_check_kernel)
		oversize = counter + 1;
		hint_check_kernel<<<(uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)di.sms * 8), 256, 0, s>>>(col->meta + first, n, stage - STAGE_PAD, oversize);
		CUDA_TRY(cudaGetLastError());
	}
	kern<<<grid, DEC_WARPS * 32, smem, s>>>(view, first, n, d_out, stage, counter, oversize);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}

int validate_device(const alpb200_column* col, int value_bytes, uint64_t* h_max_block_bytes, void* stream) {
	if (!col || (value_bytes != 8 && value_bytes != 4)) { return fail(ALPB200_EINVAL, "column_validate_device: bad argument"); }
	if (h_max_block_bytes) { *h_max_block_bytes = 0; }
	if (col->n_vectors == 0) { return ALPB200_OK; }
	if (!col->meta || !col->packed || !col->exc_val || !col->exc_pos) { return fail(ALPB200_EINVAL, "column_validate_device: null array"); }
	if ((reinterpret_cast<uintptr_t>(col->packed) & 127u) != 0 || (reinterpret_cast<uintptr_t>(col->meta) & 15u) != 0) {
		return fail(ALPB200_EINVAL, "column_validate_device: column.packed must be 128-byte aligned, column.meta 16-byte aligned");
	}
	DeviceInfo di;
	if (int rc = device_info(di)) { return rc; }
	cudaStream_t        s      = static_cast<cudaStream_t>(stream);
	unsigned long long* result = di.counters + di.next_counter;  // a pair of slots of the per-device scratch ring
	CUDA_TRY(cudaMemsetAsync(result, 0, 2 * sizeof(unsigned long long), s));
	ColView view {col->meta, col->packed, col->exc_val, col->exc_pos};
	validate_kernel<<<(uint32_t)((col->n_vectors + 255) / 256), 256, 0, s>>>(view, col->n_vectors, col->packed_capacity, col->exc_capacity,
	                                                                        (uint32_t)value_bytes, result);
	CUDA_TRY(cudaGetLastError());
	unsigned long long h[2] = {0, 0};
	CUDA_TRY(cudaMemcpyAsync(h, result, sizeof(h), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(cudaStreamSynchronize(s));
	if (h[0] != 0) { return fail(ALPB200_EINVAL, "column_validate_device: malformed vector record (field out of range, or pointing outside the column's arrays)"); }
	if (h_max_block_bytes) { *h_max_block_bytes = h[1]; }
	return ALPB200_OK;
}

template int launch_decode<double>(const alpb200_column*, uint64_t, uint64_t, double*, void*);
template int launch_decode<float>(const alpb200_column*, uint64_t, uint64_t, float*, void*);

}  // namespace alpb200
