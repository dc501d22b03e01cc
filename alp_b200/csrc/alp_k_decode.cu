// alp_k_decode.cu — launchers of the decode and fused decode+SUM kernels (one translation unit of libalp_b200.so).
#include <algorithm>

#include "alp_decode.cuh"
#include "alp_host.h"
#include "alp_scan.cuh"

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

template <typename PT>
int launch_decode_sum(const alpb200_column* col, uint64_t first, uint64_t n, double* d_sum, void* stream, uint32_t flags) {
	if (!col || !d_sum) { return fail(ALPB200_EINVAL, "decode_sum: null argument"); }
	if (first + n > col->n_vectors) { return fail(ALPB200_EINVAL, "decode_sum: vector range outside the column"); }
	if (n == 0) { return ALPB200_OK; }
	if (!col->meta || !col->packed) { return fail(ALPB200_EINVAL, "decode_sum: null argument"); }
	if ((reinterpret_cast<uintptr_t>(col->packed) & 127u) != 0) { return fail(ALPB200_EINVAL, "decode_sum: column.packed must be 128-byte aligned"); }
	if ((reinterpret_cast<uintptr_t>(col->meta) & 15u) != 0) { return fail(ALPB200_EINVAL, "decode_sum: column.meta must be 16-byte aligned"); }
	DeviceInfo di;
	if (int rc = device_info(di)) { return rc; }
	const uint32_t widest = (sizeof(PT) == 8 ? 66u : 35u) * 128u;
	uint32_t       block  = col->max_block_bytes ? (uint32_t)std::min<uint64_t>(col->max_block_bytes, widest) : widest;
	const uint32_t stage  = ((block + 127u) & ~127u) + STAGE_PAD;
	// The scan is bound by what the resident warps can unpack, so the block shape is the one that puts the most warps on
	// an SM: 8 warps per block while two stages per warp are small (narrow ALP blocks), fewer when they are wide (ALP_RD on
	// doubles: 2 x 7.3 KiB per warp would leave ONE 8-warp block per SM; 5-warp blocks fit three).
	ColView             view {col->meta, col->packed, col->exc_val, col->exc_pos};
	unsigned long long* counter = di.counters + di.next_counter;
	cudaStream_t        s       = static_cast<cudaStream_t>(stream);
	unsigned long long* oversize = nullptr;
	int                 best_w = 0, best_per_sm = 0;
	auto consider = [&](auto Wc) -> int {
		constexpr int W    = decltype(Wc)::value;
		const size_t  smem = (size_t)W * 2 * stage + W * 2 * sizeof(uint64_t);
		if (smem > (size_t)di.smem_optin) { return ALPB200_OK; }
		auto kern = decode_sum_kernel<PT, W>;
		CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		int per_sm = 0;
		CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem));
		if (per_sm * W > best_per_sm * best_w) {
			best_w      = W;
			best_per_sm = per_sm;
		}
		return ALPB200_OK;
	};
	TRY(consider(std::integral_constant<int, 8> {}));
	TRY(consider(std::integral_constant<int, 5> {}));
	TRY(consider(std::integral_constant<int, 3> {}));
	if (best_w == 0) { return fail(ALPB200_ECUDA, "decode_sum: kernel does not fit on an SM"); }
	CUDA_TRY(cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned long long), s));
	if (block < widest) {  // a hint was given: make sure it covers this call's blocks (see hint_check_kernel)
		oversize = counter + 1;
		hint_check_kernel<<<(uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)di.sms * 8), 256, 0, s>>>(col->meta + first, n, stage - STAGE_PAD, oversize);
		CUDA_TRY(cudaGetLastError());
	}
	auto launch = [&](auto Wc) -> int {
		constexpr int  W    = decltype(Wc)::value;
		const size_t   smem = (size_t)W * 2 * stage + W * 2 * sizeof(uint64_t);
		const uint64_t want = (n + W - 1) / W;
		const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)di.sms * best_per_sm);
		decode_sum_kernel<PT, W><<<grid, W * 32, smem, s>>>(view, first, n, d_sum, stage, counter, oversize, flags);
		CUDA_TRY(cudaGetLastError());
		return ALPB200_OK;
	};
	if (best_w == 8) {
		TRY(launch(std::integral_constant<int, 8> {}));
	} else if (best_w == 5) {
		TRY(launch(std::integral_constant<int, 5> {}));
	} else {
		TRY(launch(std::integral_constant<int, 3> {}));
	}
	return ALPB200_OK;
}

template <typename PT>
int launch_decode_minmax(const alpb200_column* col, uint64_t first, uint64_t n, alpb200_minmax* d_out, void* stream) {
	static_assert(sizeof(MinMaxOut) == sizeof(alpb200_minmax), "MinMaxOut mirrors alpb200_minmax");
	if (!col || !d_out) { return fail(ALPB200_EINVAL, "decode_minmax: null argument"); }
	if (first + n > col->n_vectors) { return fail(ALPB200_EINVAL, "decode_minmax: vector range outside the column"); }
	if ((reinterpret_cast<uintptr_t>(d_out) & 7u) != 0) { return fail(ALPB200_EINVAL, "decode_minmax: the result must be 8-byte aligned"); }
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	minmax_init_kernel<<<1, 1, 0, s>>>(reinterpret_cast<MinMaxOut*>(d_out));  // min = +inf, max = -inf, count = 0
	CUDA_TRY(cudaGetLastError());
	if (n == 0) { return ALPB200_OK; }
	if (!col->meta || !col->packed) { return fail(ALPB200_EINVAL, "decode_minmax: null argument"); }
	if ((reinterpret_cast<uintptr_t>(col->packed) & 127u) != 0) { return fail(ALPB200_EINVAL, "decode_minmax: column.packed must be 128-byte aligned"); }
	if ((reinterpret_cast<uintptr_t>(col->meta) & 15u) != 0) { return fail(ALPB200_EINVAL, "decode_minmax: column.meta must be 16-byte aligned"); }
	DeviceInfo di;
	if (int rc = device_info(di)) { return rc; }
	const uint32_t widest = (sizeof(PT) == 8 ? 66u : 35u) * 128u;
	uint32_t       block  = col->max_block_bytes ? (uint32_t)std::min<uint64_t>(col->max_block_bytes, widest) : widest;
	const uint32_t stage  = ((block + 127u) & ~127u) + STAGE_PAD;
	constexpr int  W      = 8;
	const size_t   smem   = (size_t)W * 2 * stage + W * 2 * sizeof(uint64_t);
	auto           kern   = decode_minmax_kernel<PT, W>;
	CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem));
	if (per_sm < 1) { return fail(ALPB200_ECUDA, "decode_minmax: kernel does not fit on an SM"); }
	unsigned long long* counter  = di.counters + di.next_counter;
	unsigned long long* oversize = nullptr;
	CUDA_TRY(cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned long long), s));
	if (block < widest) {  // a hint was given: make sure it covers this call's blocks (see hint_check_kernel)
		oversize = counter + 1;
		hint_check_kernel<<<(uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)di.sms * 8), 256, 0, s>>>(col->meta + first, n, stage - STAGE_PAD, oversize);
		CUDA_TRY(cudaGetLastError());
	}
	ColView        view {col->meta, col->packed, col->exc_val, col->exc_pos};
	const uint32_t grid = (uint32_t)std::min<uint64_t>((n + W - 1) / W, (uint64_t)di.sms * per_sm);
	kern<<<grid, W * 32, smem, s>>>(view, first, n, reinterpret_cast<MinMaxOut*>(d_out), stage, counter, oversize);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}
template int launch_decode_minmax<double>(const alpb200_column*, uint64_t, uint64_t, alpb200_minmax*, void*);
template int launch_decode_minmax<float>(const alpb200_column*, uint64_t, uint64_t, alpb200_minmax*, void*);

template <typename PT>
int launch_decode_filter(const alpb200_column* col, uint64_t first, uint64_t n, uint32_t op, double constant, uint32_t* d_bitmap,
                         uint64_t* d_selected, void* stream) {
	if (!col || !d_bitmap) { return fail(ALPB200_EINVAL, "decode_filter: null argument"); }
	if (op > ALPB200_FILTER_NE) { return fail(ALPB200_EINVAL, "decode_filter: unknown comparison"); }
	if (first + n > col->n_vectors) { return fail(ALPB200_EINVAL, "decode_filter: vector range outside the column"); }
	if ((reinterpret_cast<uintptr_t>(d_bitmap) & 3u) != 0 || (reinterpret_cast<uintptr_t>(d_selected) & 7u) != 0) {
		return fail(ALPB200_EINVAL, "decode_filter: misaligned output");
	}
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	if (d_selected) { CUDA_TRY(cudaMemsetAsync(d_selected, 0, sizeof(uint64_t), s)); }
	if (n == 0) { return ALPB200_OK; }
	if (!col->meta || !col->packed) { return fail(ALPB200_EINVAL, "decode_filter: null argument"); }
	if ((reinterpret_cast<uintptr_t>(col->packed) & 127u) != 0) { return fail(ALPB200_EINVAL, "decode_filter: column.packed must be 128-byte aligned"); }
	if ((reinterpret_cast<uintptr_t>(col->meta) & 15u) != 0) { return fail(ALPB200_EINVAL, "decode_filter: column.meta must be 16-byte aligned"); }
	DeviceInfo di;
	if (int rc = device_info(di)) { return rc; }
	const uint32_t widest = (sizeof(PT) == 8 ? 66u : 35u) * 128u;
	uint32_t       block  = col->max_block_bytes ? (uint32_t)std::min<uint64_t>(col->max_block_bytes, widest) : widest;
	const uint32_t stage  = ((block + 127u) & ~127u) + STAGE_PAD;
	constexpr int  W      = 8;
	const size_t   smem   = (size_t)W * 2 * stage + W * 2 * sizeof(uint64_t);
	auto           kern   = decode_filter_kernel<PT, W>;
	CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem));
	if (per_sm < 1) { return fail(ALPB200_ECUDA, "decode_filter: kernel does not fit on an SM"); }
	unsigned long long* counter  = di.counters + di.next_counter;
	unsigned long long* oversize = nullptr;
	CUDA_TRY(cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned long long), s));
	if (block < widest) {  // a hint was given: make sure it covers this call's blocks (see hint_check_kernel)
		oversize = counter + 1;
		hint_check_kernel<<<(uint32_t)std::min<uint64_t>((n + 255) / 256, (uint64_t)di.sms * 8), 256, 0, s>>>(col->meta + first, n, stage - STAGE_PAD, oversize);
		CUDA_TRY(cudaGetLastError());
	}
	ColView        view {col->meta, col->packed, col->exc_val, col->exc_pos};
	const uint32_t grid = (uint32_t)std::min<uint64_t>((n + W - 1) / W, (uint64_t)di.sms * per_sm);
	kern<<<grid, W * 32, smem, s>>>(view, first, n, op, constant, d_bitmap, reinterpret_cast<unsigned long long*>(d_selected), stage, counter, oversize);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}
template int launch_decode_filter<double>(const alpb200_column*, uint64_t, uint64_t, uint32_t, double, uint32_t*, uint64_t*, void*);
template int launch_decode_filter<float>(const alpb200_column*, uint64_t, uint64_t, uint32_t, double, uint32_t*, uint64_t*, void*);

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
template int launch_decode_sum<double>(const alpb200_column*, uint64_t, uint64_t, double*, void*, uint32_t);
template int launch_decode_sum<float>(const alpb200_column*, uint64_t, uint64_t, double*, void*, uint32_t);

}  // namespace alpb200
