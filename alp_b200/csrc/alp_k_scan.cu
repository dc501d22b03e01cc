// alp_k_scan.cu — launchers of the fused scan kernels: decode + SUM, decode + MIN / MAX / COUNT, decode + predicate filter
// (one translation unit of libalp_b200.so; split from alp_k_decode.cu so that nvcc compiles the two in parallel).
#include <algorithm>

#include "alp_decode.cuh"
#include "alp_host.h"
#include "alp_scan.cuh"

namespace alpb200 {

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

template int launch_decode_sum<double>(const alpb200_column*, uint64_t, uint64_t, double*, void*, uint32_t);
template int launch_decode_sum<float>(const alpb200_column*, uint64_t, uint64_t, double*, void*, uint32_t);

}  // namespace alpb200
