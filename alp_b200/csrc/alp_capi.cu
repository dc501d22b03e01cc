// alp_capi.cu — the C ABI of include/alp_b200.h: argument checks, launches, the host-buffer codec context and the
// single-vector primitive wrappers.  No CPU fallback anywhere: without a CUDA device every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "alp_b200.h"
#include "alp_decode.cuh"
#include "alp_encode.cuh"
#include "alp_init.cuh"
#include "alp_prims.cuh"
#include "alp_scan.cuh"

static_assert(sizeof(alpb200_rg_state) == 1196, "alpb200_rg_state layout");
static_assert(sizeof(alpb200_vec_meta) == 32, "alpb200_vec_meta layout");
static_assert(sizeof(alpb200_column) == 80, "alpb200_column layout");

using namespace alpb200;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
	char buf[512];
	snprintf(buf, sizeof(buf), fmt, a, b);
	g_last_error = buf;
	return code;
}

#define CUDA_TRY(expr)                                                                        \
	do {                                                                                      \
		cudaError_t err__ = (expr);                                                           \
		if (err__ != cudaSuccess) { return fail(ALPB200_ECUDA, "%s: %s", #expr, cudaGetErrorString(err__)); } \
	} while (0)

constexpr int DEC_WARPS = 8;
constexpr int ENC_WARPS = 8;

struct DeviceInfo {
	int sms = 0;
	int smem_optin = 0;
	// work-distribution counters of the decode kernel: one slot per launch, handed out round-robin, zeroed on the
	// launch's stream right before the kernel (the library allocates this once per device; nothing on the hot path)
	unsigned long long* counters = nullptr;
	uint32_t            next_counter = 0;
};
constexpr uint32_t N_COUNTERS = 4096;
DeviceInfo g_dev[64];
std::mutex g_dev_mutex;

int device_info(DeviceInfo& out) {
	int dev = 0;
	CUDA_TRY(cudaGetDevice(&dev));
	std::lock_guard<std::mutex> lock(g_dev_mutex);
	if (dev < 0 || dev >= 64) { return fail(ALPB200_EINVAL, "device index out of range"); }
	if (g_dev[dev].sms == 0) {
		CUDA_TRY(cudaDeviceGetAttribute(&g_dev[dev].sms, cudaDevAttrMultiProcessorCount, dev));
		CUDA_TRY(cudaDeviceGetAttribute(&g_dev[dev].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
		CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&g_dev[dev].counters), N_COUNTERS * sizeof(unsigned long long)));
	}
	g_dev[dev].next_counter = (g_dev[dev].next_counter + 1) % N_COUNTERS;
	out = g_dev[dev];
	return ALPB200_OK;
}

template <typename PT>
int launch_decode(const alpb200_column* col, uint64_t first, uint64_t n, PT* d_out, void* stream) {
	if (!col) { return fail(ALPB200_EINVAL, "decode: null argument"); }
	if (first + n > col->n_vectors) { return fail(ALPB200_EINVAL, "decode: vector range outside the column"); }
	if (n == 0) { return ALPB200_OK; }
	if (!d_out || !col->meta || !col->packed) { return fail(ALPB200_EINVAL, "decode: null argument"); }
	if ((reinterpret_cast<uintptr_t>(col->packed) & 127u) != 0) { return fail(ALPB200_EINVAL, "decode: column.packed must be 128-byte aligned"); }
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
	unsigned long long* counter = di.counters + di.next_counter;
	CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), static_cast<cudaStream_t>(stream)));
	kern<<<grid, DEC_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(view, first, n, d_out, stage, counter);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}

template <typename PT>
int launch_decode_sum(const alpb200_column* col, uint64_t first, uint64_t n, double* d_sum, void* stream) {
	if (!col || !d_sum) { return fail(ALPB200_EINVAL, "decode_sum: null argument"); }
	if (first + n > col->n_vectors) { return fail(ALPB200_EINVAL, "decode_sum: vector range outside the column"); }
	if (n == 0) { return ALPB200_OK; }
	if (!col->meta || !col->packed) { return fail(ALPB200_EINVAL, "decode_sum: null argument"); }
	if ((reinterpret_cast<uintptr_t>(col->packed) & 127u) != 0) { return fail(ALPB200_EINVAL, "decode_sum: column.packed must be 128-byte aligned"); }
	DeviceInfo di;
	if (int rc = device_info(di)) { return rc; }
	const uint32_t widest = (sizeof(PT) == 8 ? 66u : 35u) * 128u;
	uint32_t       block  = col->max_block_bytes ? (uint32_t)std::min<uint64_t>(col->max_block_bytes, widest) : widest;
	const uint32_t stage  = ((block + 127u) & ~127u) + STAGE_PAD;
	const size_t   smem   = (size_t)DEC_WARPS * 2 * stage + DEC_WARPS * 2 * sizeof(uint64_t);
	auto           kern   = decode_sum_kernel<PT, DEC_WARPS>;
	CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int per_sm = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DEC_WARPS * 32, smem));
	if (per_sm < 1) { return fail(ALPB200_ECUDA, "decode_sum: kernel does not fit on an SM"); }
	const uint64_t want = (n + DEC_WARPS - 1) / DEC_WARPS;
	const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)di.sms * per_sm);
	ColView        view {col->meta, col->packed, col->exc_val, col->exc_pos};
	unsigned long long* counter = di.counters + di.next_counter;
	CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), static_cast<cudaStream_t>(stream)));
	kern<<<grid, DEC_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(view, first, n, d_sum, stage, counter);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}

size_t encode_workspace_bytes(uint64_t n_vectors) {
	const uint64_t blocks = (n_vectors + ENC_WARPS - 1) / ENC_WARPS;
	return (size_t)((2 + 2 * blocks) * sizeof(uint64_t) + 255) & ~(size_t)255;
}

template <typename PT>
int launch_encode(const PT* d_in, uint64_t n, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws, void* stream) {
	if (!d_in || !d_states || !col || !ws || !col->meta || !col->packed || !col->exc_val || !col->exc_pos || !col->totals) {
		return fail(ALPB200_EINVAL, "encode: null argument");
	}
	if (n > col->n_vectors) { return fail(ALPB200_EINVAL, "encode: column.n_vectors is smaller than n_vectors"); }
	if (n > (1ull << 22)) { return fail(ALPB200_EINVAL, "encode: at most 2^22 vectors (2^32 values) per call"); }
	if ((reinterpret_cast<uintptr_t>(col->packed) & 127u) != 0) { return fail(ALPB200_EINVAL, "encode: column.packed must be 128-byte aligned"); }
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	CUDA_TRY(cudaMemsetAsync(ws, 0, encode_workspace_bytes(n), s));
	CUDA_TRY(cudaMemsetAsync(col->totals, 0, 4 * sizeof(uint64_t), s));
	if (n == 0) { return ALPB200_OK; }
	constexpr size_t smem = (size_t)ENC_WARPS * EncodeCfg<PT>::SMEM_PER_WARP;  // f64: one tile per warp; f32: one stage per warp
	auto             kern = encode_kernel<PT, ENC_WARPS>;
	CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	ColOut         out {col->meta, col->packed, col->packed_capacity, col->exc_val, col->exc_pos, col->exc_capacity, col->totals};
	const uint32_t grid = (uint32_t)((n + ENC_WARPS - 1) / ENC_WARPS);
	kern<<<grid, ENC_WARPS * 32, smem, s>>>(d_in, n, d_states, out, static_cast<uint64_t*>(ws));
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}

size_t init_workspace_bytes(uint64_t n_values) {
	const uint64_t n_vec = n_values / VEC;
	const uint64_t n_rg  = (n_vec + ALPB200_ROWGROUP_VECTORS - 1) / ALPB200_ROWGROUP_VECTORS;
	return (size_t)(n_rg * MAX_SAMPLED_VECS * sizeof(SearchResult) + 255) & ~(size_t)255;
}

template <typename PT>
int launch_init(const PT* d_in, uint64_t n_values, alpb200_rg_state* d_states, void* ws, void* stream) {
	if (!d_in || !d_states || !ws) { return fail(ALPB200_EINVAL, "rowgroup_init: null argument"); }
	if (n_values % VEC != 0) { return fail(ALPB200_EINVAL, "rowgroup_init: n_values must be a multiple of 1024"); }
	const uint64_t n_vec = n_values / VEC;
	if (n_vec == 0) { return ALPB200_OK; }
	const uint64_t n_rg = (n_vec + ALPB200_ROWGROUP_VECTORS - 1) / ALPB200_ROWGROUP_VECTORS;
	cudaStream_t   s    = static_cast<cudaStream_t>(stream);
	constexpr int  W    = 4;
	const uint64_t jobs = n_rg * MAX_SAMPLED_VECS;
	init_search_kernel<PT, W><<<(uint32_t)((jobs + W - 1) / W), W * 32, 0, s>>>(d_in, n_vec, n_rg, static_cast<SearchResult*>(ws));
	CUDA_TRY(cudaGetLastError());
	init_finalize_kernel<PT, W><<<(uint32_t)((n_rg + W - 1) / W), W * 32, 0, s>>>(d_in, n_vec, n_rg, static_cast<const SearchResult*>(ws),
	                                                                              d_states);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}

// ---- tiny RAII device buffer for the single-vector primitives ----
struct DevBuf {
	void* p = nullptr;
	~DevBuf() {
		if (p) { cudaFree(p); }
	}
	int alloc(size_t bytes) {
		CUDA_TRY(cudaMalloc(&p, bytes ? bytes : 16));
		return ALPB200_OK;
	}
	int upload(const void* h, size_t bytes) {
		if (int rc = alloc(bytes)) { return rc; }
		if (bytes) { CUDA_TRY(cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice)); }
		return ALPB200_OK;
	}
	int download(void* h, size_t bytes) const {
		if (bytes) { CUDA_TRY(cudaMemcpy(h, p, bytes, cudaMemcpyDeviceToHost)); }
		return ALPB200_OK;
	}
	template <typename T>
	T* as() const {
		return static_cast<T*>(p);
	}
};

#define TRY(expr)                    \
	do {                             \
		if (int rc__ = (expr)) { return rc__; } \
	} while (0)

int finish_kernel() {
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaDeviceSynchronize());
	return ALPB200_OK;
}

template <typename PT>
int prim_encode(const PT* h_in, const alpb200_rg_state* h_state, PT* h_exc, uint16_t* h_pos, uint16_t* h_cnt,
                typename Traits<PT>::ST* h_enc, uint8_t* e, uint8_t* f) {
	using UT = typename Traits<PT>::UT;
	if (!h_in || !h_state || !h_exc || !h_pos || !h_cnt || !h_enc || !e || !f) { return fail(ALPB200_EINVAL, "prim_encode: null argument"); }
	if (h_state->scheme != ALPB200_SCHEME_ALP || h_state->k < 1 || h_state->k > ALPB200_MAX_K) {
		return fail(ALPB200_EINVAL, "prim_encode: the state is not an ALP state with 1..5 combinations");
	}
	DevBuf in, st, exc, pos, cnt, enc, ef;
	TRY(in.upload(h_in, VEC * sizeof(PT)));
	TRY(st.upload(h_state, sizeof(*h_state)));
	TRY(exc.alloc(VEC * sizeof(PT)));
	TRY(pos.alloc(VEC * 2));
	TRY(cnt.alloc(16));
	TRY(enc.alloc(VEC * sizeof(PT)));
	TRY(ef.alloc(16));
	prim_encode_kernel<PT><<<1, 32>>>(in.as<PT>(), st.as<alpb200_rg_state>(), exc.as<UT>(), pos.as<uint16_t>(), cnt.as<uint16_t>(),
	                                  enc.as<UT>(), ef.as<uint8_t>());
	TRY(finish_kernel());
	uint8_t ef_h[2];
	TRY(cnt.download(h_cnt, 2));
	TRY(ef.download(ef_h, 2));
	TRY(enc.download(h_enc, VEC * sizeof(PT)));
	TRY(exc.download(h_exc, (size_t)h_cnt[0] * sizeof(PT)));
	TRY(pos.download(h_pos, (size_t)h_cnt[0] * 2));
	*e = ef_h[0];
	*f = ef_h[1];
	return ALPB200_OK;
}

template <typename PT>
int prim_analyze(const typename Traits<PT>::ST* h_enc, uint8_t* bw, typename Traits<PT>::ST* base) {
	using ST = typename Traits<PT>::ST;
	if (!h_enc || !bw || !base) { return fail(ALPB200_EINVAL, "prim_analyze_ffor: null argument"); }
	DevBuf enc, b, s;
	TRY(enc.upload(h_enc, VEC * sizeof(ST)));
	TRY(b.alloc(16));
	TRY(s.alloc(16));
	prim_analyze_ffor_kernel<PT><<<1, 32>>>(enc.as<ST>(), b.as<uint8_t>(), s.as<ST>());
	TRY(finish_kernel());
	TRY(b.download(bw, 1));
	TRY(s.download(base, sizeof(ST)));
	return ALPB200_OK;
}

template <typename PT>
int prim_ffor(const typename Traits<PT>::UT* h_in, typename Traits<PT>::UT* h_out, uint8_t bw, typename Traits<PT>::UT base) {
	using UT = typename Traits<PT>::UT;
	if (!h_in || !h_out) { return fail(ALPB200_EINVAL, "prim_ffor: null argument"); }
	if (bw > 8 * sizeof(UT)) { return fail(ALPB200_EINVAL, "prim_ffor: bit width exceeds the lane width"); }
	DevBuf in, out;
	TRY(in.upload(h_in, VEC * sizeof(UT)));
	TRY(out.alloc(128u * 64u));
	prim_ffor_kernel<PT><<<1, 32>>>(in.as<UT>(), out.as<uint8_t>(), bw, base);
	TRY(finish_kernel());
	TRY(out.download(h_out, 128u * bw));
	return ALPB200_OK;
}

template <typename PT>
int prim_unffor(const typename Traits<PT>::UT* h_in, typename Traits<PT>::UT* h_out, uint8_t bw, typename Traits<PT>::UT base) {
	using UT = typename Traits<PT>::UT;
	if (!h_in || !h_out) { return fail(ALPB200_EINVAL, "prim_unffor: null argument"); }
	if (bw > 8 * sizeof(UT)) { return fail(ALPB200_EINVAL, "prim_unffor: bit width exceeds the lane width"); }
	DevBuf in, out;
	TRY(in.upload(h_in, 128u * bw));
	TRY(out.alloc(VEC * sizeof(UT)));
	prim_unffor_kernel<PT><<<1, 32>>>(in.as<uint8_t>(), out.as<UT>(), bw, base);
	TRY(finish_kernel());
	TRY(out.download(h_out, VEC * sizeof(UT)));
	return ALPB200_OK;
}

template <typename PT>
int prim_falp(const typename Traits<PT>::UT* h_packed, PT* h_out, uint8_t bw, typename Traits<PT>::UT base, uint8_t f, uint8_t e) {
	using T = Traits<PT>;
	if (!h_packed || !h_out) { return fail(ALPB200_EINVAL, "prim_falp: null argument"); }
	if (bw > T::TBITS || e > T::MAX_EXP || f > e) { return fail(ALPB200_EINVAL, "prim_falp: bit width / exponent / factor out of range"); }
	DevBuf in, out;
	TRY(in.upload(h_packed, 128u * bw));
	TRY(out.alloc(VEC * sizeof(PT)));
	prim_falp_kernel<PT><<<1, 32>>>(in.as<uint8_t>(), out.as<PT>(), bw, base, f, e);
	TRY(finish_kernel());
	TRY(out.download(h_out, VEC * sizeof(PT)));
	return ALPB200_OK;
}

template <typename PT>
int prim_decode(const typename Traits<PT>::ST* h_enc, uint8_t f, uint8_t e, PT* h_out) {
	using T  = Traits<PT>;
	using ST = typename T::ST;
	if (!h_enc || !h_out) { return fail(ALPB200_EINVAL, "prim_decode: null argument"); }
	if (e > T::MAX_EXP || f > e) { return fail(ALPB200_EINVAL, "prim_decode: exponent / factor out of range"); }
	DevBuf in, out;
	TRY(in.upload(h_enc, VEC * sizeof(ST)));
	TRY(out.alloc(VEC * sizeof(PT)));
	prim_decode_kernel<PT><<<1, 32>>>(in.as<ST>(), f, e, out.as<PT>());
	TRY(finish_kernel());
	TRY(out.download(h_out, VEC * sizeof(PT)));
	return ALPB200_OK;
}

template <typename PT>
int prim_patch(PT* h_out, const PT* h_exc, const uint16_t* h_pos, uint16_t cnt) {
	using UT = typename Traits<PT>::UT;
	if (!h_out || (cnt && (!h_exc || !h_pos))) { return fail(ALPB200_EINVAL, "prim_patch: null argument"); }
	for (uint16_t i = 0; i < cnt; i++) {
		if (h_pos[i] >= VEC) { return fail(ALPB200_EINVAL, "prim_patch: exception position outside the vector"); }
	}
	DevBuf out, exc, pos;
	TRY(out.upload(h_out, VEC * sizeof(PT)));
	TRY(exc.upload(h_exc, (size_t)cnt * sizeof(PT)));
	TRY(pos.upload(h_pos, (size_t)cnt * 2));
	prim_patch_kernel<UT><<<1, 32>>>(out.as<UT>(), exc.as<UT>(), pos.as<uint16_t>(), cnt);
	TRY(finish_kernel());
	TRY(out.download(h_out, VEC * sizeof(PT)));
	return ALPB200_OK;
}

int check_rd_state(const alpb200_rg_state* st, int tbits, const char* who) {
	if (st->scheme != ALPB200_SCHEME_ALP_RD || st->right_bw >= tbits || st->right_bw < tbits - 16 || st->left_bw < 1 || st->left_bw > 3 ||
	    st->dict_size < 1 || st->dict_size > ALPB200_RD_DICT_SIZE || st->n_extra > ALPB200_MAX_SAMPLES) {
		return fail(ALPB200_EINVAL, "%s: the state is not a valid ALP_RD state", who);
	}
	return ALPB200_OK;
}

template <typename PT>
int prim_rd_encode(const PT* h_in, const alpb200_rg_state* h_state, uint16_t* h_exc, uint16_t* h_pos, uint16_t* h_cnt,
                   typename Traits<PT>::UT* h_right, uint16_t* h_left) {
	using UT = typename Traits<PT>::UT;
	if (!h_in || !h_state || !h_exc || !h_pos || !h_cnt || !h_right || !h_left) { return fail(ALPB200_EINVAL, "prim_rd_encode: null argument"); }
	TRY(check_rd_state(h_state, Traits<PT>::TBITS, "prim_rd_encode"));
	DevBuf in, st, exc, pos, cnt, right, left;
	TRY(in.upload(h_in, VEC * sizeof(PT)));
	TRY(st.upload(h_state, sizeof(*h_state)));
	TRY(exc.alloc(VEC * 2));
	TRY(pos.alloc(VEC * 2));
	TRY(cnt.alloc(16));
	TRY(right.alloc(VEC * sizeof(UT)));
	TRY(left.alloc(VEC * 2));
	prim_rd_encode_kernel<PT><<<1, 32>>>(in.as<PT>(), st.as<alpb200_rg_state>(), exc.as<uint16_t>(), pos.as<uint16_t>(), cnt.as<uint16_t>(),
	                                     right.as<UT>(), left.as<uint16_t>());
	TRY(finish_kernel());
	TRY(cnt.download(h_cnt, 2));
	TRY(right.download(h_right, VEC * sizeof(UT)));
	TRY(left.download(h_left, VEC * 2));
	TRY(exc.download(h_exc, (size_t)h_cnt[0] * 2));
	TRY(pos.download(h_pos, (size_t)h_cnt[0] * 2));
	return ALPB200_OK;
}

template <typename PT>
int prim_rd_decode(PT* h_out, const typename Traits<PT>::UT* h_right, const uint16_t* h_left, const uint16_t* h_exc, const uint16_t* h_pos,
                   uint16_t cnt, const alpb200_rg_state* h_state) {
	using UT = typename Traits<PT>::UT;
	if (!h_out || !h_right || !h_left || !h_state || (cnt && (!h_exc || !h_pos))) { return fail(ALPB200_EINVAL, "prim_rd_decode: null argument"); }
	TRY(check_rd_state(h_state, Traits<PT>::TBITS, "prim_rd_decode"));
	for (uint16_t i = 0; i < cnt; i++) {
		if (h_pos[i] >= VEC) { return fail(ALPB200_EINVAL, "prim_rd_decode: exception position outside the vector"); }
	}
	DevBuf out, right, left, exc, pos, st;
	TRY(out.alloc(VEC * sizeof(PT)));
	TRY(right.upload(h_right, VEC * sizeof(UT)));
	TRY(left.upload(h_left, VEC * 2));
	TRY(exc.upload(h_exc, (size_t)cnt * 2));
	TRY(pos.upload(h_pos, (size_t)cnt * 2));
	TRY(st.upload(h_state, sizeof(*h_state)));
	prim_rd_decode_kernel<PT><<<1, 32>>>(out.as<UT>(), right.as<UT>(), left.as<uint16_t>(), exc.as<uint16_t>(), pos.as<uint16_t>(), cnt,
	                                     st.as<alpb200_rg_state>());
	TRY(finish_kernel());
	TRY(out.download(h_out, VEC * sizeof(PT)));
	return ALPB200_OK;
}

template <typename PT>
int prim_init(const PT* h_col, uint64_t offset, uint64_t n_values, alpb200_rg_state* h_state) {
	if (!h_col || !h_state || offset >= n_values) { return fail(ALPB200_EINVAL, "prim_init: bad argument"); }
	const uint64_t span = std::min<uint64_t>(ALPB200_ROWGROUP_SIZE, n_values - offset) / VEC * VEC;
	if (span == 0) { return fail(ALPB200_EINVAL, "prim_init: the row-group holds no complete vector"); }
	DevBuf in, st, ws;
	TRY(in.upload(h_col + offset, span * sizeof(PT)));
	TRY(st.alloc(sizeof(alpb200_rg_state)));
	TRY(ws.alloc(init_workspace_bytes(span)));
	TRY(launch_init<PT>(in.as<PT>(), span, st.as<alpb200_rg_state>(), ws.p, nullptr));
	TRY(finish_kernel());
	TRY(st.download(h_state, sizeof(*h_state)));
	return ALPB200_OK;
}

}  // namespace

// =====================================================================================================================
// Host-buffer codec context
// =====================================================================================================================
struct alpb200_ctx {
	int          device      = 0;
	int          value_bytes = 8;
	uint64_t     max_vectors = 0;
	cudaStream_t streams[3]  = {nullptr, nullptr, nullptr};
	cudaEvent_t  events[3]   = {nullptr, nullptr, nullptr};
	// device-side column + value buffers, sized for max_vectors
	void*             d_values = nullptr;
	alpb200_vec_meta* d_meta   = nullptr;
	uint8_t*          d_packed = nullptr;
	void*             d_exc_val = nullptr;
	uint16_t*         d_exc_pos = nullptr;
	uint64_t*         d_totals  = nullptr;
	alpb200_rg_state* d_states  = nullptr;
	void*             d_ws_enc  = nullptr;
	void*             d_ws_init = nullptr;
	uint64_t          packed_capacity = 0, exc_capacity = 0;
	uint64_t*         h_totals = nullptr;  // pinned
};

namespace {

void ctx_release(alpb200_ctx* c) {
	if (!c) { return; }
	cudaSetDevice(c->device);
	for (int i = 0; i < 3; i++) {
		if (c->streams[i]) { cudaStreamDestroy(c->streams[i]); }
		if (c->events[i]) { cudaEventDestroy(c->events[i]); }
	}
	cudaFree(c->d_values);
	cudaFree(c->d_meta);
	cudaFree(c->d_packed);
	cudaFree(c->d_exc_val);
	cudaFree(c->d_exc_pos);
	cudaFree(c->d_totals);
	cudaFree(c->d_states);
	cudaFree(c->d_ws_enc);
	cudaFree(c->d_ws_init);
	if (c->h_totals) { cudaFreeHost(c->h_totals); }
	delete c;
}

// tail vector: the values after n_values up to the next multiple of 1024 repeat the last value, which
// keeps the vector's (e,f) choice and bit width what the real values ask for
template <typename PT>
__global__ void pad_tail_kernel(PT* __restrict__ v, uint64_t n_values, uint64_t n_padded) {
	const PT last = v[n_values - 1];
	for (uint64_t i = n_values + threadIdx.x; i < n_padded; i += blockDim.x) {
		v[i] = last;
	}
}

template <typename PT>
int compress_host(alpb200_ctx* c, const PT* h_in, uint64_t n_values_in, alpb200_column* h_col) {
	if (!c || !h_in || !h_col || !h_col->meta || !h_col->packed || !h_col->exc_val || !h_col->exc_pos || !h_col->totals) {
		return fail(ALPB200_EINVAL, "compress_host: null argument");
	}
	if (c->value_bytes != (int)sizeof(PT)) { return fail(ALPB200_EINVAL, "compress_host: context was created for another value width"); }
	const uint64_t n_vec    = (n_values_in + VEC - 1) / VEC;
	const uint64_t n_values = n_vec * VEC;  // padded length
	if (n_vec > c->max_vectors || n_vec > h_col->n_vectors) { return fail(ALPB200_EINVAL, "compress_host: column larger than the context / container"); }
	CUDA_TRY(cudaSetDevice(c->device));
	cudaStream_t s = c->streams[0];
	if (n_vec == 0) {
		h_col->n_vectors = 0;
		h_col->n_values  = 0;
		std::memset(h_col->totals, 0, 4 * sizeof(uint64_t));
		return ALPB200_OK;
	}
	CUDA_TRY(cudaMemcpyAsync(c->d_values, h_in, n_values_in * sizeof(PT), cudaMemcpyHostToDevice, s));
	if (n_values != n_values_in) {
		pad_tail_kernel<PT><<<1, 256, 0, s>>>(static_cast<PT*>(c->d_values), n_values_in, n_values);
		CUDA_TRY(cudaGetLastError());
	}
	TRY(launch_init<PT>(static_cast<const PT*>(c->d_values), n_values, c->d_states, c->d_ws_init, s));
	alpb200_column d_col {};
	d_col.n_vectors       = n_vec;
	d_col.meta            = c->d_meta;
	d_col.packed          = c->d_packed;
	d_col.packed_capacity = c->packed_capacity;
	d_col.exc_val         = c->d_exc_val;
	d_col.exc_pos         = c->d_exc_pos;
	d_col.exc_capacity    = c->exc_capacity;
	d_col.totals          = c->d_totals;
	TRY(launch_encode<PT>(static_cast<const PT*>(c->d_values), n_vec, c->d_states, &d_col, c->d_ws_enc, s));
	CUDA_TRY(cudaMemcpyAsync(c->h_totals, c->d_totals, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(cudaMemcpyAsync(h_col->meta, c->d_meta, n_vec * sizeof(alpb200_vec_meta), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(cudaStreamSynchronize(s));
	const uint64_t packed_bytes = c->h_totals[0], n_exc = c->h_totals[1];
	if (c->h_totals[2] != 0) { return fail(ALPB200_ECAPACITY, "compress_host: device staging capacity exceeded"); }
	if (packed_bytes > h_col->packed_capacity || n_exc > h_col->exc_capacity) {
		return fail(ALPB200_ECAPACITY, "compress_host: the host column container is too small");
	}
	CUDA_TRY(cudaMemcpyAsync(h_col->packed, c->d_packed, packed_bytes, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(cudaMemcpyAsync(h_col->exc_val, c->d_exc_val, n_exc * sizeof(PT), cudaMemcpyDeviceToHost, c->streams[1]));
	CUDA_TRY(cudaMemcpyAsync(h_col->exc_pos, c->d_exc_pos, n_exc * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->streams[1]));
	CUDA_TRY(cudaStreamSynchronize(s));
	CUDA_TRY(cudaStreamSynchronize(c->streams[1]));
	h_col->n_vectors       = n_vec;
	h_col->n_values        = n_values_in;
	h_col->max_block_bytes = c->h_totals[3];
	std::memcpy(h_col->totals, c->h_totals, 4 * sizeof(uint64_t));
	return ALPB200_OK;
}

// Chunked pipeline: while chunk i decodes and drains to the host on one stream, chunk i+1's compressed bytes are
// already travelling on another (PCIe is full duplex).  Chunks are whole ranges of vectors; thanks to the vector-order
// layout each chunk's packed bytes and exceptions are one contiguous range.
template <typename PT>
int decompress_host(alpb200_ctx* c, const alpb200_column* h_col, PT* h_out) {
	if (!c || !h_col || !h_out || !h_col->meta) { return fail(ALPB200_EINVAL, "decompress_host: null argument"); }
	if (c->value_bytes != (int)sizeof(PT)) { return fail(ALPB200_EINVAL, "decompress_host: context was created for another value width"); }
	const uint64_t n_vec = h_col->n_vectors;
	if (n_vec > c->max_vectors) { return fail(ALPB200_EINVAL, "decompress_host: column larger than the context"); }
	if (n_vec == 0) { return ALPB200_OK; }
	const uint64_t n_out = h_col->n_values ? h_col->n_values : n_vec * VEC;  // a padded tail is not returned
	if (n_out > n_vec * VEC || n_out + VEC <= n_vec * VEC) { return fail(ALPB200_EINVAL, "decompress_host: n_values does not match n_vectors"); }
	CUDA_TRY(cudaSetDevice(c->device));
	const alpb200_vec_meta* hm = h_col->meta;
	auto block_units = [](const alpb200_vec_meta& m) -> uint64_t { return m.scheme == ALPB200_SCHEME_ALP_RD ? (uint64_t)m.bw + m.e : m.bw; };
	const alpb200_vec_meta& last = hm[n_vec - 1];
	const uint64_t total_packed  = ((uint64_t)last.packed_off + block_units(last)) * 128ull;
	const uint64_t total_exc     = (uint64_t)last.exc_off + last.exc_cnt;
	if (total_packed > c->packed_capacity || total_exc > c->exc_capacity) { return fail(ALPB200_ECAPACITY, "decompress_host: column exceeds the context's staging capacity"); }

	alpb200_column d_col {};
	d_col.n_vectors       = n_vec;
	d_col.meta            = c->d_meta;
	d_col.packed          = c->d_packed;
	d_col.exc_val         = c->d_exc_val;
	d_col.exc_pos         = c->d_exc_pos;
	d_col.max_block_bytes = h_col->max_block_bytes;

	const uint64_t chunk = std::max<uint64_t>(1024, (n_vec + 15) / 16);  // ~16 chunks, at least 8 MiB of f64 output each
	int            si    = 0;
	for (uint64_t v0 = 0; v0 < n_vec; v0 += chunk, si = (si + 1) % 3) {
		const uint64_t v1 = std::min(n_vec, v0 + chunk);
		cudaStream_t   s  = c->streams[si];
		const uint64_t p0 = (uint64_t)hm[v0].packed_off * 128ull, e0 = hm[v0].exc_off;
		const uint64_t p1 = v1 < n_vec ? (uint64_t)hm[v1].packed_off * 128ull : total_packed;
		const uint64_t e1 = v1 < n_vec ? hm[v1].exc_off : total_exc;
		CUDA_TRY(cudaMemcpyAsync(c->d_meta + v0, hm + v0, (v1 - v0) * sizeof(alpb200_vec_meta), cudaMemcpyHostToDevice, s));
		if (p1 > p0) { CUDA_TRY(cudaMemcpyAsync(c->d_packed + p0, h_col->packed + p0, p1 - p0, cudaMemcpyHostToDevice, s)); }
		if (e1 > e0) {
			CUDA_TRY(cudaMemcpyAsync(static_cast<PT*>(c->d_exc_val) + e0, static_cast<const PT*>(h_col->exc_val) + e0, (e1 - e0) * sizeof(PT),
			                         cudaMemcpyHostToDevice, s));
			CUDA_TRY(cudaMemcpyAsync(c->d_exc_pos + e0, h_col->exc_pos + e0, (e1 - e0) * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
		}
		PT* d_out = static_cast<PT*>(c->d_values) + v0 * VEC;
		TRY(launch_decode<PT>(&d_col, v0, v1 - v0, d_out, s));
		const uint64_t n_copy = std::min<uint64_t>((v1 - v0) * VEC, n_out - v0 * VEC);
		CUDA_TRY(cudaMemcpyAsync(h_out + v0 * VEC, d_out, n_copy * sizeof(PT), cudaMemcpyDeviceToHost, s));
	}
	for (int i = 0; i < 3; i++) {
		CUDA_TRY(cudaStreamSynchronize(c->streams[i]));
	}
	return ALPB200_OK;
}

}  // namespace

// =====================================================================================================================
// extern "C"
// =====================================================================================================================
extern "C" {

int alpb200_version(void) { return 100; }

void alpb200_abi_sizes(uint32_t out[3]) {
	out[0] = sizeof(alpb200_rg_state);
	out[1] = sizeof(alpb200_vec_meta);
	out[2] = sizeof(alpb200_column);
}

const char* alpb200_last_error(void) { return g_last_error.c_str(); }

int alpb200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return fail(ALPB200_ENODEVICE, "no CUDA device is visible; alp_b200 has no CPU fallback");
	}
	return n;
}

size_t alpb200_init_workspace_bytes(uint64_t n_values) { return init_workspace_bytes(n_values); }
int alpb200_rowgroup_init_f64(const double* d_in, uint64_t n_values, alpb200_rg_state* d_states, void* ws, void* stream) {
	return launch_init<double>(d_in, n_values, d_states, ws, stream);
}
int alpb200_rowgroup_init_f32(const float* d_in, uint64_t n_values, alpb200_rg_state* d_states, void* ws, void* stream) {
	return launch_init<float>(d_in, n_values, d_states, ws, stream);
}

size_t alpb200_encode_workspace_bytes(uint64_t n_vectors) { return encode_workspace_bytes(n_vectors); }
int alpb200_encode_f64(const double* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws,
                       void* stream) {
	return launch_encode<double>(d_in, n_vectors, d_states, col, ws, stream);
}
int alpb200_encode_f32(const float* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws,
                       void* stream) {
	return launch_encode<float>(d_in, n_vectors, d_states, col, ws, stream);
}

int alpb200_decode_f64(const alpb200_column* col, uint64_t first, uint64_t n, double* d_out, void* stream) {
	// d_out receives vector `first + i` at d_out[i * 1024]
	return launch_decode<double>(col, first, n, d_out, stream);
}
int alpb200_decode_f32(const alpb200_column* col, uint64_t first, uint64_t n, float* d_out, void* stream) {
	return launch_decode<float>(col, first, n, d_out, stream);
}

int alpb200_decode_sum_f64(const alpb200_column* col, uint64_t first, uint64_t n, double* d_sum, void* stream) {
	return launch_decode_sum<double>(col, first, n, d_sum, stream);
}
int alpb200_decode_sum_f32(const alpb200_column* col, uint64_t first, uint64_t n, double* d_sum, void* stream) {
	return launch_decode_sum<float>(col, first, n, d_sum, stream);
}

int alpb200_ctx_create(alpb200_ctx** out, int device, uint64_t max_vectors, int value_bytes) {
	if (!out || max_vectors == 0 || max_vectors > (1ull << 22) || (value_bytes != 8 && value_bytes != 4)) {
		return fail(ALPB200_EINVAL, "ctx_create: bad argument (max_vectors in 1..2^22, value_bytes 8 or 4)");
	}
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
		cudaGetLastError();
		return fail(ALPB200_ENODEVICE, "no CUDA device is visible; alp_b200 has no CPU fallback");
	}
	CUDA_TRY(cudaSetDevice(device));
	alpb200_ctx* c = new (std::nothrow) alpb200_ctx();
	if (!c) { return fail(ALPB200_EINVAL, "ctx_create: out of host memory"); }
	c->device      = device;
	c->value_bytes = value_bytes;
	c->max_vectors = max_vectors;
	const uint64_t n_rg = (max_vectors + ALPB200_ROWGROUP_VECTORS - 1) / ALPB200_ROWGROUP_VECTORS;
	c->packed_capacity  = max_vectors * ((value_bytes == 8 ? 66ull : 35ull) * 128ull);
	c->exc_capacity     = max_vectors * VEC;
	cudaError_t err     = cudaSuccess;
	auto        alloc   = [&](void** p, size_t bytes) {
        if (err == cudaSuccess) { err = cudaMalloc(p, bytes); }
	};
	alloc(&c->d_values, max_vectors * VEC * value_bytes);
	alloc(reinterpret_cast<void**>(&c->d_meta), max_vectors * sizeof(alpb200_vec_meta));
	alloc(reinterpret_cast<void**>(&c->d_packed), c->packed_capacity);
	alloc(&c->d_exc_val, c->exc_capacity * value_bytes);
	alloc(reinterpret_cast<void**>(&c->d_exc_pos), c->exc_capacity * sizeof(uint16_t));
	alloc(reinterpret_cast<void**>(&c->d_totals), 4 * sizeof(uint64_t));
	alloc(reinterpret_cast<void**>(&c->d_states), n_rg * sizeof(alpb200_rg_state));
	alloc(&c->d_ws_enc, encode_workspace_bytes(max_vectors));
	alloc(&c->d_ws_init, init_workspace_bytes(max_vectors * VEC));
	if (err == cudaSuccess) { err = cudaHostAlloc(reinterpret_cast<void**>(&c->h_totals), 4 * sizeof(uint64_t), cudaHostAllocDefault); }
	for (int i = 0; i < 3 && err == cudaSuccess; i++) {
		err = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
		if (err == cudaSuccess) { err = cudaEventCreateWithFlags(&c->events[i], cudaEventDisableTiming); }
	}
	if (err != cudaSuccess) {
		ctx_release(c);
		return fail(ALPB200_ECUDA, "ctx_create: %s", cudaGetErrorString(err));
	}
	*out = c;
	return ALPB200_OK;
}
void alpb200_ctx_destroy(alpb200_ctx* ctx) { ctx_release(ctx); }

int alpb200_compress_host_f64(alpb200_ctx* ctx, const double* h_in, uint64_t n_values, alpb200_column* h_col) {
	return compress_host<double>(ctx, h_in, n_values, h_col);
}
int alpb200_compress_host_f32(alpb200_ctx* ctx, const float* h_in, uint64_t n_values, alpb200_column* h_col) {
	return compress_host<float>(ctx, h_in, n_values, h_col);
}
int alpb200_decompress_host_f64(alpb200_ctx* ctx, const alpb200_column* h_col, double* h_out) { return decompress_host<double>(ctx, h_col, h_out); }
int alpb200_decompress_host_f32(alpb200_ctx* ctx, const alpb200_column* h_col, float* h_out) { return decompress_host<float>(ctx, h_col, h_out); }

void* alpb200_host_alloc(size_t bytes) {
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
		cudaGetLastError();
		fail(ALPB200_ECUDA, "host_alloc: cudaHostAlloc failed");
		return nullptr;
	}
	return p;
}
void alpb200_host_free(void* p) {
	if (p) { cudaFreeHost(p); }
}

int alpb200_prim_encode_f64(const double* in, const alpb200_rg_state* st, double* exc, uint16_t* pos, uint16_t* cnt, int64_t* enc,
                            uint8_t* e, uint8_t* f) {
	return prim_encode<double>(in, st, exc, pos, cnt, enc, e, f);
}
int alpb200_prim_encode_f32(const float* in, const alpb200_rg_state* st, float* exc, uint16_t* pos, uint16_t* cnt, int32_t* enc,
                            uint8_t* e, uint8_t* f) {
	return prim_encode<float>(in, st, exc, pos, cnt, enc, e, f);
}
int alpb200_prim_analyze_ffor_i64(const int64_t* enc, uint8_t* bw, int64_t* base) { return prim_analyze<double>(enc, bw, base); }
int alpb200_prim_analyze_ffor_i32(const int32_t* enc, uint8_t* bw, int32_t* base) { return prim_analyze<float>(enc, bw, base); }
int alpb200_prim_ffor_u64(const uint64_t* in, uint64_t* out, uint8_t bw, uint64_t base) { return prim_ffor<double>(in, out, bw, base); }
int alpb200_prim_ffor_u32(const uint32_t* in, uint32_t* out, uint8_t bw, uint32_t base) { return prim_ffor<float>(in, out, bw, base); }
int alpb200_prim_ffor_u16(const uint16_t* h_in, uint16_t* h_out, uint8_t bw, uint16_t base) {
	if (!h_in || !h_out) { return fail(ALPB200_EINVAL, "prim_ffor: null argument"); }
	if (bw > 16) { return fail(ALPB200_EINVAL, "prim_ffor: bit width exceeds the lane width"); }
	DevBuf in, out;
	TRY(in.upload(h_in, VEC * 2));
	TRY(out.alloc(128u * 16u));
	prim_ffor16_kernel<<<1, 32>>>(in.as<uint16_t>(), out.as<uint16_t>(), bw, base);
	TRY(finish_kernel());
	TRY(out.download(h_out, 128u * bw));
	return ALPB200_OK;
}
int alpb200_prim_unffor_u64(const uint64_t* in, uint64_t* out, uint8_t bw, uint64_t base) { return prim_unffor<double>(in, out, bw, base); }
int alpb200_prim_unffor_u32(const uint32_t* in, uint32_t* out, uint8_t bw, uint32_t base) { return prim_unffor<float>(in, out, bw, base); }
int alpb200_prim_unffor_u16(const uint16_t* h_in, uint16_t* h_out, uint8_t bw, uint16_t base) {
	if (!h_in || !h_out) { return fail(ALPB200_EINVAL, "prim_unffor: null argument"); }
	if (bw > 16) { return fail(ALPB200_EINVAL, "prim_unffor: bit width exceeds the lane width"); }
	DevBuf in, out;
	TRY(in.upload(h_in, 128u * bw));
	TRY(out.alloc(VEC * 2));
	prim_unffor16_kernel<<<1, 32>>>(in.as<uint16_t>(), out.as<uint16_t>(), bw, base);
	TRY(finish_kernel());
	TRY(out.download(h_out, VEC * 2));
	return ALPB200_OK;
}
int alpb200_prim_falp_f64(const uint64_t* packed, double* out, uint8_t bw, uint64_t base, uint8_t f, uint8_t e) {
	return prim_falp<double>(packed, out, bw, base, f, e);
}
int alpb200_prim_falp_f32(const uint32_t* packed, float* out, uint8_t bw, uint32_t base, uint8_t f, uint8_t e) {
	return prim_falp<float>(packed, out, bw, base, f, e);
}
int alpb200_prim_decode_f64(const int64_t* enc, uint8_t f, uint8_t e, double* out) { return prim_decode<double>(enc, f, e, out); }
int alpb200_prim_decode_f32(const int32_t* enc, uint8_t f, uint8_t e, float* out) { return prim_decode<float>(enc, f, e, out); }
int alpb200_prim_patch_f64(double* out, const double* exc, const uint16_t* pos, uint16_t cnt) { return prim_patch<double>(out, exc, pos, cnt); }
int alpb200_prim_patch_f32(float* out, const float* exc, const uint16_t* pos, uint16_t cnt) { return prim_patch<float>(out, exc, pos, cnt); }
int alpb200_prim_rd_encode_f64(const double* in, const alpb200_rg_state* st, uint16_t* exc, uint16_t* pos, uint16_t* cnt, uint64_t* right,
                               uint16_t* left) {
	return prim_rd_encode<double>(in, st, exc, pos, cnt, right, left);
}
int alpb200_prim_rd_encode_f32(const float* in, const alpb200_rg_state* st, uint16_t* exc, uint16_t* pos, uint16_t* cnt, uint32_t* right,
                               uint16_t* left) {
	return prim_rd_encode<float>(in, st, exc, pos, cnt, right, left);
}
int alpb200_prim_rd_decode_f64(double* out, const uint64_t* right, const uint16_t* left, const uint16_t* exc, const uint16_t* pos,
                               uint16_t cnt, const alpb200_rg_state* st) {
	return prim_rd_decode<double>(out, right, left, exc, pos, cnt, st);
}
int alpb200_prim_rd_decode_f32(float* out, const uint32_t* right, const uint16_t* left, const uint16_t* exc, const uint16_t* pos,
                               uint16_t cnt, const alpb200_rg_state* st) {
	return prim_rd_decode<float>(out, right, left, exc, pos, cnt, st);
}
int alpb200_prim_init_f64(const double* col, uint64_t offset, uint64_t n_values, alpb200_rg_state* st) {
	return prim_init<double>(col, offset, n_values, st);
}
int alpb200_prim_init_f32(const float* col, uint64_t offset, uint64_t n_values, alpb200_rg_state* st) {
	return prim_init<float>(col, offset, n_values, st);
}

int alpb200_generate_f64(double* d_out, uint64_t n_values, uint64_t first_index, uint64_t seed, int kind, void* stream) {
	if (!d_out || (kind != 2 && kind != 3)) { return fail(ALPB200_EINVAL, "generate_f64: kind must be 2 or 3"); }
	if (n_values == 0) { return ALPB200_OK; }
	generate_f64_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_out, n_values, first_index, seed, kind);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}
int alpb200_generate_f32(float* d_out, uint64_t n_values, uint64_t first_index, uint64_t seed, int kind, void* stream) {
	if (!d_out || kind != 4) { return fail(ALPB200_EINVAL, "generate_f32: kind must be 4"); }
	if (n_values == 0) { return ALPB200_OK; }
	generate_f32_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_out, n_values, first_index, seed, kind);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}

}  // extern "C"
