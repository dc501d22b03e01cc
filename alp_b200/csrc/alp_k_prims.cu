// alp_k_prims.cu — row-group initialisation, the single-vector primitive wrappers (alpb200_prim_*), the tail padding and
// the synthetic column generators (one translation unit of libalp_b200.so).
#include <algorithm>

#include "alp_host.h"
#include "alp_init.cuh"
#include "alp_prims.cuh"

namespace alpb200 {

size_t init_workspace_bytes(uint64_t n_values) {
	const uint64_t n_vec = n_values / VEC;
	const uint64_t n_rg  = (n_vec + ALPB200_ROWGROUP_VECTORS - 1) / ALPB200_ROWGROUP_VECTORS;
	return (size_t)(n_rg * MAX_SAMPLED_VECS * sizeof(SearchResult) + 255) & ~(size_t)255;
}

template <typename PT>
int launch_init(const PT* d_in, uint64_t n_values, alpb200_rg_state* d_states, void* ws, void* stream, bool force_rd) {
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
	                                                                              d_states, force_rd ? 1u : 0u);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}

template int launch_init<double>(const double*, uint64_t, alpb200_rg_state*, void*, void*, bool);
template int launch_init<float>(const float*, uint64_t, alpb200_rg_state*, void*, void*, bool);

// tail vector + NULLs: see fill_invalid_kernel in alp_prims.cuh
template <typename PT>
int launch_fill_invalid(PT* d_values, uint64_t n_values, const uint8_t* d_validity, const alpb200_rg_state* d_states, void* stream) {
	if (!d_values) { return fail(ALPB200_EINVAL, "fill_invalid: null argument"); }
	if ((reinterpret_cast<uintptr_t>(d_validity) & 3u) != 0) { return fail(ALPB200_EINVAL, "fill_invalid: the validity bitmap must be 4-byte aligned"); }
	const uint64_t n_vec = (n_values + VEC - 1) / VEC;
	if (n_vec == 0) { return ALPB200_OK; }
	constexpr int  W     = 4;
	// without a bitmap only the last vector can have anything to fill
	const uint64_t first = d_validity ? 0 : n_vec - 1;
	if (!d_validity && n_values % VEC == 0) { return ALPB200_OK; }
	const uint64_t count = n_vec - first;
	fill_invalid_kernel<PT, W><<<(uint32_t)((count + W - 1) / W), W * 32, 0, static_cast<cudaStream_t>(stream)>>>(
	    d_values + first * VEC, n_values - first * VEC, reinterpret_cast<const uint32_t*>(d_validity), d_states ? d_states + first / ALPB200_ROWGROUP_VECTORS : nullptr,
	    count);
	CUDA_TRY(cudaGetLastError());
	return ALPB200_OK;
}
template int launch_fill_invalid<double>(double*, uint64_t, const uint8_t*, const alpb200_rg_state*, void*);
template int launch_fill_invalid<float>(float*, uint64_t, const uint8_t*, const alpb200_rg_state*, void*);

}  // namespace alpb200

using namespace alpb200;

namespace {

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
int prim_init(const PT* h_col, uint64_t offset, uint64_t n_values, alpb200_rg_state* h_state, bool force_rd = false) {
	if (!h_col || !h_state || offset >= n_values) { return fail(ALPB200_EINVAL, "prim_init: bad argument"); }
	const uint64_t span = std::min<uint64_t>(ALPB200_ROWGROUP_SIZE, n_values - offset) / VEC * VEC;
	if (span == 0) { return fail(ALPB200_EINVAL, "prim_init: the row-group holds no complete vector"); }
	DevBuf in, st, ws;
	TRY(in.upload(h_col + offset, span * sizeof(PT)));
	TRY(st.alloc(sizeof(alpb200_rg_state)));
	TRY(ws.alloc(init_workspace_bytes(span)));
	TRY(launch_init<PT>(in.as<PT>(), span, st.as<alpb200_rg_state>(), ws.p, nullptr, force_rd));
	TRY(finish_kernel());
	TRY(st.download(h_state, sizeof(*h_state)));
	return ALPB200_OK;
}

}  // namespace

extern "C" {

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
int alpb200_prim_ffor_u8(const uint8_t* h_in, uint8_t* h_out, uint8_t bw, uint8_t base) {
	if (!h_in || !h_out) { return fail(ALPB200_EINVAL, "prim_ffor: null argument"); }
	if (bw > 8) { return fail(ALPB200_EINVAL, "prim_ffor: bit width exceeds the lane width"); }
	DevBuf in, out;
	TRY(in.upload(h_in, VEC));
	TRY(out.alloc(128u * 8u));
	prim_ffor8_kernel<<<1, 32>>>(in.as<uint8_t>(), out.as<uint8_t>(), bw, base);
	TRY(finish_kernel());
	TRY(out.download(h_out, 128u * bw));
	return ALPB200_OK;
}
int alpb200_prim_unffor_u8(const uint8_t* h_in, uint8_t* h_out, uint8_t bw, uint8_t base) {
	if (!h_in || !h_out) { return fail(ALPB200_EINVAL, "prim_unffor: null argument"); }
	if (bw > 8) { return fail(ALPB200_EINVAL, "prim_unffor: bit width exceeds the lane width"); }
	DevBuf in, out;
	TRY(in.upload(h_in, bw ? 128u * bw : 16u));
	TRY(out.alloc(VEC));
	prim_unffor8_kernel<<<1, 32>>>(in.as<uint8_t>(), out.as<uint8_t>(), bw, base);
	TRY(finish_kernel());
	TRY(out.download(h_out, VEC));
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
int alpb200_prim_rd_init_f64(const double* col, uint64_t offset, uint64_t n_values, alpb200_rg_state* st) {
	return prim_init<double>(col, offset, n_values, st, true);
}
int alpb200_prim_rd_init_f32(const float* col, uint64_t offset, uint64_t n_values, alpb200_rg_state* st) {
	return prim_init<float>(col, offset, n_values, st, true);
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
