/*
 * alp_b200.hpp — the reference's primitive C++ API (cwida/ALP, PRIMITIVES.md) on top of the C ABI of alp_b200.h.
 *
 * A translation unit written against the reference —
 *
 *     #include "alp.hpp"
 *     alp::encoder<double>::init(col, offset, n, sample, stt);
 *     alp::encoder<double>::encode(in, exc, pos, cnt, enc, stt);
 *     alp::encoder<double>::analyze_ffor(enc, bw, base);
 *     ffor::ffor(enc, packed, bw, base);
 *     generated::falp::fallback::scalar::falp(packed, out, bw, base, stt.fac, stt.exp);
 *     alp::decoder<double>::patch_exceptions(out, exc, pos, cnt);
 *
 * — compiles unchanged against this header (include it instead of alp.hpp and link libalp_b200.so): same namespaces,
 * same names, same argument order and meaning.  Each call is a one-vector batch on the GPU (host pointers in and
 * out), so this header is the drop-in / parity surface; engines that care about throughput call the batched entry
 * points of alp_b200.h (alpb200_rowgroup_init_*, alpb200_encode_*, alpb200_decode_*) on whole columns.
 *
 * Mirrored reference declarations (file:line in the reference tree):
 *   alp::config::*                              include/alp/config.hpp:11-26
 *   alp::Scheme, bw_t, exp_c_t, exp_p_t, ...    include/alp/constants.hpp:10-14, include/alp/common.hpp:8-16
 *   alp::inner_t<PT>                            include/alp/decoder.hpp:15-30
 *   alp::state<PT>                              include/alp/encoder.hpp:35-62
 *   alp::encoder<PT>::{init,encode,analyze_ffor}   include/alp/encoder.hpp:420-427,402-418,109-120
 *   alp::decoder<PT>::{decode,patch_exceptions}    include/alp/decoder.hpp:134-149
 *   alp::rd_encoder<PT>::{init,encode,decode}      include/alp/rd.hpp:180-185,109-147,152-178
 *   ffor::ffor / unffor::unffor                 include/fastlanes/ffor.hpp:7-15, include/fastlanes/unffor.hpp:7-15
 *   generated::falp::fallback::scalar::falp     include/alp/falp.hpp:10-44
 *
 * Error behaviour: the reference's primitives return void and have undefined behaviour on bad input.  Here a failing
 * C-ABI call (no GPU, bad argument) throws alp::gpu_error; there is no CPU fallback.
 */
#ifndef ALP_B200_HPP
#define ALP_B200_HPP

// (the standard headers the reference's headers pull in — code written against alp.hpp relies on them transitively)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "alp_b200.h"

namespace alp {

struct gpu_error : std::runtime_error {
	int code;
	gpu_error(int c, const char* msg) : std::runtime_error(msg), code(c) {}
};
inline void check_(int rc) {
	if (rc < 0) { throw gpu_error(rc, alpb200_last_error()); }
}

namespace config {
inline constexpr size_t VECTOR_SIZE             = ALPB200_VECTOR_SIZE;
inline constexpr size_t N_VECTORS_PER_ROWGROUP  = ALPB200_ROWGROUP_VECTORS;
inline constexpr size_t ROWGROUP_SIZE           = ALPB200_ROWGROUP_SIZE;
inline constexpr size_t ROWGROUP_VECTOR_SAMPLES = 8;
inline constexpr size_t ROWGROUP_SAMPLES_JUMP   = (ROWGROUP_SIZE / ROWGROUP_VECTOR_SAMPLES) / VECTOR_SIZE;
inline constexpr size_t SAMPLES_PER_VECTOR      = 32;
inline constexpr size_t MAX_K_COMBINATIONS      = ALPB200_MAX_K;
inline constexpr size_t CUTTING_LIMIT           = 16;
inline constexpr size_t MAX_RD_DICT_BIT_WIDTH   = 3;
inline constexpr size_t MAX_RD_DICTIONARY_SIZE  = ALPB200_RD_DICT_SIZE;
} // namespace config

enum class Scheme : uint8_t { INVALID = ALPB200_SCHEME_INVALID, ALP_RD = ALPB200_SCHEME_ALP_RD, ALP = ALPB200_SCHEME_ALP };

using bw_t           = uint8_t;
using exp_c_t        = uint16_t;
using exp_p_t        = uint16_t;
using factor_idx_t   = uint8_t;
using exponent_idx_t = uint8_t;

template <typename T>
struct inner_t;
template <>
struct inner_t<float> {
	using ut = uint32_t;
	using st = int32_t;
};
template <>
struct inner_t<double> {
	using ut = uint64_t;
	using st = int64_t;
};

template <typename PT>
struct state {
	using UT = typename inner_t<PT>::ut;
	using ST = typename inner_t<PT>::st;

	Scheme   scheme {Scheme::INVALID};
	uint16_t vector_size {config::VECTOR_SIZE};
	uint16_t exceptions_count {0};
	size_t   sampled_values_n {0};

	// ALP
	uint16_t                         k_combinations {5};
	std::vector<std::pair<int, int>> best_k_combinations;
	uint8_t                          exp {};
	uint8_t                          fac {};
	bw_t                             bit_width {};
	ST                               for_base {};

	// ALP_RD
	bw_t                                   right_bit_width {0};
	bw_t                                   left_bit_width {0};
	UT                                     right_for_base {0};
	uint16_t                               left_for_base {0};
	uint16_t                               left_parts_dict[config::MAX_RD_DICTIONARY_SIZE] {};
	uint8_t                                actual_dictionary_size {};
	uint32_t                               actual_dictionary_size_bytes {};
	std::unordered_map<uint16_t, uint16_t> left_parts_dict_map;
};

namespace detail {

template <typename PT>
inline void to_pod(const state<PT>& stt, alpb200_rg_state& s) {
	std::memset(&s, 0, sizeof(s));
	s.scheme = static_cast<int32_t>(stt.scheme);
	s.k      = stt.k_combinations;
	for (int i = 0; i < s.k && i < ALPB200_MAX_K && i < static_cast<int>(stt.best_k_combinations.size()); i++) {
		s.combos[i][0] = static_cast<uint8_t>(stt.best_k_combinations[i].first);
		s.combos[i][1] = static_cast<uint8_t>(stt.best_k_combinations[i].second);
	}
	s.right_bw  = stt.right_bit_width;
	s.left_bw   = stt.left_bit_width;
	s.dict_size = stt.actual_dictionary_size;
	for (int i = 0; i < ALPB200_RD_DICT_SIZE; i++) {
		s.dict[i] = stt.left_parts_dict[i];
	}
	for (auto const& kv : stt.left_parts_dict_map) {
		if (kv.second >= stt.actual_dictionary_size && s.n_extra < ALPB200_MAX_SAMPLES) {
			s.extra_key[s.n_extra] = kv.first;
			s.extra_idx[s.n_extra] = kv.second;
			s.n_extra++;
		}
	}
}

template <typename PT>
inline void from_pod(const alpb200_rg_state& s, state<PT>& stt) {
	stt.scheme = static_cast<Scheme>(s.scheme);
	stt.best_k_combinations.clear();
	if (s.scheme == ALPB200_SCHEME_ALP) {
		stt.k_combinations = static_cast<uint16_t>(s.k);
		for (int i = 0; i < s.k; i++) {
			stt.best_k_combinations.emplace_back(s.combos[i][0], s.combos[i][1]);
		}
	} else {
		stt.right_bit_width              = s.right_bw;
		stt.left_bit_width               = s.left_bw;
		stt.actual_dictionary_size       = s.dict_size;
		stt.actual_dictionary_size_bytes = s.dict_size * 2u;
		stt.left_parts_dict_map.clear();
		for (int i = 0; i < ALPB200_RD_DICT_SIZE; i++) {
			stt.left_parts_dict[i] = s.dict[i];
		}
		for (int i = 0; i < s.dict_size; i++) {
			stt.left_parts_dict_map.insert({s.dict[i], static_cast<uint16_t>(i)});
		}
		for (int i = 0; i < s.n_extra; i++) {
			stt.left_parts_dict_map.insert({s.extra_key[i], s.extra_idx[i]});
		}
	}
}

// dispatch on the value type
inline int init(const double* c, uint64_t o, uint64_t n, alpb200_rg_state* s) { return alpb200_prim_init_f64(c, o, n, s); }
inline int init(const float* c, uint64_t o, uint64_t n, alpb200_rg_state* s) { return alpb200_prim_init_f32(c, o, n, s); }
inline int rd_init(const double* c, uint64_t o, uint64_t n, alpb200_rg_state* s) { return alpb200_prim_rd_init_f64(c, o, n, s); }
inline int rd_init(const float* c, uint64_t o, uint64_t n, alpb200_rg_state* s) { return alpb200_prim_rd_init_f32(c, o, n, s); }
//! The first-level sample as the caller's `sample_arr` would hold it (sampler.hpp:14-52 restricted to whole vectors, which is what
//! the device init samples): values 0, 32, ..., 992 of vectors 0, 12, 24, ... of the row-group.  Returns the number of samples.
//! (An output parameter filled from the caller's own host column; the search itself runs on the device.)
template <typename PT>
inline size_t fill_sample(const PT* column, size_t offset, size_t n_values, PT* sample_arr) {
	const size_t in_rowgroup = n_values - offset < size_t(ALPB200_ROWGROUP_SIZE) ? n_values - offset : size_t(ALPB200_ROWGROUP_SIZE);
	const size_t vectors     = in_rowgroup / ALPB200_VECTOR_SIZE;
	size_t       n           = 0;
	for (size_t v = 0; v < vectors; v += ALPB200_ROWGROUP_SAMPLES_JUMP) {
		const PT* vec = column + offset + v * ALPB200_VECTOR_SIZE;
		for (size_t i = 0; i < size_t(ALPB200_VECTOR_SIZE); i += 32) {
			if (sample_arr) { sample_arr[n] = vec[i]; }
			n++;
		}
	}
	return n;
}
inline int encode(const double* in, const alpb200_rg_state* s, double* exc, uint16_t* pos, uint16_t* cnt, int64_t* enc, uint8_t* e, uint8_t* f) {
	return alpb200_prim_encode_f64(in, s, exc, pos, cnt, enc, e, f);
}
inline int encode(const float* in, const alpb200_rg_state* s, float* exc, uint16_t* pos, uint16_t* cnt, int32_t* enc, uint8_t* e, uint8_t* f) {
	return alpb200_prim_encode_f32(in, s, exc, pos, cnt, enc, e, f);
}
inline int analyze(const int64_t* enc, uint8_t* bw, int64_t* base) { return alpb200_prim_analyze_ffor_i64(enc, bw, base); }
inline int analyze(const int32_t* enc, uint8_t* bw, int32_t* base) { return alpb200_prim_analyze_ffor_i32(enc, bw, base); }
inline int decode(const int64_t* enc, uint8_t f, uint8_t e, double* out) { return alpb200_prim_decode_f64(enc, f, e, out); }
inline int decode(const int32_t* enc, uint8_t f, uint8_t e, float* out) { return alpb200_prim_decode_f32(enc, f, e, out); }
inline int patch(double* out, const double* exc, const uint16_t* pos, uint16_t cnt) { return alpb200_prim_patch_f64(out, exc, pos, cnt); }
inline int patch(float* out, const float* exc, const uint16_t* pos, uint16_t cnt) { return alpb200_prim_patch_f32(out, exc, pos, cnt); }
inline int rd_encode(const double* in, const alpb200_rg_state* s, uint16_t* exc, uint16_t* pos, uint16_t* cnt, uint64_t* right, uint16_t* left) {
	return alpb200_prim_rd_encode_f64(in, s, exc, pos, cnt, right, left);
}
inline int rd_encode(const float* in, const alpb200_rg_state* s, uint16_t* exc, uint16_t* pos, uint16_t* cnt, uint32_t* right, uint16_t* left) {
	return alpb200_prim_rd_encode_f32(in, s, exc, pos, cnt, right, left);
}
inline int rd_decode(double* out, const uint64_t* right, const uint16_t* left, const uint16_t* exc, const uint16_t* pos, uint16_t cnt, const alpb200_rg_state* s) {
	return alpb200_prim_rd_decode_f64(out, right, left, exc, pos, cnt, s);
}
inline int rd_decode(float* out, const uint32_t* right, const uint16_t* left, const uint16_t* exc, const uint16_t* pos, uint16_t cnt, const alpb200_rg_state* s) {
	return alpb200_prim_rd_decode_f32(out, right, left, exc, pos, cnt, s);
}

} // namespace detail

template <typename PT>
struct encoder {
	using UT = typename inner_t<PT>::ut;
	using ST = typename inner_t<PT>::st;

	//! once per row-group: first-level sampling + top-k (exponent, factor) search + scheme decision.
	//! `sample_arr` receives the first-level sample like in the reference (encoder.hpp:420-427); the search reads the column on the device.
	static inline void init(const PT* data_column, const size_t column_offset, const size_t tuples_count, PT* sample_arr, state<PT>& stt) {
		alpb200_rg_state s;
		check_(detail::init(data_column, column_offset, tuples_count, &s));
		detail::from_pod(s, stt);
		stt.sampled_values_n = detail::fill_sample(data_column, column_offset, tuples_count, sample_arr);
		if (stt.scheme == Scheme::ALP_RD) {
			// the reference decides the scheme here and builds the dictionary in rd_encoder<PT>::init; the device init does
			// both at once, so only the scheme is reported until rd_encoder<PT>::init is called
			stt.left_parts_dict_map.clear();
		}
	}

	//! encode one vector: second-level sampling when k > 1, encode, exception extraction
	static inline void encode(const PT* input_vector, PT* exceptions, uint16_t* exceptions_positions, uint16_t* exceptions_count,
	                          ST* encoded_integers, state<PT>& stt) {
		alpb200_rg_state s;
		detail::to_pod(stt, s);
		s.scheme = ALPB200_SCHEME_ALP;
		check_(detail::encode(input_vector, &s, exceptions, exceptions_positions, exceptions_count, encoded_integers, &stt.exp, &stt.fac));
		stt.exceptions_count = exceptions_count[0];
	}

	//! frame-of-reference analysis: bit width of (max - min) and the base
	static inline void analyze_ffor(const ST* input_vector, bw_t& bit_width, ST* base_for) {
		check_(detail::analyze(input_vector, &bit_width, base_for));
	}
};

template <typename PT>
struct decoder {
	using UT = typename inner_t<PT>::ut;
	using ST = typename inner_t<PT>::st;

	static inline void decode(const ST* encoded_integers, const uint8_t fac_idx, const uint8_t exp_idx, PT* output) {
		check_(detail::decode(encoded_integers, fac_idx, exp_idx, output));
	}
	static inline void patch_exceptions(PT* out, const PT* exceptions, const exp_p_t* exceptions_positions, const exp_c_t* exceptions_count) {
		check_(detail::patch(out, exceptions, exceptions_positions, exceptions_count[0]));
	}
};

template <typename PT>
struct rd_encoder {
	using UT = typename inner_t<PT>::ut;

	//! rd.hpp:180-185: builds cut + dictionary for ANY row-group (callers that force ALP_RD, e.g. bench_alp_cutter_encode.cpp:110)
	static inline void init(const PT* data_column, size_t column_offset, size_t tuples_count, PT* sample_arr, state<PT>& stt) {
		alpb200_rg_state s;
		check_(detail::rd_init(data_column, column_offset, tuples_count, &s));
		detail::from_pod(s, stt);
		stt.sampled_values_n = detail::fill_sample(data_column, column_offset, tuples_count, sample_arr);
	}
	static inline void encode(const PT* dbl_arr, uint16_t* exceptions, uint16_t* exception_positions, uint16_t* exceptions_count_p,
	                          UT* right_parts, uint16_t* left_parts, state<PT>& stt) {
		alpb200_rg_state s;
		detail::to_pod(stt, s);
		s.scheme = ALPB200_SCHEME_ALP_RD;
		check_(detail::rd_encode(dbl_arr, &s, exceptions, exception_positions, exceptions_count_p, right_parts, left_parts));
		stt.exceptions_count = exceptions_count_p[0];
	}
	static inline void decode(PT* a_out, UT* unffor_right_arr, uint16_t* unffor_left_arr, uint16_t* exceptions, uint16_t* exceptions_positions,
	                          uint16_t* exceptions_count, state<PT>& stt) {
		alpb200_rg_state s;
		detail::to_pod(stt, s);
		s.scheme = ALPB200_SCHEME_ALP_RD;
		check_(detail::rd_decode(a_out, unffor_right_arr, unffor_left_arr, exceptions, exceptions_positions, exceptions_count[0], &s));
	}
};

} // namespace alp

// ---- FastLanes FFOR / UNFFOR (include/fastlanes/ffor.hpp, unffor.hpp) ------------------------------------------------
namespace fastlanes { namespace generated { namespace ffor { namespace fallback { namespace scalar {
inline void ffor(const uint64_t* in, uint64_t* out, uint8_t bw, const uint64_t* a_base_p) { alp::check_(alpb200_prim_ffor_u64(in, out, bw, *a_base_p)); }
inline void ffor(const uint32_t* in, uint32_t* out, uint8_t bw, const uint32_t* a_base_p) { alp::check_(alpb200_prim_ffor_u32(in, out, bw, *a_base_p)); }
inline void ffor(const uint16_t* in, uint16_t* out, uint8_t bw, const uint16_t* a_base_p) { alp::check_(alpb200_prim_ffor_u16(in, out, bw, *a_base_p)); }
inline void ffor(const uint8_t* in, uint8_t* out, uint8_t bw, const uint8_t* a_base_p) { alp::check_(alpb200_prim_ffor_u8(in, out, bw, *a_base_p)); }
inline void ffor(const int8_t* in, int8_t* out, uint8_t bw, const int8_t* a_base_p) {
	alp::check_(alpb200_prim_ffor_u8(reinterpret_cast<const uint8_t*>(in), reinterpret_cast<uint8_t*>(out), bw, static_cast<uint8_t>(*a_base_p)));
}
inline void ffor(const int64_t* in, int64_t* out, uint8_t bw, const int64_t* a_base_p) {
	alp::check_(alpb200_prim_ffor_u64(reinterpret_cast<const uint64_t*>(in), reinterpret_cast<uint64_t*>(out), bw, static_cast<uint64_t>(*a_base_p)));
}
inline void ffor(const int32_t* in, int32_t* out, uint8_t bw, const int32_t* a_base_p) {
	alp::check_(alpb200_prim_ffor_u32(reinterpret_cast<const uint32_t*>(in), reinterpret_cast<uint32_t*>(out), bw, static_cast<uint32_t>(*a_base_p)));
}
inline void ffor(const int16_t* in, int16_t* out, uint8_t bw, const int16_t* a_base_p) {
	alp::check_(alpb200_prim_ffor_u16(reinterpret_cast<const uint16_t*>(in), reinterpret_cast<uint16_t*>(out), bw, static_cast<uint16_t>(*a_base_p)));
}
}}}}} // namespace fastlanes::generated::ffor::fallback::scalar
namespace ffor = fastlanes::generated::ffor::fallback::scalar;

namespace fastlanes { namespace generated { namespace unffor { namespace fallback { namespace scalar {
inline void unffor(const uint64_t* in, uint64_t* out, uint8_t bw, const uint64_t* a_base_p) { alp::check_(alpb200_prim_unffor_u64(in, out, bw, *a_base_p)); }
inline void unffor(const uint32_t* in, uint32_t* out, uint8_t bw, const uint32_t* a_base_p) { alp::check_(alpb200_prim_unffor_u32(in, out, bw, *a_base_p)); }
inline void unffor(const uint16_t* in, uint16_t* out, uint8_t bw, const uint16_t* a_base_p) { alp::check_(alpb200_prim_unffor_u16(in, out, bw, *a_base_p)); }
inline void unffor(const uint8_t* in, uint8_t* out, uint8_t bw, const uint8_t* a_base_p) { alp::check_(alpb200_prim_unffor_u8(in, out, bw, *a_base_p)); }
inline void unffor(const int8_t* in, int8_t* out, uint8_t bw, const int8_t* a_base_p) {
	alp::check_(alpb200_prim_unffor_u8(reinterpret_cast<const uint8_t*>(in), reinterpret_cast<uint8_t*>(out), bw, static_cast<uint8_t>(*a_base_p)));
}
inline void unffor(const int64_t* in, int64_t* out, uint8_t bw, const int64_t* a_base_p) {
	alp::check_(alpb200_prim_unffor_u64(reinterpret_cast<const uint64_t*>(in), reinterpret_cast<uint64_t*>(out), bw, static_cast<uint64_t>(*a_base_p)));
}
inline void unffor(const int32_t* in, int32_t* out, uint8_t bw, const int32_t* a_base_p) {
	alp::check_(alpb200_prim_unffor_u32(reinterpret_cast<const uint32_t*>(in), reinterpret_cast<uint32_t*>(out), bw, static_cast<uint32_t>(*a_base_p)));
}
inline void unffor(const int16_t* in, int16_t* out, uint8_t bw, const int16_t* a_base_p) {
	alp::check_(alpb200_prim_unffor_u16(reinterpret_cast<const uint16_t*>(in), reinterpret_cast<uint16_t*>(out), bw, static_cast<uint16_t>(*a_base_p)));
}
}}}}} // namespace fastlanes::generated::unffor::fallback::scalar
namespace unffor = fastlanes::generated::unffor::fallback::scalar;

// ---- fused decode (include/alp/falp.hpp:10-44) ------------------------------------------------------------------------
namespace generated { namespace falp { namespace fallback { namespace scalar {
inline void falp(const uint64_t* in, double* out, uint8_t bw, const uint64_t* a_base_p, uint8_t factor, uint8_t exponent) {
	alp::check_(alpb200_prim_falp_f64(in, out, bw, *a_base_p, factor, exponent));
}
inline void falp(const int64_t* in, double* out, uint8_t bw, const int64_t* base, uint8_t factor, uint8_t exponent) {
	alp::check_(alpb200_prim_falp_f64(reinterpret_cast<const uint64_t*>(in), out, bw, static_cast<uint64_t>(*base), factor, exponent));
}
inline void falp(const uint32_t* in, float* out, uint8_t bw, const uint32_t* base, uint8_t factor, uint8_t exponent) {
	alp::check_(alpb200_prim_falp_f32(in, out, bw, *base, factor, exponent));
}
inline void falp(const int32_t* in, float* out, uint8_t bw, const int32_t* base, uint8_t factor, uint8_t exponent) {
	alp::check_(alpb200_prim_falp_f32(reinterpret_cast<const uint32_t*>(in), out, bw, static_cast<uint32_t>(*base), factor, exponent));
}
}}}} // namespace generated::falp::fallback::scalar

#endif // ALP_B200_HPP
