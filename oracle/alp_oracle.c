/*
 * oracle/alp_oracle.c — TEST INFRASTRUCTURE, not product code.  See alp_oracle.h for the contract and for how
 * this restatement is pinned against the reference.
 *
 * Restates, in plain C, the algorithm of cwida/ALP's per-vector path:
 *   constants/tables      include/alp/constants.hpp:16-155, include/alp/config.hpp:11-26
 *   sampling + (e,f)      include/alp/sampler.hpp:14-52, include/alp/encoder.hpp:139-305
 *   encode / analyze      include/alp/encoder.hpp:81-120,307-418
 *   decode / patch        include/alp/decoder.hpp:128-149
 *   ALP_RD                include/alp/rd.hpp:23-185
 *   FFOR / UNFFOR layout  src/fastlanes_generated_ffor.cpp:7788-7999 (3-bit/64-bit instance), :29750-30137 (dispatch)
 *                         src/fastlanes_generated_unffor.cpp:6389-6500, :22812-23211
 *   fused FALP            src/falp.cpp:1033-1060 (per-value recipe)
 *
 * Compile with -ffp-contract=off: the two multiplies and the add/subtract of the magic number must round
 * separately (encoder.hpp:83,87).
 */
#include "alp_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

const char* alpo_build_info(void) { return "alp_b200 oracle: plain-C restatement of cwida/ALP (gcc " __VERSION__ ")"; }

/* struct sizes this library was compiled with — lets the Python driver reject a stale build */
void alpo_abi_sizes(uint32_t* out) {
	out[0] = sizeof(alpb200_rg_state);
	out[1] = sizeof(alpb200_vec_meta);
	out[2] = sizeof(alpb200_column);
}

/* constants.hpp:17-18 */
#define ENC_UPPER 9223372036854774784.0
#define ENC_LOWER (-9223372036854774784.0)

/* constants.hpp:85-155 (double) */
static const double F64_EXP[24] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22, 1e23};
static const double F64_FRAC[21] = {1.0,   0.1,   0.01,  0.001, 1e-4,  1e-5,  1e-6,  1e-7,  1e-8,  1e-9, 1e-10,
                                    1e-11, 1e-12, 1e-13, 1e-14, 1e-15, 1e-16, 1e-17, 1e-18, 1e-19, 1e-20};
static const int64_t F64_FACT[19] = {1LL,
                                     10LL,
                                     100LL,
                                     1000LL,
                                     10000LL,
                                     100000LL,
                                     1000000LL,
                                     10000000LL,
                                     100000000LL,
                                     1000000000LL,
                                     10000000000LL,
                                     100000000000LL,
                                     1000000000000LL,
                                     10000000000000LL,
                                     100000000000000LL,
                                     1000000000000000LL,
                                     10000000000000000LL,
                                     100000000000000000LL,
                                     1000000000000000000LL};

/* constants.hpp:48-63 (float).  The reference's FACT_ARR has 10 entries but MAX_EXPONENT is 10, so the pair
 * (e=10, f=10) reads one element past the array (decoder.hpp:129) — undefined behaviour.  Entry [10] here is 0,
 * which is what that read returns in the g++ build of the reference (oracle/_ref; pinned by
 * tests/test_oracle_vs_reference.py::test_float_fact_out_of_bounds): with it the pair (10,10) can only ever
 * round-trip zeros, and an all-zero float vector does select it (largest e, then largest f win ties). */
static const float   F32_EXP[11]  = {1e0f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};
static const float   F32_FRAC[11] = {1.0f,     0.1f,      0.01f,      0.001f,      0.0001f,     0.00001f,
                                     0.000001f, 0.0000001f, 0.00000001f, 0.000000001f, 0.0000000001f};
static const int32_t F32_FACT[11] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000, 0};

/* x86 cvttsd2si / cvttss2si: NaN and out-of-range inputs produce the "integer indefinite" value */
static inline int64_t cast_x86_i64(double t) {
	return (t >= -9223372036854775808.0 && t < 9223372036854775808.0) ? (int64_t)t : INT64_MIN;
}
static inline int32_t cast_x86_i32(float t) { return (t >= -2147483648.0f && t < 2147483648.0f) ? (int32_t)t : INT32_MIN; }

/* ---- FFOR / UNFFOR, interleaved FastLanes layout ----
 * T-bit lanes: L = 1024/T lanes, T rows; value index v = L*row + lane; row `row` of a lane sits at bits
 * [row*bw, row*bw+bw) of that lane's bit stream; word w of the stream is stored at out[L*w + lane]. */
#define DEFINE_FFOR(T, U)                                                                                              \
	void alpo_ffor_u##T(const U* in, U* out, uint8_t bw, U base) {                                                     \
		const unsigned L = 1024u / T;                                                                                  \
		if (bw == 0 || bw > T) { return; } /* bw=0 writes nothing (ffor.cpp:4); bw>T: switch has no default */         \
		const U mask = bw == T ? (U) ~(U)0 : (U)((((U)1) << bw) - 1);                                                  \
		memset(out, 0, (size_t)bw * 128u);                                                                             \
		for (unsigned lane = 0; lane < L; lane++) {                                                                    \
			for (unsigned row = 0; row < T; row++) {                                                                   \
				U        d   = (U)((U)(in[L * row + lane] - base) & mask);                                             \
				unsigned bit = row * bw, w = bit / T, sh = bit % T;                                                    \
				out[L * w + lane] = (U)(out[L * w + lane] | (U)(d << sh));                                             \
				if (sh + bw > T) { out[L * (w + 1) + lane] = (U)(out[L * (w + 1) + lane] | (U)(d >> (T - sh))); }      \
			}                                                                                                          \
		}                                                                                                              \
	}                                                                                                                  \
	void alpo_unffor_u##T(const U* in, U* out, uint8_t bw, U base) {                                                   \
		const unsigned L = 1024u / T;                                                                                  \
		if (bw > T) { return; }                                                                                        \
		if (bw == 0) { /* unffor.cpp:4-22: broadcast the base */                                                       \
			for (unsigned i = 0; i < 1024u; i++) {                                                                     \
				out[i] = base;                                                                                         \
			}                                                                                                          \
			return;                                                                                                    \
		}                                                                                                              \
		const U mask = bw == T ? (U) ~(U)0 : (U)((((U)1) << bw) - 1);                                                  \
		for (unsigned lane = 0; lane < L; lane++) {                                                                    \
			for (unsigned row = 0; row < T; row++) {                                                                   \
				unsigned bit = row * bw, w = bit / T, sh = bit % T;                                                    \
				U        d   = (U)(in[L * w + lane] >> sh);                                                            \
				if (sh + bw > T) { d = (U)(d | (U)(in[L * (w + 1) + lane] << (T - sh))); }                             \
				out[L * row + lane] = (U)((U)(d & mask) + base);                                                       \
			}                                                                                                          \
		}                                                                                                              \
	}
DEFINE_FFOR(64, uint64_t)
DEFINE_FFOR(32, uint32_t)
DEFINE_FFOR(16, uint16_t)
DEFINE_FFOR(8, uint8_t) /* include/fastlanes/ffor.hpp:10, unffor.hpp:10 (src/fastlanes_generated_ffor.cpp:4-300) */

/* ---- double instance ---- */
#define PT double
#define UT uint64_t
#define ST int64_t
#define TBITS 64u
#define SFX f64
#define ISFX i64
#define USFX u64
#define MAX_EXP 18
#define MAGIC 6755399441055744.0 /* constants.hpp:70: 2^52 + 2^51 */
#define RD_LIMIT (48u * 32u)     /* constants.hpp:69 */
#define EXC_BITS 64u             /* constants.hpp:71 */
#define ST_MIN INT64_MIN
#define ST_MAX INT64_MAX
#define EXP_T F64_EXP
#define FRAC_T F64_FRAC
#define FACT_T F64_FACT
#define CAST_X86(t) cast_x86_i64(t)
#define SAFE_SENTINEL ((int64_t)9223372036854774784LL) /* (int64_t)ENCODING_UPPER_LIMIT, exactly representable */
/* encoder.hpp:326-331 with Constants<double>: EXPONENTIAL_BITS_MASK is a 65-digit literal (constants.hpp:82-83)
 * whose value is 0xFFE0000000000000, so the NaN/Inf half of the test can never fire; only -0.0 is replaced.
 * NaN and ±Inf still end up as exceptions through the decoded != original compare. */
#define IS_SPECIAL(bits) ((bits) == 0x8000000000000000ULL)
#include "alp_oracle_impl.inc"
#undef PT
#undef UT
#undef ST
#undef TBITS
#undef SFX
#undef ISFX
#undef USFX
#undef MAX_EXP
#undef MAGIC
#undef RD_LIMIT
#undef EXC_BITS
#undef ST_MIN
#undef ST_MAX
#undef EXP_T
#undef FRAC_T
#undef FACT_T
#undef CAST_X86
#undef SAFE_SENTINEL
#undef IS_SPECIAL

/* ---- float instance ---- */
#define PT float
#define UT uint32_t
#define ST int32_t
#define TBITS 32u
#define SFX f32
#define ISFX i32
#define USFX u32
#define MAX_EXP 10
#define MAGIC 12582912.0f    /* constants.hpp:34: 2^23 + 2^22 */
#define RD_LIMIT (22u * 32u) /* constants.hpp:33 */
#define EXC_BITS 32u         /* constants.hpp:35 */
#define ST_MIN INT32_MIN
#define ST_MAX INT32_MAX
#define EXP_T F32_EXP
#define FRAC_T F32_FRAC
#define FACT_T F32_FACT
#define CAST_X86(t) cast_x86_i32(t)
/* encoder.hpp:85 `return ENCODING_UPPER_LIMIT;` with ST = int32_t converts the double constant 2^63-1024 at compile
 * time; g++ folds the out-of-range conversion by saturation (pinned against oracle/_ref in
 * tests/test_oracle_vs_reference.py::test_safe_sentinel) */
#define SAFE_SENTINEL INT32_MAX
/* encoder.hpp:326-331 with Constants<float> (constants.hpp:41-46): NaN, ±Inf and -0.0 */
#define IS_SPECIAL(bits) ((((bits) & 0x7FFFFFFFu) >= 0x7F800000u) || (bits) == 0x80000000u)
#include "alp_oracle_impl.inc"
#undef PT
#undef UT
#undef ST

/* ---- synthetic columns ---- */
static inline uint64_t splitmix64(uint64_t seed, uint64_t i) {
	uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
	z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z          = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

void alpo_generate_f64(double* out, uint64_t n_values, uint64_t first_index, uint64_t seed, int kind) {
	static const double DIV[4] = {1.0, 10.0, 100.0, 1000.0};
	for (uint64_t j = 0; j < n_values; j++) {
		uint64_t i = first_index + j;
		uint64_t r = splitmix64(seed, i);
		if (kind == 3) { /* latitude-like, full 53-bit mantissas → ALP_RD */
			double u = (double)(r >> 11) * 0x1.0p-53;
			double t = u * 180.0;
			out[j]   = t - 90.0;
		} else { /* kind 2: ≤3 decimals, the number of decimals constant per row-group */
			out[j] = (double)(r % 1000000ULL) / DIV[(i / ALPB200_ROWGROUP_SIZE) % 4];
		}
	}
}

void alpo_generate_f32(float* out, uint64_t n_values, uint64_t first_index, uint64_t seed, int kind) {
	(void)kind;
	for (uint64_t j = 0; j < n_values; j++) {
		uint64_t r = splitmix64(seed, first_index + j);
		if (r % 100 >= 5) {
			out[j] = (float)((r >> 8) % 10000ULL) / 100.0f;
		} else {
			uint32_t b = (uint32_t)(r >> 32);
			b          = (b & 0x807FFFFFu) | ((20u + ((b >> 23) % 200u)) << 23);
			memcpy(&out[j], &b, sizeof(b));
		}
	}
}
