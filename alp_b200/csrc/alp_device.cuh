// alp_device.cuh — device-side building blocks of the B200-native ALP codec (sm_100a).
//
// Everything here restates the arithmetic contract of the reference (cwida/ALP) in CUDA terms; the layout and
// arithmetic are documented in SURVEY.md appendix A and DESIGN.md.  Reference citations are file:line in the
// reference tree.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "alp_b200.h"

namespace alpb200 {

constexpr int      WARP          = 32;
constexpr uint32_t FULL          = 0xFFFFFFFFu;
constexpr int      VEC           = 1024;
constexpr uint32_t STAGE_PAD     = 128;   // unpackers may read one word-row past the packed block
constexpr uint32_t MAX_BLOCK     = 8192 + 3 * 128;  // largest packed block: ALP bw=64, or ALP_RD 63 right + 3 left bits... (capped below)

// ---------------------------------------------------------------------------------------------------------------
// Tables (include/alp/constants.hpp:48-63 float, :85-155 double).  Decode multiplies by FRAC[e]; it does not divide.
// F32 FACT[10] is the reference's out-of-bounds read (decoder.hpp:129 with MAX_EXPONENT 10, constants.hpp:39,63);
// 0 is what the g++ build of the reference returns there (see oracle/alp_oracle.c).
// ---------------------------------------------------------------------------------------------------------------
static __constant__ double  C_F64_EXP[24]  = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                       1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22, 1e23};
static __constant__ double  C_F64_FRAC[21] = {1.0,   0.1,   0.01,  0.001, 1e-4,  1e-5,  1e-6,  1e-7,  1e-8,  1e-9, 1e-10,
                                       1e-11, 1e-12, 1e-13, 1e-14, 1e-15, 1e-16, 1e-17, 1e-18, 1e-19, 1e-20};
static __constant__ int64_t C_F64_FACT[19] = {1LL,
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
static __constant__ float   C_F32_EXP[11]  = {1e0f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};
static __constant__ float   C_F32_FRAC[11] = {1.0f,      0.1f,       0.01f,       0.001f,       0.0001f,      0.00001f,
                                       0.000001f, 0.0000001f, 0.00000001f, 0.000000001f, 0.0000000001f};
static __constant__ int32_t C_F32_FACT[11] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000, 0};

// ---------------------------------------------------------------------------------------------------------------
// Per-type traits.  All floating-point steps use the explicitly rounded intrinsics so that nvcc can never contract
// `*FRAC + MAGIC` into an FMA (encoder.hpp:83,87 round separately).
// ---------------------------------------------------------------------------------------------------------------
template <typename PT>
struct Traits;

template <>
struct Traits<double> {
	using UT = uint64_t;
	using ST = int64_t;
	static constexpr int      TBITS    = 64;
	static constexpr int      LANES    = 16;   // 1024 / 64 FastLanes lanes
	static constexpr int      MAX_EXP  = 18;   // constants.hpp:73
	static constexpr int      N_COMBOS = 190;  // pairs (e, f <= e)
	static constexpr uint32_t EXC_BITS = 64;   // constants.hpp:71
	static constexpr uint32_t RD_LIMIT = 48 * 32;  // constants.hpp:69
	static constexpr int64_t  ST_MIN   = INT64_MIN;
	static constexpr int64_t  ST_MAX   = INT64_MAX;
	static constexpr int64_t  SAFE_SENTINEL = 9223372036854774784LL;  // (int64)ENCODING_UPPER_LIMIT, encoder.hpp:85

	__device__ static __forceinline__ double exp10(int e) { return C_F64_EXP[e]; }
	__device__ static __forceinline__ double frac10(int e) { return C_F64_FRAC[e]; }
	__device__ static __forceinline__ int64_t fact10(int f) { return C_F64_FACT[f]; }
	__device__ static __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
	// encoder.hpp:87 with MAGIC_NUMBER = 2^52 + 2^51 (constants.hpp:70)
	__device__ static __forceinline__ double magic_round(double t) {
		return __dsub_rn(__dadd_rn(t, 6755399441055744.0), 6755399441055744.0);
	}
	// x86 cvttsd2si: NaN / out-of-range → 0x8000000000000000 (PTX cvt.rzi would saturate and map NaN to 0)
	__device__ static __forceinline__ int64_t cast_x86(double t) {
		// |t| < 2^63 is false for NaN and for t == -2^63, whose conversion is INT64_MIN anyway
		return fabs(t) < 9223372036854775808.0 ? __double2ll_rz(t) : INT64_MIN;
	}
	__device__ static __forceinline__ double to_float(int64_t x) { return __ll2double_rn(x); }
	__device__ static __forceinline__ uint64_t bits(double v) { return (uint64_t)__double_as_longlong(v); }
	__device__ static __forceinline__ double from_bits(uint64_t b) { return __longlong_as_double((long long)b); }
	// encoder.hpp:326-331 with Constants<double>: the NaN/Inf half of the test can never fire (constants.hpp:82-83,
	// the mask literal has 65 digits); only -0.0 is pre-replaced.  NaN/±Inf become exceptions through the compare.
	__device__ static __forceinline__ bool is_special(uint64_t b) { return b == 0x8000000000000000ULL; }
	__device__ static __forceinline__ double upper_limit() { return 9223372036854774784.0; }  // constants.hpp:17
	__device__ static __forceinline__ int bitlen(uint64_t x) { return 64 - __clzll((long long)x); }
};

template <>
struct Traits<float> {
	using UT = uint32_t;
	using ST = int32_t;
	static constexpr int      TBITS    = 32;
	static constexpr int      LANES    = 32;
	static constexpr int      MAX_EXP  = 10;   // constants.hpp:37
	static constexpr int      N_COMBOS = 66;
	static constexpr uint32_t EXC_BITS = 32;   // constants.hpp:35
	static constexpr uint32_t RD_LIMIT = 22 * 32;  // constants.hpp:33
	static constexpr int32_t  ST_MIN   = INT32_MIN;
	static constexpr int32_t  ST_MAX   = INT32_MAX;
	// encoder.hpp:85 converts the double constant 2^63-1024 to int32 at compile time; g++ folds that by saturation
	static constexpr int32_t  SAFE_SENTINEL = INT32_MAX;

	__device__ static __forceinline__ float exp10(int e) { return C_F32_EXP[e]; }
	__device__ static __forceinline__ float frac10(int e) { return C_F32_FRAC[e]; }
	__device__ static __forceinline__ int32_t fact10(int f) { return C_F32_FACT[f]; }
	__device__ static __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
	// MAGIC_NUMBER = 2^23 + 2^22 (constants.hpp:34)
	__device__ static __forceinline__ float magic_round(float t) { return __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f); }
	// x86 cvttss2si (32-bit destination)
	__device__ static __forceinline__ int32_t cast_x86(float t) {
		return fabsf(t) < 2147483648.0f ? __float2int_rz(t) : INT32_MIN;
	}
	__device__ static __forceinline__ float to_float(int32_t x) { return __int2float_rn(x); }
	__device__ static __forceinline__ uint32_t bits(float v) { return __float_as_uint(v); }
	__device__ static __forceinline__ float from_bits(uint32_t b) { return __uint_as_float(b); }
	// encoder.hpp:326-331 with Constants<float> (constants.hpp:41-46): NaN, ±Inf, -0.0
	__device__ static __forceinline__ bool is_special(uint32_t b) { return (b & 0x7FFFFFFFu) >= 0x7F800000u || b == 0x80000000u; }
	__device__ static __forceinline__ float upper_limit() { return 9223372036854775808.0f; }  // (float)ENCODING_UPPER_LIMIT
	__device__ static __forceinline__ int bitlen(uint32_t x) { return 32 - __clz((int)x); }
};

// encoder.hpp:74-78 is_impossible_to_encode, evaluated on the scaled value
template <typename PT>
__device__ __forceinline__ bool impossible_to_encode(PT t) {
	const double d = (double)t;
	return !isfinite(t) || d > 9223372036854774784.0 || d < -9223372036854774784.0 || (t == (PT)0 && signbit(t));
}

// encoder.hpp:81-89 encode_value<SAFE>
template <typename PT, bool SAFE>
__device__ __forceinline__ typename Traits<PT>::ST encode_value(PT v, PT exp10, PT frac10) {
	using T = Traits<PT>;
	PT t    = T::mul(T::mul(v, exp10), frac10);
	if (SAFE) {
		if (impossible_to_encode<PT>(t)) { return T::SAFE_SENTINEL; }
	}
	return T::cast_x86(T::magic_round(t));
}

// decoder.hpp:128-131 decode_value: wrapping integer multiply, RN convert, one RN multiply
template <typename PT>
__device__ __forceinline__ PT decode_value(typename Traits<PT>::ST enc, typename Traits<PT>::ST fact, PT frac) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	using ST = typename T::ST;
	return T::mul(T::to_float((ST)((UT)enc * (UT)fact)), frac);
}

// Helpers of the floating-point fast path of the exception test (derivation: alp_encode.cuh, "Two implementations of
// the per-value step"): key = |product| as an unsigned integer pattern (high word for doubles), BIG = the pattern of
// 2^63 | 2^31, fact_fp = 10^f as an exactly representable PT, cast_sat = the plain (saturating) conversion.
template <typename PT>
struct FastLimits;
template <>
struct FastLimits<double> {
	static constexpr uint32_t BIG = 0x43E00000u;  // high word of 2^63
	__device__ static __forceinline__ uint32_t key(double pd) { return (uint32_t)((uint64_t)__double_as_longlong(pd) >> 32) & 0x7FFFFFFFu; }
	__device__ static __forceinline__ double fact_fp(int f) { return C_F64_EXP[f]; }  // 10^f, exact
	__device__ static __forceinline__ int64_t cast_sat(double tr) { return __double2ll_rz(tr); }
};
template <>
struct FastLimits<float> {
	static constexpr uint32_t BIG = 0x4F000000u;  // 2^31
	__device__ static __forceinline__ uint32_t key(float pd) { return __float_as_uint(pd) & 0x7FFFFFFFu; }
	__device__ static __forceinline__ float fact_fp(int f) { return f < 10 ? C_F32_EXP[f] : 0.0f; }  // FACT[10] = 0 (alp_device.cuh)
	__device__ static __forceinline__ int32_t cast_sat(float tr) { return __float2int_rz(tr); }
};

// encoder.hpp:91-107 count_bits(max, min)
template <typename PT>
__device__ __forceinline__ int bits_of_range(typename Traits<PT>::ST mx, typename Traits<PT>::ST mn) {
	using UT   = typename Traits<PT>::UT;
	const UT d = (UT)mx - (UT)mn;
	return d == 0 ? 0 : Traits<PT>::bitlen(d);
}

// ---------------------------------------------------------------------------------------------------------------
// Thread -> value mapping.  64-bit lanes: thread (lane = t&15, half = t>>4) owns rows 32*half .. 32*half+31 of FastLanes
// lane `lane`, i.e. values 16*(32*half + r) + lane.  32-bit lanes: thread t owns lane t, values 32*r + t.
// ---------------------------------------------------------------------------------------------------------------
template <typename PT>
struct Map;
template <>
struct Map<double> {
	__device__ static __forceinline__ int index(int t, int r) { return 512 * (t >> 4) + 16 * r + (t & 15); }
	// thread and row that own value index v
	__device__ static __forceinline__ int thread_of(int v) { return (v & 15) + 16 * (v >> 9); }
	__device__ static __forceinline__ int row_of(int v) { return (v >> 4) & 31; }
};
template <>
struct Map<float> {
	__device__ static __forceinline__ int index(int t, int r) { return 32 * r + t; }
	__device__ static __forceinline__ int thread_of(int v) { return v & 31; }
	__device__ static __forceinline__ int row_of(int v) { return v >> 5; }
};

// ---------------------------------------------------------------------------------------------------------------
// Warp helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t shfl_xor_i64(int64_t v, int m) {
	int lo = __shfl_xor_sync(FULL, (int)(uint32_t)v, m);
	int hi = __shfl_xor_sync(FULL, (int)(uint32_t)((uint64_t)v >> 32), m);
	return (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
}
__device__ __forceinline__ int32_t shfl_xor_i64(int32_t v, int m) { return __shfl_xor_sync(FULL, v, m); }

// warp-wide signed min / max on the hardware reduction (REDUX): 32-bit directly; 64-bit as high word first, then the low
// words of the lanes that hold the winning high word (4 instructions instead of 5 shuffle + compare/select rounds)
__device__ __forceinline__ int32_t warp_min_impl(int32_t v) { return __reduce_min_sync(FULL, v); }
__device__ __forceinline__ int32_t warp_max_impl(int32_t v) { return __reduce_max_sync(FULL, v); }
__device__ __forceinline__ int64_t warp_min_impl(int64_t v) {
	const int32_t  hi  = (int32_t)(v >> 32);
	const uint32_t lo  = (uint32_t)v;
	const int32_t  mhi = __reduce_min_sync(FULL, hi);
	const uint32_t mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xFFFFFFFFu);
	return (int64_t)(((uint64_t)(uint32_t)mhi << 32) | mlo);
}
__device__ __forceinline__ int64_t warp_max_impl(int64_t v) {
	const int32_t  hi  = (int32_t)(v >> 32);
	const uint32_t lo  = (uint32_t)v;
	const int32_t  mhi = __reduce_max_sync(FULL, hi);
	const uint32_t mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
	return (int64_t)(((uint64_t)(uint32_t)mhi << 32) | mlo);
}
template <typename ST>
__device__ __forceinline__ ST warp_min(ST v) {
	return warp_min_impl(v);
}
template <typename ST>
__device__ __forceinline__ ST warp_max(ST v) {
	return warp_max_impl(v);
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
	uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src);
	uint32_t hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
	return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total) {
	uint32_t x = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t y = __shfl_up_sync(FULL, x, d);
		if (lane >= d) { x += y; }
	}
	total = __shfl_sync(FULL, x, 31);
	return x - v;
}

// ---------------------------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (TMA 1-D, no tensor map): SASS UBLKCP / SYNCS
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "LAB_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
	    "@P1 bra DONE;\n"
	    "bra LAB_WAIT;\n"
	    "DONE:\n"
	    "}" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
// generic-proxy writes to shared memory must be fenced before an async-proxy (bulk) read of them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// FastLanes interleaved layout, read side (src/fastlanes_generated_unffor.cpp:6389-6500).
// `blk` is a packed block copied verbatim into shared memory: word w of lane l of a T-bit-lane stream sits at
// element LANES*w + l.  extract() returns the bw-bit field of row `row` (no base added).  It may touch the
// word-row after the block (STAGE_PAD), whose contents are masked off.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t extract64(const uint64_t* blk, int lane, uint32_t bit, uint64_t mask) {
	const uint32_t w  = bit >> 6;
	const uint32_t sh = bit & 63;
	const uint64_t lo = blk[16 * w + lane];
	const uint64_t hi = blk[16 * (w + 1) + lane];
	return ((lo >> sh) | ((hi << 1) << (63 - sh))) & mask;
}
__device__ __forceinline__ uint32_t extract32(const uint32_t* blk, int lane, uint32_t bit, uint32_t mask) {
	const uint32_t w  = bit >> 5;
	const uint32_t sh = bit & 31;
	return __funnelshift_r(blk[32 * w + lane], blk[32 * (w + 1) + lane], sh) & mask;
}
__device__ __forceinline__ uint32_t extract16(const uint16_t* blk, int lane, uint32_t bit, uint32_t mask) {
	const uint32_t w  = bit >> 4;
	const uint32_t sh = bit & 15;
	const uint32_t lo = blk[64 * w + lane];
	const uint32_t hi = blk[64 * (w + 1) + lane];
	return ((lo | (hi << 16)) >> sh) & mask;
}

template <typename UT>
__device__ __forceinline__ UT low_mask(int bw) {
	return bw >= (int)(8 * sizeof(UT)) ? (UT) ~(UT)0 : (UT)((((UT)1) << bw) - 1);
}

// dictionary lookup in the 16-byte rd_dict image held in four 32-bit registers (rd.hpp:166)
__device__ __forceinline__ uint32_t dict_lookup(const uint4& d, uint32_t idx) {
	const uint32_t w = (idx & 4) ? ((idx & 2) ? d.w : d.z) : ((idx & 2) ? d.y : d.x);
	return (idx & 1) ? (w >> 16) : (w & 0xFFFFu);
}

}  // namespace alpb200
