// alp_scan.cuh — fused decode + SUM: the compressed column is decoded in registers and aggregated; nothing is written
// back.  This is the reference's scan primitive `alp_func` (AVX-512 falp + patch_exceptions into a thread-private
// 8 KiB buffer) followed by `aggr_plus` (publication/source_code/bench_end_to_end/src/benchmarks/alp/queries/q1.cpp:63-102)
// as one kernel: the write-bound decode (8 B/value out) becomes a read-bound scan (~bw/8 B/value in).
//
// Same machinery as decode_kernel: persistent warps, chunks of vectors drawn from a global counter, packed blocks
// fetched by bulk-async copies (TMA 1-D) into a double-buffered stage, width-specialised unpack.  Exceptions: the
// vector is summed with its fill values first, then every exception adds (true value - decoded fill value).  Floating-
// point addition order is not fixed (per-thread partial sums, warp tree, one atomicAdd per warp).
#pragma once

#include "alp_decode.cuh"
#include "alp_encode.cuh"

namespace alpb200 {

// decoded value at position p of an ALP vector, recomputed from the stage (run-time width)
__device__ __forceinline__ double alp_value_at(const uint8_t* stage, const MetaRegs& m, uint32_t p, double, bool /*decimal*/ = false) {
	using T              = Traits<double>;
	const uint32_t bw    = m.bw();
	const uint64_t d     = bw ? extract64(reinterpret_cast<const uint64_t*>(stage), p & 15, (p >> 4) * bw, low_mask<uint64_t>(bw)) : 0;
	return decode_value<double>((int64_t)(d + m.base()), T::fact10(m.f()), T::frac10(m.e()));
}
__device__ __forceinline__ double alp_value_at(const uint8_t* stage, const MetaRegs& m, uint32_t p, float, bool decimal = false) {
	using T           = Traits<float>;
	const uint32_t bw = m.bw();
	const uint32_t d  = bw ? extract32(reinterpret_cast<const uint32_t*>(stage), p & 31, (p >> 5) * bw, low_mask<uint32_t>(bw)) : 0;
	if (decimal) {  // the slot's value as sum_alp_vector's decimal path counted it
		return __dmul_rn(__dmul_rn((double)(int32_t)(d + m.a.x), __ll2double_rn(Traits<double>::fact10(m.f()))), Traits<double>::frac10(m.e()));
	}
	return (double)decode_value<float>((int32_t)(d + m.a.x), T::fact10(m.f()), T::frac10(m.e()));
}
// May a float vector take the decimal path?  Every slot's integer times 10^f must stay inside int32 (then the reference's
// wrapping 32-bit product is the real product) and the FACT table entry must be the real power of ten (f <= 9).
__device__ __forceinline__ bool sum_decimal_ok(const MetaRegs& m) {
	const uint32_t bw = m.bw(), f = m.f();
	if (bw > 31 || f > 9) { return false; }
	const int64_t b   = (int64_t)(int32_t)m.a.x;
	const int64_t mag = (b < 0 ? -b : b) + (1ll << bw);
	return mag * Traits<double>::fact10(f) < (1ll << 31);
}

// Sum of the thread's 32 rows of an ALP vector, exception slots counted with their fill values.
//
// f64 fast path.  decode_value(x) = fl(fl(x * 10^f) * 10^-e) per value; summing those doubles in any order has an error
// bound of order n * eps * sum|x|.  The thread instead sums its 32 INTEGERS exactly (x_r = d_r + base) and converts
// once: fl(fl(fl(X) * 10^f) * 10^-e) with X = sum of x_r — three roundings for 32 values instead of two per value plus
// 32 additions, i.e. at least as accurate, and ~3 instructions per value (field extract + integer add) instead of ~12
// (64-bit add, 64-bit multiply, I2F.F64.S64, DMUL, DADD).  Taken when neither the integer sum nor the conversion can
// overflow: bw <= 57 and |base| < 2^56 (=> |X| < 2^62).  (10^f, f <= 18, is exact in double.)
// The float path keeps the per-value decode: a float column's SUM is the sum of its FLOAT values, and fl32 rounding
// of each value is visible at double precision.
__device__ __forceinline__ double sum_alp_vector(const uint8_t* stage, const MetaRegs& m, int t, double, bool /*decimal*/ = false) {
	using T             = Traits<double>;
	const int64_t  fact = T::fact10(m.f());
	const double   frac = T::frac10(m.e());
	const uint64_t base = m.base();
	const uint32_t bw   = m.bw();
	const int64_t  sb   = (int64_t)base;
	const bool     fast = bw <= 57 && sb < (1ll << 56) && sb > -(1ll << 56);
	if (fast) {
		uint64_t total = 0;
		dispatch_width<0, 57>(bw, [&](auto W) {
			constexpr int BW = decltype(W)::value;
			if constexpr (BW <= 27) {  // 32 fields of <= 27 bits: the sum fits 32 bits
				uint32_t acc = 0;
				unpack64_rows<BW>(stage, t & 15, t >> 4, [&](int, uint32_t lo, uint32_t) { acc += lo; });
				total = acc;
			} else if constexpr (BW <= 32) {  // two 16-bit halves, each sum fits 32 bits
				uint32_t acc_lo = 0, acc_hi = 0;
				unpack64_rows<BW>(stage, t & 15, t >> 4, [&](int, uint32_t lo, uint32_t) {
					acc_lo += lo & 0xFFFFu;
					acc_hi += lo >> 16;
				});
				total = (uint64_t)acc_lo + ((uint64_t)acc_hi << 16);
			} else {
				uint64_t acc = 0;
				unpack64_rows<BW>(stage, t & 15, t >> 4, [&](int, uint32_t lo, uint32_t hi) { acc += ((uint64_t)hi << 32) | lo; });
				total = acc;
			}
		});
		const int64_t X = (int64_t)total + 32 * sb;
		return __dmul_rn(__dmul_rn(__ll2double_rn(X), __ll2double_rn(fact)), frac);
	}
	// rare: very wide fields or a huge base — per-value decode with run-time extraction (small code, not fast)
	double         acc[4] = {0.0, 0.0, 0.0, 0.0};
	const uint64_t msk    = low_mask<uint64_t>((int)bw);
#pragma unroll 4
	for (int r = 0; r < 32; r++) {
		const uint32_t p = (uint32_t)Map<double>::index(t, r);
		const uint64_t d = bw ? extract64(reinterpret_cast<const uint64_t*>(stage), p & 15, (p >> 4) * bw, msk) : 0;
		acc[r & 3] += decode_value<double>((int64_t)(d + base), fact, frac);
	}
	return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}
// Floats, two semantics (alpb200_decode_sum_ex_f32):
//   per value (default)   the sum, in double, of the FLOAT values the decoder produces — what decode + add gives, bit for bit up to
//                         the order of the additions.  ~8 instructions per value (wrapping multiply, I2F, FMUL, F2F, DADD).
//   ALPB200_SUM_DECIMAL   the sum of the DECIMALS the floats stand for: the thread adds its 32 integers exactly and converts once,
//                         X * 10^f * 10^-e in double (the f64 recipe).  Differs from the per-value sum by the floats' own rounding:
//                         |difference| <= 2^-23 * sum |x_i| (each float is within 2^-24 of its decimal, relatively, and the float
//                         constant 10^-e within 2^-24 of the real one).  ~3 instructions per value.
__device__ __forceinline__ double sum_alp_vector(const uint8_t* stage, const MetaRegs& m, int t, float, bool decimal = false) {
	using T             = Traits<float>;
	const int32_t  fact = T::fact10(m.f());
	const float    frac = T::frac10(m.e());
	const uint32_t base = m.a.x;
	if (decimal) {
		uint64_t total = 0;
		dispatch_width<0, 31>(m.bw(), [&](auto W) {
			constexpr int BW = decltype(W)::value;
			if constexpr (BW <= 27) {  // 32 fields of <= 27 bits: the sum fits 32 bits
				uint32_t acc = 0;
				unpack32_rows<BW>(stage, t, [&](int, uint32_t d) { acc += d; });
				total = acc;
			} else {
				uint32_t acc_lo = 0, acc_hi = 0;
				unpack32_rows<BW>(stage, t, [&](int, uint32_t d) {
					acc_lo += d & 0xFFFFu;
					acc_hi += d >> 16;
				});
				total = (uint64_t)acc_lo + ((uint64_t)acc_hi << 16);
			}
		});
		const int64_t X = (int64_t)total + 32 * (int64_t)(int32_t)base;
		return __dmul_rn(__dmul_rn(__ll2double_rn(X), __ll2double_rn(Traits<double>::fact10(m.f()))), Traits<double>::frac10(m.e()));
	}
	double         acc[4] = {0.0, 0.0, 0.0, 0.0};
	dispatch_width<0, 32>(m.bw(), [&](auto W) {
		constexpr int BW = decltype(W)::value;
		unpack32_rows<BW>(stage, t, [&](int r, uint32_t d) { acc[r & 3] += (double)decode_value<float>((int32_t)(d + base), fact, frac); });
	});
	return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// ALP_RD value at position p: (left << right_bw) | right, left from the dictionary or, for exceptions, given
template <typename PT>
__device__ __forceinline__ double rd_value(const uint8_t* stage, const MetaRegs& m, uint32_t p, bool use_left, uint32_t left_part) {
	using T   = Traits<PT>;
	using UT  = typename T::UT;
	const uint32_t rbw = m.bw(), lbw = m.e();
	const UT       right = rd_right_at(stage, rbw, p, UT());
	if (!use_left) {
		const uint32_t idx = extract16(reinterpret_cast<const uint16_t*>(stage + 128u * rbw), p & 63, (p >> 6) * lbw, (1u << lbw) - 1);
		left_part          = dict_lookup(m.a, idx);
	}
	return (double)T::from_bits(((UT)left_part << rbw) | right);
}
// sum of the thread's 32 rows of an ALP_RD vector, exception slots counted with their dictionary values
template <typename PT>
__device__ __forceinline__ double sum_rd_vector(const uint8_t* stage, const MetaRegs& m, int t) {
	using UT      = typename Traits<PT>::UT;
	double acc[4] = {0.0, 0.0, 0.0, 0.0};
	rd_unpack_rows(stage, m, t, PT(), [&](int r, UT bits) { acc[r & 3] += (double)Traits<PT>::from_bits(bits); });
	return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// the thread's share of a vector's sum straight from global memory (slow generic path, see hint_check_kernel)
template <typename PT>
__device__ __forceinline__ double sum_vector_slow(const ColView& col, const MetaRegs& m, int t) {
	using UT           = typename Traits<PT>::UT;
	const uint8_t* blk = col.packed + (uint64_t)m.packed_off() * 128u;
	double         acc = 0.0;
#pragma unroll 1
	for (int i = t; i < VEC; i += 32) {
		acc += (double)Traits<PT>::from_bits(value_bits_slow<PT>(blk, m, (uint32_t)i));
	}
	const UT*       ev = static_cast<const UT*>(col.exc_val) + m.exc_off();
	const uint16_t* ep = col.exc_pos + m.exc_off();
	const bool      rd = m.scheme() != ALPB200_SCHEME_ALP;
#pragma unroll 1
	for (uint32_t i = t; i < m.exc_cnt(); i += 32) {
		const uint32_t p    = ep[i];
		const UT       fill = value_bits_slow<PT>(blk, m, p);
		UT             v    = ev[i];
		if (rd) { v = (UT)(((v & 0xFFFFu) << m.bw()) | (fill & low_mask<UT>((int)m.bw()))); }
		acc += (double)Traits<PT>::from_bits(v) - (double)Traits<PT>::from_bits(fill);
	}
	return acc;
}

// The first 32 K exceptions of a vector (K per lane: lane t holds ranks t, t + 32, ...), fetched one vector ahead so that the
// exception loop never waits for memory.
template <typename UT, int K>
struct ExcRegsK {
	UT       val[K];
	uint16_t raw_pos[K];  // (nothing is computed on the loaded words here: a use would make the warp wait for the load right away)
	__device__ __forceinline__ uint32_t pos(int k) const { return raw_pos[k]; }
};
template <typename UT, int K>
__device__ __forceinline__ ExcRegsK<UT, K> load_exceptions_k(const ColView& col, const MetaRegs& m, int t) {
	ExcRegsK<UT, K> x;
	const uint32_t  cnt = m.exc_cnt();
	const UT*       ev  = static_cast<const UT*>(col.exc_val) + m.exc_off();
	const uint16_t* ep  = col.exc_pos + m.exc_off();
#pragma unroll
	for (int k = 0; k < K; k++) {
		const uint32_t i = (uint32_t)t + 32u * k;
		x.val[k]         = 0;
		x.raw_pos[k]     = 0;
		if (i < cnt) {
			x.val[k]     = __ldg(ev + i);
			x.raw_pos[k] = __ldg(ep + i);
		}
	}
	return x;
}
template <typename PT>
struct SumCfg {
	static constexpr int EXC_K = sizeof(PT) == 4 ? 4 : 1;  // floats: exception-heavy columns are the norm (6 bytes per exception)
};

// register allocation limited for 32 resident warps per SM (64 registers, no spills): against 24 x 80 registers the f64
// scan gains 4-8 % (0.293 -> 0.280 ms per 2^29 values on config 2)
template <typename PT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 32 / WARPS) decode_sum_kernel(ColView col, uint64_t first_vector, uint64_t n_vectors,
                                                                  double* __restrict__ sum, uint32_t stage_bytes,
                                                                  unsigned long long* __restrict__ counter,
                                                                  const unsigned long long* __restrict__ oversize, uint32_t flags) {
	using UT = typename Traits<PT>::UT;
	constexpr int K = SumCfg<PT>::EXC_K;
	using XR        = ExcRegsK<UT, K>;
	extern __shared__ __align__(128) uint8_t smem[];
	const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
	if (oversize != nullptr && *oversize != 0) {  // the hint was too small for this call (hint_check_kernel): slow, correct
		const uint64_t n_warps = (uint64_t)gridDim.x * WARPS;
		double         part    = 0.0;
		for (uint64_t w = (uint64_t)blockIdx.x * WARPS + warp; w < n_vectors; w += n_warps) {
			part += sum_vector_slow<PT>(col, load_meta(col.meta + first_vector + w), t);
		}
#pragma unroll
		for (int m = 16; m > 0; m >>= 1) {
			part += __longlong_as_double((long long)shfl_xor_i64((int64_t)__double_as_longlong(part), m));
		}
		if (t == 0) { atomicAdd(sum, part); }
		return;
	}
	uint8_t*  stage = smem + (size_t)warp * 2 * stage_bytes;
	uint64_t* bars  = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * 2 * stage_bytes) + 2 * warp;
	if (t == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
		fence_mbar_init();
	}
	__syncwarp();

	constexpr uint32_t CHUNK = 16, REFILL_AT = 6;
	auto draw = [&]() -> uint64_t {
		unsigned long long b = 0;
		if (t == 0) { b = atomicAdd(counter, (unsigned long long)CHUNK); }
		return shfl_u64(b, 0);
	};
	uint64_t chunk_base = draw(), next_base = 0;
	uint32_t chunk_used = 0;
	auto     take       = [&]() -> uint64_t {
        if (chunk_used == CHUNK) {
            chunk_base = next_base;
            chunk_used = 0;
        }
        const uint64_t idx = chunk_base + chunk_used++;
        if (chunk_used == CHUNK - REFILL_AT) { next_base = draw(); }
        return idx;
	};
	uint64_t v = take(), v_next = take();
	if (v >= n_vectors) { return; }
	const alpb200_vec_meta* meta = col.meta + first_vector;
	auto issue = [&](const MetaRegs& m, int s) {
		const uint32_t bytes = m.block_bytes();
		if (t == 0 && bytes != 0) {
			mbar_arrive_expect_tx(&bars[s], bytes);
			bulk_g2s(stage + (size_t)s * stage_bytes, col.packed + (uint64_t)m.packed_off() * 128u, bytes, &bars[s]);
		}
	};
	MetaRegs cur = load_meta(meta + v);
	bool     has_next = v_next < n_vectors;
	MetaRegs nxt      = cur;
	if (has_next) { nxt = load_meta(meta + v_next); }
	issue(cur, 0);
	XR       xcur  = load_exceptions_k<UT, K>(col, cur, t);
	uint32_t phase = 0;
	double   acc   = 0.0;
	for (int s = 0;; s ^= 1) {
		XR xnxt = xcur;
		if (has_next) {
			issue(nxt, s ^ 1);
			xnxt = load_exceptions_k<UT, K>(col, nxt, t);
			if (nxt.exc_cnt() > 32u * K) { prefetch_exception_tail(col, nxt, t, sizeof(UT)); }
		}
		const uint64_t v_nn   = has_next ? take() : v_next;
		const bool     has_nn = has_next && v_nn < n_vectors;
		MetaRegs       nn     = nxt;
		if (has_nn) { nn = load_meta(meta + v_nn); }
		const uint8_t* stg = stage + (size_t)s * stage_bytes;
		if (cur.block_bytes() != 0) {
			mbar_wait(&bars[s], (phase >> s) & 1u);
			phase ^= 1u << s;
		}
		const uint32_t  cnt = cur.exc_cnt();
		const UT*       ev  = static_cast<const UT*>(col.exc_val) + cur.exc_off();
		const uint16_t* ep  = col.exc_pos + cur.exc_off();
		if (cur.scheme() == ALPB200_SCHEME_ALP) {
			bool decimal = false;
			if constexpr (sizeof(PT) == 4) { decimal = (flags & ALPB200_SUM_DECIMAL) != 0 && sum_decimal_ok(cur); }
			acc += sum_alp_vector(stg, cur, t, PT(), decimal);
			// exception: + true value - what the slot was counted as
#pragma unroll
			for (int k = 0; k < K; k++) {
				if ((uint32_t)t + 32u * k < cnt) { acc += (double)Traits<PT>::from_bits(xcur.val[k]) - alp_value_at(stg, cur, xcur.pos(k), PT(), decimal); }
			}
			for (uint32_t i = (uint32_t)t + 32u * K; i < cnt; i += 32) {
				acc += (double)Traits<PT>::from_bits(ev[i]) - alp_value_at(stg, cur, ep[i], PT(), decimal);
			}
		} else {
			acc += sum_rd_vector<PT>(stg, cur, t);
#pragma unroll
			for (int k = 0; k < K; k++) {
				if ((uint32_t)t + 32u * k < cnt) {
					const uint32_t p = xcur.pos(k);
					acc += rd_value<PT>(stg, cur, p, true, (uint32_t)(xcur.val[k] & 0xFFFFu)) - rd_value<PT>(stg, cur, p, false, 0);
				}
			}
			for (uint32_t i = (uint32_t)t + 32u * K; i < cnt; i += 32) {
				const uint32_t p = ep[i];
				acc += rd_value<PT>(stg, cur, p, true, (uint32_t)(ev[i] & 0xFFFFu)) - rd_value<PT>(stg, cur, p, false, 0);
			}
		}
		__syncwarp();
		if (!has_next) { break; }
		cur      = nxt;
		nxt      = nn;
		xcur     = xnxt;
		has_next = has_nn;
		v        = v_next;
		v_next   = v_nn;
	}
#pragma unroll
	for (int m = 16; m > 0; m >>= 1) {
		acc += __longlong_as_double((long long)shfl_xor_i64((int64_t)__double_as_longlong(acc), m));
	}
	if (t == 0) { atomicAdd(sum, acc); }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Fused decode + MIN / MAX / COUNT (alpb200_decode_minmax_*): the scan-side siblings of SUM that an engine's zone maps and
// filters ask for.  The reference ships no such scan (its scan query is alp_func + aggr_plus, q1.cpp:63-102); the semantics
// are those of decode + reduce: MIN / MAX over the decoded values with NaNs ignored, COUNT = number of non-NaN values.
// Nothing depends on what a writer left in exception slots: the exceptions' rows are marked in a per-warp bitmap (one atomicOr
// per exception) and skipped by their owners, the exceptions' true values are taken from the exception arrays.
// Nothing is written back to HBM: read-bound like SUM, the same pipeline (persistent warps, chunks, TMA stages).
// ---------------------------------------------------------------------------------------------------------------------------
struct MinMaxOut {  // = alpb200_minmax (include/alp_b200.h)
	double             min, max;
	unsigned long long count;
};
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
	unsigned long long* a   = reinterpret_cast<unsigned long long*>(addr);
	unsigned long long  old = *a;
	while (v < __longlong_as_double((long long)old)) {
		const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
		if (seen == old) { break; }
		old = seen;
	}
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
	unsigned long long* a   = reinterpret_cast<unsigned long long*>(addr);
	unsigned long long  old = *a;
	while (v > __longlong_as_double((long long)old)) {
		const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
		if (seen == old) { break; }
		old = seen;
	}
}
static __global__ void minmax_init_kernel(MinMaxOut* out) {
	out->min   = __longlong_as_double(0x7FF0000000000000ll);   // +inf
	out->max   = __longlong_as_double((long long)0xFFF0000000000000ull);  // -inf
	out->count = 0;
}

// One vector's MIN / MAX / COUNT into the thread's running (mn, mx, cnt).  `excrows`: bit r = the thread's row r is an exception
// slot (skipped here; the exceptions' true values are added by the caller).
// Decoding is monotonic in the encoded integer (multiplications by positive constants, round-to-nearest) as long as no integer
// product wraps, so the thread takes the unsigned min / max of its FIELDS (2-4 instructions per value) and decodes just those
// two with the reference's recipe — bit-identical to decoding all 32 and comparing.  Wrapping is judged on the two candidates
// themselves (if neither base + field nor its product with 10^f wraps for the smallest and the largest integer, nothing in
// between does); a vector with a candidate that might wrap — possible only in columns whose non-exception slots do not
// round-trip — is decoded slot by slot instead.
template <typename PT>
__device__ __forceinline__ void minmax_alp_vector(const uint8_t* stage, const MetaRegs& m, int t, uint32_t excrows, double& mn, double& mx, uint32_t& cnt) {
	using T            = Traits<PT>;
	using UT           = typename T::UT;
	using ST           = typename T::ST;
	const uint32_t bw  = m.bw();
	const uint32_t own = ~excrows;
	cnt += __popc(own);  // a decoded ALP slot is never NaN
	UT lo_f = (UT) ~(UT)0, hi_f = 0;
	if constexpr (sizeof(PT) == 8) {
		dispatch_width<0, 32>(bw > 32 ? 0u : bw, [&](auto W) {  // (wider fields: slot by slot below; their unpack windows cost too many registers)
			constexpr int BW = decltype(W)::value;
			uint32_t      a = 0xFFFFFFFFu, b = 0;
			unpack64_rows<BW>(stage, t & 15, t >> 4, [&](int r, uint32_t lo, uint32_t) {
				if ((own >> r) & 1u) {
					a = min(a, lo);
					b = max(b, lo);
				}
			});
			lo_f = a;
			hi_f = b;
		});
	} else {
		dispatch_width<0, 32>(bw, [&](auto W) {
			constexpr int BW = decltype(W)::value;
			uint32_t      a = 0xFFFFFFFFu, b = 0;
			unpack32_rows<BW>(stage, t, [&](int r, uint32_t d) {
				if ((own >> r) & 1u) {
					a = min(a, d);
					b = max(b, d);
				}
			});
			lo_f = a;
			hi_f = b;
		});
	}
	const UT base = sizeof(PT) == 8 ? (UT)m.base() : (UT)m.a.x;
	const ST x_lo = (ST)(lo_f + base), x_hi = (ST)(hi_f + base), sbase = (ST)base;
	bool     ok   = sizeof(PT) == 4 || bw <= 32;
	if (own && ok) {
		ok = x_lo >= sbase && x_hi >= x_lo;  // base + field did not leave the signed range
		if constexpr (sizeof(PT) == 8) {
			// |x| * 10^f < 2^63, exactly (128-bit product): config 2 has vectors whose largest non-exception sits 0.0001 % below the limit
			const uint64_t f10 = (uint64_t)T::fact10(m.f());
			const uint64_t alo = (uint64_t)(x_lo < 0 ? -x_lo : x_lo), ahi = (uint64_t)(x_hi < 0 ? -x_hi : x_hi);
			ok = ok && __umul64hi(alo, f10) == 0 && (alo * f10) < (1ull << 63) && __umul64hi(ahi, f10) == 0 && (ahi * f10) < (1ull << 63);
		} else {
			const int64_t f10 = m.f() <= 9 ? Traits<double>::fact10(m.f()) : (1ll << 40);  // (float FACT[10] is the reference's 0: slot by slot)
			const int64_t a = (int64_t)x_lo * f10, b = (int64_t)x_hi * f10;
			ok = ok && a > -(1ll << 31) && a < (1ll << 31) && b > -(1ll << 31) && b < (1ll << 31);
		}
	}
	if (__all_sync(FULL, ok)) {
		if (own) {
			const double dlo = (double)decode_value<PT>(x_lo, T::fact10(m.f()), T::frac10(m.e()));
			const double dhi = (double)decode_value<PT>(x_hi, T::fact10(m.f()), T::frac10(m.e()));
			mn               = dlo < mn ? dlo : mn;
			mx               = dhi > mx ? dhi : mx;
		}
		return;
	}
	// rare: integer products that may wrap — every slot decoded with run-time extraction
#pragma unroll 1
	for (int r = 0; r < 32; r++) {
		if ((own >> r) & 1u) {
			const double x = alp_value_at(stage, m, (uint32_t)Map<PT>::index(t, r), PT());
			mn             = x < mn ? x : mn;
			mx             = x > mx ? x : mx;
		}
	}
}
// ALP_RD: any bit pattern can occur, so every slot is glued and compared.  Not inlined: the width-specialised unpack's window
// (up to 63 words) gets its own register allocation instead of competing with the kernel's pipeline state.
struct MinMaxAcc {
	double   mn, mx;
	uint32_t cnt;
};
template <typename PT>
__device__ __noinline__ MinMaxAcc minmax_rd_vector(const uint8_t* stage, MetaRegs m, int t, uint32_t excrows, MinMaxAcc in) {
	using UT   = typename Traits<PT>::UT;
	double   a = in.mn, b = in.mx;
	uint32_t c = 0;
	rd_unpack_rows(stage, m, t, PT(), [&](int r, UT bits) {
		const double x = (double)Traits<PT>::from_bits(bits);
		if (!((excrows >> r) & 1u) && x == x) {
			c++;
			a = x < a ? x : a;
			b = x > b ? x : b;
		}
	});
	in.mn = a;
	in.mx = b;
	in.cnt += c;
	return in;
}

template <typename PT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 32 / WARPS) decode_minmax_kernel(ColView col, uint64_t first_vector, uint64_t n_vectors,
                                                                     MinMaxOut* __restrict__ result, uint32_t stage_bytes,
                                                                     unsigned long long* __restrict__ counter,
                                                                     const unsigned long long* __restrict__ oversize) {
	using UT = typename Traits<PT>::UT;
	constexpr int K = SumCfg<PT>::EXC_K;
	using XR        = ExcRegsK<UT, K>;
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ uint32_t s_rows[WARPS][32];  // per vector: bit r of word j = row r of thread j is an exception slot
	const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
	double    mn = __longlong_as_double(0x7FF0000000000000ll), mx = -mn;
	uint32_t  cnt = 0;
	const alpb200_vec_meta* meta = col.meta + first_vector;
	auto take_value = [&](UT bits) {  // one true value (an exception's)
		const double x = (double)Traits<PT>::from_bits(bits);
		if (x == x) {
			cnt++;
			mn = x < mn ? x : mn;
			mx = x > mx ? x : mx;
		}
	};
	if (oversize != nullptr && *oversize != 0) {  // the hint was too small for this call (hint_check_kernel): slow, correct
		const uint64_t n_warps = (uint64_t)gridDim.x * WARPS;
		for (uint64_t w = (uint64_t)blockIdx.x * WARPS + warp; w < n_vectors; w += n_warps) {
			const MetaRegs  m   = load_meta(meta + w);
			const uint8_t*  blk = col.packed + (uint64_t)m.packed_off() * 128u;
			const UT*       ev  = static_cast<const UT*>(col.exc_val) + m.exc_off();
			const uint16_t* ep  = col.exc_pos + m.exc_off();
			const bool      rd  = m.scheme() != ALPB200_SCHEME_ALP;
			s_rows[warp][t]     = 0;
			__syncwarp();
			for (uint32_t i = t; i < m.exc_cnt(); i += 32) {
				const uint32_t p = ep[i];
				atomicOr(&s_rows[warp][p & 31], 1u << (p >> 5));  // (slow path: value i belongs to lane i % 32, row i / 32)
				UT v = ev[i];
				if (rd) { v = (UT)(((v & 0xFFFFu) << m.bw()) | (value_bits_slow<PT>(blk, m, p) & low_mask<UT>((int)m.bw()))); }
				take_value(v);
			}
			__syncwarp();
			const uint32_t skip = s_rows[warp][t];
#pragma unroll 1
			for (int r = 0; r < 32; r++) {
				if (!((skip >> r) & 1u)) { take_value(value_bits_slow<PT>(blk, m, (uint32_t)(32 * r + t))); }
			}
			__syncwarp();
		}
	} else {
		uint8_t*  stage = smem + (size_t)warp * 2 * stage_bytes;
		uint64_t* bars  = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * 2 * stage_bytes) + 2 * warp;
		if (t == 0) {
			mbar_init(&bars[0], 1);
			mbar_init(&bars[1], 1);
			fence_mbar_init();
		}
		__syncwarp();
		constexpr uint32_t CHUNK = 16, REFILL_AT = 6;
		auto draw = [&]() -> uint64_t {
			unsigned long long b = 0;
			if (t == 0) { b = atomicAdd(counter, (unsigned long long)CHUNK); }
			return shfl_u64(b, 0);
		};
		uint64_t chunk_base = draw(), next_base = 0;
		uint32_t chunk_used = 0;
		auto     take       = [&]() -> uint64_t {
            if (chunk_used == CHUNK) {
                chunk_base = next_base;
                chunk_used = 0;
            }
            const uint64_t idx = chunk_base + chunk_used++;
            if (chunk_used == CHUNK - REFILL_AT) { next_base = draw(); }
            return idx;
		};
		uint64_t v = take(), v_next = take();
		if (v < n_vectors) {
			auto issue = [&](const MetaRegs& m, int s) {
				const uint32_t bytes = m.block_bytes();
				if (t == 0 && bytes != 0) {
					mbar_arrive_expect_tx(&bars[s], bytes);
					bulk_g2s(stage + (size_t)s * stage_bytes, col.packed + (uint64_t)m.packed_off() * 128u, bytes, &bars[s]);
				}
			};
			MetaRegs cur      = load_meta(meta + v);
			bool     has_next = v_next < n_vectors;
			MetaRegs nxt      = cur;
			if (has_next) { nxt = load_meta(meta + v_next); }
			issue(cur, 0);
			XR       xcur  = load_exceptions_k<UT, K>(col, cur, t);
			uint32_t phase = 0;
			for (int s = 0;; s ^= 1) {
				XR xnxt = xcur;
				if (has_next) {
					issue(nxt, s ^ 1);
					xnxt = load_exceptions_k<UT, K>(col, nxt, t);
					if (nxt.exc_cnt() > 32u * K) { prefetch_exception_tail(col, nxt, t, sizeof(UT)); }
				}
				const uint64_t v_nn   = has_next ? take() : v_next;
				const bool     has_nn = has_next && v_nn < n_vectors;
				MetaRegs       nn     = nxt;
				if (has_nn) { nn = load_meta(meta + v_nn); }
				const uint8_t* stg = stage + (size_t)s * stage_bytes;
				if (cur.block_bytes() != 0) {
					mbar_wait(&bars[s], (phase >> s) & 1u);
					phase ^= 1u << s;
				}
				const uint32_t  n_exc = cur.exc_cnt();
				const UT*       ev    = static_cast<const UT*>(col.exc_val) + cur.exc_off();
				const uint16_t* ep    = col.exc_pos + cur.exc_off();
				const bool      rd    = cur.scheme() != ALPB200_SCHEME_ALP;
				const uint32_t  rbw   = cur.bw();
				// the exceptions: their rows are marked for the owners to skip, their true values are taken here
				uint32_t excrows = 0;
				if (n_exc != 0) {
					s_rows[warp][t] = 0;
					__syncwarp();
					auto one = [&](uint32_t p, UT val) {
						atomicOr(&s_rows[warp][Map<PT>::thread_of((int)p)], 1u << Map<PT>::row_of((int)p));
						take_value(rd ? (UT)(((val & 0xFFFFu) << rbw) | rd_right_at(stg, rbw, p, UT())) : val);
					};
#pragma unroll
					for (int k = 0; k < K; k++) {
						if ((uint32_t)t + 32u * k < n_exc) { one(xcur.pos(k), xcur.val[k]); }
					}
					for (uint32_t i = (uint32_t)t + 32u * K; i < n_exc; i += 32) {
						one(ep[i], ev[i]);
					}
					__syncwarp();
					excrows = s_rows[warp][t];
				}
				if (!rd) {
					minmax_alp_vector<PT>(stg, cur, t, excrows, mn, mx, cnt);
				} else {
					MinMaxAcc acc {mn, mx, cnt};
					acc = minmax_rd_vector<PT>(stg, cur, t, excrows, acc);
					mn  = acc.mn;
					mx  = acc.mx;
					cnt = acc.cnt;
				}
				__syncwarp();
				if (!has_next) { break; }
				cur      = nxt;
				nxt      = nn;
				xcur     = xnxt;
				has_next = has_nn;
				v        = v_next;
				v_next   = v_nn;
			}
		}
	}
#pragma unroll
	for (int m = 16; m > 0; m >>= 1) {
		const double omn = __longlong_as_double((long long)shfl_xor_i64((int64_t)__double_as_longlong(mn), m));
		const double omx = __longlong_as_double((long long)shfl_xor_i64((int64_t)__double_as_longlong(mx), m));
		mn               = omn < mn ? omn : mn;
		mx               = omx > mx ? omx : mx;
	}
	cnt = __reduce_add_sync(FULL, cnt);
	if (t == 0 && cnt != 0) {
		atomic_min_double(&result->min, mn);
		atomic_max_double(&result->max, mx);
		atomicAdd(&result->count, (unsigned long long)cnt);
	}
}

// ---------------------------------------------------------------------------------------------------------------------------
// Fused decode + predicate filter (alpb200_decode_filter_*): a selection bitmap instead of the decoded column — 1 bit per value
// out (128 bytes per vector) instead of 8 / 4 bytes.  Semantics = decode + compare, exact by construction: every slot is decoded
// with the reference's recipe IN REGISTERS (no tile: 32 resident warps like SUM) and compared, the warp's ballots are the bitmap
// words (bit j of word w of a vector = value 32 w + j), and the exceptions' bits are then corrected from their true values in a
// 128-byte shared-memory copy of the words; IEEE comparisons, i.e. a NaN satisfies only NE.  The reference ships no filter scan
// (its scan query is SUM, q1.cpp:63-102).
// ---------------------------------------------------------------------------------------------------------------------------
#ifndef ALPB200_FILTER_BLOCKS
#define ALPB200_FILTER_BLOCKS 3  // resident 8-warp blocks per SM the register allocation is limited for
#endif
// which of (less, equal, greater, unordered) satisfy the comparison: bits 0..3
__device__ __forceinline__ uint32_t filter_mask(uint32_t op) {
	switch (op) {
	case ALPB200_FILTER_LT: return 1u;
	case ALPB200_FILTER_LE: return 3u;
	case ALPB200_FILTER_GT: return 4u;
	case ALPB200_FILTER_GE: return 6u;
	case ALPB200_FILTER_EQ: return 2u;
	default: return 13u;  // NE: less, greater or unordered (NaN)
	}
}
__device__ __forceinline__ bool filter_test(double x, double c, uint32_t mask) {  // branch-free: the comparison is a per-launch constant
	const bool lt = x < c, eq = x == c, gt = x > c;
	return ((mask & 1u) && lt) || ((mask & 2u) && eq) || ((mask & 4u) && gt) || ((mask & 8u) && !(lt || eq || gt));
}
// per-thread results -> bitmap words.  A thread collects the outcomes of its 32 rows in one register (bit r = row r; no warp-wide
// step per row), a 32x32 bit-matrix transpose (transpose32, alp_encode.cuh: 5 shuffle rounds) turns that into "lane r holds the
// ballot of row r", and the layout does the rest: a thread's row r is value Map<PT>::index(t, r) — floats 32 r + t, so the ballot
// of row r IS word r; doubles 512 half + 16 r + lane, so the low / high halves of the ballots of rows 2 w and 2 w + 1 make words w
// and 16 + w (two shuffles).  Lane w ends up holding word w.
template <typename PT>
struct BitWords {
	uint32_t mask = 0;
	__device__ __forceinline__ void row(int r, bool pass, int) { mask |= (uint32_t)pass << r; }
	__device__ __forceinline__ uint32_t words(int t) const {
		const uint32_t x = transpose32(mask, t);
		if constexpr (sizeof(PT) == 4) {
			return x;
		} else {
			const uint32_t a = __shfl_sync(FULL, x, 2 * (t & 15)), b = __shfl_sync(FULL, x, 2 * (t & 15) + 1);
			return t < 16 ? ((a & 0xFFFFu) | (b << 16)) : ((a >> 16) | (b & 0xFFFF0000u));
		}
	}
};
// ALP_RD vectors (inlined: at the filter kernel's 80 registers a call costs more than the window's spills, 4.2 vs 2.3 ms per 2^30
// values of config 3 — the opposite of the MIN / MAX kernel at 64 registers)
template <typename PT>
__device__ __forceinline__ uint32_t filter_rd_vector(const uint8_t* stage, MetaRegs m, int t, double constant, uint32_t fmask) {
	using UT = typename Traits<PT>::UT;
	BitWords<PT> w;
	rd_unpack_rows(stage, m, t, PT(), [&](int r, UT bits) { w.row(r, filter_test((double)Traits<PT>::from_bits(bits), constant, fmask), t); });
	return w.words(t);
}

template <typename PT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, ALPB200_FILTER_BLOCKS) decode_filter_kernel(ColView col, uint64_t first_vector, uint64_t n_vectors, uint32_t op,
                                                                     double constant, uint32_t* __restrict__ bitmap,
                                                                     unsigned long long* __restrict__ selected, uint32_t stage_bytes,
                                                                     unsigned long long* __restrict__ counter,
                                                                     const unsigned long long* __restrict__ oversize) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	using ST = typename T::ST;
	constexpr int K = SumCfg<PT>::EXC_K;
	using XR        = ExcRegsK<UT, K>;
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ uint32_t s_bits[WARPS][32];  // a vector's bitmap words while its exceptions' bits are corrected
	const int      warp = threadIdx.x >> 5, t = threadIdx.x & 31;
	const uint32_t fmask = filter_mask(op);
	uint32_t       hits  = 0;
	const alpb200_vec_meta* meta = col.meta + first_vector;
	if (oversize != nullptr && *oversize != 0) {  // the hint was too small for this call (hint_check_kernel): slow, correct
		const uint64_t n_warps = (uint64_t)gridDim.x * WARPS;
		for (uint64_t w = (uint64_t)blockIdx.x * WARPS + warp; w < n_vectors; w += n_warps) {
			const MetaRegs  m   = load_meta(meta + w);
			const uint8_t*  blk = col.packed + (uint64_t)m.packed_off() * 128u;
			const UT*       ev  = static_cast<const UT*>(col.exc_val) + m.exc_off();
			const uint16_t* ep  = col.exc_pos + m.exc_off();
			const bool      rd  = m.scheme() != ALPB200_SCHEME_ALP;
			uint32_t        word = 0;
#pragma unroll 1
			for (int r = 0; r < 32; r++) {  // (here value 32 r + t belongs to lane t, row r — for doubles too)
				const uint32_t b = __ballot_sync(FULL, filter_test((double)T::from_bits(value_bits_slow<PT>(blk, m, (uint32_t)(32 * r + t))), constant, fmask));
				if (t == r) { word = b; }
			}
			s_bits[warp][t] = word;
			__syncwarp();
			for (uint32_t i = t; i < m.exc_cnt(); i += 32) {
				const uint32_t p = ep[i];
				UT             v = ev[i];
				if (rd) { v = (UT)(((v & 0xFFFFu) << m.bw()) | (value_bits_slow<PT>(blk, m, p) & low_mask<UT>((int)m.bw()))); }
				atomicAnd(&s_bits[warp][p >> 5], ~(1u << (p & 31)));
				if (filter_test((double)T::from_bits(v), constant, fmask)) { atomicOr(&s_bits[warp][p >> 5], 1u << (p & 31)); }
			}
			__syncwarp();
			word               = s_bits[warp][t];
			bitmap[w * 32 + t] = word;
			hits += __popc(word);
			__syncwarp();
		}
	} else {
		uint8_t*  stage = smem + (size_t)warp * 2 * stage_bytes;
		uint64_t* bars  = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * 2 * stage_bytes) + 2 * warp;
		if (t == 0) {
			mbar_init(&bars[0], 1);
			mbar_init(&bars[1], 1);
			fence_mbar_init();
		}
		__syncwarp();
		constexpr uint32_t CHUNK = 16, REFILL_AT = 6;
		auto draw = [&]() -> uint64_t {
			unsigned long long b = 0;
			if (t == 0) { b = atomicAdd(counter, (unsigned long long)CHUNK); }
			return shfl_u64(b, 0);
		};
		uint64_t chunk_base = draw(), next_base = 0;
		uint32_t chunk_used = 0;
		auto     take       = [&]() -> uint64_t {
            if (chunk_used == CHUNK) {
                chunk_base = next_base;
                chunk_used = 0;
            }
            const uint64_t idx = chunk_base + chunk_used++;
            if (chunk_used == CHUNK - REFILL_AT) { next_base = draw(); }
            return idx;
		};
		uint64_t v = take(), v_next = take();
		if (v < n_vectors) {
			auto issue = [&](const MetaRegs& m, int s) {
				const uint32_t bytes = m.block_bytes();
				if (t == 0 && bytes != 0) {
					mbar_arrive_expect_tx(&bars[s], bytes);
					bulk_g2s(stage + (size_t)s * stage_bytes, col.packed + (uint64_t)m.packed_off() * 128u, bytes, &bars[s]);
				}
			};
			MetaRegs cur      = load_meta(meta + v);
			bool     has_next = v_next < n_vectors;
			MetaRegs nxt      = cur;
			if (has_next) { nxt = load_meta(meta + v_next); }
			issue(cur, 0);
			XR       xcur  = load_exceptions_k<UT, K>(col, cur, t);
			uint32_t phase = 0;
			for (int s = 0;; s ^= 1) {
				XR xnxt = xcur;
				if (has_next) {
					issue(nxt, s ^ 1);
					xnxt = load_exceptions_k<UT, K>(col, nxt, t);
					if (nxt.exc_cnt() > 32u * K) { prefetch_exception_tail(col, nxt, t, sizeof(UT)); }
				}
				const uint64_t v_nn   = has_next ? take() : v_next;
				const bool     has_nn = has_next && v_nn < n_vectors;
				MetaRegs       nn     = nxt;
				if (has_nn) { nn = load_meta(meta + v_nn); }
				const uint8_t* stg = stage + (size_t)s * stage_bytes;
				if (cur.block_bytes() != 0) {
					mbar_wait(&bars[s], (phase >> s) & 1u);
					phase ^= 1u << s;
				}
				const uint32_t  n_exc = cur.exc_cnt();
				const UT*       ev    = static_cast<const UT*>(col.exc_val) + cur.exc_off();
				const uint16_t* ep    = col.exc_pos + cur.exc_off();
				const bool      rd    = cur.scheme() != ALPB200_SCHEME_ALP;
				const uint32_t  rbw   = cur.bw();
				uint32_t        word;
				if (rd) {
					word = filter_rd_vector<PT>(stg, cur, t, constant, fmask);
				} else {
					// every slot decoded with the reference's recipe in registers and compared: exact by construction
					const ST bw_ok = (ST)0;
					(void)bw_ok;
					const auto fact = T::fact10(cur.f());
					const auto frac = T::frac10(cur.e());
					BitWords<PT> w;
					if constexpr (sizeof(PT) == 8) {
						const uint64_t base = cur.base();
						if (cur.bw() <= 32) {
							dispatch_width<0, 32>(cur.bw(), [&](auto Wc) {
								constexpr int BW = decltype(Wc)::value;
								unpack64_rows<BW>(stg, t & 15, t >> 4, [&](int r, uint32_t lo, uint32_t) {
									w.row(r, filter_test(decode_value<double>((int64_t)((uint64_t)lo + base), fact, frac), constant, fmask), t);
								});
							});
						} else {  // wide fields (rare): run-time extraction
#pragma unroll 1
							for (int r = 0; r < 32; r++) {
								w.row(r, filter_test(alp_value_at(stg, cur, (uint32_t)Map<PT>::index(t, r), PT()), constant, fmask), t);
							}
						}
					} else {
						const uint32_t base = cur.a.x;
						dispatch_width<0, 32>(cur.bw(), [&](auto Wc) {
							constexpr int BW = decltype(Wc)::value;
							unpack32_rows<BW>(stg, t, [&](int r, uint32_t d) {
								w.row(r, filter_test((double)decode_value<float>((int32_t)(d + base), fact, frac), constant, fmask), t);
							});
						});
					}
					word = w.words(t);
				}
				if (n_exc != 0) {  // the exceptions' bits, from their true values
					s_bits[warp][t] = word;
					__syncwarp();
					auto one = [&](uint32_t p, UT val) {
						const UT bits = rd ? (UT)(((val & 0xFFFFu) << rbw) | rd_right_at(stg, rbw, p, UT())) : val;
						atomicAnd(&s_bits[warp][p >> 5], ~(1u << (p & 31)));
						if (filter_test((double)T::from_bits(bits), constant, fmask)) { atomicOr(&s_bits[warp][p >> 5], 1u << (p & 31)); }
					};
#pragma unroll
					for (int k = 0; k < K; k++) {
						if ((uint32_t)t + 32u * k < n_exc) { one(xcur.pos(k), xcur.val[k]); }
					}
					for (uint32_t i = (uint32_t)t + 32u * K; i < n_exc; i += 32) {
						one(ep[i], ev[i]);
					}
					__syncwarp();
					word = s_bits[warp][t];
				}
				bitmap[v * 32 + t] = word;
				hits += __popc(word);
				__syncwarp();  // stage s and s_bits are free again
				if (!has_next) { break; }
				cur      = nxt;
				nxt      = nn;
				xcur     = xnxt;
				has_next = has_nn;
				v        = v_next;
				v_next   = v_nn;
			}
		}
	}
	hits = __reduce_add_sync(FULL, hits);
	if (t == 0 && hits != 0 && selected != nullptr) { atomicAdd(selected, (unsigned long long)hits); }
}

}  // namespace alpb200
