// alp_encode.cuh — fused ALP / ALP_RD vector encoder, one warp per 1024-value vector.
//
// Replaces, per vector, alp::encoder<PT>::encode (include/alp/encoder.hpp:402-418: second-level sampling :241-305 +
// encode_simdized :307-400) + analyze_ffor (:109-120) + ffor::ffor (src/fastlanes_generated_ffor.cpp:29939), or for
// ALP_RD row-groups alp::rd_encoder<PT>::encode (include/alp/rd.hpp:109-147) + two ffor::ffor calls
// (test/test_alp_sample.cpp:141-145,164-166).
//
// Register-resident: a thread keeps its 32 encoded integers in registers from the input load to the bit-packer.
// 64-bit lanes: thread (lane = t&15, half = t>>4) owns rows 32*half .. 32*half+31 of FastLanes lane `lane`, i.e. values
// 16*(32*half + r) + lane — exactly half of that lane's bit stream, which is bw whole 32-bit words.
// 32-bit lanes: thread t owns lane t, values 32*r + t.
// The packed block is assembled in shared memory in its final (verbatim) layout and leaves with one bulk-async
// store (TMA 1-D).  Output offsets come from a single-pass decoupled look-back over thread blocks, so the column is
// written contiguously, in vector order, in the same pass that reads the input.
#pragma once

#include "alp_device.cuh"

namespace alpb200 {

// thread -> value mapping of the two lane widths
template <typename PT>
struct Map;
template <>
struct Map<double> {
	__device__ static __forceinline__ int index(int t, int r) { return 512 * (t >> 4) + 16 * r + (t & 15); }
};
template <>
struct Map<float> {
	__device__ static __forceinline__ int index(int t, int r) { return 32 * r + t; }
};

// the head of alpb200_rg_state (44 bytes), loaded once per warp
struct StateRegs {
	uint32_t scheme, k, c0, c1, c2, ds;
	uint4    dict;
	uint32_t n_extra;
	__device__ __forceinline__ int exp_of(int i) const { return (int)((i < 2 ? c0 : (i < 4 ? c1 : c2)) >> (16 * (i & 1)) & 0xFF); }
	__device__ __forceinline__ int fac_of(int i) const { return (int)((i < 2 ? c0 : (i < 4 ? c1 : c2)) >> (16 * (i & 1) + 8) & 0xFF); }
	__device__ __forceinline__ uint32_t right_bw() const { return (c2 >> 16) & 0xFF; }
	__device__ __forceinline__ uint32_t left_bw() const { return c2 >> 24; }
	__device__ __forceinline__ uint32_t dict_size() const { return ds & 0xFF; }
};
__device__ __forceinline__ StateRegs load_state(const alpb200_rg_state* s) {
	const uint32_t* p = reinterpret_cast<const uint32_t*>(s);
	StateRegs       r;
	r.scheme  = __ldg(p + 0);
	r.k       = __ldg(p + 1);
	r.c0      = __ldg(p + 2);
	r.c1      = __ldg(p + 3);
	r.c2      = __ldg(p + 4);
	r.ds      = __ldg(p + 5);
	r.dict    = make_uint4(__ldg(p + 6), __ldg(p + 7), __ldg(p + 8), __ldg(p + 9));
	r.n_extra = __ldg(p + 10) & 0xFFFFu;
	return r;
}

// what a warp knows about its vector after the analysis phase
template <typename PT>
struct Analysis {
	typename Traits<PT>::UT payload[32];  // ALP: encoded integers (exceptions already overwritten by the fill value);
	                                      // ALP_RD: right parts
	uint32_t left_nib[4];                 // ALP_RD: dictionary index of row r in nibble r (already masked to left_bw)
	uint32_t rowmask;                     // lane r: ballot of "is exception" over row r
	uint32_t cnt;                         // exceptions in the vector
	uint32_t bw;                          // ALP: FFOR width; ALP_RD: right width
	uint32_t e, f;                        // ALP: exponent/factor;  ALP_RD: left width / dictionary size
	typename Traits<PT>::ST base;         // ALP FOR base (0 for ALP_RD)
};

// ---- second-level sampling: encoder.hpp:241-305 --------------------------------------------------------------------
template <typename PT>
__device__ __forceinline__ void choose_exponent_factor(const PT* __restrict__ in_vec, const StateRegs& st, int t, int& e_out,
                                                       int& f_out) {
	using T  = Traits<PT>;
	using ST = typename T::ST;
	if (st.k <= 1) {  // encoder.hpp:409-412
		e_out = st.exp_of(0);
		f_out = st.fac_of(0);
		return;
	}
	const PT xs      = in_vec[32 * t];  // samples 0, 32, ..., 992 (encoder.hpp:253-254,266)
	int      best_e  = 0, best_f = 0, worse = 0;
	uint32_t best_sz = 0;
	for (int k = 0; k < (int)st.k; k++) {
		const int e = st.exp_of(k), f = st.fac_of(k);
		const ST  enc = encode_value<PT, true>(xs, T::exp10(e), T::frac10(f));
		const PT  dec = decode_value<PT>(enc, T::fact10(f), T::frac10(e));
		const bool ok = dec == xs;
		const uint32_t n_exc = 32 - __popc(__ballot_sync(FULL, ok));
		const ST       mx    = warp_max<ST>(ok ? enc : T::ST_MIN);
		const ST       mn    = warp_min<ST>(ok ? enc : T::ST_MAX);
		const uint32_t sz    = 32u * bits_of_range<PT>(mx, mn) + n_exc * (T::EXC_BITS + 16);
		if (k == 0) {
			best_sz = sz;
			best_e  = e;
			best_f  = f;
			continue;
		}
		if (sz >= best_sz) {
			if (++worse == 2) { break; }  // SAMPLING_EARLY_EXIT_THRESHOLD, constants.hpp:16
			continue;
		}
		best_sz = sz;
		best_e  = e;
		best_f  = f;
		worse   = 0;
	}
	e_out = best_e;
	f_out = best_f;
}

// ---- ALP analysis: encoder.hpp:307-400 + :109-120 ------------------------------------------------------------------
template <typename PT>
__device__ __forceinline__ void analyze_alp(const PT* __restrict__ in_vec, const StateRegs& st, int t, Analysis<PT>& a) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	using ST = typename T::ST;
	PT x[32];
#pragma unroll
	for (int r = 0; r < 32; r++) {
		x[r] = in_vec[Map<PT>::index(t, r)];
	}
	int e, f;
	choose_exponent_factor<PT>(in_vec, st, t, e, f);
	const PT ex = T::exp10(e), frf = T::frac10(f), fre = T::frac10(e);
	const ST fa = T::fact10(f);

	uint32_t myexc = 0, rowmask = 0;
#pragma unroll
	for (int r = 0; r < 32; r++) {
		const PT   v   = T::is_special(T::bits(x[r])) ? T::upper_limit() : x[r];  // encoder.hpp:326-338
		const ST   enc = encode_value<PT, false>(v, ex, frf);                     // :345
		const PT   dec = decode_value<PT>(enc, fa, fre);                          // :347
		const bool exc = dec != v;                                                // :374-379
		a.payload[r]   = (UT)enc;
		myexc |= (uint32_t)exc << r;
		const uint32_t m = __ballot_sync(FULL, exc);
		if (t == r) { rowmask = m; }
	}
	// fill value = encoded integer at the first non-exception position, 0 if there is none (encoder.hpp:382-388)
	uint32_t cand = 0xFFFFu;
	if (sizeof(PT) == 8) {
		const uint32_t lo = ~rowmask & 0xFFFFu, hi = (~rowmask) >> 16;
		if (lo) { cand = 16u * t + (__ffs(lo) - 1); }
		if (hi) { cand = min(cand, 16u * (32 + t) + (__ffs(hi) - 1)); }
	} else {
		if (~rowmask) { cand = 32u * t + (__ffs(~rowmask) - 1); }
	}
	cand    = __reduce_min_sync(FULL, cand);
	ST fill = 0;
	if (cand != 0xFFFFu) { fill = encode_value<PT, false>(in_vec[cand], ex, frf); }
	ST mn = T::ST_MAX, mx = T::ST_MIN;
#pragma unroll
	for (int r = 0; r < 32; r++) {
		ST v = (ST)a.payload[r];
		if ((myexc >> r) & 1u) { v = fill; }  // encoder.hpp:393
		a.payload[r] = (UT)v;
		mn           = v < mn ? v : mn;
		mx           = v > mx ? v : mx;
	}
	mn        = warp_min<ST>(mn);  // analyze_ffor, encoder.hpp:109-120
	mx        = warp_max<ST>(mx);
	a.rowmask = rowmask;
	a.cnt     = __reduce_add_sync(FULL, (uint32_t)__popc(rowmask));
	a.bw      = (uint32_t)bits_of_range<PT>(mx, mn);
	a.base    = mn;
	a.e       = (uint32_t)e;
	a.f       = (uint32_t)f;
}

// ---- ALP_RD analysis: rd.hpp:109-147 --------------------------------------------------------------------------------
// on_index(r, idx) reports the unmasked dictionary index of row r (what rd.hpp:136 stores before FFOR masks it).
template <typename PT, typename OnIndex>
__device__ __forceinline__ void analyze_rd(const PT* __restrict__ in_vec, const alpb200_rg_state* state, const StateRegs& st,
                                           int t, Analysis<PT>& a, OnIndex&& on_index) {
	using T              = Traits<PT>;
	using UT             = typename T::UT;
	const uint32_t rbw   = st.right_bw(), lbw = st.left_bw(), ds = st.dict_size();
	const UT       rmask = low_mask<UT>(rbw);
	const uint32_t lmask = (1u << lbw) - 1;
	uint32_t       rowmask = 0;
	a.left_nib[0] = a.left_nib[1] = a.left_nib[2] = a.left_nib[3] = 0;
#pragma unroll
	for (int r = 0; r < 32; r++) {
		const UT       bits = T::bits(in_vec[Map<PT>::index(t, r)]);
		const uint32_t left = (uint32_t)(bits >> rbw);
		a.payload[r]        = bits & rmask;
		uint32_t idx = ds;  // rd.hpp:129-131: a left part nobody has seen gets the smallest non-dictionary index
		bool     hit = false;
#pragma unroll
		for (uint32_t d = 0; d < ALPB200_RD_DICT_SIZE; d++) {
			if (d < ds && !hit && dict_lookup(st.dict, d) == left) {
				idx = d;
				hit = true;
			}
		}
		if (st.n_extra != 0 && __any_sync(FULL, !hit)) {  // rd.hpp:73-77,133: sampled left parts outside the dictionary
			if (!hit) {
				for (uint32_t x = 0; x < st.n_extra; x++) {
					if (state->extra_key[x] == left) {
						idx = state->extra_idx[x];
						break;
					}
				}
			}
		}
		const bool     exc = idx >= ds;  // rd.hpp:138-142
		const uint32_t m   = __ballot_sync(FULL, exc);
		if (t == r) { rowmask = m; }
		on_index(r, idx);
		a.left_nib[r >> 3] |= (idx & lmask) << (4 * (r & 7));  // FFOR masks the stored index to left_bw bits
	}
	a.rowmask = rowmask;
	a.cnt     = __reduce_add_sync(FULL, (uint32_t)__popc(rowmask));
	a.bw      = rbw;
	a.e       = lbw;
	a.f       = ds;
	a.base    = 0;
}

// ---- FFOR bit packer (write side of SURVEY.md appendix A.1; src/fastlanes_generated_ffor.cpp:7788-7999) ----------
// A thread appends fields of at most 32 bits to its private stream and emits whole 32-bit words.  Every thread of
// the warp has the same (bw-determined) sequence of `nb`, so all branches are warp-uniform.
struct BitSink {
	uint64_t acc;
	uint32_t nb;
	__device__ __forceinline__ BitSink() : acc(0), nb(0) {}
	template <typename Emit>
	__device__ __forceinline__ void push(uint32_t x, uint32_t n, Emit&& emit) {
		acc |= (uint64_t)x << nb;
		nb += n;
		if (nb >= 32) {
			emit((uint32_t)acc);
			acc >>= 32;
			nb -= 32;
		}
	}
};

// pack a thread's 32 rows of 64-bit-lane values ((payload - base) & mask) into the verbatim block image `blk`
__device__ __forceinline__ void pack_rows(const uint64_t (&payload)[32], uint64_t base, uint32_t bw, int t, uint8_t* blk) {
	if (bw == 0) { return; }  // ffor bw=0 writes nothing (src/fastlanes_generated_ffor.cpp:4)
	const int      lane = t & 15, half = t >> 4;
	const uint64_t mask = low_mask<uint64_t>(bw);
	uint32_t*      w32  = reinterpret_cast<uint32_t*>(blk);
	uint32_t       j    = half * bw;  // 32-bit word index inside the lane's stream
	auto           emit = [&](uint32_t w) {
        w32[32 * (j >> 1) + 2 * lane + (j & 1)] = w;  // 64-bit word (j>>1) of lane `lane` lives at element 16*(j>>1)+lane
        j++;
	};
	BitSink        sink;
	const uint32_t n_lo = bw < 32 ? bw : 32, n_hi = bw - n_lo;
#pragma unroll
	for (int r = 0; r < 32; r++) {
		const uint64_t d = (payload[r] - base) & mask;
		sink.push((uint32_t)d, n_lo, emit);
		if (n_hi) { sink.push((uint32_t)(d >> 32), n_hi, emit); }
	}
}
__device__ __forceinline__ void pack_rows(const uint32_t (&payload)[32], uint32_t base, uint32_t bw, int t, uint8_t* blk) {
	if (bw == 0) { return; }
	const uint32_t mask = low_mask<uint32_t>(bw);
	uint32_t*      w32  = reinterpret_cast<uint32_t*>(blk);
	uint32_t       j    = 0;
	auto           emit = [&](uint32_t w) {
        w32[32 * j + t] = w;
        j++;
	};
	BitSink sink;
#pragma unroll
	for (int r = 0; r < 32; r++) {
		sink.push((payload[r] - base) & mask, bw, emit);
	}
}

__device__ __forceinline__ uint32_t nib(const uint32_t (&n)[4], int r) { return (n[r >> 3] >> (4 * (r & 7))) & 0xFu; }

// pack the ALP_RD dictionary indices on 16-bit lanes (64 lanes x 16 rows): value v = 64*row16 + lane16
__device__ __forceinline__ void pack_left(const uint32_t (&left_nib)[4], uint32_t lbw, int t, uint16_t* blk, double /*tag*/) {
	// thread (lane, half) holds values 16*(32*half + r) + lane: lane16 = lane + 16*(r&3), row16 = 8*half + (r>>2)
	const int lane = t & 15, half = t >> 4;
#pragma unroll
	for (int q = 0; q < 4; q++) {
		uint32_t chunk = 0;  // rows 8*half .. 8*half+7 of lane16 = lane + 16q: 8*lbw <= 24 bits
#pragma unroll
		for (int i = 0; i < 8; i++) {
			chunk |= nib(left_nib, 4 * i + q) << (i * lbw);
		}
		const uint32_t other  = __shfl_xor_sync(FULL, chunk, 16);
		const uint64_t stream = half ? ((uint64_t)other | ((uint64_t)chunk << (8 * lbw))) : ((uint64_t)chunk | ((uint64_t)other << (8 * lbw)));
		if (half == 0) {
			for (uint32_t w = 0; w < lbw; w++) {
				blk[64 * w + lane + 16 * q] = (uint16_t)(stream >> (16 * w));
			}
		}
	}
}
__device__ __forceinline__ void pack_left(const uint32_t (&left_nib)[4], uint32_t lbw, int t, uint16_t* blk, float /*tag*/) {
	// thread t holds values 32*r + t: lane16 = t + 32*(r&1), row16 = r>>1 — two complete 16-row streams
#pragma unroll
	for (int q = 0; q < 2; q++) {
		uint64_t stream = 0;
#pragma unroll
		for (int i = 0; i < 16; i++) {
			stream |= (uint64_t)nib(left_nib, 2 * i + q) << (i * lbw);
		}
		for (uint32_t w = 0; w < lbw; w++) {
			blk[64 * w + t + 32 * q] = (uint16_t)(stream >> (16 * w));
		}
	}
}

// ---- exception emission in position order (encoder.hpp:390-397 / rd.hpp:138-142) ------------------------------------
// value_of(p) returns what is stored for position p (the original value for ALP, the left part for ALP_RD).
template <typename PT, typename ValueOf, typename Store>
__device__ __forceinline__ void emit_exceptions(uint32_t rowmask, int t, ValueOf&& value_of, Store&& store) {
	uint32_t rows = __ballot_sync(FULL, rowmask != 0);
	if (rows == 0) { return; }
	if (sizeof(PT) == 8) {
		const int lane = t & 15, half = t >> 4;
		uint32_t  tot_lo, tot_hi;
		const uint32_t p_lo = warp_excl_scan(__popc(rowmask & 0xFFFFu), t, tot_lo);
		const uint32_t p_hi = warp_excl_scan(__popc(rowmask >> 16), t, tot_hi) + tot_lo;
		while (rows) {
			const int r = __ffs(rows) - 1;
			rows &= rows - 1;
			const uint32_t m  = __shfl_sync(FULL, rowmask, r);
			const uint32_t pl = __shfl_sync(FULL, p_lo, r), ph = __shfl_sync(FULL, p_hi, r);
			if ((m >> t) & 1u) {
				const uint32_t hm   = half ? (m >> 16) : (m & 0xFFFFu);
				const uint32_t rank = (half ? ph : pl) + __popc(hm & ((1u << lane) - 1));
				const uint32_t p    = 16u * (32 * half + r) + lane;
				store(rank, p, value_of(p));
			}
		}
	} else {
		uint32_t       tot;
		const uint32_t pre = warp_excl_scan(__popc(rowmask), t, tot);
		while (rows) {
			const int r = __ffs(rows) - 1;
			rows &= rows - 1;
			const uint32_t m  = __shfl_sync(FULL, rowmask, r);
			const uint32_t pr = __shfl_sync(FULL, pre, r);
			if ((m >> t) & 1u) {
				const uint32_t rank = pr + __popc(m & ((1u << t) - 1));
				const uint32_t p    = 32u * r + t;
				store(rank, p, value_of(p));
			}
		}
	}
}

// ---- decoupled look-back over thread blocks ---------------------------------------------------------------------------
// status word: [63:62] flag (0 empty, 1 block aggregate, 2 inclusive prefix) | [61:32] packed size in 128-byte units |
// [31:0] exception slots
constexpr uint64_t LB_AGG = 1ull << 62, LB_PRE = 2ull << 62, LB_VAL = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t* p) {
	uint64_t v;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t* p, uint64_t v) {
	asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
	for (int m = 16; m > 0; m >>= 1) {
		v += (uint64_t)shfl_xor_i64((int64_t)v, m);
	}
	return v;
}
// called by one full warp of block `bid`; returns the exclusive prefix (packed units << 32 | exceptions)
__device__ __forceinline__ uint64_t lookback(uint64_t* status, uint32_t bid, uint64_t aggregate, int t) {
	if (bid == 0) {
		if (t == 0) { st_volatile_u64(&status[0], LB_PRE | aggregate); }
		return 0;
	}
	if (t == 0) { st_volatile_u64(&status[bid], LB_AGG | aggregate); }
	uint64_t excl = 0;
	int64_t  look = (int64_t)bid - 1;
	for (;;) {
		const int64_t idx = look - t;
		uint64_t      val = LB_PRE;  // before block 0: an inclusive prefix of zero
		if (idx >= 0) { val = ld_volatile_u64(&status[idx]); }
		if (__any_sync(FULL, (val >> 62) == 0)) { continue; }  // someone has not published yet: look again
		const uint32_t pre = __ballot_sync(FULL, (val >> 62) == 2);
		if (pre) {
			const int first = __ffs(pre) - 1;  // nearest predecessor with an inclusive prefix
			excl += warp_sum_u64(t <= first ? (val & LB_VAL) : 0);
			break;
		}
		excl += warp_sum_u64(val & LB_VAL);
		look -= 32;
	}
	if (t == 0) { st_volatile_u64(&status[bid], LB_PRE | (excl + aggregate)); }
	return excl;
}

struct ColOut {
	alpb200_vec_meta* meta;
	uint8_t*          packed;
	uint64_t          packed_capacity;
	void*             exc_val;
	uint16_t*         exc_pos;
	uint64_t          exc_capacity;
	uint64_t*         totals;
};

// workspace layout: [0] ticket counter, [1] reserved, [2..] one status word per thread block
template <typename PT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) encode_kernel(const PT* __restrict__ in, uint64_t n_vectors,
                                                            const alpb200_rg_state* __restrict__ states, ColOut col,
                                                            uint64_t* workspace) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	constexpr uint32_t STAGE = (sizeof(PT) == 8 ? 66u : 35u) * 128u;  // widest block: 63+3 bits (f64 RD) / 32+3 bits
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ uint32_t s_bid;
	__shared__ uint32_t s_units[WARPS], s_cnt[WARPS];
	__shared__ uint64_t s_excl;

	const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
	if (threadIdx.x == 0) { s_bid = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long*>(workspace), 1ull); }
	__syncthreads();
	const uint32_t bid    = s_bid;
	const uint64_t v      = (uint64_t)bid * WARPS + warp;
	const bool     active = v < n_vectors;
	uint8_t*       blk    = smem + (size_t)warp * STAGE;

	Analysis<PT> a;
	a.cnt = 0;
	a.bw = a.e = a.f = 0;
	a.rowmask        = 0;
	a.base           = 0;
	StateRegs st;
	st.scheme                      = ALPB200_SCHEME_INVALID;
	const PT*               in_vec = in + v * (uint64_t)VEC;
	const alpb200_rg_state* state  = states + (active ? v / ALPB200_ROWGROUP_VECTORS : 0);
	uint32_t                units  = 0;
	if (active) {
		st = load_state(state);
		if (st.scheme == ALPB200_SCHEME_ALP_RD) {
			analyze_rd<PT>(in_vec, state, st, t, a, [](int, uint32_t) {});
			units = a.bw + a.e;
		} else {
			analyze_alp<PT>(in_vec, st, t, a);
			units = a.bw;
		}
	}
	if (t == 0) {
		s_units[warp] = units;
		s_cnt[warp]   = a.cnt;
	}
	__syncthreads();
	if (warp == 0) {
		uint64_t mine = 0;
		if (t < WARPS) { mine = ((uint64_t)s_units[t] << 32) | s_cnt[t]; }
		const uint64_t agg  = warp_sum_u64(mine);
		const uint32_t wide = __reduce_max_sync(FULL, (uint32_t)(mine >> 32));
		const uint64_t excl = lookback(workspace + 2, bid, agg, t);
		if (t == 0) {
			s_excl = excl;
			atomicMax(reinterpret_cast<unsigned long long*>(&col.totals[3]), (unsigned long long)wide * 128ull);
			if ((uint64_t)(bid + 1) * WARPS >= n_vectors) {  // last block: publish the column totals
				const uint64_t incl = excl + agg;
				col.totals[0]       = (incl >> 32) * 128ull;
				col.totals[1]       = incl & 0xFFFFFFFFull;
			}
		}
	}
	__syncthreads();
	if (!active) { return; }
	uint64_t units_off = s_excl >> 32, exc_off = s_excl & 0xFFFFFFFFull;
	for (int w = 0; w < warp; w++) {
		units_off += s_units[w];
		exc_off += s_cnt[w];
	}
	const uint32_t bytes = units * 128u;
	if (units_off * 128ull + bytes > col.packed_capacity || exc_off + a.cnt > col.exc_capacity) {
		if (t == 0) { atomicExch(reinterpret_cast<unsigned long long*>(&col.totals[2]), 1ull); }
		return;
	}

	// ---- pack into the shared-memory image of the block, then one bulk store ----
	const bool rd = st.scheme == ALPB200_SCHEME_ALP_RD;
	pack_rows(a.payload, (UT)a.base, a.bw, t, blk);
	if (rd) { pack_left(a.left_nib, a.e, t, reinterpret_cast<uint16_t*>(blk + 128u * a.bw), PT()); }
	if (bytes) {
		fence_proxy_async_smem();
		__syncwarp();
		if (t == 0) {
			bulk_s2g(col.packed + units_off * 128ull, blk, bytes);
			bulk_commit();
		}
	}
	// ---- exceptions, in position order ----
	UT*       ev  = static_cast<UT*>(col.exc_val) + exc_off;
	uint16_t* ep  = col.exc_pos + exc_off;
	const uint32_t rbw = a.bw;
	emit_exceptions<PT>(
	    a.rowmask, t,
	    [&](uint32_t p) -> UT {
		    const UT bits = T::bits(in_vec[p]);
		    return rd ? (UT)(bits >> rbw) : bits;
	    },
	    [&](uint32_t rank, uint32_t p, UT val) {
		    ev[rank] = val;
		    ep[rank] = (uint16_t)p;
	    });
	// ---- the 32-byte record ----
	if (t == 0) {
		uint4 ra, rb;
		if (rd) {
			ra = st.dict;
		} else {
			const int64_t b = (int64_t)a.base;
			ra              = make_uint4((uint32_t)b, (uint32_t)((uint64_t)b >> 32), 0u, 0u);
		}
		rb.x = (uint32_t)units_off;
		rb.y = (uint32_t)exc_off;
		rb.z = a.cnt | (st.scheme << 16) | (a.bw << 24);
		rb.w = a.e | (a.f << 8);
		uint4* dst = reinterpret_cast<uint4*>(col.meta + v);
		dst[0]     = ra;
		dst[1]     = rb;
		if (bytes) { bulk_wait_read_all(); }  // the stage must outlive the bulk store's read of it
	}
}

}  // namespace alpb200
