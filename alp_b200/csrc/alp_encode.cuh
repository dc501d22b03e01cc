// alp_encode.cuh — fused ALP / ALP_RD vector encoder, one warp per 1024-value vector.
//
// Replaces, per vector, alp::encoder<PT>::encode (include/alp/encoder.hpp:402-418: second-level sampling :241-305 +
// encode_simdized :307-400) + analyze_ffor (:109-120) + ffor::ffor (src/fastlanes_generated_ffor.cpp:29939), or for
// ALP_RD row-groups alp::rd_encoder<PT>::encode (include/alp/rd.hpp:109-147) + two ffor::ffor calls
// (test/test_alp_sample.cpp:141-145,164-166).
//
// Thread -> value mapping.  64-bit lanes: thread (lane = t&15, half = t>>4) owns rows 32*half .. 32*half+31 of FastLanes
// lane `lane`, i.e. values 16*(32*half + r) + lane — exactly half of that lane's bit stream, which is bw whole 32-bit
// words.  32-bit lanes: thread t owns lane t, values 32*r + t.  Every warp load instruction reads full 128-byte lines.
//
// Phases of a warp:
//   1 analysis   stream the vector once (double-buffered batches of 8 rows), encode + decode + compare per value,
//                track min/max of the non-exceptions and a per-thread exception bitmap; the encoded integers are
//                parked in a per-warp shared-memory tile ([row][thread], conflict-free) instead of 64 registers
//   2 placement  the block's packed size and exception count join a single-pass decoupled look-back over thread
//                blocks (tickets are handed out in launch order), so the column comes out contiguous and in vector order
//   3 packing    FFOR straight from the tile into the final interleaved layout: each thread assembles whole 64-bit
//                words of its lane's stream and the warp stores full 128-byte lines; then exceptions in position order
//                and the 32-byte record
#pragma once

#include "alp_device.cuh"
#include "alp_ffor.cuh"

namespace alpb200 {

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

// What a warp knows about its vector after the analysis phase.  The payload (ALP: encoded integers, ALP_RD: right
// parts) sits in the warp's shared-memory tile: element [r*32 + t] belongs to thread t, row r.
template <typename PT>
struct Analysis {
	uint32_t left_nib[4];          // ALP_RD: dictionary index of row r in nibble r (already masked to left_bw)
	uint32_t myexc;                // bit r: this thread's value in row r is an exception
	uint32_t cnt;                  // exceptions in the vector
	uint32_t bw;                   // ALP: FFOR width; ALP_RD: right width
	uint32_t e, f;                 // ALP: exponent/factor;  ALP_RD: left width / dictionary size
	typename Traits<PT>::ST base;  // ALP FOR base (0 for ALP_RD)
	typename Traits<PT>::ST fill;  // ALP: value stored at exception positions (encoder.hpp:382-393)
};

// 32x32 bit-matrix transpose across the warp: in = bit r of lane t, out = bit t of lane r.
// Turns the per-thread exception bitmaps into per-row ballots (what the position-ordered emission needs).
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int t) {
	uint32_t m = 0x0000FFFFu;
#pragma unroll
	for (int j = 16; j > 0; j >>= 1) {
		const uint32_t y = __shfl_xor_sync(FULL, x, j);
		x                = (t & j) ? ((x & ~m) | ((y & ~m) >> j)) : ((x & m) | ((y & m) << j));
		m ^= m << (j >> 1);
	}
	return x;
}

// ---- second-level sampling: encoder.hpp:241-305 --------------------------------------------------------------------
// xs = this lane's sample, value 32*lane of the vector (samples 0, 32, ..., 992; encoder.hpp:253-254,266)
template <typename PT>
__device__ __forceinline__ void choose_exponent_factor(PT xs, const StateRegs& st, int& e_out, int& f_out) {
	using T  = Traits<PT>;
	using ST = typename T::ST;
	int      best_e  = st.exp_of(0), best_f = st.fac_of(0), worse = 0;
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

// ---- value access policy ----------------------------------------------------------------------------------------------
// The analysis and packing code reads row r of the thread's 32 rows through an IO policy with a compile-time row
// number.  TileIO works on a shared-memory copy of the vector (value order) and stores the encoded integers back in
// place — the batched kernel's tile (filled by a bulk-async copy) and the single-vector primitives' tile alike.
// KEEP_EXC (the streaming encoder, alp_encode_stream.cuh): exception slots keep the ORIGINAL bits (ALP_RD: every slot does),
// so the exceptions' values can be gathered from the tile after the analysis instead of from global memory.
template <typename PT, bool KEEP = false>
struct TileIO {
	using UT = typename Traits<PT>::UT;
	static constexpr bool KEEP_EXC = KEEP;
	UT* tile;
	int t;
	__device__ __forceinline__ TileIO(UT* tile_, int lane_id) : tile(tile_), t(lane_id) {}
	template <int R>
	__device__ __forceinline__ UT load(std::integral_constant<int, R>) {
		return tile[Map<PT>::index(t, R)];
	}
	template <int R>
	__device__ __forceinline__ void store(std::integral_constant<int, R>, UT v) {
		tile[Map<PT>::index(t, R)] = v;
	}
	__device__ __forceinline__ UT   load_row(int r) { return tile[Map<PT>::index(t, r)]; }
	__device__ __forceinline__ void store_row(int r, UT v) { tile[Map<PT>::index(t, r)] = v; }
	// value i of the vector (second-level sampling; before the analysis overwrites the tile): no trip to global memory
	__device__ __forceinline__ PT sample(const PT*, int i) const { return Traits<PT>::from_bits(tile[i]); }
	// start over: the tile was overwritten with encoded integers; every thread restores ITS OWN slots from global memory
	__device__ __forceinline__ void rewind(const PT* in_vec) {
#pragma unroll 8
		for (int r = 0; r < 32; r++) {
			tile[Map<PT>::index(t, r)] = Traits<PT>::bits(in_vec[Map<PT>::index(t, r)]);
		}
	}
	// the encoded integer stored for value `position`
	__device__ __forceinline__ UT kept(int position) const { return tile[position]; }
};

// ---- ALP analysis: encoder.hpp:307-400 + :109-120 ------------------------------------------------------------------
// Exception test.  The reference pre-replaces special values (double: -0.0; float: NaN, ±Inf, -0.0; encoder.hpp:326-338)
// by ENCODING_UPPER_LIMIT, which never round-trips, and then flags `decoded != value` (:374-379).  Comparing the BIT
// PATTERNS of decoded and original value gives the same set in one integer compare: the decoded value is always
// finite and never -0.0 (an integer times a positive power of ten), so NaN, ±Inf and -0.0 differ from it in bits, and
// for every other value equal numbers have equal bits.
//
// min/max of the non-exceptions (analyze_ffor, encoder.hpp:109-120; exception slots hold `fill`, itself a
// non-exception).  64-bit integers have no native min/max (2 ISETP + 2 SEL each), so the common case — all high
// words equal, e.g. every encoded integer in [0, 2^32) — tracks unsigned min/max of the low words plus AND/OR of the
// high words, and falls back to a second pass over the input otherwise.
//
// Two implementations of the per-value step, bit-identical in outcome:
//
//  EXACT  the reference's recipe literally: enc = x86-cast(t_r), dec = (PT)(enc * FACT[f]) * FRAC[e]  (t_r = the
//         magic-rounded scaled value).  For doubles that is a range test + two selects around F2I, a 64-bit integer
//         multiply (5 instructions) and an I2F.F64.S64.
//  FAST   stays in floating point: Pd = t_r * 10^f, dec = Pd * FRAC[e].  Why it is the same number: t_r is
//         integer-valued, so while |t_r| < 2^63 the cast is exact (enc == t_r) and the integer product P = enc * 10^f is
//         the real product of two exactly representable numbers (10^f <= 10^18 is exact in double, 10^9 in float); as
//         long as |P| < 2^63 it does not wrap, the reference's integer-to-float conversion returns RN(P), and so does
//         the floating-point multiply: Pd == (PT)(enc * FACT[f]).
//         When the product does NOT fit (|P| >= 2^63 | 2^31, or t_r itself is out of range / NaN and the cast returns
//         ST_MIN) the reference's value is always an exception: x * 10^e lies within 10^f / 2 <= |P| / 2 of P, i.e. it
//         has P's sign and at least half its magnitude, whereas the wrapped product has the opposite sign for
//         2^63 <= |P| < 2^64 and at most half the magnitude beyond — the decoded value cannot equal x.  So FAST forces
//         an exception whenever |Pd| > 2^63 (2^31), judged on the bit pattern of Pd (`key`: |Pd| as an unsigned
//         integer, high word for doubles; NaN and Inf compare above).  Only |Pd| == 2^63 leaves the question open
//         (P within rounding distance of the boundary; -2^63 itself is representable): the warp then abandons FAST and
//         redoes the vector with EXACT.  For doubles the test looks at the high word only, which widens that case to
//         2^63 <= |Pd| < 2^63 * (1 + 2^-20) — still a one-in-a-million sliver.
//         (float, f = 10: FACT[10] is the reference's out-of-bounds 0, so P = 0 and nothing wraps; 10^f is taken as 0.)
// (FastLimits<PT>: alp_device.cuh)

#ifndef ALPB200_ANALYZE_UNROLL
#define ALPB200_ANALYZE_UNROLL 8  // rows per iteration of the analysis loop (32 = fully unrolled)
#endif
// what the 32 rows of a thread accumulate
template <typename PT>
struct RowAcc {
	using ST = typename Traits<PT>::ST;
	uint32_t myexc  = 0;
	uint32_t lo_min = 0xFFFFFFFFu, lo_max = 0, hi_and = 0xFFFFFFFFu, hi_or = 0;  // f64
	ST       mn = Traits<PT>::ST_MAX, mx = Traits<PT>::ST_MIN;                   // f32
};

// One pass over the thread's 32 rows.  Returns false (FAST only) when some lane met a value on the boundary described above.
template <typename PT, bool FAST, typename IO>
__device__ __forceinline__ bool analyze_rows(IO& io, int e, int f, RowAcc<PT>& acc) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	using ST = typename T::ST;
	using FL = FastLimits<PT>;
	const PT       ex = T::exp10(e), frf = T::frac10(f), fre = T::frac10(e);
	const ST       fa  = T::fact10(f);
	const PT       fap = FL::fact_fp(f);
	bool           suspicious = false;
	// ANALYZE_UNROLL rows per loop iteration (not all 32 unrolled): the kernel's instruction footprint is what the warps of an
	// SM share in the instruction caches, and this loop is the largest piece of it
#pragma unroll 1
	for (int r0 = 0; r0 < 32; r0 += ALPB200_ANALYZE_UNROLL) {
		uint32_t excbits = 0;
		static_for<0, ALPB200_ANALYZE_UNROLL>([&](auto R) {
			constexpr int i  = decltype(R)::value;
			const int     r  = r0 + i;
			const UT      xb = io.load_row(r);
			ST            enc;
			bool          exc;
			if constexpr (FAST) {
				const PT       tr  = T::magic_round(T::mul(T::mul(T::from_bits(xb), ex), frf));  // encoder.hpp:83,87
				const PT       pd  = T::mul(tr, fap);
				const uint32_t key = FL::key(pd);
				enc                = FL::cast_sat(tr);
				suspicious |= key == FL::BIG;
				const PT dec = T::mul(pd, fre);
				exc          = (T::bits(dec) != xb) | (key > FL::BIG);
			} else {
				enc          = encode_value<PT, false>(T::from_bits(xb), ex, frf);  // encoder.hpp:345
				const PT dec = decode_value<PT>(enc, fa, fre);                       // :347
				exc          = T::bits(dec) != xb;
			}
			if constexpr (IO::KEEP_EXC) {
				if (!exc) { io.store_row(r, (UT)enc); }  // (a predicated store: no extra instruction)
			} else {
				io.store_row(r, (UT)enc);
			}
			if (exc) { excbits |= 1u << i; }
			if (!exc) {  // exceptions take no part in min / max (predicated, no branch)
				if constexpr (sizeof(PT) == 8) {
					const uint32_t lo = (uint32_t)(uint64_t)enc, hi = (uint32_t)((uint64_t)enc >> 32);
					acc.lo_min = min(acc.lo_min, lo);
					acc.lo_max = max(acc.lo_max, lo);
					acc.hi_and &= hi;
					acc.hi_or |= hi;
				} else {
					acc.mn = min(acc.mn, enc);
					acc.mx = max(acc.mx, enc);
				}
			}
		});
		acc.myexc |= excbits << r0;
	}
	if constexpr (FAST) { return !__any_sync(FULL, suspicious); }
	return true;
}

template <typename PT, typename IO>
__device__ __forceinline__ void analyze_alp(const PT* __restrict__ in_vec, const StateRegs& st, int t, IO& io, Analysis<PT>& a) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	using ST = typename T::ST;
	int e = st.exp_of(0), f = st.fac_of(0);
	if (st.k > 1) { choose_exponent_factor<PT>(io.sample(in_vec, 32 * t), st, e, f); }  // encoder.hpp:409-412

	RowAcc<PT> acc;
	if (!analyze_rows<PT, true>(io, e, f, acc)) {
		acc = RowAcc<PT>();
		io.rewind(in_vec);
		analyze_rows<PT, false>(io, e, f, acc);
	}
	const uint32_t myexc = acc.myexc;
	ST             mn = acc.mn, mx = acc.mx;
	a.myexc = myexc;
	a.cnt   = __reduce_add_sync(FULL, (uint32_t)__popc(myexc));
	// fill value = encoded integer at the first non-exception position, 0 if there is none (encoder.hpp:382-388)
	uint32_t cand = 0xFFFFu;
	if (~myexc) { cand = (uint32_t)Map<PT>::index(t, __ffs((int)~myexc) - 1); }
	cand   = __reduce_min_sync(FULL, cand);
	a.fill = 0;
	if (cand != 0xFFFFu) {
		__syncwarp();  // another lane stored it
		a.fill = (ST)io.kept((int)cand);
		if constexpr (sizeof(PT) == 8) {
			const uint32_t hi_and = __reduce_and_sync(FULL, acc.hi_and);
			const uint32_t hi_or  = __reduce_or_sync(FULL, acc.hi_or);
			if (hi_and == hi_or) {  // one common high word: order is decided by the low words
				mn = (ST)(((uint64_t)hi_or << 32) | __reduce_min_sync(FULL, acc.lo_min));
				mx = (ST)(((uint64_t)hi_or << 32) | __reduce_max_sync(FULL, acc.lo_max));
			} else {  // wide range: full 64-bit pass, re-encoding from the input
				const PT ex = T::exp10(e), frf = T::frac10(f);
#pragma unroll 4
				for (int r = 0; r < 32; r++) {
					const ST v = encode_value<PT, false>(in_vec[Map<PT>::index(t, r)], ex, frf);
					if (!((myexc >> r) & 1u)) {
						mn = v < mn ? v : mn;
						mx = v > mx ? v : mx;
					}
				}
				mn = warp_min<ST>(mn);
				mx = warp_max<ST>(mx);
			}
		} else {
			mn     = warp_min<ST>(mn);
			mx     = warp_max<ST>(mx);
		}
	} else {
		mn = mx = 0;
	}
	a.bw   = (uint32_t)bits_of_range<PT>(mx, mn);
	a.base = mn;
	a.e    = (uint32_t)e;
	a.f    = (uint32_t)f;
}

// ---- ALP_RD analysis: rd.hpp:109-147 --------------------------------------------------------------------------------
// on_index(r, idx) reports the unmasked dictionary index of row r (what rd.hpp:136 stores before FFOR masks it).
template <typename PT, typename IO, typename OnIndex>
__device__ __forceinline__ void analyze_rd(const alpb200_rg_state* state, const StateRegs& st, int t, IO& io, Analysis<PT>& a,
                                           OnIndex&& on_index) {
	using T              = Traits<PT>;
	using UT             = typename T::UT;
	const uint32_t rbw   = st.right_bw(), lbw = st.left_bw(), ds = st.dict_size();
	const UT       rmask = low_mask<UT>(rbw);
	const uint32_t lmask = (1u << lbw) - 1;
	uint32_t       myexc = 0;
	uint32_t       nib0 = 0, nib1 = 0, nib2 = 0, nib3 = 0;
	// the dictionary as eight 32-bit compare operands; slots beyond dict_size can never match (left parts are < 2^16)
	uint32_t dict[ALPB200_RD_DICT_SIZE];
#pragma unroll
	for (uint32_t d = 0; d < ALPB200_RD_DICT_SIZE; d++) {
		dict[d] = d < ds ? dict_lookup(st.dict, d) : 0xFFFFFFFFu;
	}
	static_for<0, 32>([&](auto R) {
		constexpr int  r    = decltype(R)::value;
		const UT       bits = io.load(R);
		const uint32_t left = (uint32_t)(bits >> rbw);
		if constexpr (!IO::KEEP_EXC) { io.store(R, bits & rmask); }  // (KEEP_EXC: the packer masks to right_bw bits itself)
		uint32_t idx = ds;  // rd.hpp:129-131: a left part nobody has seen gets the smallest non-dictionary index
		// (entries are distinct — rd.hpp:56-66 takes them from a map's keys; descending, so that the lowest slot would win)
#pragma unroll
		for (int d = ALPB200_RD_DICT_SIZE - 1; d >= 0; d--) {
			idx = left == dict[d] ? (uint32_t)d : idx;
		}
		const bool hit = idx != ds;
		if (st.n_extra != 0 && !hit) {  // rd.hpp:73-77,133: sampled left parts outside the dictionary
			for (uint32_t x = 0; x < st.n_extra; x++) {
				if (state->extra_key[x] == left) {
					idx = state->extra_idx[x];
					break;
				}
			}
		}
		on_index(r, idx);
		myexc |= (uint32_t)(idx >= ds) << r;  // rd.hpp:138-142
		const uint32_t n = (idx & lmask) << (4 * (r & 7));  // FFOR masks the stored index to left_bw bits
		if constexpr (r < 8) {
			nib0 |= n;
		} else if constexpr (r < 16) {
			nib1 |= n;
		} else if constexpr (r < 24) {
			nib2 |= n;
		} else {
			nib3 |= n;
		}
	});
	a.left_nib[0] = nib0;
	a.left_nib[1] = nib1;
	a.left_nib[2] = nib2;
	a.left_nib[3] = nib3;
	a.myexc       = myexc;
	a.cnt         = __reduce_add_sync(FULL, (uint32_t)__popc(myexc));
	a.bw          = rbw;
	a.e           = lbw;
	a.f           = ds;
	a.base        = 0;
	a.fill        = 0;
}

// ---- FFOR bit packer (write side of SURVEY.md appendix A.1; src/fastlanes_generated_ffor.cpp:7788-7999) ----------
// Width-specialised (alp_ffor.cuh): one `switch (bw)` per vector, then every shift and register index is a constant.
// 64-bit lanes: a thread builds its BW 32-bit words, pairs them into 64-bit elements (element 16*w + lane of the block)
// and a half-warp stores one full 128-byte line per instruction; for odd BW the element shared by the two halves of a
// lane is completed with one shuffle.  Exception slots take `fill` (encoder.hpp:393).
__device__ __forceinline__ void pack_rows(const uint64_t* __restrict__ tile, uint32_t myexc, uint64_t fill, uint64_t base,
                                          uint32_t bw, int t, uint8_t* __restrict__ dst) {
	const int lane = t & 15, half = t >> 4;
	dispatch_width<0, 64>(bw, [&](auto W) {
		constexpr int BW = decltype(W)::value;
		if constexpr (BW > 0) {  // ffor bw=0 writes nothing (src/fastlanes_generated_ffor.cpp:4)
			pack64_rows<BW>(lane, half, reinterpret_cast<uint64_t*>(dst), [&](auto R, uint32_t& lo, uint32_t& hi) {
				constexpr int r = decltype(R)::value;
				uint64_t      v = tile[Map<double>::index(t, r)];
				if ((myexc >> r) & 1u) { v = fill; }
				const uint64_t d = v - base;  // masked to BW bits by the packer
				lo               = (uint32_t)d;
				hi               = (uint32_t)(d >> 32);
			});
		}
	});
}
// 32-bit lanes: word j of lane t at element 32*j + t — every store instruction writes one full 128-byte line
__device__ __forceinline__ void pack_rows(const uint32_t* tile, uint32_t myexc, uint32_t fill, uint32_t base, uint32_t bw, int t,
                                          uint8_t* dst) {
	dispatch_width<0, 32>(bw, [&](auto W) {
		constexpr int BW = decltype(W)::value;
		if constexpr (BW > 0) {
			pack32_rows<BW>(t, reinterpret_cast<uint32_t*>(dst), [&](auto R) -> uint32_t {
				constexpr int r = decltype(R)::value;
				uint32_t      v = tile[Map<float>::index(t, r)];
				if ((myexc >> r) & 1u) { v = fill; }
				return v - base;
			});
		}
	});
}

// FFOR IN PLACE: the tile of encoded integers becomes the packed block image (bytes [0, 128*bw) of the tile), ready to
// leave with one bulk-async store.  Runs BEFORE the block's output offset is known, i.e. it overlaps the placement wait.
// 64-bit lanes: bw <= 32 keeps every word in registers until the warp has read the whole tile and reads only the low
// words ((v - base) mod 2^32 is all a field of <= 32 bits needs); wider blocks defer just the words that would land on
// unread rows (alp_ffor.cuh, PACK_INPLACE_WIDE).
__device__ __forceinline__ void pack_rows_inplace(uint64_t* tile, uint32_t myexc, uint64_t fill, uint64_t base, uint32_t bw, int t) {
	const int       lane = t & 15, half = t >> 4;
	const uint32_t* lo32 = reinterpret_cast<const uint32_t*>(tile);
	dispatch_width<0, 64>(bw, [&](auto W) {
		constexpr int BW = decltype(W)::value;
		if constexpr (BW == 0) {
			return;
		} else if constexpr (BW <= 32) {
			pack64_rows<BW, PACK_INPLACE_NARROW>(lane, half, tile, [&](auto R, uint32_t& lo, uint32_t& hi) {
				constexpr int r = decltype(R)::value;
				uint32_t      v = lo32[2 * Map<double>::index(t, r)];
				if ((myexc >> r) & 1u) { v = (uint32_t)fill; }
				lo = v - (uint32_t)base;
				hi = 0;
			});
		} else {
			pack64_rows<BW, PACK_INPLACE_WIDE>(lane, half, tile, [&](auto R, uint32_t& lo, uint32_t& hi) {
				constexpr int r = decltype(R)::value;
				uint64_t      v = tile[Map<double>::index(t, r)];
				if ((myexc >> r) & 1u) { v = fill; }
				const uint64_t d = v - base;  // masked to BW bits by the packer
				lo               = (uint32_t)d;
				hi               = (uint32_t)(d >> 32);
			});
		}
	});
}
// 32-bit lanes: pack32_rows is in-place safe as it is (alp_ffor.cuh)
__device__ __forceinline__ void pack_rows_inplace(uint32_t* tile, uint32_t myexc, uint32_t fill, uint32_t base, uint32_t bw, int t) {
	pack_rows(tile, myexc, fill, base, bw, t, reinterpret_cast<uint8_t*>(tile));
}

__device__ __forceinline__ uint32_t nib(const uint32_t (&n)[4], int r) { return (n[r >> 3] >> (4 * (r & 7))) & 0xFu; }

// pack the ALP_RD dictionary indices on 16-bit lanes (64 lanes x 16 rows): value v = 64*row16 + lane16
__device__ __forceinline__ void pack_left(const uint32_t (&left_nib)[4], uint32_t lbw, int t, uint16_t* __restrict__ blk, double /*tag*/) {
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
__device__ __forceinline__ void pack_left(const uint32_t (&left_nib)[4], uint32_t lbw, int t, uint16_t* __restrict__ blk, float /*tag*/) {
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
// plan_exceptions computes the ranks: a 32x32 bit-matrix transpose of the per-thread bitmaps and two warp scans — nothing
// that needs the output offset, so the batched encoder does it before it waits for that offset.
//   value_of(p) returns what is stored for position p (the original value for ALP, the left part for ALP_RD).
struct ExcPlan {
	uint32_t rowmask;  // lane r: ballot of "is exception" over the threads' row r
	uint32_t pre;      // lane r: exceptions before row r (64-bit lanes: low / high 16 bits for the two halves of the warp)
	bool     any;
};
template <typename PT>
__device__ __forceinline__ ExcPlan plan_exceptions(uint32_t myexc, int t) {
	ExcPlan pl;
	pl.any     = __any_sync(FULL, myexc != 0);
	pl.rowmask = 0;
	pl.pre     = 0;
	if (!pl.any) { return pl; }
	pl.rowmask = transpose32(myexc, t);
	if (sizeof(PT) == 8) {
		uint32_t       tot_lo, tot_hi;
		const uint32_t p_lo = warp_excl_scan(__popc(pl.rowmask & 0xFFFFu), t, tot_lo);
		const uint32_t p_hi = warp_excl_scan(__popc(pl.rowmask >> 16), t, tot_hi) + tot_lo;
		pl.pre              = p_lo | (p_hi << 16);  // both at most 1024
	} else {
		uint32_t tot;
		pl.pre = warp_excl_scan(__popc(pl.rowmask), t, tot);
	}
	return pl;
}
// Every thread walks ITS OWN exceptions and fetches the rank of each one's row from lane r: the loop runs
// max-exceptions-per-thread times (2-4 for a typical vector) instead of once per row that holds an exception.
// visit(rank, position) is called once per exception.
template <typename PT, typename Visit>
__device__ __forceinline__ void for_each_exception(const ExcPlan& pl, uint32_t myexc, int t, Visit&& visit) {
	if (!pl.any) { return; }
	uint32_t m = myexc;
	while (__any_sync(FULL, m != 0)) {
		const int      r  = m ? __ffs((int)m) - 1 : 0;
		const uint32_t rm = __shfl_sync(FULL, pl.rowmask, r);
		const uint32_t pr = __shfl_sync(FULL, pl.pre, r);
		if (m) {
			uint32_t rank;
			if (sizeof(PT) == 8) {
				const int      lane = t & 15, half = t >> 4;
				const uint32_t hm   = half ? (rm >> 16) : (rm & 0xFFFFu);
				rank                = (half ? (pr >> 16) : (pr & 0xFFFFu)) + __popc(hm & ((1u << lane) - 1));
			} else {
				rank = pr + __popc(rm & ((1u << t) - 1));
			}
			visit(rank, (uint32_t)Map<PT>::index(t, r));
		}
		m &= m - 1;
	}
}
template <typename PT, typename ValueOf, typename Store>
__device__ __forceinline__ void emit_exceptions(uint32_t myexc, int t, ValueOf&& value_of, Store&& store) {
	const ExcPlan pl = plan_exceptions<PT>(myexc, t);
	for_each_exception<PT>(pl, myexc, t, [&](uint32_t rank, uint32_t p) { store(rank, p, value_of(p)); });
}

#ifndef ALPB200_ENC_LOOKBACK
#define ALPB200_ENC_LOOKBACK 1  // 1: blocks resolve their prefix by look-back from the scanner's anchors; 0: per-block prefixes from the scanner
#endif
#ifndef ALPB200_ENC_EARLY_ANCHOR
#define ALPB200_ENC_EARLY_ANCHOR 1  // the look-back's anchor words are requested before the pack (one L2 round trip off the wait)
#endif
#ifndef ALPB200_ENC_SPIN_NS
#define ALPB200_ENC_SPIN_NS 100  // back-off of the thread that polls for its block's prefix (frees issue slots and L2 bandwidth); 0 = none
#endif
__device__ __forceinline__ void enc_backoff() {
#if ALPB200_ENC_SPIN_NS > 0
	__nanosleep(ALPB200_ENC_SPIN_NS);
#endif
}

// ---- placement: in-order prefix sums over thread blocks ---------------------------------------------------------------
// Every block publishes its aggregate (packed 128-byte units << AGG_SHIFT | exception slots) and then waits for its exclusive
// prefix.  The prefixes are produced by ONE scanner warp — the last warp of the block that drew ticket 0 turns into it
// once its own vector is written — which walks the aggregates in ticket order, 128 per step (one 32-byte sector per
// lane), and publishes the running sums.  The scanner advances 128 blocks per L2 round trip, several times faster than
// blocks are produced, so a block waits about one round trip after its slowest predecessor has published.  (A classic
// decoupled look-back with a 32-wide window per round trip cannot keep up here: ~450 blocks are in flight and the prefix
// frontier would advance only 32 blocks per round trip.)  Tickets are drawn when a block starts, so every predecessor
// of a waiting block is already resident: no deadlock.
// status word: [63] valid | [61:33] packed size in 128-byte units | [32:0] exception slots
constexpr uint64_t SCAN_VALID = 1ull << 63, SCAN_VAL = (1ull << 62) - 1;
// aggregate / prefix value: packed 128-byte units << AGG_SHIFT | exception slots.  33 bits of exception slots: a call of 2^22
// vectors whose every value is an exception (2^32 slots) must not carry into the units half; 29 bits of units cover 2^22
// vectors at the widest block (66 units).
constexpr int      AGG_SHIFT    = 33;
constexpr uint64_t AGG_EXC_MASK = (1ull << AGG_SHIFT) - 1;

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
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d) {
	const uint32_t lo = __shfl_up_sync(FULL, (uint32_t)v, d), hi = __shfl_up_sync(FULL, (uint32_t)(v >> 32), d);
	return ((uint64_t)hi << 32) | lo;
}
// run by one warp: aggregates[0..n_blocks) -> prefixes[0..n_blocks) (exclusive), in order.  Each step looks at the
// next 128 aggregates and publishes prefixes for the leading run that is already valid, so a block never waits for
// blocks behind it.
__device__ __forceinline__ void scan_blocks(const uint64_t* aggregates, uint64_t* prefixes, uint32_t n_blocks, int t, uint64_t start) {
	uint64_t running = start;
	uint32_t base    = 0;
	while (base < n_blocks) {
		uint64_t v[4];
		uint32_t n_ok = 0;  // leading valid entries among this lane's four
		bool     run  = true;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t idx = base + 4 * t + k;
			v[k]               = idx < n_blocks ? ld_volatile_u64(&aggregates[idx]) : 0;
		}
#pragma unroll
		for (int k = 0; k < 4; k++) {
			run = run && (v[k] & SCAN_VALID);
			n_ok += run;
		}
		// number of leading valid entries in the 128-window: lanes before the first incomplete lane are full
		const uint32_t incomplete = __ballot_sync(FULL, n_ok < 4);
		const int      first      = incomplete ? __ffs(incomplete) - 1 : 32;
		const uint32_t n_first    = __shfl_sync(FULL, n_ok, first & 31);
		const uint32_t n_lead     = incomplete ? 4u * first + n_first : 128u;
		if (n_lead == 0) { continue; }
		uint64_t mine = 0;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			if ((uint32_t)(4 * t + k) < n_lead) { mine += v[k] & SCAN_VAL; }
		}
		uint64_t incl = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint64_t o = shfl_up_u64(incl, d);
			if (t >= d) { incl += o; }
		}
		uint64_t excl = running + incl - mine;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			if ((uint32_t)(4 * t + k) < n_lead) {
				st_volatile_u64(&prefixes[base + 4 * t + k], SCAN_VALID | excl);
				excl += v[k] & SCAN_VAL;
			}
		}
		running += shfl_u64(incl, 31);
		base += n_lead;
	}
}

// ---- placement, second scheme: adaptive look-back on top of ANCHORS -------------------------------------------------
// With per-block prefixes from the scanner a waiting block sits behind three L2 hops (its last predecessor's aggregate ->
// scanner -> prefix -> waiter): ~2.5 us per block, 20-30 % of the vector-order kernel.  Here the scanner only publishes an
// ANCHOR every 32 blocks (the exclusive prefix of block 32 G), and a block resolves its own prefix: it takes the nearest
// anchor that is already there — typically 3-4 groups back, the scanner's lag — and adds the aggregates from that anchor up
// to its predecessor, which it reads itself, 32 per load, all loads in flight at once.  Only the newest few of them can
// still be missing; each lane re-polls just its own missing entries.  The critical path shrinks to ONE hop (the last
// predecessor's aggregate becoming visible), the scanner is off it as long as it stays within LB_DEPTH groups of the
// frontier, and the polling traffic is ~2 KB per block once plus 8 bytes per pending entry per poll.
// (Round 1 tried a fixed 32-block window on top of per-block prefixes — the window's start still waited for the scanner —
// and a 128-wide window re-read in full on every poll, ~1 TB/s of L2 traffic; both lost.  The adaptive anchor is what
// takes the scanner off the critical path, the per-entry re-poll what keeps the traffic down.)
#ifndef ALPB200_ENC_LB_DEPTH
#define ALPB200_ENC_LB_DEPTH 4  // groups of 32 blocks a look-back reaches back (A/B on B200: 3-4 beat 8 by 1.5-4 %: the nearest anchor is 1-2 groups away)
#endif
constexpr uint32_t LB_GROUP = 32, LB_DEPTH = ALPB200_ENC_LB_DEPTH;

// run by one warp: anchors[G + 1] = exclusive prefix of block 32 (G + 1), for every complete group of 32 blocks, in order
// (anchors[0] = the call's start offset, written by encode_prepare_kernel)
__device__ __forceinline__ void scan_anchors(const uint64_t* aggregates, uint64_t* anchors, uint32_t n_blocks, int t, uint64_t start) {
	uint64_t       running  = start;
	const uint32_t n_groups = n_blocks / LB_GROUP;  // complete groups; a ragged last group needs no anchor behind it
	uint32_t       G        = 0;
	while (G < n_groups) {
		uint64_t v[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			v[k] = G + k < n_groups ? ld_volatile_u64(&aggregates[(size_t)LB_GROUP * (G + k) + t]) : 0;
		}
		uint32_t done = 0;  // leading complete groups among the four
		bool     run  = true;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			run = run && __all_sync(FULL, (v[k] & SCAN_VALID) != 0);
			if (run) {
				running += warp_sum_u64(v[k] & SCAN_VAL);
				if (t == 0) { st_volatile_u64(&anchors[G + k + 1], SCAN_VALID | running); }
				done++;
			}
		}
		G += done;
		if (done == 0) { enc_backoff(); }
	}
}

// the anchor words a look-back starts from (lane j: anchor g - j).  They can be loaded ahead of time: a block issues this right
// after publishing its aggregate and packs its vector while the words are on their way (ALPB200_ENC_EARLY_ANCHOR).
__device__ __forceinline__ uint64_t lookback_anchor_word(const uint64_t* anchors, uint32_t bid, int t) {
	const uint32_t g = bid / LB_GROUP;
	return ((uint32_t)t < LB_DEPTH && (uint32_t)t <= g) ? ld_volatile_u64(&anchors[g - t]) : 0ull;
}
// run by one warp of a block that has published its aggregate: the exclusive prefix of block `bid`
__device__ __forceinline__ uint64_t lookback_prefix(const uint64_t* aggregates, const uint64_t* anchors, uint32_t bid, int t,
                                                    uint64_t early_anchor = 0) {
	const uint32_t g = bid / LB_GROUP;
	// the nearest anchor that is already published, at most LB_DEPTH - 1 groups back (lane j looks at anchor g - j)
	uint64_t base = 0;
	uint32_t G    = 0;
	for (bool first = true;; first = false) {
		uint64_t a = early_anchor;
		if (!first || !__any_sync(FULL, (early_anchor & SCAN_VALID) != 0)) { a = lookback_anchor_word(anchors, bid, t); }
		const uint32_t ok = __ballot_sync(FULL, (a & SCAN_VALID) != 0);
		if (ok) {
			const int j = __ffs((int)ok) - 1;
			base        = shfl_u64(a, j) & SCAN_VAL;
			G           = g - (uint32_t)j;
			break;
		}
		enc_backoff();  // the scanner is more than LB_DEPTH groups behind: wait for it
	}
	// aggregates of blocks [32 G, bid): chunk k = blocks 32 (G + k) .. 32 (G + k) + 31, one per lane; all loads go out at once
	uint64_t v[LB_DEPTH];
	bool     pending = false;
#pragma unroll
	for (uint32_t k = 0; k < LB_DEPTH; k++) {
		const uint32_t i = LB_GROUP * (G + k) + (uint32_t)t;
		v[k]             = i < bid ? ld_volatile_u64(&aggregates[i]) : SCAN_VALID;  // (beyond the predecessor: nothing to add)
	}
#pragma unroll
	for (uint32_t k = 0; k < LB_DEPTH; k++) {
		pending = pending || !(v[k] & SCAN_VALID);
	}
	while (__any_sync(FULL, pending)) {  // the newest predecessors are still analysing: re-poll only what is missing
		enc_backoff();
		pending = false;
#pragma unroll
		for (uint32_t k = 0; k < LB_DEPTH; k++) {
			if (!(v[k] & SCAN_VALID)) {
				v[k]    = ld_volatile_u64(&aggregates[LB_GROUP * (G + k) + (uint32_t)t]);
				pending = pending || !(v[k] & SCAN_VALID);
			}
		}
	}
	uint64_t sum = 0;
#pragma unroll
	for (uint32_t k = 0; k < LB_DEPTH; k++) {
		sum += v[k] & SCAN_VAL;
	}
	return base + warp_sum_u64(sum);
}

__device__ __forceinline__ void run_scanner(const uint64_t* aggregates, uint64_t* prefixes, uint32_t n_blocks, int t, uint64_t start) {
#if ALPB200_ENC_LOOKBACK
	scan_anchors(aggregates, prefixes, n_blocks, t, start);
#else
	scan_blocks(aggregates, prefixes, n_blocks, t, start);
#endif
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

// workspace layout: [0] ticket counter, [1] where this call's output starts (packed units << AGG_SHIFT | exception slots: 0, or
// the column's running totals when appending — set by encode_prepare_kernel; completion order allocates from it with
// atomics), [2 .. 2+n_blocks) block aggregates, [2+n_blocks .. 2+2*n_blocks) exclusive prefixes.
//
// One CTA = WARPS vectors, one warp each.  Per warp:
//   1  the vector arrives in a per-warp shared-memory tile with one bulk-async copy (TMA 1-D)
//   2  analysis in place: the tile now holds the encoded integers (ALP) / right parts (ALP_RD)
//   3  the CTA publishes its aggregate (packed units, exceptions) and the scanner starts working out its offsets
//   4  WHILE the offsets are on their way: FFOR in place — the head of the tile becomes the packed block image
//   5  once the offsets are known the block leaves with one bulk-async store (TMA 1-D); exceptions and the record follow
// Wide 64-bit-lane blocks (bw > 32: ALP_RD on doubles, huge integers) cannot be packed in place within the register
// budget; they are FFOR-ed from the tile straight into the column after step 3's wait (full 128-byte line stores).
//
// ORDERED = true  (alpb200_encode_*): blocks and exceptions land in VECTOR ORDER — the bytes of a column are a pure function
//                  of its values, any vector range is one contiguous byte range.  Price: a block's CTA waits until every
//                  predecessor has published its size (measured: 25-30 % of the kernel on B200).
// ORDERED = false (alpb200_encode_unordered_*): one atomicAdd hands out the space, so blocks land in COMPLETION order:
//                  the same blocks, the same records (offsets differ), dense, but not sorted by vector.
//
// vectors (= warps) per thread block: build knobs (tools/build_variant.sh)
// ALPB200_ENC_HELPER: the vector-order kernel's blocks carry one extra warp WITHOUT a vector.  It sums the block's sizes,
// publishes the aggregate and resolves the block's prefix while the other warps pack — the look-back's L2 round trips leave the
// path of the warps that hold tiles (with the work on warp 0, its pack and its look-back were in series: ~2 us of a ~10 us round).
#ifndef ALPB200_ENC_HELPER
#define ALPB200_ENC_HELPER 0
#endif
#ifndef ALPB200_ENC_WPS64
#define ALPB200_ENC_WPS64 27  // resident warps per SM the f64 kernels' registers are limited for
#endif
#ifndef ALPB200_ENC_WARPS64
#define ALPB200_ENC_WARPS64 9
#endif
#ifndef ALPB200_ENC_WARPS32
#define ALPB200_ENC_WARPS32 8
#endif
#ifndef ALPB200_ENC_EXC_REGS64
#define ALPB200_ENC_EXC_REGS64 2
#endif
#ifndef ALPB200_ENC_EXC_REGS32
#define ALPB200_ENC_EXC_REGS32 4
#endif
template <typename PT>
struct EncodeCfg;
template <>
struct EncodeCfg<double> {
	static constexpr uint32_t SMEM_PER_WARP = VEC * sizeof(double);  // the tile
	static constexpr uint32_t INPLACE_MAX   = SMEM_PER_WARP;         // largest block (bytes) packed in place: whatever fits the tile
	// 9 warps x 3 blocks = 27 warps per SM: what 227 KiB of shared memory hold at 8 KiB a vector (72 registers per thread).
	// Measured against 8 x 3 (80 registers): 2-3 % faster.
	static constexpr int      EXC_REGS      = ALPB200_ENC_EXC_REGS64;  // exceptions held in registers per lane across the placement wait (x 32 per vector)
	static constexpr int      WARPS         = ALPB200_ENC_WARPS64;
	static constexpr int      WARPS_PER_SM  = ALPB200_ENC_WPS64;
};
template <>
struct EncodeCfg<float> {
	static constexpr uint32_t SMEM_PER_WARP = 35 * 128u;  // tile (32 units); every f32 block fits: 32 bits, or ALP_RD 31 + 3
	static constexpr uint32_t INPLACE_MAX   = SMEM_PER_WARP;
	static constexpr int      EXC_REGS      = ALPB200_ENC_EXC_REGS32;
	static constexpr int      WARPS         = ALPB200_ENC_WARPS32;
	static constexpr int      WARPS_PER_SM  = 32;  // 4.4 KiB and 64 registers
};

// runs after the workspace was zeroed: appending calls continue at the column's running totals
static __global__ void encode_prepare_kernel(uint64_t* workspace, const uint64_t* totals, int append, uint32_t n_blocks) {
	const uint64_t start    = append ? (((totals[0] / 128ull) << AGG_SHIFT) | (totals[1] & AGG_EXC_MASK)) : 0ull;
	workspace[1]            = start;
	workspace[2 + n_blocks] = SCAN_VALID | start;  // anchors[0] (look-back placement; the slot is prefixes[0] otherwise and rewritten)
}

template <bool ORDERED>
__host__ __device__ constexpr int enc_helper_warps() {
	return ORDERED && ALPB200_ENC_HELPER ? 1 : 0;
}
template <typename PT, int WARPS, bool ORDERED>
__global__ void __launch_bounds__((WARPS + enc_helper_warps<ORDERED>()) * 32, EncodeCfg<PT>::WARPS_PER_SM / (WARPS + enc_helper_warps<ORDERED>())) encode_kernel(const PT* __restrict__ in, uint64_t n_vectors,
                                                                                    const alpb200_rg_state* __restrict__ states,
                                                                                    ColOut col, uint64_t* workspace) {
	using T   = Traits<PT>;
	using UT  = typename T::UT;
	using Cfg = EncodeCfg<PT>;
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ uint32_t s_bid;
	__shared__ uint32_t s_units[WARPS], s_cnt[WARPS];
	__shared__ uint64_t s_excl, s_pre[WARPS];
	__shared__ __align__(8) uint64_t s_bar[WARPS];
	constexpr int EXC_REGS = Cfg::EXC_REGS, EXC_CAP = 32 * EXC_REGS;
	__shared__ uint16_t s_pos[WARPS][EXC_CAP > 0 ? EXC_CAP : 1];  // positions of a vector's first EXC_CAP exceptions, by rank

	const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
	// ORDERED: tickets are drawn when a block starts, so every predecessor of a waiting block is resident (no deadlock)
	// and ticket order = start order.  Completion order needs neither: the block index will do, one L2 round trip saved.
	uint32_t bid = blockIdx.x;
	if constexpr (ORDERED) {
		if (threadIdx.x == 0) { s_bid = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long*>(workspace), 1ull); }
		__syncthreads();
		bid = s_bid;
	}
	constexpr int  PUB    = enc_helper_warps<ORDERED>() ? WARPS : 0;  // the warp that publishes the block's sizes and resolves its prefix
	const bool     helper = enc_helper_warps<ORDERED>() && warp == WARPS;
	const uint64_t v      = (uint64_t)bid * WARPS + warp;
	const bool     active = !helper && v < n_vectors;
	uint8_t*       mine   = smem + (size_t)warp * Cfg::SMEM_PER_WARP;  // the tile
	UT*            tile   = reinterpret_cast<UT*>(mine);

	Analysis<PT> a;
	a.cnt = a.bw = a.e = a.f = a.myexc = 0;
	a.base = a.fill = 0;
	StateRegs st;
	st.scheme                      = ALPB200_SCHEME_INVALID;
	const PT*               in_vec = in + v * (uint64_t)VEC;
	const alpb200_rg_state* state  = states + (active ? v / ALPB200_ROWGROUP_VECTORS : 0);
	uint32_t                units  = 0;
	bool                    rd     = false;
	if (active) {
		// the whole input vector (contiguous) arrives in the tile with one bulk-async copy (TMA 1-D)
		if (t == 0) {
			mbar_init(&s_bar[warp], 1);
			fence_mbar_init();
			mbar_arrive_expect_tx(&s_bar[warp], VEC * sizeof(PT));
			bulk_g2s(tile, in_vec, VEC * sizeof(PT), &s_bar[warp]);
		}
		st = load_state(state);
		rd = st.scheme == ALPB200_SCHEME_ALP_RD;
		__syncwarp();
		mbar_wait(&s_bar[warp], 0);
		TileIO<PT, true> io(tile, t);  // exception slots keep their original bits (ALP_RD: every slot)
		if (rd) {
			analyze_rd<PT>(state, st, t, io, a, [](int, uint32_t) {});
		} else {
			analyze_alp<PT>(in_vec, st, t, io, a);
		}
		units = rd ? a.bw + a.e : a.bw;
	}
	if (t == 0 && !helper) {
		s_units[warp] = units;
		s_cnt[warp]   = a.cnt;
	}
	__syncthreads();
	uint64_t* aggregates = workspace + 2;
	uint64_t* prefixes   = aggregates + gridDim.x;
	uint64_t  agg        = 0, early_excl = 0, early_anchor = 0;
	if (warp == PUB) {
		uint64_t mine_agg = 0;
		if (t < WARPS) { mine_agg = ((uint64_t)s_units[t] << AGG_SHIFT) | s_cnt[t]; }
		agg                 = warp_sum_u64(mine_agg);
		const uint32_t wide = __reduce_max_sync(FULL, (uint32_t)(mine_agg >> AGG_SHIFT));
		uint64_t       incl = mine_agg;  // offsets of the warps inside the block
#pragma unroll
		for (int d = 1; d < WARPS; d <<= 1) {
			const uint64_t o = shfl_up_u64(incl, d);
			if (t >= d) { incl += o; }
		}
		if (t < WARPS) { s_pre[t] = incl - mine_agg; }
		if (t == 0) {
			if constexpr (ORDERED) {
				st_volatile_u64(&aggregates[bid], SCAN_VALID | agg);
			} else {
				// completion order: one atomic hands out the block's space (its round trip overlaps the packing below);
				// the running totals are the column totals
				early_excl = (uint64_t)atomicAdd(reinterpret_cast<unsigned long long*>(workspace + 1), (unsigned long long)agg);
				atomicAdd(reinterpret_cast<unsigned long long*>(&col.totals[0]), (unsigned long long)(agg >> AGG_SHIFT) * 128ull);
				atomicAdd(reinterpret_cast<unsigned long long*>(&col.totals[1]), (unsigned long long)(agg & AGG_EXC_MASK));
			}
			atomicMax(reinterpret_cast<unsigned long long*>(&col.totals[3]), (unsigned long long)wide * 128ull);
		}
#if ALPB200_ENC_LOOKBACK && ALPB200_ENC_EARLY_ANCHOR
		if (ORDERED && bid != 0) { early_anchor = lookback_anchor_word(prefixes, bid, t); }  // consumed after the pack below
#endif
	}
	const uint32_t bytes = units * 128u;
	// ---- exceptions, first part: ranks, and the ORIGINALS of up to EXC_CAP exceptions into registers -------------------------
	// The tile still holds the original bits at exception slots (KEEP_EXC), so the values are gathered from shared memory
	// now — before the pack overwrites the tile and before the placement wait — instead of being re-read from global memory
	// (one dependent L2 round trip per vector) once the output offset is known.  What remains after the wait is stores.
	const uint32_t rbw = a.bw;
	ExcPlan        plan;
	plan.any     = false;
	plan.rowmask = plan.pre = 0;
	UT       ev_reg[EXC_REGS > 0 ? EXC_REGS : 1];
	uint32_t ep_reg[EXC_REGS > 0 ? EXC_REGS : 1];
	bool     captured = false;  // warp-uniform
	if (active) {
		plan     = plan_exceptions<PT>(a.myexc, t);
		captured = EXC_REGS > 0 && plan.any && a.cnt <= (uint32_t)EXC_CAP;
		if (captured) {
			for_each_exception<PT>(plan, a.myexc, t, [&](uint32_t rank, uint32_t p) { s_pos[warp][rank] = (uint16_t)p; });
			__syncwarp();  // positions filed; the tile is complete (analysis stored to it lane by lane)
#pragma unroll
			for (int k = 0; k < EXC_REGS; k++) {
				const uint32_t i = (uint32_t)t + 32u * k;
				ev_reg[k] = 0;
				ep_reg[k] = 0;
				if (i < a.cnt) {
					const uint32_t p    = s_pos[warp][i];
					const UT       bits = tile[p];
					ep_reg[k]           = p;
					ev_reg[k]           = rd ? (UT)(bits >> rbw) : bits;
				}
			}
		}
	}
	// ---- step 4: build the packed block image in shared memory while the scanner works out this block's offsets ----
	bool staged = false;  // warp-uniform: the block sits at `mine`, ready for a bulk store
	// (An ALP_RD block of 63 + 3 or 62 + 3 units on doubles would outgrow the tile.  Completion order has no wait to fill:
	// there wide blocks are better off with the direct line stores — 1.74 vs 1.80 ms per 2^29 on the ALP_RD column; with
	// the wait, packing in place first wins, 1.99 vs 2.05 ms.)
	if (active && bytes <= Cfg::INPLACE_MAX && (ORDERED || a.bw <= 32)) {
		__syncwarp();  // the tile is complete (analysis stored to it lane by lane)
		pack_rows_inplace(tile, rd ? 0u : a.myexc, (UT)a.fill, (UT)a.base, a.bw, t);
		if (rd) {  // the index block follows the right parts
			__syncwarp();
			pack_left(a.left_nib, a.e, t, reinterpret_cast<uint16_t*>(mine + 128u * a.bw), PT());
		}
		staged = true;
	}
	if (staged) { fence_proxy_async_smem(); }  // generic-proxy writes to the block image -> visible to the bulk-copy engine
	// Exceptions.  Their ranks (position order) come from a bit-matrix transpose of the per-thread bitmaps and two warp
	// scans; walking them is a loop over the exceptions of the busiest thread.  When the part of the tile behind the block
	// image has room (2 bytes per exception), every thread FILES the positions of its exceptions at their ranks there —
	// shuffles and shared-memory stores only, and all of it before the output offset is needed — and the emission after
	// the wait is a coalesced loop: 32 consecutive ranks per step, 32 independent loads of the original values in flight,
	// contiguous stores.  (Walking the exceptions thread by thread with one dependent L2 load per step was what kept
	// exception-heavy columns — 90-150 per vector — at 0.42-0.49 of the roofline.)  Otherwise the values are fetched and stored inside that walk.
	auto           exc_value = [&](uint32_t p) -> UT {
        const UT bits = T::bits(in_vec[p]);
        return rd ? (UT)(bits >> rbw) : bits;
	};
	uint16_t* pos_list = reinterpret_cast<uint16_t*>(mine + bytes);
	bool      listed   = false;  // warp-uniform
	if (active && !captured) {
		listed = staged && plan.any && Cfg::SMEM_PER_WARP - bytes >= 2u * a.cnt;
		if (listed) {
			__syncwarp();  // every lane is done reading its rows of the tile (32-bit lanes pack without a warp-wide sync)
			for_each_exception<PT>(plan, a.myexc, t, [&](uint32_t rank, uint32_t p) { pos_list[rank] = (uint16_t)p; });
			__syncwarp();
		}
	}
	if (warp == PUB) {
		uint64_t excl = 0;
		if constexpr (!ORDERED) {
			excl = early_excl;
		} else if (bid == 0) {
			if (t == 0) { excl = workspace[1]; }  // where this call's output starts (0 unless appending)
		} else {
#if ALPB200_ENC_LOOKBACK
			excl = lookback_prefix(aggregates, prefixes, bid, t, early_anchor);  // (`prefixes` holds the anchors in this scheme)
#else
			if (t == 0) {
				while (!((excl = ld_volatile_u64(&prefixes[bid])) & SCAN_VALID)) { enc_backoff(); }
				excl &= SCAN_VAL;
			}
#endif
		}
		if (t == 0) {
			s_excl = excl;
			if (ORDERED && (uint64_t)(bid + 1) * WARPS >= n_vectors) {  // last block: publish the column totals
				const uint64_t incl = excl + agg;
				col.totals[0]       = (incl >> AGG_SHIFT) * 128ull;
				col.totals[1]       = incl & AGG_EXC_MASK;
			}
		}
	}
	__syncthreads();
	// this warp produces every block's prefix once its own work is done (the helper warp when there is one)
	const bool scanner = ORDERED && bid == 0 && warp == (enc_helper_warps<ORDERED>() ? WARPS : WARPS - 1);
	if (!active) {
		if (scanner) { run_scanner(aggregates, prefixes, gridDim.x, t, workspace[1]); }
		return;
	}
	const uint64_t my_excl   = s_excl + s_pre[warp];  // both halves add without carry into each other (sizes checked below)
	const uint64_t units_off = my_excl >> AGG_SHIFT, exc_off = my_excl & AGG_EXC_MASK;
	if (units_off * 128ull + bytes > col.packed_capacity || exc_off + a.cnt > col.exc_capacity) {
		if (t == 0) { atomicExch(reinterpret_cast<unsigned long long*>(&col.totals[2]), 1ull); }
		if (scanner) { run_scanner(aggregates, prefixes, gridDim.x, t, workspace[1]); }
		return;
	}
	uint8_t* dst = col.packed + units_off * 128ull;
	if (staged) {
		if (t == 0 && bytes) {
			bulk_s2g(dst, mine, bytes);  // one contiguous write of the whole block (TMA 1-D)
			bulk_commit();
		}
	} else {
		// wide block: FFOR from the tile straight into the column (ALP_RD exceptions concern the left parts only)
		pack_rows(tile, rd ? 0u : a.myexc, (UT)a.fill, (UT)a.base, a.bw, t, dst);
		if (rd) { pack_left(a.left_nib, a.e, t, reinterpret_cast<uint16_t*>(dst + 128u * a.bw), PT()); }
	}
	// ---- exceptions, in position order ----
	UT*       ev = static_cast<UT*>(col.exc_val) + exc_off;
	uint16_t* ep = col.exc_pos + exc_off;
	if (captured) {
#pragma unroll
		for (int k = 0; k < EXC_REGS; k++) {
			const uint32_t i = (uint32_t)t + 32u * k;
			if (i < a.cnt) {
				ev[i] = ev_reg[k];
				ep[i] = (uint16_t)ep_reg[k];
			}
		}
	} else if (listed) {
		for (uint32_t i = t; i < a.cnt; i += 32) {
			const uint32_t p = pos_list[i];
			ev[i]            = exc_value(p);
			ep[i]            = (uint16_t)p;
		}
	} else {
		for_each_exception<PT>(plan, a.myexc, t, [&](uint32_t rank, uint32_t p) {
			ev[rank] = exc_value(p);
			ep[rank] = (uint16_t)p;
		});
	}
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
		uint4* out = reinterpret_cast<uint4*>(col.meta + v);
		out[0]     = ra;
		out[1]     = rb;
		if (staged && bytes) { bulk_wait_all(); }  // the block image must outlive the bulk store
	}
	if (scanner) {
		__syncwarp();
		run_scanner(aggregates, prefixes, gridDim.x, t, workspace[1]);
	}
}

}  // namespace alpb200
