// alp_prims.cuh — single-vector kernels behind the alpb200_prim_* entry points: the reference's primitive API
// (PRIMITIVES.md) one vector at a time, built from the same device functions as the batched kernels.
// One warp per launch; these exist for drop-in use and parity tests, not for speed.
#pragma once

#include "alp_decode.cuh"
#include "alp_encode.cuh"

namespace alpb200 {

// alp::encoder<PT>::encode (encoder.hpp:402-418): encoded integers + exceptions + positions + count + chosen (e,f)
template <typename PT>
__global__ void __launch_bounds__(32) prim_encode_kernel(const PT* __restrict__ in, const alpb200_rg_state* __restrict__ state,
                                                         typename Traits<PT>::UT* __restrict__ exc, uint16_t* __restrict__ pos,
                                                         uint16_t* __restrict__ cnt, typename Traits<PT>::UT* __restrict__ enc,
                                                         uint8_t* __restrict__ ef) {
	using UT = typename Traits<PT>::UT;
	__shared__ __align__(128) UT tile[VEC];
	const int    t  = threadIdx.x;
	StateRegs    st = load_state(state);
	Analysis<PT> a;
	for (int i = t; i < VEC; i += 32) {
		tile[i] = Traits<PT>::bits(in[i]);
	}
	__syncwarp();
	TileIO<PT> io(tile, t);
	analyze_alp<PT>(in, st, t, io, a);
	__syncwarp();
	for (int r = 0; r < 32; r++) {
		enc[Map<PT>::index(t, r)] = ((a.myexc >> r) & 1u) ? (UT)a.fill : tile[Map<PT>::index(t, r)];  // encoder.hpp:393
	}
	emit_exceptions<PT>(
	    a.myexc, t, [&](uint32_t p) -> UT { return Traits<PT>::bits(in[p]); },
	    [&](uint32_t rank, uint32_t p, UT val) {
		    exc[rank] = val;
		    pos[rank] = (uint16_t)p;
	    });
	if (t == 0) {
		cnt[0] = (uint16_t)a.cnt;
		ef[0]  = (uint8_t)a.e;
		ef[1]  = (uint8_t)a.f;
	}
}

// alp::encoder<PT>::analyze_ffor (encoder.hpp:109-120)
template <typename PT>
__global__ void __launch_bounds__(32) prim_analyze_ffor_kernel(const typename Traits<PT>::ST* __restrict__ enc, uint8_t* __restrict__ bw,
                                                               typename Traits<PT>::ST* __restrict__ base) {
	using ST    = typename Traits<PT>::ST;
	const int t = threadIdx.x;
	ST        mn = Traits<PT>::ST_MAX, mx = Traits<PT>::ST_MIN;
	for (int i = t; i < VEC; i += 32) {
		const ST v = enc[i];
		mn         = v < mn ? v : mn;
		mx         = v > mx ? v : mx;
	}
	mn = warp_min<ST>(mn);
	mx = warp_max<ST>(mx);
	if (t == 0) {
		bw[0]   = (uint8_t)bits_of_range<PT>(mx, mn);
		base[0] = mn;
	}
}

// ffor::ffor on 64- and 32-bit lanes (include/fastlanes/ffor.hpp:7-8)
template <typename PT>
__global__ void __launch_bounds__(32) prim_ffor_kernel(const typename Traits<PT>::UT* __restrict__ in, uint8_t* __restrict__ out,
                                                       uint32_t bw, typename Traits<PT>::UT base) {
	using UT = typename Traits<PT>::UT;
	__shared__ __align__(128) UT tile[VEC];
	const int t = threadIdx.x;
	for (int i = t; i < VEC; i += 32) {
		tile[i] = in[i];
	}
	__syncwarp();
	pack_rows(tile, 0u, (UT)0, base, bw, t, out);
}

// ffor / unffor on 16-bit lanes (ffor.hpp:9): 64 lanes x 16 rows, value v = 64*row + lane.  Plain loops.
__global__ void __launch_bounds__(32) prim_ffor16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, uint32_t bw,
                                                         uint16_t base) {
	const int t = threadIdx.x;
	if (bw == 0) { return; }
	const uint32_t mask = bw >= 16 ? 0xFFFFu : ((1u << bw) - 1);
	for (int lane = t; lane < 64; lane += 32) {
		uint64_t acc = 0;
		uint32_t nb = 0, w = 0;
		for (int row = 0; row < 16; row++) {
			const uint32_t d = ((uint32_t)(uint16_t)(in[64 * row + lane] - base)) & mask;
			acc |= (uint64_t)d << nb;
			nb += bw;
			while (nb >= 16) {
				out[64 * w + lane] = (uint16_t)acc;
				acc >>= 16;
				nb -= 16;
				w++;
			}
		}
	}
}
__global__ void __launch_bounds__(32) prim_unffor16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, uint32_t bw,
                                                           uint16_t base) {
	__shared__ __align__(128) uint16_t blk[17 * 64];
	const int t = threadIdx.x;
	for (uint32_t i = t; i < 17 * 64; i += 32) {
		blk[i] = i < bw * 64u ? in[i] : (uint16_t)0;
	}
	__syncwarp();
	const uint32_t mask = bw >= 16 ? 0xFFFFu : ((1u << bw) - 1);
	for (int v = t; v < VEC; v += 32) {
		const uint32_t d = bw == 0 ? 0u : extract16(blk, v & 63, (v >> 6) * bw, mask);
		out[v]           = (uint16_t)(d + base);
	}
}

// ffor / unffor on 8-bit lanes (ffor.hpp:10, unffor.hpp:10; src/fastlanes_generated_ffor.cpp:4-300): 128 lanes x 8 rows, value
// v = 128*row + lane, stream word w of a lane at byte 128*w + lane.  A lane's whole stream is 8*bw <= 64 bits: one register.
__global__ void __launch_bounds__(32) prim_ffor8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t bw, uint8_t base) {
	const int t = threadIdx.x;
	if (bw == 0) { return; }
	const uint32_t mask = (1u << bw) - 1;  // bw <= 8
	for (int lane = t; lane < 128; lane += 32) {
		uint64_t acc = 0;
		for (int row = 0; row < 8; row++) {
			acc |= (uint64_t)(((uint32_t)(uint8_t)(in[128 * row + lane] - base)) & mask) << (row * bw);
		}
		for (uint32_t w = 0; w < bw; w++) {
			out[128 * w + lane] = (uint8_t)(acc >> (8 * w));
		}
	}
}
__global__ void __launch_bounds__(32) prim_unffor8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t bw, uint8_t base) {
	const int      t    = threadIdx.x;
	const uint32_t mask = (1u << bw) - 1;
	for (int lane = t; lane < 128; lane += 32) {
		uint64_t acc = 0;
		for (uint32_t w = 0; w < bw; w++) {
			acc |= (uint64_t)in[128 * w + lane] << (8 * w);
		}
		for (int row = 0; row < 8; row++) {
			out[128 * row + lane] = (uint8_t)(((uint32_t)(acc >> (row * bw)) & mask) + base);  // bw = 0: the base (unffor.cpp:4-22)
		}
	}
}

// unffor::unffor on 64- and 32-bit lanes (include/fastlanes/unffor.hpp:7-8)
template <typename PT>
__global__ void __launch_bounds__(32) prim_unffor_kernel(const uint8_t* __restrict__ in, typename Traits<PT>::UT* __restrict__ out,
                                                         uint32_t bw, typename Traits<PT>::UT base) {
	using UT = typename Traits<PT>::UT;
	__shared__ __align__(128) uint8_t blk[64 * 128 + STAGE_PAD];
	const int t = threadIdx.x;
	for (uint32_t i = t; i < bw * 8u; i += 32) {
		reinterpret_cast<uint4*>(blk)[i] = reinterpret_cast<const uint4*>(in)[i];
	}
	__syncwarp();
	const UT mask = low_mask<UT>(bw);
	for (int r = 0; r < 32; r++) {
		UT d = 0;
		if (bw != 0) {
			if (sizeof(PT) == 8) {
				d = (UT)extract64(reinterpret_cast<const uint64_t*>(blk), t & 15, (32u * (t >> 4) + r) * bw, (uint64_t)mask);
			} else {
				d = (UT)extract32(reinterpret_cast<const uint32_t*>(blk), t, r * bw, (uint32_t)mask);
			}
		}
		out[Map<PT>::index(t, r)] = d + base;
	}
}

// generated::falp::...::falp (include/alp/falp.hpp:10-44) with unfused semantics at every width
template <typename PT>
__global__ void __launch_bounds__(32) prim_falp_kernel(const uint8_t* __restrict__ in, PT* __restrict__ out, uint32_t bw,
                                                       typename Traits<PT>::UT base, uint32_t f, uint32_t e) {
	__shared__ __align__(128) uint8_t blk[64 * 128 + STAGE_PAD];
	const int t = threadIdx.x;
	for (uint32_t i = t; i < bw * 8u; i += 32) {
		reinterpret_cast<uint4*>(blk)[i] = reinterpret_cast<const uint4*>(in)[i];
	}
	__syncwarp();
	MetaRegs m;
	const uint64_t b64 = (uint64_t)base;
	m.a                = make_uint4((uint32_t)b64, (uint32_t)(b64 >> 32), 0, 0);
	m.b                = make_uint4(0, 0, ((uint32_t)ALPB200_SCHEME_ALP << 16) | (bw << 24), e | (f << 8));
	decode_alp_vector(blk, m, out, t);
}

// alp::decoder<PT>::decode (decoder.hpp:134-138)
template <typename PT>
__global__ void __launch_bounds__(32) prim_decode_kernel(const typename Traits<PT>::ST* __restrict__ enc, uint32_t f, uint32_t e,
                                                         PT* __restrict__ out) {
	using T = Traits<PT>;
	for (int i = threadIdx.x; i < VEC; i += 32) {
		out[i] = decode_value<PT>(enc[i], T::fact10(f), T::frac10(e));
	}
}

// alp::decoder<PT>::patch_exceptions (decoder.hpp:141-149)
template <typename UT>
__global__ void __launch_bounds__(32) prim_patch_kernel(UT* __restrict__ out, const UT* __restrict__ exc, const uint16_t* __restrict__ pos,
                                                        uint32_t cnt) {
	for (uint32_t i = threadIdx.x; i < cnt; i += 32) {
		out[pos[i]] = exc[i];
	}
}

// alp::rd_encoder<PT>::encode (rd.hpp:109-147): right parts, dictionary indices (unmasked), left-part exceptions
template <typename PT>
__global__ void __launch_bounds__(32) prim_rd_encode_kernel(const PT* __restrict__ in, const alpb200_rg_state* __restrict__ state,
                                                            uint16_t* __restrict__ exc, uint16_t* __restrict__ pos,
                                                            uint16_t* __restrict__ cnt, typename Traits<PT>::UT* __restrict__ right,
                                                            uint16_t* __restrict__ left) {
	using UT = typename Traits<PT>::UT;
	__shared__ __align__(128) UT tile[VEC];
	const int    t  = threadIdx.x;
	StateRegs    st = load_state(state);
	Analysis<PT> a;
	for (int i = t; i < VEC; i += 32) {
		tile[i] = Traits<PT>::bits(in[i]);
	}
	__syncwarp();
	TileIO<PT> io(tile, t);
	analyze_rd<PT>(state, st, t, io, a, [&](int r, uint32_t idx) { left[Map<PT>::index(t, r)] = (uint16_t)idx; });
	__syncwarp();
	for (int i = t; i < VEC; i += 32) {
		right[i] = tile[i];
	}
	const uint32_t rbw = a.bw;
	emit_exceptions<PT>(
	    a.myexc, t, [&](uint32_t p) -> UT { return (UT)(Traits<PT>::bits(in[p]) >> rbw); },
	    [&](uint32_t rank, uint32_t p, UT val) {
		    exc[rank] = (uint16_t)val;
		    pos[rank] = (uint16_t)p;
	    });
	if (t == 0) { cnt[0] = (uint16_t)a.cnt; }
}

// alp::rd_encoder<PT>::decode (rd.hpp:152-178) on unpacked right parts / dictionary indices
template <typename PT>
__global__ void __launch_bounds__(32) prim_rd_decode_kernel(typename Traits<PT>::UT* __restrict__ out,
                                                            const typename Traits<PT>::UT* __restrict__ right,
                                                            const uint16_t* __restrict__ left, const uint16_t* __restrict__ exc,
                                                            const uint16_t* __restrict__ pos, uint32_t cnt,
                                                            const alpb200_rg_state* __restrict__ state) {
	using UT    = typename Traits<PT>::UT;
	const int t = threadIdx.x;
	StateRegs st = load_state(state);
	const uint32_t rbw = st.right_bw();
	for (int i = t; i < VEC; i += 32) {
		out[i] = ((UT)dict_lookup(st.dict, left[i] & 7u) << rbw) | right[i];
	}
	__syncwarp();
	for (uint32_t i = t; i < cnt; i += 32) {
		out[pos[i]] = ((UT)exc[i] << rbw) | right[pos[i]];
	}
}

// ---- tail vector and NULLs (PRIMITIVES.md:141-144 "Last Vector Encoding") -----------------------------
// The codec works on whole vectors of 1024 REAL values.  Slots that hold no data — NULLs, and the slots behind the
// column's last value up to the next multiple of 1024 — are given a filler before the encoder sees them, so that they are
// ordinary values to everybody downstream (and to the reference, fed the same buffer: byte-identical output).
// One warp per vector; a vector without such slots costs one 128-byte read of the validity bitmap and nothing else.
//   states == nullptr  first strategy:  the filler is the vector's first valid value (0 if it has none)
//   states given       second strategy: the filler is the vector's first valid NON-EXCEPTION value under the (e,f) the
//                      encoder will pick for it (ALP), or the first valid value whose left part is in the dictionary
//                      (ALP_RD) — the slot then costs no exception and cannot widen the bit width.  Run after the first
//                      strategy + row-group init (the sampling must already see real values everywhere).
template <typename PT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) fill_invalid_kernel(PT* __restrict__ values, uint64_t n_values, const uint32_t* __restrict__ validity,
                                                                  const alpb200_rg_state* __restrict__ states, uint64_t n_vectors) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	using ST = typename T::ST;
	const int      t = threadIdx.x & 31;
	const uint64_t v = (uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
	if (v >= n_vectors) { return; }
	PT* vec = values + v * (uint64_t)VEC;
	// bit r of `ok`: this lane's slot in row r (position 32 r + t) holds a real value
	uint32_t ok = 0;
	for (int r = 0; r < 32; r++) {
		const uint64_t i     = v * VEC + 32u * r + t;
		bool           valid = i < n_values;
		if (valid && validity != nullptr) { valid = (__ldg(validity + 32 * v + r) >> t) & 1u; }  // one word per row: broadcast
		ok |= (uint32_t)valid << r;
	}
	if (__all_sync(FULL, ok == 0xFFFFFFFFu)) { return; }  // nothing to fill
	// first position (in value order) that qualifies as the filler's source
	auto first_position = [&](uint32_t mask) -> uint32_t {  // mask: bit r = this lane's slot of row r qualifies; 1024 = none
		uint32_t rows = 0;  // lane r: ballot of row r
		for (int r = 0; r < 32; r++) {
			const uint32_t b = __ballot_sync(FULL, (mask >> r) & 1u);
			if (t == r) { rows = b; }
		}
		const uint32_t have = __ballot_sync(FULL, rows != 0);
		if (have == 0) { return 1024u; }
		const int r0 = __ffs((int)have) - 1;
		return 32u * r0 + (uint32_t)(__ffs((int)__shfl_sync(FULL, rows, r0)) - 1);
	};
	uint32_t src = 1024;
	if (states == nullptr) {
		src = first_position(ok);
	} else {
		StateRegs st = load_state(states + v / ALPB200_ROWGROUP_VECTORS);
		uint32_t  good = 0;  // valid and not an exception
		if (st.scheme == ALPB200_SCHEME_ALP_RD) {
			const uint32_t rbw = st.right_bw(), ds = st.dict_size();
			for (int r = 0; r < 32; r++) {
				const uint32_t left = (uint32_t)(T::bits(vec[32 * r + t]) >> rbw);
				bool           hit  = false;
				for (uint32_t d = 0; d < ds; d++) {
					hit = hit || dict_lookup(st.dict, d) == left;
				}
				good |= (uint32_t)(hit && ((ok >> r) & 1u)) << r;
			}
		} else {
			int e = st.exp_of(0), f = st.fac_of(0);
			if (st.k > 1) { choose_exponent_factor<PT>(vec[32 * t], st, e, f); }  // encoder.hpp:409-412 (samples 0, 32, ..., 992)
			const PT ex = T::exp10(e), frf = T::frac10(f), fre = T::frac10(e);
			const ST fa = T::fact10(f);
			for (int r = 0; r < 32; r++) {
				const PT   x   = vec[32 * r + t];
				const ST   enc = encode_value<PT, false>(x, ex, frf);
				const bool exc = T::bits(decode_value<PT>(enc, fa, fre)) != T::bits(x);  // (specials never round-trip: alp_encode.cuh)
				good |= (uint32_t)(!exc && ((ok >> r) & 1u)) << r;
			}
		}
		src = first_position(good);
		if (src == 1024u) { return; }  // no such value: the first strategy's filler stays
	}
	const PT filler = src < 1024u ? vec[src] : (PT)0;
	__syncwarp();
	for (int r = 0; r < 32; r++) {
		if (!((ok >> r) & 1u)) { vec[32 * r + t] = filler; }
	}
}

}  // namespace alpb200
