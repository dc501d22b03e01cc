// alp_init.cuh — row-group initialisation on the device.
//
// Replaces alp::encoder<PT>::init (include/alp/encoder.hpp:420-427) = sampler::first_level_sample
// (include/alp/sampler.hpp:14-52) + find_top_k_combinations (encoder.hpp:139-235), and for row-groups that fall to
// ALP_RD alp::rd_encoder<PT>::init = find_best_dictionary (include/alp/rd.hpp:89-104,180-185).
//
// Two kernels:
//   init_search_kernel    one warp per (row-group, sampled vector): lanes own (e,f) pairs, loop over the 32 samples
//                         (broadcast from shared memory) -> best pair and its estimated size for that vector
//   init_finalize_kernel  one warp per row-group: histogram of the <=9 winners -> top-5 list, ALP/ALP_RD decision and,
//                         for ALP_RD, the cut position + left-part dictionary from the <=288 samples
//
// n_values is a multiple of 1024 here, so every sampled vector is complete and yields exactly 32 samples
// (values 0,32,...,992 of vectors 0,12,24,... of the row-group).
#pragma once

#include "alp_device.cuh"

namespace alpb200 {

constexpr int SAMPLE_JUMP        = 12;  // config.hpp:17-19: (102400 / 8) / 1024
constexpr int MAX_SAMPLED_VECS   = 9;   // vectors 0,12,...,96

struct SearchResult {
	uint32_t e, f, size, pad;
};

__host__ __device__ constexpr int sampled_vectors(uint64_t rg_vectors) { return (int)((rg_vectors + SAMPLE_JUMP - 1) / SAMPLE_JUMP); }

// enumeration order of encoder.hpp:157-158: e = MAX..0, f = e..0; combo c -> (e,f)
template <int MAX_EXP>
__device__ __forceinline__ void combo_of(int c, int& e, int& f) {
	// row e (counting down from MAX_EXP) has e+1 entries
	int ee = MAX_EXP, start = 0;
	while (c >= start + ee + 1) {
		start += ee + 1;
		ee--;
	}
	e = ee;
	f = ee - (c - start);
}

template <typename PT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) init_search_kernel(const PT* __restrict__ in, uint64_t n_vectors, uint64_t n_rowgroups,
                                                                 SearchResult* __restrict__ results) {
	using T  = Traits<PT>;
	using ST = typename T::ST;
	using FL = FastLimits<PT>;
	__shared__ PT s_smp[WARPS][32];
	const int      warp = threadIdx.x >> 5, t = threadIdx.x & 31;
	const uint64_t job  = (uint64_t)blockIdx.x * WARPS + warp;
	if (job >= n_rowgroups * MAX_SAMPLED_VECS) { return; }
	const uint64_t rg   = job / MAX_SAMPLED_VECS;
	const int      slot = (int)(job % MAX_SAMPLED_VECS);
	const uint64_t rgv  = min((uint64_t)ALPB200_ROWGROUP_VECTORS, n_vectors - rg * ALPB200_ROWGROUP_VECTORS);
	if (slot >= sampled_vectors(rgv)) { return; }
	const PT* vec  = in + (rg * ALPB200_ROWGROUP_VECTORS + (uint64_t)slot * SAMPLE_JUMP) * VEC;
	s_smp[warp][t] = vec[32 * t];
	__syncwarp();

	constexpr uint32_t WORST = 32u * (T::EXC_BITS + 16) + 32u * T::EXC_BITS;  // encoder.hpp:151-153
	uint32_t           best  = 0xFFFFFFFFu;                                  // (size << 8) | combo index
	for (int c = t; c < T::N_COMBOS; c += 32) {
		int e, f;
		combo_of<T::MAX_EXP>(c, e, f);
		const PT ex = T::exp10(e), frf = T::frac10(f), fre = T::frac10(e);
		const ST fa  = T::fact10(f);
		const PT fap = FL::fact_fp(f);
		uint32_t n_ok = 0;
		ST       mx = T::ST_MIN, mn = T::ST_MAX;                   // exact path (rare)
		PT       fmx = -T::upper_limit(), fmn = T::upper_limit();  // the encoded integers as exactly representable PT values
		// encode_value<SAFE> turns every "impossible" scaled value (not finite, beyond +-(2^63 - 1024), or -0.0) into one
		// sentinel integer: what that decodes to is a constant of the (e,f) pair.  Lanes hold different pairs, and for the
		// large exponents most samples ARE impossible, so this case must not cost a divergent branch.
		const PT sent_fp  = (PT)T::SAFE_SENTINEL;
		const PT dec_sent = decode_value<PT>(T::SAFE_SENTINEL, fa, fre);
#pragma unroll 4
		for (int i = 0; i < 32; i++) {
			const PT x = s_smp[warp][i];
			// Same reasoning as analyze_rows<FAST> in alp_encode.cuh: while the scaled value is encodable and its product
			// with 10^f stays below 2^63 | 2^31, (PT)(enc * FACT[f]) == t_r * 10^f; a product beyond that never decodes
			// back; only a product of exactly that magnitude takes the literal recipe (a branch that is almost never taken).
			const PT       t        = T::mul(T::mul(x, ex), frf);
			const bool     possible = fabs((double)t) <= 9223372036854774784.0 && T::bits(t) != T::bits((PT)-0.0);
			const PT       tr       = T::magic_round(t);
			const PT       pd       = T::mul(tr, fap);
			const uint32_t key      = FL::key(pd);
			if (possible && key == FL::BIG) {
				const ST enc = encode_value<PT, true>(x, ex, frf);
				const PT dec = decode_value<PT>(enc, fa, fre);
				if (dec == x) {
					n_ok++;
					mx = enc > mx ? enc : mx;
					mn = enc < mn ? enc : mn;
				}
			} else {
				const PT   dec = possible ? T::mul(pd, fre) : dec_sent;
				const bool ok  = dec == x && (!possible || key < FL::BIG);
				const PT   v   = possible ? tr : sent_fp;
				if (ok) {
					n_ok++;
					fmx = v > fmx ? v : fmx;  // (no NaNs here: plain compares, fmax() would be an 8-instruction emulation)
					fmn = v < fmn ? v : fmn;
				}
			}
		}
		if (n_ok < 2) { continue; }  // encoder.hpp:183
		if (fmx >= fmn) {  // exact in PT: |t_r| < 2^63 | 2^31; the sentinel converts back to itself (float: by saturation)
			const ST imx = FL::cast_sat(fmx), imn = FL::cast_sat(fmn);
			mx           = imx > mx ? imx : mx;
			mn           = imn < mn ? imn : mn;
		}
		const uint32_t size = 32u * bits_of_range<PT>(mx, mn) + (32u - n_ok) * (T::EXC_BITS + 16);
		best                = min(best, (size << 8) | (uint32_t)c);
	}
	// Smallest size wins; among equal sizes the earliest pair in enumeration order (largest e, then largest f), which is
	// what the update rule of encoder.hpp:191-199 amounts to.  A valid pair's size is always below WORST.
	best = __reduce_min_sync(FULL, best);
	if (t == 0) {
		SearchResult r;
		if (best == 0xFFFFFFFFu) {
			r.e = r.f = 0;
			r.size    = WORST;
		} else {
			int e, f;
			combo_of<T::MAX_EXP>((int)(best & 0xFF), e, f);
			r.e    = (uint32_t)e;
			r.f    = (uint32_t)f;
			r.size = best >> 8;
		}
		r.pad                                 = 0;
		results[rg * MAX_SAMPLED_VECS + slot] = r;
	}
}

// ---- ALP_RD: rd.hpp:33-104 on one warp -------------------------------------------------------------------------------
// Entries of equal frequency are ordered by smaller left part first (the reference leaves this to the STL; see
// oracle/alp_oracle_impl.inc rd_build_dict).  packed key = count << 16 | (0xFFFF - left): larger is better.
//
// The 16 candidate cuts keep the top 1..16 bits of a value: the left part at cut i is a PREFIX of the left part at cut
// i+1.  So the samples' top 16 bits are sorted ONCE (rank by counting, 288^2 / 32 compares per lane); in that order the equal
// left parts of every cut are runs of neighbours, and a cut costs one pass over 9 neighbours per lane (run starts and
// lengths) plus 8 rounds of REDUX.MAX for the dictionary — instead of the O(n^2) occurrence count per cut of round 1
// (2.9 ms for the 5243 row-groups of 2^29 high-precision doubles; the whole ALP search takes 0.42 ms).
constexpr int RD_PER_LANE = MAX_SAMPLED_VECS;  // 288 samples = 32 lanes x 9 consecutive positions of the sorted order

template <typename UT>
__device__ void rd_find_best_dictionary(const UT* s_bits, uint16_t* s_keys, int n, int t, alpb200_rg_state* out) {
	constexpr int TBITS = 8 * sizeof(UT);
	// ---- sort the top 16 bits of the samples (ascending; equal keys in sample order) ----
	{
		uint32_t key[RD_PER_LANE], rank[RD_PER_LANE];
#pragma unroll
		for (int q = 0; q < RD_PER_LANE; q++) {
			const int j = t + 32 * q;
			key[q]      = j < n ? (uint32_t)(s_bits[j] >> (TBITS - 16)) : 0xFFFFFFFFu;
			rank[q]     = 0;
		}
		for (int m = 0; m < n; m++) {
			const uint32_t km = (uint32_t)(s_bits[m] >> (TBITS - 16));  // (one address for the whole warp: broadcast)
#pragma unroll
			for (int q = 0; q < RD_PER_LANE; q++) {
				rank[q] += (km < key[q]) || (km == key[q] && m < t + 32 * q);
			}
		}
#pragma unroll
		for (int q = 0; q < RD_PER_LANE; q++) {
			if (t + 32 * q < n) { s_keys[rank[q]] = (uint16_t)key[q]; }
		}
		__syncwarp();
	}
	// this lane's stretch of the sorted order, plus the key before it
	uint32_t K[RD_PER_LANE];
#pragma unroll
	for (int k = 0; k < RD_PER_LANE; k++) {
		const int p = RD_PER_LANE * t + k;
		K[k]        = p < n ? s_keys[p] : 0u;
	}
	const uint32_t K_before = t > 0 && RD_PER_LANE * t - 1 < n ? s_keys[RD_PER_LANE * t - 1] : 0u;

	uint32_t best_rbw = 0;
	double   best_est = 1.7976931348623157e308;
	for (int pass = 0; pass < 2; pass++) {
		const int i0 = pass == 0 ? 1 : (int)(TBITS - best_rbw), i1 = pass == 0 ? 16 : i0;  // config.hpp:23 CUTTING_LIMIT
		for (int i = i0; i <= i1; i++) {
			const uint32_t rbw = TBITS - i, sh = 16 - i;
			// run starts in this lane's stretch; first_start = position of the lane's first one (n if none)
			uint32_t starts = 0;
#pragma unroll
			for (int k = 0; k < RD_PER_LANE; k++) {
				const int      p    = RD_PER_LANE * t + k;
				const uint32_t prev = k == 0 ? K_before : K[k - 1];
				if (p < n && (p == 0 || (K[k] >> sh) != (prev >> sh))) { starts |= 1u << k; }
			}
			uint32_t next = starts ? (uint32_t)(RD_PER_LANE * t + __ffs((int)starts) - 1) : (uint32_t)n;  // becomes: first start AFTER this lane
			{
				uint32_t m = next;  // suffix minimum over the lanes behind this one (exclusive)
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t o = __shfl_down_sync(FULL, m, d);
					if (t + d < 32) { m = min(m, o); }
				}
				next = __shfl_down_sync(FULL, m, 1);
				if (t == 31) { next = (uint32_t)n; }
			}
			// candidates: one per run that starts here, packed (count << 16 | 0xFFFF - left part)
			uint32_t cand[RD_PER_LANE];
#pragma unroll
			for (int k = RD_PER_LANE - 1; k >= 0; k--) {
				const uint32_t p = (uint32_t)(RD_PER_LANE * t + k);
				cand[k]          = 0;
				if ((starts >> k) & 1u) {
					cand[k] = ((next - p) << 16) | (0xFFFFu - (K[k] >> sh));
					next    = p;
				}
			}
			// selection in rank order: rank 0..7 -> dictionary; rank dict_size is skipped and ranks above it are
			// remembered with their rank as index (rd.hpp:63-77)
			uint32_t in_dict = 0, ds = 0, rank = 0;
			for (;; rank++) {
				uint32_t mine = 0;
#pragma unroll
				for (int k = 0; k < RD_PER_LANE; k++) {
					mine = max(mine, cand[k]);
				}
				const uint32_t top = __reduce_max_sync(FULL, mine);
				if (top == 0) { break; }
#pragma unroll
				for (int k = 0; k < RD_PER_LANE; k++) {
					if (cand[k] == top) { cand[k] = 0; }  // (packed keys are distinct: exactly one lane, one slot)
				}
				const uint32_t key = 0xFFFFu - (top & 0xFFFFu);
				if (rank < ALPB200_RD_DICT_SIZE) {
					in_dict += top >> 16;
					ds = rank + 1;
					if (pass == 1 && t == 0) { out->dict[rank] = (uint16_t)key; }
				} else if (pass == 0) {
					break;  // the estimate only needs the dictionary part
				} else if (rank > ds && t == 0) {
					const uint32_t x  = out->n_extra;
					out->extra_key[x] = (uint16_t)key;
					out->extra_idx[x] = (uint16_t)rank;
					out->n_extra      = (uint16_t)(x + 1);
				}
			}
			uint32_t lbw = 1;  // max(1, ceil(log2(dict_size))), rd.hpp:61
			while ((1u << lbw) < ds) {
				lbw++;
			}
			if (pass == 0) {
				const double exc_bits = (double)(n - (int)in_dict) * 32.0;  // rd.hpp:26
				const double est      = (double)rbw + (double)lbw + exc_bits / (double)n;
				if (est < best_est) {
					best_est = est;
					best_rbw = rbw;
				}
			} else if (t == 0) {
				out->right_bw  = (uint8_t)rbw;
				out->left_bw   = (uint8_t)lbw;
				out->dict_size = (uint8_t)ds;
			}
		}
	}
}

template <typename PT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) init_finalize_kernel(const PT* __restrict__ in, uint64_t n_vectors, uint64_t n_rowgroups,
                                                                   const SearchResult* __restrict__ results,
                                                                   alpb200_rg_state* __restrict__ states, uint32_t force_rd) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	__shared__ UT       s_bits[WARPS][ALPB200_MAX_SAMPLES];
	__shared__ uint16_t s_keys[WARPS][ALPB200_MAX_SAMPLES];
	const int      warp = threadIdx.x >> 5, t = threadIdx.x & 31;
	const uint64_t rg   = (uint64_t)blockIdx.x * WARPS + warp;
	if (rg >= n_rowgroups) { return; }
	const uint64_t rgv = min((uint64_t)ALPB200_ROWGROUP_VECTORS, n_vectors - rg * ALPB200_ROWGROUP_VECTORS);
	const int      nsv = sampled_vectors(rgv);
	alpb200_rg_state* out = states + rg;

	// zero the record (1196 bytes = 299 words)
	uint32_t* w = reinterpret_cast<uint32_t*>(out);
	for (int i = t; i < (int)(sizeof(alpb200_rg_state) / 4); i += 32) {
		w[i] = 0;
	}
	__syncwarp();

	// lane j holds the winner of sampled vector j
	SearchResult r;
	r.e = r.f = 0;
	r.size    = 0xFFFFFFFFu;
	if (t < nsv) { r = results[rg * MAX_SAMPLED_VECS + t]; }
	const uint32_t best_overall = __reduce_min_sync(FULL, r.size);
	if (best_overall >= T::RD_LIMIT || force_rd) {  // encoder.hpp:213-216; forced: rd_encoder<PT>::init on any row-group (rd.hpp:180-185)
		const int n = 32 * nsv;
		for (int j = t; j < n; j += 32) {
			const PT* vec   = in + (rg * ALPB200_ROWGROUP_VECTORS + (uint64_t)(j >> 5) * SAMPLE_JUMP) * VEC;
			s_bits[warp][j] = T::bits(vec[32 * (j & 31)]);
		}
		__syncwarp();
		if (t == 0) { out->scheme = ALPB200_SCHEME_ALP_RD; }
		rd_find_best_dictionary<UT>(s_bits[warp], s_keys[warp], n, t, out);
		return;
	}
	// histogram of winners, ranked by (occurrences desc, e desc, f desc): encoder.hpp:126-131,218-234
	const uint32_t mine = t < nsv ? ((r.e << 8) | r.f) : 0xFFFFFFFFu;
	uint32_t       occ  = 0;
	bool           first = t < nsv;
	for (int j = 0; j < nsv; j++) {
		const uint32_t other = __shfl_sync(FULL, mine, j);
		if (t < nsv && other == mine) {
			occ++;
			if (j < t) { first = false; }
		}
	}
	uint32_t key = first ? ((occ << 16) | mine) : 0;  // unique per distinct pair; larger is better
	uint32_t k   = 0;
	for (; k < ALPB200_MAX_K; k++) {
		const uint32_t top = __reduce_max_sync(FULL, key);
		if (top == 0) { break; }
		if (key == top) { key = 0; }
		if (t == 0) {
			out->combos[k][0] = (uint8_t)((top >> 8) & 0xFF);
			out->combos[k][1] = (uint8_t)(top & 0xFF);
		}
	}
	if (t == 0) {
		out->scheme = ALPB200_SCHEME_ALP;
		out->k      = (int32_t)k;
	}
}

// ---- synthetic columns: stateless splitmix64 per index ---------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t seed, uint64_t i) {
	uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
	z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z          = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

__global__ void generate_f64_kernel(double* __restrict__ out, uint64_t n, uint64_t first, uint64_t seed, int kind) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
		const uint64_t i = first + j;
		const uint64_t r = splitmix64(seed, i);
		double         x;
		if (kind == 3) {  // latitude-like, full 53-bit mantissas
			const double u = __dmul_rn((double)(r >> 11), 0x1.0p-53);
			x              = __dsub_rn(__dmul_rn(u, 180.0), 90.0);
		} else {  // <= 3 decimals, the number of decimals constant per row-group
			const uint32_t d   = (uint32_t)((i / ALPB200_ROWGROUP_SIZE) % 4);
			const double   div = d == 0 ? 1.0 : (d == 1 ? 10.0 : (d == 2 ? 100.0 : 1000.0));
			x                  = __ddiv_rn((double)(r % 1000000ULL), div);
		}
		out[j] = x;
	}
}

__global__ void generate_f32_kernel(float* __restrict__ out, uint64_t n, uint64_t first, uint64_t seed, int kind) {
	(void)kind;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
		const uint64_t r = splitmix64(seed, first + j);
		float          x;
		if (r % 100 >= 5) {
			x = __fdiv_rn((float)((r >> 8) % 10000ULL), 100.0f);
		} else {
			uint32_t b = (uint32_t)(r >> 32);
			b          = (b & 0x807FFFFFu) | ((20u + ((b >> 23) % 200u)) << 23);
			x          = __uint_as_float(b);
		}
		out[j] = x;
	}
}

}  // namespace alpb200
