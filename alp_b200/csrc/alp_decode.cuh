// alp_decode.cuh — fused UNFFOR + ALP decode + exception patch, and ALP_RD decode, one warp per 1024-value vector.
//
// Replaces generated::falp::fallback::scalar::falp (src/falp.cpp:42440,42644; per-value recipe :1049-1060) +
// alp::decoder<PT>::patch_exceptions (include/alp/decoder.hpp:141-149), and for ALP_RD vectors
// 2x unffor::unffor + alp::rd_encoder<PT>::decode (include/alp/rd.hpp:152-178).
//
// Data movement per vector (DESIGN.md "decode kernel"):
//   * one 32-byte metadata record, read by every lane from the same address (one sector, broadcast)
//   * the packed block (128*bw bytes, 128-byte aligned) is fetched by ONE bulk-async copy (TMA 1-D,
//     cp.async.bulk ... mbarrier::complete_tx) into this warp's shared-memory stage, double buffered: the copy
//     for vector i+1 is in flight while vector i is unpacked
//   * unpack straight out of the verbatim block image in shared memory: 64-bit lanes -> thread (lane = t&15,
//     half = t>>4) owns rows 32*half..32*half+31 of its lane; 32-bit lanes -> thread t owns lane t
//   * every warp store instruction writes full 128-byte lines: out[16*row + lane] for 16 lanes x 2 halves
//     (f64: two lines) or out[32*row + lane] (f32: one line)
//   * exceptions are patched by the same warp after a __syncwarp(), i.e. while the lines are still in L2
#pragma once

#include "alp_device.cuh"
#include "alp_ffor.cuh"

// build-time knobs (tools/decode_probe.py builds variants to compare them on the GPU)
#ifndef ALPB200_DEC_MINBLOCKS
#define ALPB200_DEC_MINBLOCKS 2  // resident CTAs per SM the register allocation is limited for
#endif
#ifndef ALPB200_DEC_MINBLOCKS_F32
#define ALPB200_DEC_MINBLOCKS_F32 2  // (three CTAs of 80 registers were tried for floats: spills, 0.304 vs 0.285 ms per 2^28 values of config 4)
#endif
#ifndef ALPB200_DEC_STREAMING_STORES
#define ALPB200_DEC_STREAMING_STORES 0  // 1: st.global.cs for the decoded values (written once, never re-read)
#endif

namespace alpb200 {

template <typename V>
__device__ __forceinline__ void store_out(V* p, V v) {
#if ALPB200_DEC_STREAMING_STORES
	__stcs(p, v);
#else
	*p = v;
#endif
}

struct ColView {
	const alpb200_vec_meta* meta;
	const uint8_t*          packed;
	const void*             exc_val;
	const uint16_t*         exc_pos;
};

// the 32-byte alpb200_vec_meta record in two 16-byte registers
struct MetaRegs {
	uint4 a;  // ALP: base (x,y) | reserved;  ALP_RD: rd_dict[8]
	uint4 b;  // packed_off | exc_off | exc_cnt,scheme,bw | e,f,reserved
	__device__ __forceinline__ uint32_t packed_off() const { return b.x; }
	__device__ __forceinline__ uint32_t exc_off() const { return b.y; }
	__device__ __forceinline__ uint32_t exc_cnt() const { return b.z & 0xFFFFu; }
	__device__ __forceinline__ uint32_t scheme() const { return (b.z >> 16) & 0xFFu; }
	__device__ __forceinline__ uint32_t bw() const { return b.z >> 24; }
	__device__ __forceinline__ uint32_t e() const { return b.w & 0xFFu; }
	__device__ __forceinline__ uint32_t f() const { return (b.w >> 8) & 0xFFu; }
	__device__ __forceinline__ uint64_t base() const { return ((uint64_t)a.y << 32) | a.x; }
	__device__ __forceinline__ uint32_t block_bytes() const {
		return 128u * (scheme() == ALPB200_SCHEME_ALP_RD ? bw() + e() : bw());
	}
};

__device__ __forceinline__ MetaRegs load_meta(const alpb200_vec_meta* m) {
	const uint4* p = reinterpret_cast<const uint4*>(m);
	MetaRegs     r;
	r.a = __ldg(p);
	r.b = __ldg(p + 1);
	return r;
}

// ---- slow generic path ---------------------------------------------------------------------------------------------------
// The hot loops below size each warp's shared-memory stage from the caller's hint column.max_block_bytes and bulk-copy
// every block into it unchecked.  A hint smaller than a real block (a stale value after re-encoding into the same
// container, a hand-filled struct) would overrun the stage — so whenever a hint is given, the launchers first run
// hint_check_kernel over the records of the call (one 32-byte record per thread, ~7 us per 2^20 vectors, which also pulls
// the records into L2); if any block outgrows the stage, every thread block of the main kernel takes decode_slow /
// sum_slow instead of its hot loop: straight from global memory, run-time widths, correct for any well-formed column.
// The hot loops themselves stay exactly as they were (the f32 instances lose 3-14 % to any per-vector check).
static __global__ void hint_check_kernel(const alpb200_vec_meta* __restrict__ meta, uint64_t n_vectors, uint32_t stage_cap,
                                         unsigned long long* __restrict__ oversize) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	bool           big    = false;
	for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vectors; v += stride) {
		const uint4    b     = __ldg(reinterpret_cast<const uint4*>(meta + v) + 1);
		const uint32_t bw    = b.z >> 24, e = b.w & 0xFFu;
		const uint32_t bytes = 128u * (((b.z >> 16) & 0xFFu) == ALPB200_SCHEME_ALP_RD ? bw + e : bw);
		big                  = big || bytes > stage_cap;
	}
	if (big) { *oversize = 1ull; }
}

// the bw-bit field at bit offset `bit` of lane `lane` of a T-bit-lane block in GLOBAL memory; never reads past the block
template <typename UT>
__device__ __forceinline__ UT field_at_global(const UT* blk, int lane, uint32_t bit, uint32_t bw) {
	constexpr uint32_t TB = 8 * sizeof(UT), L = 1024 / TB;
	if (bw == 0) { return 0; }
	const uint32_t w = bit / TB, sh = bit % TB;
	UT             v = (UT)(blk[L * w + lane] >> sh);
	if (sh + bw > TB) { v = (UT)(v | (UT)(blk[L * (w + 1) + lane] << (TB - sh))); }
	return (UT)(v & low_mask<UT>((int)bw));
}
// decoded bit pattern of value p of a vector, exceptions NOT applied
template <typename PT>
__device__ __forceinline__ typename Traits<PT>::UT value_bits_slow(const uint8_t* blk, const MetaRegs& m, uint32_t p) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	using ST = typename T::ST;
	const UT d = field_at_global<UT>(reinterpret_cast<const UT*>(blk), p % T::LANES, (p / T::LANES) * m.bw(), m.bw());
	if (m.scheme() == ALPB200_SCHEME_ALP) {
		const UT base = sizeof(PT) == 8 ? (UT)m.base() : (UT)m.a.x;
		return T::bits(decode_value<PT>((ST)(UT)(d + base), T::fact10(m.f()), T::frac10(m.e())));
	}
	const uint32_t idx = field_at_global<uint16_t>(reinterpret_cast<const uint16_t*>(blk + 128u * m.bw()), p & 63, (p >> 6) * m.e(), m.e());
	return (UT)(((UT)dict_lookup(m.a, idx) << m.bw()) | d);
}
// one warp decodes + patches vector v straight from global memory
template <typename PT>
__device__ __forceinline__ void decode_vector_slow(const ColView& col, const MetaRegs& m, PT* out_vec, int t) {
	using UT           = typename Traits<PT>::UT;
	const uint8_t* blk = col.packed + (uint64_t)m.packed_off() * 128u;
	UT*            ov  = reinterpret_cast<UT*>(out_vec);
#pragma unroll 1
	for (int i = t; i < VEC; i += 32) {
		ov[i] = value_bits_slow<PT>(blk, m, (uint32_t)i);
	}
	__syncwarp();
	const UT*       ev = static_cast<const UT*>(col.exc_val) + m.exc_off();
	const uint16_t* ep = col.exc_pos + m.exc_off();
	const bool      rd = m.scheme() != ALPB200_SCHEME_ALP;
#pragma unroll 1
	for (uint32_t i = t; i < m.exc_cnt(); i += 32) {
		const uint32_t p = ep[i];
		UT             v = ev[i];
		if (rd) {  // the true left part replaces the dictionary entry (rd.hpp:172-177)
			const UT right = field_at_global<UT>(reinterpret_cast<const UT*>(blk), p % Traits<PT>::LANES, (p / Traits<PT>::LANES) * m.bw(), m.bw());
			v              = (UT)(((v & 0xFFFFu) << m.bw()) | right);
		}
		ov[p] = v;
	}
	__syncwarp();
}

// ---- ALP, 64-bit lanes -------------------------------------------------------------------------------------------
// One `switch (bw)` per vector (warp-uniform) into a template instance whose shifts, masks and offsets are constants.
__device__ __forceinline__ void decode_alp_vector(const uint8_t* stage, const MetaRegs& m, double* __restrict__ out_vec, int t) {
	using T             = Traits<double>;
	const int      lane = t & 15, half = t >> 4;
	const int64_t  fact = T::fact10(m.f());
	const double   frac = T::frac10(m.e());
	const uint64_t base = m.base();
	double*        o    = out_vec + 512 * half + lane;  // value index 16*(32*half + r) + lane
	dispatch_width<0, 64>(m.bw(), [&](auto W) {
		constexpr int BW = decltype(W)::value;
		unpack64_rows<BW>(stage, lane, half, [&](int r, uint32_t lo, uint32_t hi) {
			const uint64_t d = BW <= 32 ? (uint64_t)lo : ((uint64_t)hi << 32) | lo;
			store_out(&o[16 * r], decode_value<double>((int64_t)(d + base), fact, frac));  // src/falp.cpp:1049-1056
		});
	});
}

// ---- ALP, 32-bit lanes -------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_alp_vector(const uint8_t* stage, const MetaRegs& m, float* __restrict__ out_vec, int t) {
	using T             = Traits<float>;
	const int32_t  fact = T::fact10(m.f());
	const float    frac = T::frac10(m.e());
	const uint32_t base = m.a.x;
	float*         o    = out_vec + t;  // value index 32*r + lane
	dispatch_width<0, 32>(m.bw(), [&](auto W) {
		constexpr int BW = decltype(W)::value;
		unpack32_rows<BW>(stage, t, [&](int r, uint32_t d) { store_out(&o[32 * r], decode_value<float>((int32_t)(d + base), fact, frac)); });
	});
}

// ---- exceptions ------------------------------------------------------------------------------------------------------
// The first 32 exceptions of a vector (one per lane) are fetched one vector ahead, so that the patch never waits for
// DRAM; longer runs fall back to a loop.
template <typename UT>
struct ExcRegs {
	uint32_t pos;
	UT       val;
};
template <typename UT>
__device__ __forceinline__ ExcRegs<UT> load_exceptions(const ColView& col, const MetaRegs& m, int t) {
	ExcRegs<UT> x;
	x.pos = 0;
	x.val = 0;
	if ((uint32_t)t < m.exc_cnt()) {
		x.pos = __ldg(col.exc_pos + m.exc_off() + t);
		x.val = __ldg(static_cast<const UT*>(col.exc_val) + m.exc_off() + t);
	}
	return x;
}
// Exceptions 32 .. 32 (1 + TAIL) - 1 of an ALP vector (lane t: ranks t + 32, t + 64, ...), also fetched one vector ahead.  Float
// columns are where exception-heavy vectors are the norm (6 bytes per exception: BASELINE config 4 has 90 per vector); with the
// tail in registers their patch is stores only.  Nothing is computed on the loaded words here (a use would make the warp wait).
template <typename PT>
struct DecCfg {
	static constexpr int EXC_TAIL  = sizeof(PT) == 4 ? 3 : 0;  // floats: 128 exceptions per vector in registers
	static constexpr int MINBLOCKS = sizeof(PT) == 4 ? ALPB200_DEC_MINBLOCKS_F32 : ALPB200_DEC_MINBLOCKS;
};
template <typename UT, int N>
struct ExcTail {
	UT       val[N > 0 ? N : 1];
	uint16_t pos[N > 0 ? N : 1];
};
template <typename UT, int N>
__device__ __forceinline__ ExcTail<UT, N> load_exception_tail(const ColView& col, const MetaRegs& m, int t) {
	ExcTail<UT, N> x;
	if constexpr (N > 0) {
		const uint32_t  cnt = m.exc_cnt();
		const UT*       ev  = static_cast<const UT*>(col.exc_val) + m.exc_off();
		const uint16_t* ep  = col.exc_pos + m.exc_off();
#pragma unroll
		for (int k = 0; k < N; k++) {
			const uint32_t i = (uint32_t)t + 32u * (k + 1);
			x.val[k]         = 0;
			x.pos[k]         = 0;
			if (i < cnt && m.scheme() == ALPB200_SCHEME_ALP) {
				x.val[k] = __ldg(ev + i);
				x.pos[k] = __ldg(ep + i);
			}
		}
	}
	return x;
}
// exceptions 32.. of a vector: pull their cache lines towards the SM one vector ahead (lane i takes line i)
__device__ __forceinline__ void prefetch_exception_tail(const ColView& col, const MetaRegs& m, int t, uint32_t value_bytes) {
	const uint32_t cnt = m.exc_cnt();
	if (cnt <= 32) { return; }
	const uint64_t first = (uint64_t)m.exc_off() + 32, last = (uint64_t)m.exc_off() + cnt - 1;
	const char*    v0    = static_cast<const char*>(col.exc_val) + ((first * value_bytes) & ~127ull);
	const char*    v1    = static_cast<const char*>(col.exc_val) + last * value_bytes;
	const char*    p0    = reinterpret_cast<const char*>(col.exc_pos) + ((first * 2) & ~127ull);
	const char*    p1    = reinterpret_cast<const char*>(col.exc_pos) + last * 2;
	for (const char* a = v0 + 128 * t; a <= v1; a += 128 * 32) {
		asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
	}
	if (p0 + 128 * t <= p1) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + 128 * t)); }
}

// ALP exception patch (decoder.hpp:141-149)
template <typename PT, int TAIL = 0>
__device__ __forceinline__ void patch_alp(const ColView& col, const MetaRegs& m, const ExcRegs<typename Traits<PT>::UT>& x,
                                          PT* __restrict__ out_vec, int t,
                                          const ExcTail<typename Traits<PT>::UT, TAIL>* tail = nullptr) {
	using UT           = typename Traits<PT>::UT;
	const uint32_t cnt = m.exc_cnt();
	UT*            ov  = reinterpret_cast<UT*>(out_vec);
	if ((uint32_t)t < cnt) { ov[x.pos] = x.val; }
	uint32_t from = 32;
	if constexpr (TAIL > 0) {
		if (tail != nullptr) {
#pragma unroll
			for (int k = 0; k < TAIL; k++) {
				if ((uint32_t)t + 32u * (k + 1) < cnt) { ov[tail->pos[k]] = tail->val[k]; }
			}
			from = 32u * (TAIL + 1);
		}
	}
	if (cnt > from) {
		const UT*       ev = static_cast<const UT*>(col.exc_val) + m.exc_off();
		const uint16_t* ep = col.exc_pos + m.exc_off();
		for (uint32_t i = t + from; i < cnt; i += 96) {  // three independent (position, value) loads in flight per lane
			const uint32_t i1 = i + 32, i2 = i + 64;
			uint32_t       p0 = ep[i], p1 = 0, p2 = 0;
			UT             v0 = ev[i], v1 = 0, v2 = 0;
			if (i1 < cnt) {
				p1 = ep[i1];
				v1 = ev[i1];
			}
			if (i2 < cnt) {
				p2 = ep[i2];
				v2 = ev[i2];
			}
			ov[p0] = v0;
			if (i1 < cnt) { ov[p1] = v1; }
			if (i2 < cnt) { ov[p2] = v2; }
		}
	}
}

// ---- ALP_RD (rd.hpp:152-178): right parts on T-bit lanes, dictionary indices on 16-bit lanes -----------------------
// rd_unpack_rows: emit(r, bits) receives the glued bit pattern (dict[idx] << right_bw) | right of the thread's row r
// (value index Map<PT>::index(t, r)); exceptions are NOT applied.  Shared by the decode and the decode+SUM kernels.
template <typename Emit>
__device__ __forceinline__ void rd_unpack_rows(const uint8_t* stage, const MetaRegs& m, int t, double /*tag*/, Emit&& emit) {
	const uint32_t  rbw = m.bw(), lbw = m.e();
	const int       lane = t & 15, half = t >> 4;
	const uint16_t* lblk  = reinterpret_cast<const uint16_t*>(stage + 128u * rbw);
	const uint32_t  lmask = (1u << lbw) - 1;
	dispatch_width<48, 63>(rbw < 48 ? 48u : rbw, [&](auto W) {  // right_bit_width = 64 - cut, cut in 1..16 (rd.hpp:95)
		constexpr int BW = decltype(W)::value;
		unpack64_rows<BW>(stage, lane, half, [&](int r, uint32_t lo, uint32_t hi) {
			const uint64_t right = ((uint64_t)hi << 32) | lo;
			const uint32_t v     = 16u * (32u * half + r) + lane;  // 16-bit-lane coordinates of the same value
			const uint32_t idx   = extract16(lblk, v & 63, (v >> 6) * lbw, lmask);
			emit(r, ((uint64_t)dict_lookup(m.a, idx) << BW) | right);
		});
	});
}
template <typename Emit>
__device__ __forceinline__ void rd_unpack_rows(const uint8_t* stage, const MetaRegs& m, int t, float /*tag*/, Emit&& emit) {
	const uint32_t  rbw = m.bw(), lbw = m.e();
	const uint16_t* lblk  = reinterpret_cast<const uint16_t*>(stage + 128u * rbw);
	const uint32_t  lmask = (1u << lbw) - 1;
	dispatch_width<16, 31>(rbw < 16 ? 16u : rbw, [&](auto W) {
		constexpr int BW = decltype(W)::value;
		unpack32_rows<BW>(stage, t, [&](int r, uint32_t right) {
			const uint32_t v   = 32u * r + t;
			const uint32_t idx = extract16(lblk, v & 63, (v >> 6) * lbw, lmask);
			emit(r, (dict_lookup(m.a, idx) << BW) | right);
		});
	});
}
// right part of position p, re-extracted from the stage (exception patch)
__device__ __forceinline__ uint64_t rd_right_at(const uint8_t* stage, uint32_t rbw, uint32_t p, uint64_t /*tag*/) {
	return extract64(reinterpret_cast<const uint64_t*>(stage), p & 15, (p >> 4) * rbw, low_mask<uint64_t>(rbw));
}
__device__ __forceinline__ uint32_t rd_right_at(const uint8_t* stage, uint32_t rbw, uint32_t p, uint32_t /*tag*/) {
	return extract32(reinterpret_cast<const uint32_t*>(stage), p & 31, (p >> 5) * rbw, low_mask<uint32_t>(rbw));
}

template <typename PT>
__device__ __forceinline__ void decode_rd_vector(const uint8_t* stage, const ColView& col, const MetaRegs& m,
                                                 const ExcRegs<typename Traits<PT>::UT>& x, PT* __restrict__ out_vec, int t) {
	using UT           = typename Traits<PT>::UT;
	const uint32_t rbw = m.bw();
	UT*            ov  = reinterpret_cast<UT*>(out_vec);
	UT*            o   = ov + Map<PT>::index(t, 0);
	constexpr int  ROW = Traits<PT>::LANES;  // distance between consecutive rows of a thread
	rd_unpack_rows(stage, m, t, PT(), [&](int r, UT bits) { store_out(&o[ROW * r], bits); });
	__syncwarp();
	// exceptions: the true left part replaces the dictionary entry (rd.hpp:172-177)
	const uint32_t cnt = m.exc_cnt();
	if ((uint32_t)t < cnt) { ov[x.pos] = ((x.val & 0xFFFFu) << rbw) | rd_right_at(stage, rbw, x.pos, UT()); }
	if (cnt > 32) {
		const UT*       ev = static_cast<const UT*>(col.exc_val) + m.exc_off();
		const uint16_t* ep = col.exc_pos + m.exc_off();
		for (uint32_t i = t + 32; i < cnt; i += 32) {
			const uint32_t p = ep[i];
			ov[p]            = ((ev[i] & 0xFFFFu) << rbw) | rd_right_at(stage, rbw, p, UT());
		}
	}
}

// ---- the kernel -----------------------------------------------------------------------------------------------------
// Persistent warps: warp g of the grid decodes vectors g, g + G, g + 2G, ...  Each warp owns two shared-memory
// stages of `stage_bytes` and two mbarriers.  Software pipeline per warp: records are read two vectors ahead, the
// packed block and the first 32 exceptions one vector ahead.
// OUT_TILE: decode into a per-warp shared-memory tile, patch exceptions there, and write the vector with ONE contiguous
// 8 / 4 KiB bulk-async store (TMA 1-D) — measured +4..7 % on write-heavy ALP columns over per-row line stores because
// DRAM sees dense sequential writes.  Without it (wide blocks, e.g. ALP_RD, where the tile would halve occupancy) the
// warp stores full 128-byte lines directly and patches exceptions afterwards while the lines are still in L2.
template <typename PT, int WARPS, bool OUT_TILE>
__global__ void __launch_bounds__(WARPS * 32, DecCfg<PT>::MINBLOCKS) decode_kernel(ColView col, uint64_t first_vector, uint64_t n_vectors,
                                                               PT* __restrict__ out, uint32_t stage_bytes,
                                                               unsigned long long* __restrict__ counter,
                                                               const unsigned long long* __restrict__ oversize) {
	using UT = typename Traits<PT>::UT;
	extern __shared__ __align__(128) uint8_t smem[];
	const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
	if (oversize != nullptr && *oversize != 0) {  // the hint was too small for this call (see hint_check_kernel): slow, correct
		const uint64_t n_warps = (uint64_t)gridDim.x * WARPS;
		for (uint64_t w = (uint64_t)blockIdx.x * WARPS + warp; w < n_vectors; w += n_warps) {
			decode_vector_slow<PT>(col, load_meta(col.meta + first_vector + w), out + w * (uint64_t)VEC, t);
		}
		return;
	}
	// per warp: [decoded vector tile (OUT_TILE) | packed stage 0 | packed stage 1]
	constexpr uint32_t TILE  = OUT_TILE ? VEC * sizeof(PT) : 0;
	uint8_t*           mine  = smem + (size_t)warp * (TILE + 2 * stage_bytes);
	PT*                tile  = reinterpret_cast<PT*>(mine);
	uint8_t*           stage = mine + TILE;
	uint64_t*          bars  = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * (TILE + 2 * stage_bytes)) + 2 * warp;
	if (t == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
		fence_mbar_init();
	}
	__syncwarp();

	// Dynamic work distribution: a warp draws chunks of CHUNK consecutive vectors from a global counter (the next chunk
	// is requested while a few vectors of the current one are left, so the atomic's latency never stalls the pipeline).
	// SMs differ in their distance to memory; with a static split the fast ones idle at the end (ncu: SM active 85 %).
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
	constexpr int NTAIL = DecCfg<PT>::EXC_TAIL;
	using XT            = ExcTail<UT, NTAIL>;
	ExcRegs<UT> xcur  = load_exceptions<UT>(col, cur, t);
	XT          tcur  = load_exception_tail<UT, NTAIL>(col, cur, t);
	uint32_t    phase = 0;  // bit s: parity the next wait on stage s must see
	for (int s = 0;; s ^= 1) {
		ExcRegs<UT> xnxt = xcur;
		XT          tnxt = tcur;
		if (has_next) {
			issue(nxt, s ^ 1);  // stage s^1 was drained one iteration ago (see __syncwarp below)
			xnxt = load_exceptions<UT>(col, nxt, t);
			tnxt = load_exception_tail<UT, NTAIL>(col, nxt, t);
			if (NTAIL == 0 || nxt.exc_cnt() > 32u * (NTAIL + 1) || nxt.scheme() != ALPB200_SCHEME_ALP) { prefetch_exception_tail(col, nxt, t, sizeof(UT)); }
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
		PT* out_vec = out + v * (uint64_t)VEC;
		if constexpr (OUT_TILE) {
			// the previous vector's bulk store must have finished READING the tile before it is overwritten
			if (t == 0) { bulk_wait_read_all(); }
			__syncwarp();
			out_vec = tile;
		}
		if (cur.scheme() == ALPB200_SCHEME_ALP) {
			decode_alp_vector(stg, cur, out_vec, t);
			__syncwarp();  // orders the patch stores after the lane-interleaved main stores
			patch_alp<PT, NTAIL>(col, cur, xcur, out_vec, t, &tcur);
		} else {
			decode_rd_vector<PT>(stg, col, cur, xcur, out_vec, t);
		}
		if constexpr (OUT_TILE) {
			fence_proxy_async_smem();  // generic-proxy writes to the tile -> visible to the bulk-copy engine
		}
		__syncwarp();  // every lane is done reading stage s before lane 0 refills it two iterations from now
		if constexpr (OUT_TILE) {
			if (t == 0) {
				bulk_s2g(out + v * (uint64_t)VEC, tile, TILE);  // one contiguous 8 / 4 KiB write per vector (TMA 1-D)
				bulk_commit();
			}
		}
		if (!has_next) { break; }
		cur      = nxt;
		nxt      = nn;
		xcur     = xnxt;
		tcur     = tnxt;
		has_next = has_nn;
		v        = v_next;
		v_next   = v_nn;
	}
	if constexpr (OUT_TILE) {
		if (t == 0) { bulk_wait_all(); }  // the last store must be complete before the tile's CTA goes away
	}
}

// ---- column validation on the device (alpb200_column_validate_device) ------------------------------------------------
// One thread per record: scheme / widths / exponent / factor / exception count in range, block and exception run inside
// the arrays, every exception position inside the vector.  result[0] |= 1 when a record fails; result[1] = widest block.
static __global__ void validate_kernel(ColView col, uint64_t n_vectors, uint64_t packed_capacity, uint64_t exc_capacity, uint32_t value_bytes,
                                       unsigned long long* __restrict__ result) {
	const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n_vectors) { return; }
	const MetaRegs m  = load_meta(col.meta + v);
	const uint32_t T  = 8u * value_bytes, max_exp = value_bytes == 8 ? 18u : 10u;
	const bool     rd = m.scheme() == ALPB200_SCHEME_ALP_RD;
	bool           ok = m.exc_cnt() <= VEC && (rd || m.scheme() == ALPB200_SCHEME_ALP);
	if (rd) {
		ok = ok && m.bw() < T && m.bw() + 16u >= T && m.e() >= 1 && m.e() <= 3 && m.f() >= 1 && m.f() <= ALPB200_RD_DICT_SIZE;
	} else {
		ok = ok && m.bw() <= T && m.e() <= max_exp && m.f() <= m.e();
	}
	const uint64_t bytes = m.block_bytes();
	ok = ok && (uint64_t)m.packed_off() * 128ull + bytes <= packed_capacity && (uint64_t)m.exc_off() + m.exc_cnt() <= exc_capacity;
	if (ok) {
		const uint16_t* ep = col.exc_pos + m.exc_off();
		for (uint32_t i = 0; i < m.exc_cnt(); i++) {
			ok = ok && ep[i] < VEC;
		}
	}
	if (!ok) { atomicOr(&result[0], 1ull); }
	if (ok) { atomicMax(&result[1], (unsigned long long)bytes); }
}

}  // namespace alpb200
