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

namespace alpb200 {

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

// ---- ALP, 64-bit lanes -------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_alp_vector(const uint8_t* stage, const MetaRegs& m, double* __restrict__ out_vec, int t) {
	using T             = Traits<double>;
	const uint32_t bw   = m.bw();
	const int      lane = t & 15, half = t >> 4;
	const int64_t  fact = T::fact10(m.f());
	const double   frac = T::frac10(m.e());
	const uint64_t base = m.base();
	double*        o    = out_vec + 512 * half + lane;  // value index 16*(32*half + r) + lane
	if (bw == 0) {  // unffor bw=0 broadcasts the base (src/fastlanes_generated_unffor.cpp:4-22)
		const double v = decode_value<double>((int64_t)base, fact, frac);
#pragma unroll 8
		for (int r = 0; r < 32; r++) {
			o[16 * r] = v;
		}
		return;
	}
	const uint64_t* blk  = reinterpret_cast<const uint64_t*>(stage);
	const uint64_t  mask = low_mask<uint64_t>(bw);
	uint32_t        bit  = 32u * bw * half;
#pragma unroll 4
	for (int r = 0; r < 32; r++, bit += bw) {
		const uint64_t d = extract64(blk, lane, bit, mask);
		o[16 * r]        = decode_value<double>((int64_t)(d + base), fact, frac);  // src/falp.cpp:1049-1056
	}
}

// ---- ALP, 32-bit lanes -------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_alp_vector(const uint8_t* stage, const MetaRegs& m, float* __restrict__ out_vec, int t) {
	using T             = Traits<float>;
	const uint32_t bw   = m.bw();
	const int32_t  fact = T::fact10(m.f());
	const float    frac = T::frac10(m.e());
	const uint32_t base = m.a.x;
	float*         o    = out_vec + t;  // value index 32*r + lane
	if (bw == 0) {
		const float v = decode_value<float>((int32_t)base, fact, frac);
#pragma unroll 8
		for (int r = 0; r < 32; r++) {
			o[32 * r] = v;
		}
		return;
	}
	const uint32_t* blk  = reinterpret_cast<const uint32_t*>(stage);
	const uint32_t  mask = low_mask<uint32_t>(bw);
	uint32_t        bit  = 0;
#pragma unroll 8
	for (int r = 0; r < 32; r++, bit += bw) {
		const uint32_t d = extract32(blk, t, bit, mask);
		o[32 * r]        = decode_value<float>((int32_t)(d + base), fact, frac);
	}
}

// ---- ALP exception patch (decoder.hpp:141-149) ----------------------------------------------------------------
template <typename PT>
__device__ __forceinline__ void patch_alp(const ColView& col, const MetaRegs& m, PT* __restrict__ out_vec, int t) {
	const uint32_t  cnt = m.exc_cnt();
	const PT*       ev  = static_cast<const PT*>(col.exc_val) + m.exc_off();
	const uint16_t* ep  = col.exc_pos + m.exc_off();
	for (uint32_t i = t; i < cnt; i += 32) {
		out_vec[ep[i]] = ev[i];
	}
}

// ---- ALP_RD (rd.hpp:152-178): right parts on T-bit lanes, dictionary indices on 16-bit lanes -----------------------
__device__ __forceinline__ void decode_rd_vector(const uint8_t* stage, const ColView& col, const MetaRegs& m,
                                                 double* __restrict__ out_vec, int t) {
	const uint32_t  rbw = m.bw(), lbw = m.e();
	const int       lane = t & 15, half = t >> 4;
	const uint64_t* rblk  = reinterpret_cast<const uint64_t*>(stage);
	const uint16_t* lblk  = reinterpret_cast<const uint16_t*>(stage + 128u * rbw);
	const uint64_t  rmask = low_mask<uint64_t>(rbw);
	const uint32_t  lmask = (1u << lbw) - 1;
	uint64_t*       o     = reinterpret_cast<uint64_t*>(out_vec) + 512 * half + lane;
	uint32_t        bit   = 32u * rbw * half;
#pragma unroll 4
	for (int r = 0; r < 32; r++, bit += rbw) {
		const uint64_t right = extract64(rblk, lane, bit, rmask);
		const uint32_t v     = 16u * (32u * half + r) + lane;  // 16-bit-lane coordinates of the same value
		const uint32_t idx   = extract16(lblk, v & 63, (v >> 6) * lbw, lmask);
		o[16 * r]            = ((uint64_t)dict_lookup(m.a, idx) << rbw) | right;
	}
	__syncwarp();
	// exceptions: the true left part replaces the dictionary entry (rd.hpp:172-177)
	const uint32_t  cnt = m.exc_cnt();
	const uint64_t* ev  = static_cast<const uint64_t*>(col.exc_val) + m.exc_off();
	const uint16_t* ep  = col.exc_pos + m.exc_off();
	uint64_t*       ov  = reinterpret_cast<uint64_t*>(out_vec);
	for (uint32_t i = t; i < cnt; i += 32) {
		const uint32_t p     = ep[i];
		const uint64_t right = extract64(rblk, p & 15, (p >> 4) * rbw, rmask);
		ov[p]                = ((ev[i] & 0xFFFFu) << rbw) | right;
	}
}

__device__ __forceinline__ void decode_rd_vector(const uint8_t* stage, const ColView& col, const MetaRegs& m,
                                                 float* __restrict__ out_vec, int t) {
	const uint32_t  rbw = m.bw(), lbw = m.e();
	const uint32_t* rblk  = reinterpret_cast<const uint32_t*>(stage);
	const uint16_t* lblk  = reinterpret_cast<const uint16_t*>(stage + 128u * rbw);
	const uint32_t  rmask = low_mask<uint32_t>(rbw);
	const uint32_t  lmask = (1u << lbw) - 1;
	uint32_t*       o     = reinterpret_cast<uint32_t*>(out_vec) + t;
	uint32_t        bit   = 0;
#pragma unroll 4
	for (int r = 0; r < 32; r++, bit += rbw) {
		const uint32_t right = extract32(rblk, t, bit, rmask);
		const uint32_t v     = 32u * r + t;
		const uint32_t idx   = extract16(lblk, v & 63, (v >> 6) * lbw, lmask);
		o[32 * r]            = (dict_lookup(m.a, idx) << rbw) | right;
	}
	__syncwarp();
	const uint32_t  cnt = m.exc_cnt();
	const uint32_t* ev  = static_cast<const uint32_t*>(col.exc_val) + m.exc_off();
	const uint16_t* ep  = col.exc_pos + m.exc_off();
	uint32_t*       ov  = reinterpret_cast<uint32_t*>(out_vec);
	for (uint32_t i = t; i < cnt; i += 32) {
		const uint32_t p     = ep[i];
		const uint32_t right = extract32(rblk, p & 31, (p >> 5) * rbw, rmask);
		ov[p]                = ((ev[i] & 0xFFFFu) << rbw) | right;
	}
}

// ---- the kernel -----------------------------------------------------------------------------------------------------
// Persistent warps: warp g of the grid decodes vectors g, g + G, g + 2G, ...  Each warp owns two shared-memory
// stages of `stage_bytes` and two mbarriers.
template <typename PT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) decode_kernel(ColView col, uint64_t first_vector, uint64_t n_vectors,
                                                            PT* __restrict__ out, uint32_t stage_bytes) {
	extern __shared__ __align__(128) uint8_t smem[];
	const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
	uint8_t*  stage = smem + (size_t)warp * 2 * stage_bytes;
	uint64_t* bars  = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * 2 * stage_bytes) + 2 * warp;
	if (t == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
		fence_mbar_init();
	}
	__syncwarp();

	const uint64_t stride = (uint64_t)gridDim.x * WARPS;
	uint64_t       v      = (uint64_t)blockIdx.x * WARPS + warp;
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
	bool     has_next = v + stride < n_vectors;
	MetaRegs nxt      = cur;
	if (has_next) { nxt = load_meta(meta + v + stride); }
	issue(cur, 0);
	uint32_t phase = 0;  // bit s: parity the next wait on stage s must see
	for (int s = 0;; s ^= 1) {
		if (has_next) { issue(nxt, s ^ 1); }  // stage s^1 was drained one iteration ago (see __syncwarp below)
		const bool has_nn = v + 2 * stride < n_vectors;
		MetaRegs   nn     = nxt;
		if (has_nn) { nn = load_meta(meta + v + 2 * stride); }

		PT*            out_vec = out + v * (uint64_t)VEC;
		const uint8_t* stg     = stage + (size_t)s * stage_bytes;
		if (cur.block_bytes() != 0) {
			mbar_wait(&bars[s], (phase >> s) & 1u);
			phase ^= 1u << s;
		}
		if (cur.scheme() == ALPB200_SCHEME_ALP) {
			decode_alp_vector(stg, cur, out_vec, t);
			__syncwarp();  // orders the patch stores after the lane-interleaved main stores
			patch_alp<PT>(col, cur, out_vec, t);
		} else {
			decode_rd_vector(stg, col, cur, out_vec, t);
		}
		__syncwarp();  // every lane is done reading stage s before lane 0 refills it two iterations from now
		if (!has_next) { break; }
		cur      = nxt;
		nxt      = nn;
		has_next = has_nn;
		v += stride;
	}
}

}  // namespace alpb200
