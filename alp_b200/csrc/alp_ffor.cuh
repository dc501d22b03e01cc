// alp_ffor.cuh — bit-width-specialised FastLanes FFOR / UNFFOR for one warp.
//
// The reference ships one fully unrolled scalar function per (bit width, lane width) and dispatches with a
// `switch (bw)` (src/fastlanes_generated_unffor.cpp:22812-23211, src/fastlanes_generated_ffor.cpp:29750-30137).
// The same idea on a warp: the bit width is uniform per vector, so one `switch` per vector selects a template
// instance in which every shift, mask, shared-memory offset and register index is a compile-time constant —
// 2 instructions per 32-bit field instead of ~15 for a run-time width.
//
// Layout: T-bit lanes, L = 1024/T lanes, value v = L*row + lane, row `row` of a lane at bits
// [row*bw, row*bw+bw) of the lane's stream, stream word w stored at element L*w + lane of the block.
//   64-bit lanes: thread (lane = t&15, half = t>>4) owns rows 32*half .. 32*half+31 = stream bits
//                 [32*bw*half, 32*bw*(half+1)) = the BW 32-bit words j = half*BW .. half*BW+BW-1 of the lane's stream
//   32-bit lanes: thread t owns lane t, rows 0..31 = BW 32-bit words
#pragma once

#include <type_traits>
#include <utility>

#include "alp_device.cuh"

namespace alpb200 {

// f(std::integral_constant<int, I>) for I = 0 .. N-1, unrolled at compile time
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
	if constexpr (I < N) {
		f(std::integral_constant<int, I> {});
		static_for<I + 1, N>(f);
	}
}

// ---- the thread's BW-word window, read from a verbatim block image in shared memory -------------------------------
// 64-bit lanes.  32-bit word j of lane `lane` sits at byte 4*(32*(j>>1) + 2*lane + (j&1)).
// BW even: word pairs are 64-bit aligned for both halves -> BW/2 LDS.64 (a half-warp reads 128 contiguous bytes).
// BW odd : the two halves have opposite word parity -> BW LDS.32 whose banks interleave (half 0 even banks, half 1 odd
//          banks, or vice versa): conflict-free.  Two base addresses cover even / odd local word indices.
template <int BW>
struct Window64 {
	uint32_t w[BW > 0 ? BW : 1];
	__device__ __forceinline__ void load(const uint8_t* blk, int lane, int half) {
		if constexpr (BW == 0) {
			return;
		} else if constexpr ((BW & 1) == 0) {
			const uint64_t* p = reinterpret_cast<const uint64_t*>(blk) + 16 * (half * (BW / 2)) + lane;
#pragma unroll
			for (int m = 0; m < BW / 2; m++) {
				const uint64_t v = p[16 * m];
				w[2 * m]         = (uint32_t)v;
				w[2 * m + 1]     = (uint32_t)(v >> 32);
			}
		} else {
			const uint8_t* even = blk + 8 * lane + (half ? 64 * (BW - 1) + 4 : 0);
			const uint8_t* odd  = blk + 8 * lane + (half ? 64 * (BW - 1) + 128 : 4);
#pragma unroll
			for (int i = 0; i < BW; i++) {
				w[i] = *reinterpret_cast<const uint32_t*>(((i & 1) ? odd : even) + 128 * (i >> 1));
			}
		}
	}
};

// bits [bit, bit+n) of the window, n <= 32, as compile-time shifts
template <int BW, int BIT, int N>
__device__ __forceinline__ uint32_t window_field(const uint32_t (&w)[BW > 0 ? BW : 1]) {
	constexpr int      I = BIT >> 5, S = BIT & 31;
	constexpr uint32_t M = N >= 32 ? 0xFFFFFFFFu : ((1u << N) - 1u);
	if constexpr (S + N <= 32) {
		if constexpr (S + N == 32) {
			return w[I] >> S;
		} else {
			return (w[I] >> S) & M;
		}
	} else {
		return __funnelshift_r(w[I], w[I + 1], S) & M;
	}
}

// unpack the thread's 32 rows of a 64-bit-lane block; consume(r, lo, hi) receives the BW-bit field of row r
template <int BW, typename Consume>
__device__ __forceinline__ void unpack64_rows(const uint8_t* blk, int lane, int half, Consume&& consume) {
	Window64<BW> win;
	win.load(blk, lane, half);
	auto row = [&](auto R) {
		constexpr int r = decltype(R)::value;
		if constexpr (BW == 0) {
			consume(r, 0u, 0u);
		} else if constexpr (BW <= 32) {
			consume(r, window_field<BW, r * BW, BW>(win.w), 0u);
		} else {
			consume(r, window_field<BW, r * BW, 32>(win.w), window_field<BW, r * BW + 32, BW - 32>(win.w));
		}
	};
	static_for<0, 32>(row);  // 32 rows, unrolled with compile-time row numbers
}

// 32-bit lanes: word i of lane t at element 32*i + t
template <int BW, typename Consume>
__device__ __forceinline__ void unpack32_rows(const uint8_t* blk, int t, Consume&& consume) {
	uint32_t w[BW > 0 ? BW : 1];
	const uint32_t* p = reinterpret_cast<const uint32_t*>(blk) + t;
#pragma unroll
	for (int i = 0; i < BW; i++) {
		w[i] = p[32 * i];
	}
	auto row = [&](auto R) {
		constexpr int r = decltype(R)::value;
		if constexpr (BW == 0) {
			consume(r, 0u);
		} else {
			consume(r, window_field<BW, r * BW, BW>(w));
		}
	};
	static_for<0, 32>(row);
}

// ---- pack: the thread's 32 rows -> BW 32-bit words ----------------------------------------------------------------
// OR the (masked) field of row r into the word array at a compile-time position
template <int NW, int BIT, int N, bool MASKED = true>
__device__ __forceinline__ void window_put(uint32_t (&w)[NW], uint32_t x) {
	constexpr int      I = BIT >> 5, S = BIT & 31;
	constexpr uint32_t M = N >= 32 ? 0xFFFFFFFFu : ((1u << N) - 1u);
	if constexpr (N < 32 && MASKED) { x &= M; }  // !MASKED: the caller guarantees x < 2^N
	if constexpr (S == 0) {
		w[I] = x;  // first field of a word
	} else {
		w[I] |= x << S;
		if constexpr (S + N > 32) { w[I + 1] = x >> (32 - S); }  // spill-over always opens the next word
	}
}

// 64-bit lanes.  produce(R, lo, hi) yields the (unmasked) field of row R::value; dst = the block as 64-bit elements.
// MODE 0 (PACK_DIRECT): words are stored as soon as they are complete (all positions are compile-time), so only a few
//                       stay live.  dst must not alias what produce() reads.
// MODE 1 (PACK_INPLACE_NARROW, BW <= 32): packing IN PLACE — dst is the [row][lane] tile produce() reads.  The block
//                       image of half 1 lands on rows half 0 has not read yet, so every word stays in registers until
//                       all 32 rows have been produced and the whole warp has passed a __syncwarp().
// MODE 2 (PACK_INPLACE_WIDE, BW > 32): in place with at most 2*(32 - BW/2) words deferred.  Half 0 writes element
//                       16*m + lane, the slot of its own row m, which it has consumed by the time pair m is complete
//                       (BW <= 64): immediate.  Half 1 writes element 16*(B1 + m) + lane, B1 = ceil(BW/2): below
//                       element 512 that is a row of half 0 (deferred until after the __syncwarp()), from 512 on it is
//                       its own row B1 + m - 32 <= m, already consumed: immediate.
// ROW: distance between consecutive 16-element rows of the block image in dst, in 64-bit elements (16 = dense).
// MASKED = false: produce() already yields values below 2^BW (no AND per field).
constexpr int PACK_DIRECT = 0, PACK_INPLACE_NARROW = 1, PACK_INPLACE_WIDE = 2;
template <int BW, int MODE = PACK_DIRECT, int ROW = 16, bool MASKED = true, typename Produce>
__device__ __forceinline__ void pack64_rows(int lane, int half, uint64_t* dst, Produce&& produce) {
	static_assert(MODE != PACK_INPLACE_NARROW || BW <= 32, "narrow in-place packing keeps every word in registers");
	uint32_t w[BW + 1] = {};  // (every word is assigned before it is read; the initialiser only tells the front end so)
	// BW even: pair m = words (2m, 2m+1) -> element 16*(half*BW/2 + m) + lane.
	// BW odd : half 0 owns stream words 0..BW-1, half 1 words BW..2BW-1, so pairs start one word later for half 1:
	//          pair m = half ? (2m+1, 2m+2) : (2m, 2m+1) -> element 16*((half ? (BW+1)/2 : 0) + m) + lane, and element
	//          (BW-1)/2 is shared: last word of half 0 (low) + first word of half 1 (high), completed with one shuffle.
	constexpr int B1      = (BW + 1) / 2;                                   // first element row of half 1's pairs
	constexpr int N_PAIRS = (BW & 1) ? (BW - 1) / 2 : BW / 2;               // per half
	constexpr int N_LATE  = MODE == PACK_INPLACE_NARROW ? N_PAIRS : (MODE == PACK_INPLACE_WIDE ? (32 - B1 > 0 ? (32 - B1 < N_PAIRS ? 32 - B1 : N_PAIRS) : 0) : 0);
	uint64_t*     p       = dst + ROW * ((BW & 1) ? (half ? B1 : 0) : half * (BW / 2)) + lane;
	auto store_pair = [&](auto Mc) {
		constexpr int m = decltype(Mc)::value;
		if constexpr ((BW & 1) == 0) {
			p[ROW * m] = (uint64_t)w[2 * m] | ((uint64_t)w[2 * m + 1] << 32);
		} else {
			const uint32_t a = half ? w[2 * m + 1] : w[2 * m], b = half ? w[2 * m + 2] : w[2 * m + 1];
			p[ROW * m]       = (uint64_t)a | ((uint64_t)b << 32);
		}
	};
	static_for<0, 32>([&](auto R) {
		constexpr int r = decltype(R)::value;
		uint32_t      lo, hi;
		produce(R, lo, hi);
		if constexpr (BW <= 32) {
			window_put<BW + 1, r * BW, BW, MASKED>(w, lo);
		} else {
			window_put<BW + 1, r * BW, 32>(w, lo);
			window_put<BW + 1, r * BW + 32, BW - 32>(w, hi);
		}
		if constexpr (MODE != PACK_INPLACE_NARROW) {
			constexpr int done_before = (r * BW) >> 5, done_now = ((r + 1) * BW) >> 5;  // complete words
			constexpr int pairs_before = (BW & 1) ? (done_before > 0 ? (done_before - 1) / 2 : 0) : done_before / 2;
			constexpr int pairs_now    = (BW & 1) ? (done_now > 0 ? (done_now - 1) / 2 : 0) : done_now / 2;
			static_for<pairs_before, pairs_now>([&](auto Mc) {
				constexpr int m = decltype(Mc)::value;
				if constexpr (MODE == PACK_INPLACE_WIDE && m < N_LATE) {
					if (half == 0) { store_pair(Mc); }  // half 1's pair m would land on a row half 0 still has to read
				} else {
					store_pair(Mc);
				}
			});
		}
	});
	if constexpr (MODE != PACK_DIRECT) {
		__syncwarp();  // every lane has read all of its rows
		if constexpr (MODE == PACK_INPLACE_NARROW) {
			static_for<0, N_LATE>(store_pair);
		} else {
			static_for<0, N_LATE>([&](auto Mc) {
				if (half == 1) { store_pair(Mc); }
			});
		}
	}
	if constexpr (BW & 1) {
		const uint32_t other = __shfl_xor_sync(FULL, half ? w[0] : w[BW - 1], 16);
		if (half) { dst[ROW * ((BW - 1) / 2) + lane] = (uint64_t)other | ((uint64_t)w[0] << 32); }
	}
}

// 32-bit lanes.  produce(R) yields the (unmasked) field of row R::value; word i of lane t goes to element 32*i + t.
// Safe IN PLACE (dst aliasing a [row][lane] tile that produce() reads): word i is complete only after row i has been
// produced, and element 32*i + t is the slot of this very thread's row i.
template <int BW, typename Produce>
__device__ __forceinline__ void pack32_rows(int t, uint32_t* dst, Produce&& produce) {
	uint32_t w[BW + 1] = {};  // (every word is assigned before it is read; the initialiser only tells the front end so)
	static_for<0, 32>([&](auto R) {
		constexpr int r = decltype(R)::value;
		window_put<BW + 1, r * BW, BW>(w, produce(R));
		constexpr int done_before = (r * BW) >> 5, done_now = ((r + 1) * BW) >> 5;
		static_for<done_before, done_now>([&](auto Ic) {
			constexpr int i = decltype(Ic)::value;
			dst[32 * i + t] = w[i];
		});
	});
}

// ---- pack into REGISTERS ---------------------------------------------------------------------------------------------
// The pipelined encoder (alp_encode_pipe.cuh) builds a narrow block image (BW <= 32) in the thread's registers, hands its
// shared-memory tile to the NEXT vector's bulk load and writes the image to the column once the block's offset is known.
// fill: the thread's 32 rows -> its BW 32-bit words (w[BW] is scratch).  store: the words -> the block (dense rows).
template <int BW, typename Produce>
__device__ __forceinline__ void pack64_fill(uint32_t (&w)[BW + 1], Produce&& produce) {
	static_assert(BW <= 32, "a register image holds at most 32 words per thread");
	static_for<0, 32>([&](auto R) {
		constexpr int r = decltype(R)::value;
		uint32_t      lo, hi;
		produce(R, lo, hi);
		window_put<BW + 1, r * BW, BW>(w, lo);
	});
}
// 64-bit lanes: thread (lane, half) holds stream words half*BW .. half*BW+BW-1 of its lane (see pack64_rows); a half-warp
// stores one full 128-byte line per instruction.  dst = the block as 64-bit elements (global memory).
template <int BW>
__device__ __forceinline__ void pack64_store(const uint32_t (&w)[BW + 1], int lane, int half, uint64_t* __restrict__ dst) {
	constexpr int B1      = (BW + 1) / 2;
	constexpr int N_PAIRS = (BW & 1) ? (BW - 1) / 2 : BW / 2;
	uint64_t*     p       = dst + 16 * ((BW & 1) ? (half ? B1 : 0) : half * (BW / 2)) + lane;
	static_for<0, N_PAIRS>([&](auto Mc) {
		constexpr int m = decltype(Mc)::value;
		if constexpr ((BW & 1) == 0) {
			p[16 * m] = (uint64_t)w[2 * m] | ((uint64_t)w[2 * m + 1] << 32);
		} else {
			const uint32_t a = half ? w[2 * m + 1] : w[2 * m], b = half ? w[2 * m + 2] : w[2 * m + 1];
			p[16 * m]        = (uint64_t)a | ((uint64_t)b << 32);
		}
	});
	if constexpr (BW & 1) {  // the element the two halves of a lane share
		const uint32_t other = __shfl_xor_sync(FULL, half ? w[0] : w[BW - 1], 16);
		if (half) { dst[16 * ((BW - 1) / 2) + lane] = (uint64_t)other | ((uint64_t)w[0] << 32); }
	}
}
// 32-bit lanes: thread t holds words 0..BW-1 of lane t; every store instruction writes one full 128-byte line
template <int BW, typename Produce>
__device__ __forceinline__ void pack32_fill(uint32_t (&w)[BW + 1], Produce&& produce) {
	static_for<0, 32>([&](auto R) {
		constexpr int r = decltype(R)::value;
		window_put<BW + 1, r * BW, BW>(w, produce(R));
	});
}
template <int BW>
__device__ __forceinline__ void pack32_store(const uint32_t (&w)[BW + 1], int t, uint32_t* __restrict__ dst) {
	static_for<0, BW>([&](auto Ic) {
		constexpr int i = decltype(Ic)::value;
		dst[32 * i + t] = w[i];
	});
}

// ---- switch (bw) -----------------------------------------------------------------------------------------------------
// call f(std::integral_constant<int, BW>) for the run-time width bw in [LO, HI]
template <int LO, int HI, typename F>
__device__ __forceinline__ void dispatch_width(uint32_t bw, F&& f) {
	if constexpr (LO == HI) {
		f(std::integral_constant<int, LO> {});
	} else {
		constexpr int MID = (LO + HI) / 2;
		if ((int)bw <= MID) {
			dispatch_width<LO, MID>(bw, f);
		} else {
			dispatch_width<MID + 1, HI>(bw, f);
		}
	}
}

}  // namespace alpb200
