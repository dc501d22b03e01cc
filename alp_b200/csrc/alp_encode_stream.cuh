// alp_encode_stream.cuh — the vector-order encoder as a persistent, warp-specialised pipeline (one CTA per SM).
//
// Same per-vector algorithm and the same bytes as encode_kernel (alp_encode.cuh; reference: encoder.hpp:402-418 + :109-120 +
// ffor, rd.hpp:109-147).  What changes is who waits for what.  In encode_kernel a thread block loads 9 vectors, analyses them,
// publishes its sizes and then WAITS for its output offset with 8 KiB of shared memory per vector pinned down: throughput =
// tiles per SM / lifetime of a tile, and the lifetime is load latency + analysis + in-order wait + emission (27 tiles / ~10 us).
// Here the three latencies are taken off the 8 KiB tile:
//
//   loader   (1 warp, 1 thread)  draws tickets (batches of B consecutive vectors, in order), streams the vectors into a ring of
//                                NT input tiles with bulk-async copies (TMA 1-D) as tiles come free — loads run ahead of the
//                                analysis instead of in front of it
//   compute  (W warps)           take the next landed tile (CTA-local sequence number), analyse it in place, report (units,
//                                exceptions) to their batch — the LAST reporter publishes the batch aggregate for the other
//                                CTAs' look-back — then allocate an entry in the CTA's output ring (in sequence order), gather
//                                the exceptions' original values from the tile, FFOR the block into the entry, and hand the
//                                tile back to the loader.  They never wait for an output offset.
//   placer   (1 warp)            per batch: look-back over the other CTAs' aggregates (lookback_prefix, alp_encode.cuh) for the
//                                batch's exclusive prefix, then per vector one bulk-async store of the block image (TMA 1-D),
//                                coalesced copies of the exception lists and the 32-byte record; frees the ring entries
//   scanner  (1 warp)            in the CTA that drew ticket 0: publishes an anchor every 32 batches (scan_anchors)
//
// While a vector waits for its offset it occupies its packed size (block + 10 bytes per exception), not 8 KiB.
// Deadlock freedom: a batch's aggregate needs only the ANALYSIS of its vectors (reported before the ring allocation), tiles
// are handed out in ticket order, and every wait of a ticket is for smaller tickets; the ring holds at least one entry of the
// maximum size and the placer drains entries one at a time when a successor's allocation is what it is waiting for.
#pragma once

#include "alp_encode.cuh"

namespace alpb200 {

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
	    "selp.u32 %0, 1, 0, P1;\n"
	    "}"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
	return ok != 0;
}
// one hardware-suspended wait of at most ~ns nanoseconds (no issue slots burnt while waiting)
__device__ __forceinline__ bool mbar_try_wait_ns(uint64_t* bar, uint32_t parity, uint32_t ns) {
	uint32_t ok;
#ifdef ALPB200_STREAM_HINT
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
	    "selp.u32 %0, 1, 0, P1;\n"
	    "}"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
	    : "memory");
#else
	(void)ns;  // the plain form waits up to a system-dependent limit
	asm volatile(
	    "{\n"
	    ".reg .pred P1;\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
	    "selp.u32 %0, 1, 0, P1;\n"
	    "}"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
#endif
	return ok != 0;
}
__device__ __forceinline__ uint32_t lds_volatile(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void     sts_volatile(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

// build knobs (development A/B: tools/build_variant.sh)
#ifndef ALPB200_STREAM_W
#define ALPB200_STREAM_W 16
#endif
#ifndef ALPB200_STREAM_NT64
#define ALPB200_STREAM_NT64 20
#endif
#ifndef ALPB200_STREAM_RING64
#define ALPB200_STREAM_RING64 58
#endif
#ifndef ALPB200_STREAM_NT32
#define ALPB200_STREAM_NT32 24
#endif
#ifndef ALPB200_STREAM_RING32
#define ALPB200_STREAM_RING32 64
#endif
#ifndef ALPB200_STREAM_B
#define ALPB200_STREAM_B 8
#endif
#ifndef ALPB200_STREAM_P
#define ALPB200_STREAM_P 3
#endif
// development profile (tools/build_variant.sh prof -DALPB200_STREAM_PROFILE=1): cycles per role and phase, summed over all warps
#ifdef ALPB200_STREAM_PROFILE
static __device__ unsigned long long g_stream_prof[32];
__device__ __forceinline__ long long stream_clock() {
	long long c;
	asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
	return c;
}
struct StreamProf {
	long long t0;
	__device__ __forceinline__ StreamProf() : t0(stream_clock()) {}
	__device__ __forceinline__ void lap(int slot, int t) {
		const long long now = stream_clock();
		if (t == 0) { atomicAdd(&g_stream_prof[slot], (unsigned long long)(now - t0)); }
		t0 = now;
	}
};
#else
struct StreamProf {
	__device__ __forceinline__ void lap(int, int) {}
};
#endif
template <typename PT>
struct StreamCfg;
template <>
struct StreamCfg<double> {
	static constexpr int      W    = ALPB200_STREAM_W;              // compute warps
	static constexpr int      NT   = ALPB200_STREAM_NT64;           // input tiles (8 KiB each)
	static constexpr uint32_t RING = ALPB200_STREAM_RING64 * 1024;  // output ring; >= twice the largest entry (66 units + 1024 x 10 bytes = 18.7 KB)
};
template <>
struct StreamCfg<float> {
	static constexpr int      W    = ALPB200_STREAM_W;
	static constexpr int      NT   = ALPB200_STREAM_NT32;  // 4 KiB each
	static constexpr uint32_t RING = ALPB200_STREAM_RING32 * 1024;
};
constexpr int STREAM_B   = ALPB200_STREAM_B;  // vectors per batch (= unit of the cross-CTA prefix; the workspace is sized for >= ENC_MIN_WARPS); <= 32
constexpr int STREAM_P   = ALPB200_STREAM_P;  // placer warps
constexpr int STREAM_TURNS = 32;
constexpr int STREAM_NV  = 64;  // per-vector records in flight between compute warps and placers
constexpr int STREAM_NBS = 16;  // batch slots: >= (NT + NV) / B + 2

struct StreamVRec {        // what the placer needs to emit one vector
	uint4    ra;           // first half of the 32-byte record: FOR base (ALP) / dictionary (ALP_RD)
	uint64_t v;            // vector index in this call
	uint32_t z, w;         // record words: cnt | scheme << 16 | bw << 24;  e | f << 8
	uint32_t units, cnt;   // block size in 128-byte units, exceptions
	uint32_t off, end;     // entry: physical offset in the ring, virtual end (what out_tail becomes when it is freed)
	uint32_t active, pad;
};

template <typename PT>
struct StreamShared {
	using C = StreamCfg<PT>;
	uint64_t   full[C::NT], empty[C::NT], done[STREAM_NV], sized[STREAM_NBS];
	uint64_t   turn[STREAM_TURNS];  // allocation order: sequence number s allocates after s - 1 (a chain of barriers, no polling)
	uint64_t   slot_vec[C::NT];
	uint64_t   b_agg[STREAM_NBS];
	StreamVRec vrec[STREAM_NV];
	uint32_t   b_ticket[STREAM_NBS], b_arrived[STREAM_NBS];
	uint32_t   b_info[STREAM_NBS][STREAM_B];  // units << 16 | exceptions, per vector of the batch
	uint32_t   next_seq, alloc_seq, alloc_head, out_tail, retired, retired_batch, alloc_blocked, loader_batches, loader_done, scanner_go;
};
constexpr uint64_t STREAM_POISON = ~0ull, STREAM_TAIL = ~0ull - 1;

template <typename PT>
constexpr size_t stream_smem_bytes() {
	return (size_t)StreamCfg<PT>::NT * VEC * sizeof(PT) + StreamCfg<PT>::RING + sizeof(StreamShared<PT>);
}
static_assert(stream_smem_bytes<double>() <= 232448 && stream_smem_bytes<float>() <= 232448, "227 KiB of shared memory per thread block");
static_assert(STREAM_B <= 32 && (StreamCfg<double>::W + STREAM_P + 2) * 32 <= 1024, "one lane per vector of a batch; at most 1024 threads");

// FFOR from the analysed tile into the ring entry (never in place: no deferred words, no fill select — exception slots were
// patched with the fill value when their originals were gathered)
__device__ __forceinline__ void stream_pack(const uint64_t* tile, uint64_t base, uint32_t bw, int t, uint8_t* dst) {
	const int       lane = t & 15, half = t >> 4;
	const uint32_t* lo32 = reinterpret_cast<const uint32_t*>(tile);
	dispatch_width<0, 64>(bw, [&](auto Wc) {
		constexpr int BW = decltype(Wc)::value;
		if constexpr (BW == 0) {
			return;
		} else if constexpr (BW <= 32) {
			pack64_rows<BW>(lane, half, reinterpret_cast<uint64_t*>(dst), [&](auto R, uint32_t& lo, uint32_t& hi) {
				constexpr int r = decltype(R)::value;
				lo              = lo32[2 * Map<double>::index(t, r)] - (uint32_t)base;  // (v - base) mod 2^32 is all a <= 32-bit field needs
				hi              = 0;
			});
		} else {
			pack64_rows<BW>(lane, half, reinterpret_cast<uint64_t*>(dst), [&](auto R, uint32_t& lo, uint32_t& hi) {
				constexpr int  r = decltype(R)::value;
				const uint64_t d = tile[Map<double>::index(t, r)] - base;  // masked to BW bits by the packer
				lo               = (uint32_t)d;
				hi               = (uint32_t)(d >> 32);
			});
		}
	});
}
__device__ __forceinline__ void stream_pack(const uint32_t* tile, uint32_t base, uint32_t bw, int t, uint8_t* dst) {
	dispatch_width<0, 32>(bw, [&](auto Wc) {
		constexpr int BW = decltype(Wc)::value;
		if constexpr (BW > 0) {
			pack32_rows<BW>(t, reinterpret_cast<uint32_t*>(dst), [&](auto R) -> uint32_t {
				constexpr int r = decltype(R)::value;
				return tile[Map<float>::index(t, r)] - base;
			});
		}
	});
}

template <typename PT, bool ORDERED>
__global__ void __launch_bounds__((StreamCfg<PT>::W + STREAM_P + 2) * 32, 1)
    encode_stream_kernel(const PT* __restrict__ in, uint64_t n_vectors, const alpb200_rg_state* __restrict__ states, ColOut col,
                         uint64_t* workspace) {
	using T  = Traits<PT>;
	using UT = typename T::UT;
	using C  = StreamCfg<PT>;
	constexpr int      W = C::W, NT = C::NT, B = STREAM_B, P = STREAM_P, NV = STREAM_NV, NBS = STREAM_NBS;
	constexpr uint32_t TILE = VEC * sizeof(PT), RING = C::RING;
	static_assert(NBS >= (NT + NV) / B + 2, "batch slots must outlive every sequence number in flight");
	// (twice: an entry that does not fit behind the head skips to the start of the ring, and the skipped bytes count as taken)
	static_assert(RING % 128 == 0 && RING >= 2 * (66 * 128 + 1024 * (sizeof(PT) + 2) + 128), "the ring must hold the largest entry twice");
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t*          tiles = smem;
	uint8_t*          ring  = smem + (size_t)NT * TILE;
	StreamShared<PT>& sh    = *reinterpret_cast<StreamShared<PT>*>(ring + RING);

	const int      warp      = threadIdx.x >> 5, t = threadIdx.x & 31;
	const uint64_t n_batches = (n_vectors + B - 1) / B;
	uint64_t*      aggregates = workspace + 2;
	uint64_t*      anchors    = aggregates + n_batches;

	if (threadIdx.x == 0) {
		for (int i = 0; i < NT; i++) {
			mbar_init(&sh.full[i], 1);
			mbar_init(&sh.empty[i], 1);
		}
		for (int i = 0; i < NV; i++) { mbar_init(&sh.done[i], 1); }
		for (int i = 0; i < STREAM_TURNS; i++) { mbar_init(&sh.turn[i], 1); }
		for (int i = 0; i < NBS; i++) {
			mbar_init(&sh.sized[i], 1);
			sh.b_arrived[i] = 0;
		}
		sh.next_seq = sh.alloc_seq = sh.alloc_head = sh.out_tail = sh.retired = sh.retired_batch = sh.alloc_blocked = sh.loader_batches = sh.loader_done = sh.scanner_go = 0;
		fence_mbar_init();
		mbar_arrive(&sh.turn[0]);  // sequence number 0 may allocate
	}
	__syncthreads();

	// ================================================================================================== loader
	if (warp == W) {
		if (t != 0) { return; }
		uint32_t seq    = 0, b = 0;
		uint64_t ticket = (uint64_t)atomicAdd(reinterpret_cast<unsigned long long*>(workspace), 1ull);
		sts_volatile(&sh.scanner_go, ORDERED && ticket == 0 ? 1u : 2u);
		while (ticket < n_batches) {
			// the next ticket is requested now: its round trip to L2 hides behind this batch's tile waits
			const uint64_t next = (uint64_t)atomicAdd(reinterpret_cast<unsigned long long*>(workspace), 1ull);
			sh.b_ticket[b % NBS] = (uint32_t)ticket;
			__threadfence_block();
			sts_volatile(&sh.loader_batches, b + 1);  // (the placers start their look-back on the ticket alone)
			for (int j = 0; j < B; j++, seq++) {
				const uint32_t slot = seq % NT;
				if (seq >= (uint32_t)NT) { mbar_wait(&sh.empty[slot], (seq / NT - 1) & 1); }
				const uint64_t v = ticket * B + j;
				if (v < n_vectors) {
					sh.slot_vec[slot] = v;
					__threadfence_block();
					mbar_arrive_expect_tx(&sh.full[slot], TILE);
					bulk_g2s(tiles + (size_t)slot * TILE, in + v * (uint64_t)VEC, TILE, &sh.full[slot]);
				} else {  // the ragged end of the last batch: a sequence number without a vector
					sh.slot_vec[slot] = STREAM_TAIL;
					__threadfence_block();
					mbar_arrive(&sh.full[slot]);
				}
			}
			b++;
			ticket = next;
		}
		for (int w = 0; w < W; w++, seq++) {  // one poison entry per compute warp
			const uint32_t slot = seq % NT;
			if (seq >= (uint32_t)NT) { mbar_wait(&sh.empty[slot], (seq / NT - 1) & 1); }
			sh.slot_vec[slot] = STREAM_POISON;
			__threadfence_block();
			mbar_arrive(&sh.full[slot]);
		}
		__threadfence_block();
		sts_volatile(&sh.loader_done, 1u);
		return;
	}
	// ================================================================================================== scanner
	if (warp == W + 1 + P) {
		if constexpr (ORDERED) {
			uint32_t go;
			while ((go = lds_volatile(&sh.scanner_go)) == 0) { __nanosleep(200); }
			if (go == 1) { scan_anchors(aggregates, anchors, (uint32_t)n_batches, t, workspace[1]); }
		}
		return;
	}
	// ================================================================================================== placers
	// Placer p serves the CTA's batches p, p + P, ...: the look-backs of consecutive batches overlap.  Within a batch lane j
	// serves vector j (its record, its bulk store, its 32-byte record); the exception lists are copied by the whole warp.
	// Ring entries are freed in batch order (retired_batch).
	if (warp > W) {
		const uint32_t p      = (uint32_t)(warp - (W + 1));
		uint32_t       widest = 0;
		StreamProf prof;
		for (uint32_t b = p;; b += P) {
			for (;;) {  // does batch b exist?
				if (lds_volatile(&sh.loader_batches) > b) { break; }
				if (lds_volatile(&sh.loader_done)) {
					if (lds_volatile(&sh.loader_batches) > b) { break; }
					goto finish;
				}
				__nanosleep(100);
			}
			{
				prof.lap(8, t);
				const uint32_t bslot  = b % NBS;
				const uint64_t ticket = lds_volatile(&sh.b_ticket[bslot]);
				// The look-back needs the ticket, not this batch's own sizes: it starts as soon as the batch is loaded and
				// polls the predecessors while the compute warps are still analysing this one.
				uint64_t excl = 0;
				if constexpr (ORDERED) { excl = ticket == 0 ? workspace[1] : lookback_prefix(aggregates, anchors, (uint32_t)ticket, t); }
				prof.lap(9, t);
				mbar_wait(&sh.sized[bslot], (b / NBS) & 1);
				prof.lap(10, t);
				const uint64_t agg = sh.b_agg[bslot];
				if constexpr (ORDERED) {
					if (ticket + 1 == n_batches && t == 0) {  // last batch: the column totals
						const uint64_t incl = excl + agg;
						col.totals[0]       = (incl >> AGG_SHIFT) * 128ull;
						col.totals[1]       = incl & AGG_EXC_MASK;
					}
				} else {
					if (t == 0) {
						excl = (uint64_t)atomicAdd(reinterpret_cast<unsigned long long*>(workspace + 1), (unsigned long long)agg);
						atomicAdd(reinterpret_cast<unsigned long long*>(&col.totals[0]), (unsigned long long)(agg >> AGG_SHIFT) * 128ull);
						atomicAdd(reinterpret_cast<unsigned long long*>(&col.totals[1]), (unsigned long long)(agg & AGG_EXC_MASK));
					}
					excl = shfl_u64(excl, 0);
				}
				uint64_t run  = excl;
				uint32_t from = 0;  // vectors of this batch already emitted
				while (from < (uint32_t)B) {
					// lane j waits for vector j.  Normally the whole batch is emitted at once; only when a vector of THIS batch
					// cannot get ring space (alloc_blocked) are the vectors before it emitted and freed first.
					const uint32_t seq  = b * B + (uint32_t)t, ds = seq % NV, par = (seq / NV) & 1;
					const bool     mine = (uint32_t)t >= from && t < B;
					uint32_t       upto = (uint32_t)B;
					for (;;) {
						const bool     ready = !mine || mbar_try_wait_ns(&sh.done[ds], par, 2000);
						const uint32_t late  = __ballot_sync(FULL, !ready);
						if (!late) { break; }
						const uint32_t blocked = lds_volatile(&sh.alloc_blocked);  // sequence number + 1 of an allocation waiting for space
						if (blocked > b * B + from + 1 && blocked <= b * B + (uint32_t)B) {
							upto = blocked - 1 - b * B;  // vectors [from, upto) hold entries and finish without anybody's help
							if (mine && (uint32_t)t < upto && !ready) { mbar_wait(&sh.done[ds], par); }
							__syncwarp();
							break;
						}
					}
					prof.lap(11, t);
					const bool     sel    = (uint32_t)t >= from && (uint32_t)t < upto;
					const uint32_t units  = sel ? sh.vrec[ds].units : 0u, cnt = sel ? sh.vrec[ds].cnt : 0u;
					const uint32_t off    = sel ? sh.vrec[ds].off : 0u, end = sel ? sh.vrec[ds].end : 0u;
					const bool     active = sel && sh.vrec[ds].active != 0;
					const uint64_t mine_agg = active ? (((uint64_t)units << AGG_SHIFT) | cnt) : 0ull;
					uint64_t       incl     = mine_agg;
#pragma unroll
					for (int d = 1; d < B; d <<= 1) {
						const uint64_t o = shfl_up_u64(incl, d);
						if (t >= d) { incl += o; }
					}
					const uint64_t my    = run + incl - mine_agg;
					run                  = run + shfl_u64(incl, B - 1);
					const uint32_t bytes = units * 128u;
					const uint64_t units_off = my >> AGG_SHIFT, exc_off = my & AGG_EXC_MASK;
					bool           ok    = active;
					if (active && (units_off * 128ull + bytes > col.packed_capacity || exc_off + cnt > col.exc_capacity)) {
						atomicExch(reinterpret_cast<unsigned long long*>(&col.totals[2]), 1ull);
						ok = false;
					}
					if (ok) {
						widest = max(widest, units);
						if (bytes) {
							bulk_s2g(col.packed + units_off * 128ull, ring + off, bytes);  // one contiguous write of the whole block (TMA 1-D)
							bulk_commit();
						}
						uint4* out = reinterpret_cast<uint4*>(col.meta + sh.vrec[ds].v);
						out[0]     = sh.vrec[ds].ra;
						out[1]     = make_uint4((uint32_t)units_off, (uint32_t)exc_off, sh.vrec[ds].z, sh.vrec[ds].w);
					}
					// the exception lists, vector by vector, 32 slots per step
					const uint32_t has_exc = __ballot_sync(FULL, ok && cnt != 0);
					for (uint32_t m = has_exc; m; m &= m - 1) {
						const int       jj    = __ffs((int)m) - 1;
						const uint32_t  c     = __shfl_sync(FULL, cnt, jj), by = __shfl_sync(FULL, bytes, jj), o = __shfl_sync(FULL, off, jj);
						const uint64_t  eo    = shfl_u64(exc_off, jj);
						const uint8_t*  entry = ring + o;
						const UT*       vals  = reinterpret_cast<const UT*>(entry + by);
						const uint16_t* poss  = reinterpret_cast<const uint16_t*>(entry + by + ((c * (uint32_t)sizeof(UT) + 15u) & ~15u));
						UT*             ev    = static_cast<UT*>(col.exc_val) + eo;
						uint16_t*       ep    = col.exc_pos + eo;
#pragma unroll 2
						for (uint32_t i = t; i < c; i += 32) {
							ev[i] = vals[i];
							ep[i] = poss[i];
						}
					}
					prof.lap(12, t);
					if (ok && bytes) { bulk_wait_read_all(); }  // the block image has been read
					__syncwarp();
					prof.lap(13, t);
					const uint32_t last_end = __shfl_sync(FULL, end, (int)upto - 1);
					__syncwarp();  // every lane has read its share of the exception lists
					if (t == 0) {  // free the entries, in batch order
						while (lds_volatile(&sh.retired_batch) != b) {}
						prof.lap(14, t);
						sts_volatile(&sh.out_tail, last_end);
						sts_volatile(&sh.retired, b * B + upto);
						if (upto == (uint32_t)B) {
							__threadfence_block();
							sts_volatile(&sh.retired_batch, b + 1);
						}
					}
					__syncwarp();
					from = upto;
				}
			}
		}
	finish:
		bulk_wait_all();
		widest = __reduce_max_sync(FULL, widest);
		if (t == 0 && widest) { atomicMax(reinterpret_cast<unsigned long long*>(&col.totals[3]), (unsigned long long)widest * 128ull); }
		return;
	}
	// ================================================================================================== compute warps
	StateRegs st;
	st.scheme      = ALPB200_SCHEME_INVALID;
	uint64_t st_rg = ~0ull;
	StreamProf prof;
	for (;;) {
		uint32_t seq = 0;
		prof.lap(0, t);
		if (t == 0) { seq = atomicAdd(&sh.next_seq, 1u); }
		seq                 = __shfl_sync(FULL, seq, 0);
		const uint32_t slot = seq % NT;
		mbar_wait(&sh.full[slot], (seq / NT) & 1);
		const uint64_t v = *reinterpret_cast<const volatile uint64_t*>(&sh.slot_vec[slot]);
		prof.lap(1, t);
		if (v == STREAM_POISON) { return; }
		const uint32_t b = seq / B, j = seq % B, bslot = b % NBS;
		const bool     active = v != STREAM_TAIL;
		UT*            tile   = reinterpret_cast<UT*>(tiles + (size_t)slot * TILE);

		Analysis<PT> a;
		a.cnt = a.bw = a.e = a.f = a.myexc = 0;
		a.base = a.fill = 0;
		uint32_t units  = 0;
		bool     rd     = false;
		if (active) {
			const uint64_t rg = v / ALPB200_ROWGROUP_VECTORS;
			if (rg != st_rg) {
				st    = load_state(states + rg);
				st_rg = rg;
			}
			rd = st.scheme == ALPB200_SCHEME_ALP_RD;
			TileIO<PT, true> io(tile, t);
			if (rd) {
				analyze_rd<PT>(states + rg, st, t, io, a, [](int, uint32_t) {});
			} else {
				analyze_alp<PT>(in + v * (uint64_t)VEC, st, t, io, a);
			}
			units = rd ? a.bw + a.e : a.bw;
		}
		prof.lap(2, t);
		// ---- report the sizes; the last reporter of a batch publishes its aggregate ----
		bool last = false;
		if (t == 0) {
			sh.b_info[bslot][j] = (units << 16) | a.cnt;
			__threadfence_block();
			last = atomicAdd(&sh.b_arrived[bslot], 1u) == (uint32_t)(B - 1);
		}
		last = __shfl_sync(FULL, (int)last, 0) != 0;
		if (last) {
			__threadfence_block();
			uint64_t mine_agg = 0;
			if (t < B) {
				const uint32_t info = lds_volatile(&sh.b_info[bslot][t]);
				mine_agg            = ((uint64_t)(info >> 16) << AGG_SHIFT) | (info & 0xFFFFu);
			}
			const uint64_t agg = warp_sum_u64(mine_agg);
			if (t == 0) {
				if constexpr (ORDERED) { st_volatile_u64(&aggregates[lds_volatile(&sh.b_ticket[bslot])], SCAN_VALID | agg); }
				sh.b_agg[bslot]     = agg;
				sh.b_arrived[bslot] = 0;
				mbar_arrive(&sh.sized[bslot]);
			}
		}
		// ---- the entry in the output ring: [block image][exception values][exception positions], in sequence order ----
		const uint32_t bytes   = units * 128u;
		const uint32_t val_len = (a.cnt * (uint32_t)sizeof(UT) + 15u) & ~15u;
		const uint32_t E       = (bytes + val_len + 2u * a.cnt + 127u) & ~127u;
		uint32_t       off = 0, end = 0;
		prof.lap(3, t);
		if (t == 0) {
			mbar_wait(&sh.turn[seq % STREAM_TURNS], (seq / STREAM_TURNS) & 1);
			prof.lap(4, t);
			uint32_t head = lds_volatile(&sh.alloc_head);
			uint32_t phys = head % RING;
			if (phys + E > RING) {
				head += RING - phys;
				phys = 0;
			}
			if (head + E - lds_volatile(&sh.out_tail) > RING || seq - lds_volatile(&sh.retired) >= (uint32_t)NV) {
				sts_volatile(&sh.alloc_blocked, seq + 1);  // tells the placers that entries before this one must be freed first
				while (head + E - lds_volatile(&sh.out_tail) > RING || seq - lds_volatile(&sh.retired) >= (uint32_t)NV) {}
				sts_volatile(&sh.alloc_blocked, 0u);
			}
			off = phys;
			end = head + E;
			sts_volatile(&sh.alloc_head, end);
			mbar_arrive(&sh.turn[(seq + 1) % STREAM_TURNS]);  // (release: the next allocation sees alloc_head)
		}
		off            = __shfl_sync(FULL, off, 0);
		prof.lap(5, t);
		uint8_t* entry = ring + off;
		if (active) {
			// exceptions: ranks from the bit-matrix transpose; positions filed at their ranks, then the originals gathered from
			// the tile (ALP_RD: the left parts) and, for ALP, the slots patched with the fill value for the packer
			const ExcPlan plan = plan_exceptions<PT>(a.myexc, t);
			if (plan.any) {
				UT*       vals = reinterpret_cast<UT*>(entry + bytes);
				uint16_t* poss = reinterpret_cast<uint16_t*>(entry + bytes + val_len);
				for_each_exception<PT>(plan, a.myexc, t, [&](uint32_t rank, uint32_t p) { poss[rank] = (uint16_t)p; });
				__syncwarp();
				const uint32_t rbw = a.bw;
				for (uint32_t i = t; i < a.cnt; i += 32) {
					const uint32_t p    = poss[i];
					const UT       bits = tile[p];
					vals[i]             = rd ? (UT)(bits >> rbw) : bits;
					if (!rd) { tile[p] = (UT)a.fill; }
				}
			}
			__syncwarp();  // the tile is complete: encoded integers / patched slots were stored lane by lane
			stream_pack(tile, (UT)a.base, a.bw, t, entry);
			if (rd) { pack_left(a.left_nib, a.e, t, reinterpret_cast<uint16_t*>(entry + 128u * a.bw), PT()); }
		}
		fence_proxy_async_smem();  // the entry (generic-proxy writes) is read, the tile overwritten, by the bulk-copy engine
		__syncwarp();
		prof.lap(6, t);
		if (t == 0) {
			StreamVRec& rec = sh.vrec[seq % NV];
			if (rd) {
				rec.ra = st.dict;
			} else {
				const int64_t bb = (int64_t)a.base;
				rec.ra           = make_uint4((uint32_t)bb, (uint32_t)((uint64_t)bb >> 32), 0u, 0u);
			}
			rec.v      = v;
			rec.z      = a.cnt | (st.scheme << 16) | (a.bw << 24);
			rec.w      = a.e | (a.f << 8);
			rec.units  = units;
			rec.cnt    = a.cnt;
			rec.off    = off;
			rec.end    = end;
			rec.active = active ? 1u : 0u;
			mbar_arrive(&sh.empty[slot]);
			mbar_arrive(&sh.done[seq % NV]);
		}
	}
}

}  // namespace alpb200
