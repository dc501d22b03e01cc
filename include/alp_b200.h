/*
 * alp_b200.h — C ABI of the B200-native ALP / ALP_RD column codec.
 *
 * This is the drop-in boundary.  The reference (cwida/ALP) has no C ABI:
 * it is a set of header-only C++ templates plus libALP.so with C++-mangled ffor/unffor/falp
 * overloads.  Each entry point below names the reference primitive(s) it replaces
 * (file:line relative to the reference tree).  `include/alp_b200.hpp` re-creates the
 * reference's template signatures on top of this ABI.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - `d_` pointers are device memory, `h_` pointers are host memory
 *   - every function returns 0 on success or a negative ALPB200_E* code; the message of the
 *     last failure on the calling thread is available from alpb200_last_error()
 *   - the library never allocates on the batched hot path: the caller owns all buffers
 *   - vector = 1024 values, row-group = 100 vectors (reference include/alp/config.hpp:11-15)
 */
#ifndef ALP_B200_H
#define ALP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALPB200_VECTOR_SIZE 1024u          /* config.hpp:11 VECTOR_SIZE */
#define ALPB200_ROWGROUP_VECTORS 100u      /* config.hpp:13 N_VECTORS_PER_ROWGROUP */
#define ALPB200_ROWGROUP_SIZE 102400u      /* config.hpp:15 ROWGROUP_SIZE */
#define ALPB200_ROWGROUP_SAMPLES_JUMP 12u  /* config.hpp:17-19: (ROWGROUP_SIZE / 8) / VECTOR_SIZE: every 12th vector is sampled */
#define ALPB200_MAX_K 5                    /* config.hpp:22 MAX_K_COMBINATIONS */
#define ALPB200_RD_DICT_SIZE 8             /* config.hpp:25 MAX_RD_DICTIONARY_SIZE */
#define ALPB200_MAX_SAMPLES 288            /* 9 sampled vectors x 32 values (sampler.hpp:14-52) */

/* alp::Scheme (constants.hpp:10-14) — same numeric values */
#define ALPB200_SCHEME_INVALID 0
#define ALPB200_SCHEME_ALP_RD 1
#define ALPB200_SCHEME_ALP 2

#define ALPB200_OK 0
#define ALPB200_EINVAL (-1)   /* bad argument (null pointer, bw > lane width, misaligned buffer, ...) */
#define ALPB200_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define ALPB200_ECAPACITY (-3)/* an output buffer of the column container is too small */
#define ALPB200_ENODEVICE (-4)/* no CUDA device: there is no CPU fallback */

/*
 * Row-group state: the POD image of alp::state<PT> (encoder.hpp:35-62) that the encoder needs.
 * Produced by alpb200_rowgroup_init_* (device) or by the caller (e.g. from the reference's own init).
 */
typedef struct alpb200_rg_state {
	int32_t  scheme;                         /* ALPB200_SCHEME_* */
	int32_t  k;                              /* state.k_combinations, 1..5 (ALP) */
	uint8_t  combos[ALPB200_MAX_K][2];       /* state.best_k_combinations: {exponent, factor} */
	uint8_t  right_bw;                       /* state.right_bit_width (ALP_RD) */
	uint8_t  left_bw;                        /* state.left_bit_width  (ALP_RD) */
	uint8_t  dict_size;                      /* state.actual_dictionary_size */
	uint8_t  reserved0[3];
	uint16_t dict[ALPB200_RD_DICT_SIZE];     /* state.left_parts_dict */
	/* state.left_parts_dict_map entries that are NOT in the dictionary (rd.hpp:73-77): the index the
	 * reference stores for such a left part before FFOR masks it to left_bw bits.  Only needed for byte
	 * parity of the packed left stream at exception positions; n_extra = 0 is always valid
	 * (then the stored index is dict_size, as rd.hpp:129-131 does for unseen keys). */
	uint16_t n_extra;
	uint16_t reserved1;
	uint16_t extra_key[ALPB200_MAX_SAMPLES];
	uint16_t extra_idx[ALPB200_MAX_SAMPLES];
} alpb200_rg_state;

/*
 * Per-vector record of the column container (32 bytes, one 32-B sector per vector).
 * It carries what the reference's callers keep per vector — bw, base, factor, exponent, exception
 * count (test/test_alp_sample.cpp:6-7; publication/.../bench_end_to_end/include/encoding/helper.hpp:36-67
 * `alp_m`) — plus the offsets of the vector's packed block and exception run.
 */
typedef struct alpb200_vec_meta {
	union {
		struct {
			int64_t  base;                   /* FOR base (sign-extended for f32) — analyze_ffor, encoder.hpp:109-120 */
			uint64_t reserved;
		} alp;
		uint16_t rd_dict[ALPB200_RD_DICT_SIZE]; /* ALP_RD: the row-group's left-part dictionary (rd.hpp:63-71) */
	} u;
	uint32_t packed_off;   /* offset of the packed block in `packed`, in units of 128 bytes.
	                          ALP: 128*bw bytes.  ALP_RD: right block (128*bw bytes) then left block (128*e bytes) */
	uint32_t exc_off;      /* first exception slot of this vector in exc_val / exc_pos */
	uint16_t exc_cnt;      /* exceptions_count */
	uint8_t  scheme;       /* ALPB200_SCHEME_ALP or ALPB200_SCHEME_ALP_RD */
	uint8_t  bw;           /* ALP: FFOR bit width.  ALP_RD: right_bit_width */
	uint8_t  e;            /* ALP: exponent index.  ALP_RD: left_bit_width */
	uint8_t  f;            /* ALP: factor index.    ALP_RD: actual_dictionary_size */
	uint8_t  reserved[2];
} alpb200_vec_meta;

/*
 * Column container (struct of arrays; device pointers for the batched entry points, host pointers for
 * the *_host entry points and for the CPU oracle).  A vector's packed block starts at packed + 128 * packed_off,
 * its exceptions are exc_val[exc_off + i] / exc_pos[exc_off + i], i < exc_cnt, positions ascending.  Blocks and
 * exception runs are dense; alpb200_encode_* places them in VECTOR ORDER (any vector range is one byte range),
 * alpb200_encode_unordered_* in completion order — readers must go through the offsets, and every reader in this
 * library does.  exc_val elements have the width of the column
 * type (8 bytes f64, 4 bytes f32); an ALP_RD vector stores its 16-bit left-part exceptions zero-extended
 * in the same slots.
 */
typedef struct alpb200_column {
	uint64_t          n_vectors;
	alpb200_vec_meta* meta;            /* [n_vectors] */
	uint8_t*          packed;          /* 128-byte aligned */
	uint64_t          packed_capacity; /* bytes */
	void*             exc_val;         /* [exc_capacity] of the column's value width */
	uint16_t*         exc_pos;         /* [exc_capacity] */
	uint64_t          exc_capacity;    /* exception slots */
	/* [0] packed bytes used, [1] exception slots used, [2] non-zero if a capacity was exceeded, [3] largest packed
	 * block of any vector in bytes (written by encode; device memory for the device entry points) */
	uint64_t*         totals;
	/* decode hint: the largest packed block of any vector in bytes (totals[3] after encode), or 0 when unknown —
	 * the decoder then sizes its shared-memory stages for the widest possible block */
	uint64_t          max_block_bytes;
	/* number of values the column holds; when it is not a multiple of 1024 the last vector was padded by the encoder
	 * (alpb200_compress_host_*) and only n_values are returned by alpb200_decompress_host_*.  0 = n_vectors * 1024.
	 * (The reference's drivers drop the tail, benchmarks/benchmark.cpp:191; PRIMITIVES.md:141-144 suggests padding.) */
	uint64_t          n_values;
} alpb200_column;

int         alpb200_version(void);
/* sizeof(alpb200_rg_state), sizeof(alpb200_vec_meta), sizeof(alpb200_column) as compiled into the library */
void        alpb200_abi_sizes(uint32_t out[3]);
const char* alpb200_last_error(void);
/* number of CUDA devices visible to the library; <0 on error.  There is no CPU fallback. */
int         alpb200_device_count(void);

/* ------------------------------------------------------------------------------------------------
 * Batched device entry points (the hot path).  One warp per 1024-value vector.
 * ---------------------------------------------------------------------------------------------- */

/* Row-group init: first-level sampling + top-k (e,f) search + ALP/ALP_RD decision + RD dictionary.
 * Replaces alp::encoder<PT>::init (encoder.hpp:420-427), sampler::first_level_sample (sampler.hpp:14-52),
 * find_top_k_combinations (encoder.hpp:139-235) and alp::rd_encoder<PT>::init / find_best_dictionary
 * (rd.hpp:89-104,180-185).  n_values must be a multiple of 1024; d_states has ceil(n_vectors/100) entries. */
int alpb200_rowgroup_init_f64(const double* d_in, uint64_t n_values, alpb200_rg_state* d_states, void* d_workspace,
                              void* stream);
int alpb200_rowgroup_init_f32(const float* d_in, uint64_t n_values, alpb200_rg_state* d_states, void* d_workspace,
                              void* stream);
/* Bytes of scratch alpb200_rowgroup_init_* needs for n_values (per-sampled-vector search results). */
size_t alpb200_init_workspace_bytes(uint64_t n_values);

/* Tail vector and NULLs (reference: prose only, PRIMITIVES.md:141-144 "Last Vector Encoding"; its drivers drop the tail,
 * benchmarks/benchmark.cpp:191).  The codec works on whole vectors of 1024 real values; slots that hold no data get a
 * filler on the device before the encoder sees them: the slots behind n_values up to the next multiple of 1024 (d_values
 * must have room for them) and, with a validity bitmap (Arrow layout: bit i & 7 of byte i >> 3 set = value i is valid;
 * 4-byte aligned; NULL = no NULLs), every NULL slot.
 *   d_states == NULL  first strategy: the vector's first valid value (0.0 for a vector without any).  Call it BEFORE
 *                     alpb200_rowgroup_init_*, so that the sampling only sees real values.
 *   d_states given    second strategy: the vector's first valid NON-EXCEPTION value under the state the encoder will use —
 *                     the slot then costs no exception and cannot widen the block.  Call it between init and encode.
 * After it the buffer is an ordinary column of ceil(n_values / 1024) vectors for alpb200_rowgroup_init_* / alpb200_encode_*
 * (and for the reference, which fed the same buffer produces the same bytes); alpb200_decode_values_* reads exactly n_values
 * back.  alpb200_compress_host_* does all of this for a column's partial last vector. */
int alpb200_fill_invalid_f64(double* d_values, uint64_t n_values, const uint8_t* d_validity, const alpb200_rg_state* d_states,
                             void* stream);
int alpb200_fill_invalid_f32(float* d_values, uint64_t n_values, const uint8_t* d_validity, const alpb200_rg_state* d_states,
                             void* stream);

/* Bytes of scratch alpb200_encode_* needs for n_vectors (decoupled look-back state). */
size_t alpb200_encode_workspace_bytes(uint64_t n_vectors);

/* Encode n_vectors vectors of d_in into the column (vector i uses d_states[i / 100]).
 * Replaces, per vector, alp::encoder<PT>::encode (encoder.hpp:402-418) + analyze_ffor (encoder.hpp:109-120)
 * + ffor::ffor (src/fastlanes_generated_ffor.cpp:29939), or for ALP_RD row-groups
 * alp::rd_encoder<PT>::encode (rd.hpp:109-147) + 2x ffor::ffor — i.e. test/test_alp_sample.cpp:141-145,164-166. */
int alpb200_encode_f64(const double* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states,
                       const alpb200_column* col, void* d_workspace, void* stream);
int alpb200_encode_f32(const float* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states,
                       const alpb200_column* col, void* d_workspace, void* stream);
/* The general form.  flags: ALPB200_ENCODE_UNORDERED (completion-order layout, see alpb200_encode_unordered_*) and
 * ALPB200_ENCODE_APPEND: the call CONTINUES a column — col->meta points at the record of the call's first vector (which
 * must start a row-group; d_in and d_states point at that vector's values and row-group state), col->packed / exc_val /
 * exc_pos / totals are those of the whole column, and the output starts where col->totals (device memory, left by the
 * previous call on the same stream) says the column ends; totals are updated.  Appending calls give byte for byte the
 * column one call over all vectors gives (vector-order layout) — this is how alpb200_compress_host_* overlaps the
 * encode of one chunk with the transfers of its neighbours. */
#define ALPB200_ENCODE_UNORDERED 1u
#define ALPB200_ENCODE_APPEND 2u
int alpb200_encode_ex_f64(const double* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states,
                          const alpb200_column* col, void* d_workspace, void* stream, uint32_t flags);
int alpb200_encode_ex_f32(const float* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states,
                          const alpb200_column* col, void* d_workspace, void* stream, uint32_t flags);

/* Same encoder, COMPLETION-ORDER layout: the same per-vector blocks, exception runs and records, dense in the same
 * arrays, but handed out with one atomic per thread block instead of an in-order prefix, so blocks sit in the order
 * their thread blocks finished (roughly, not exactly, vector order) and the bytes of the column are not a pure function
 * of the input.  Every consumer in this library reads columns through the per-vector offsets and accepts both layouts.
 * Measured on B200 (2^29 f64 values): 1.15 ms vs 1.54 ms — the in-order wait is 25-30 % of the ordered kernel.
 * The reference has no container at all (its callers place blocks wherever they like, benchmarks/benchmark.cpp:200-285),
 * so neither layout is "the reference's"; alpb200_encode_* is the default because reproducible bytes and contiguous
 * vector ranges are worth having. */
int alpb200_encode_unordered_f64(const double* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states,
                                 const alpb200_column* col, void* d_workspace, void* stream);
int alpb200_encode_unordered_f32(const float* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states,
                                 const alpb200_column* col, void* d_workspace, void* stream);

/* Decode vectors [first_vector, first_vector + n_vectors) of the column into d_out (1024 values each).
 * Replaces generated::falp::fallback::scalar::falp (src/falp.cpp:42440,42644) + alp::decoder<PT>::patch_exceptions
 * (decoder.hpp:141-149), or for ALP_RD vectors 2x unffor::unffor + alp::rd_encoder<PT>::decode (rd.hpp:152-178)
 * — i.e. test/test_alp_sample.cpp:148-151,169-170. */
int alpb200_decode_f64(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, double* d_out,
                       void* stream);
int alpb200_decode_f32(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, float* d_out,
                       void* stream);

/* Decode exactly n_values values starting at vector first_vector: whole vectors go straight to d_out (which then needs
 * room for n_values values only), a partial last vector is decoded into d_scratch (1024 values; may be NULL when n_values is
 * a multiple of 1024) and its first n_values % 1024 values are copied behind them.  This is how a column whose length is
 * not a multiple of 1024 (see alpb200_fill_invalid_*) is read back without its padding. */
int alpb200_decode_values_f64(const alpb200_column* col, uint64_t first_vector, uint64_t n_values, double* d_out, double* d_scratch,
                              void* stream);
int alpb200_decode_values_f32(const alpb200_column* col, uint64_t first_vector, uint64_t n_values, float* d_out, float* d_scratch,
                              void* stream);

/* The batched decoders assume a well-formed column (as the reference's primitives do: undefined behaviour on bad input,
 * SURVEY.md section 8b).  What they guarantee regardless: a stale / wrong max_block_bytes hint is harmless — whenever a hint is
 * given, the records of the call are checked against it on the device first and the call falls back to a slow, generic
 * path if a block outgrows the shared-memory stage the hint sized.  A column of unknown provenance is checked with this
 * call first: one pass over the records on the device (scheme, widths, exponent / factor, exception
 * count, block and exception run inside the arrays given by packed_capacity / exc_capacity, every exception position
 * < 1024).  Synchronises `stream`.  ALPB200_EINVAL for a malformed column; on success *h_max_block_bytes (may be NULL)
 * receives the widest block, i.e. the decode hint. */
int alpb200_column_validate_device(const alpb200_column* col, int value_bytes, uint64_t* h_max_block_bytes, void* stream);

/* Fused decode + SUM (no decoded column is written): *d_sum += sum of the decoded values of vectors
 * [first_vector, first_vector + n_vectors), accumulated in double.  The caller zeroes *d_sum.  Floating-point
 * addition order is not fixed (per-thread partial sums, one atomic add per warp).  Mirrors the reference's scan
 * primitive `alp_func` + `aggr_plus`
 * (publication/source_code/bench_end_to_end/src/benchmarks/alp/queries/q1.cpp:63-102). */
int alpb200_decode_sum_f64(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, double* d_sum,
                           void* stream);
int alpb200_decode_sum_f32(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, double* d_sum,
                           void* stream);
/* The same with flags.  ALPB200_SUM_DECIMAL (floats; doubles always work this way, where it is at least as accurate as adding
 * the decoded doubles): a vector's non-exception slots are added as INTEGERS and converted once, X * 10^f * 10^-e in double —
 * the sum of the decimals the floats stand for instead of the sum of the floats.  The two differ by the floats' own rounding:
 * |difference| <= 2^-23 * sum |x_i|.  About 2x faster on float columns (3 instead of 8 instructions per value).  Vectors whose
 * integers times 10^f could leave int32 (where the reference's 32-bit product wraps) keep the per-value path. */
#define ALPB200_SUM_DECIMAL 1u
int alpb200_decode_sum_ex_f64(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, double* d_sum, uint32_t flags,
                              void* stream);
int alpb200_decode_sum_ex_f32(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, double* d_sum, uint32_t flags,
                              void* stream);

/* Fused decode + MIN / MAX / COUNT (no decoded column is written): the scan-side siblings of SUM (zone maps, filters).
 * *d_out (device memory, overwritten) receives the minimum and maximum of the decoded values of vectors [first_vector,
 * first_vector + n_vectors) as doubles, NaNs ignored, and the number of non-NaN values; an empty range or an all-NaN range
 * gives min = +inf, max = -inf, count = 0.  Semantics = decode + reduce (the reference ships no such scan; its scan query is
 * SUM only, q1.cpp:63-102): every vector is decoded and patched like alpb200_decode_* does, in shared memory. */
typedef struct alpb200_minmax {
	double   min;
	double   max;
	uint64_t count;
} alpb200_minmax;
int alpb200_decode_minmax_f64(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, alpb200_minmax* d_out,
                              void* stream);
int alpb200_decode_minmax_f32(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, alpb200_minmax* d_out,
                              void* stream);

/* Fused decode + predicate filter: a selection bitmap instead of the decoded column.  Bit j of d_bitmap[32 v + w] (v counted
 * from first_vector) is set iff value 32 w + j of vector first_vector + v satisfies `value op constant` (IEEE comparison in
 * double: a NaN satisfies only ALPB200_FILTER_NE; float columns compare (double)value).  d_bitmap: 32 * n_vectors words
 * (128 bytes per vector), 4-byte aligned; *d_selected (device memory, may be NULL, overwritten) receives the number of set
 * bits.  Semantics = decode + compare: every vector is decoded and patched in shared memory like alpb200_decode_* does. */
#define ALPB200_FILTER_LT 0u
#define ALPB200_FILTER_LE 1u
#define ALPB200_FILTER_GT 2u
#define ALPB200_FILTER_GE 3u
#define ALPB200_FILTER_EQ 4u
#define ALPB200_FILTER_NE 5u
int alpb200_decode_filter_f64(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, uint32_t op, double constant,
                              uint32_t* d_bitmap, uint64_t* d_selected, void* stream);
int alpb200_decode_filter_f32(const alpb200_column* col, uint64_t first_vector, uint64_t n_vectors, uint32_t op, double constant,
                              uint32_t* d_bitmap, uint64_t* d_selected, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer entry points (what a host engine calls; copies are part of the call).
 * A codec context owns device staging buffers, a small pinned area and three streams so that repeated
 * calls do not allocate; every call pipelines the column in ~16 chunks (transfers of neighbouring chunks overlap the
 * kernels of the current one) and returns when the result is complete.  A context serves ONE call at a time: use one
 * context per host thread (they may share a device).
 * ---------------------------------------------------------------------------------------------- */
typedef struct alpb200_ctx alpb200_ctx;

/* max_vectors: the largest column (in vectors) a call will pass (at most 2^22 = 2^32 values); value_bytes: 8 or 4.
 * The context's device staging is sized for the worst case a column can reach (every vector ALP_RD at full width with 1024
 * exceptions: ~27 bytes per f64 value).  alpb200_ctx_create_ex takes the caller's own bound on the compressed column instead
 * — packed bytes and exception slots, 0 = worst case; a column that outgrows them fails with ALPB200_ECAPACITY. */
int  alpb200_ctx_create(alpb200_ctx** out, int device, uint64_t max_vectors, int value_bytes);
int  alpb200_ctx_create_ex(alpb200_ctx** out, int device, uint64_t max_vectors, int value_bytes, uint64_t packed_capacity,
                           uint64_t exc_capacity);
void alpb200_ctx_destroy(alpb200_ctx* ctx);
/* Context options.  ALPB200_OPT_UNORDERED (0 | 1, default 0): alpb200_compress_host_* encodes with the
 * completion-order layout (alpb200_encode_unordered_*).  ALPB200_OPT_CHUNKS (1..16, default 16): how many chunks the
 * host entry points cut a column into (transfers of neighbouring chunks overlap the kernels of the current one). */
#define ALPB200_OPT_UNORDERED 1
#define ALPB200_OPT_CHUNKS 2
int  alpb200_ctx_set_option(alpb200_ctx* ctx, int option, int value);

/* Compress a host column of n_values values (any length: the slots of a partial last vector are filled on the device with
 * that vector's first non-exception value, alpb200_fill_invalid_*, and n_values is recorded in the container) into a host column container whose arrays the
 * caller allocated (capacities in h_col; n_vectors >= ceil(n_values / 1024)).  H2D of the values, init, encode,
 * D2H of the compressed arrays.  h_col->totals[0..1] receive packed bytes / exception slots used. */
int alpb200_compress_host_f64(alpb200_ctx* ctx, const double* h_in, uint64_t n_values, alpb200_column* h_col);
int alpb200_compress_host_f32(alpb200_ctx* ctx, const float* h_in, uint64_t n_values, alpb200_column* h_col);
/* Decompress a host column container into h_out: H2D of the compressed arrays, decode, D2H of the values,
 * pipelined in chunks of vectors over two streams. */
int alpb200_decompress_host_f64(alpb200_ctx* ctx, const alpb200_column* h_col, double* h_out);
int alpb200_decompress_host_f32(alpb200_ctx* ctx, const alpb200_column* h_col, float* h_out);
/* SUM of a host column container without materialising it anywhere: H2D of the compressed arrays, fused decode + SUM
 * on the device (alpb200_decode_sum_*), D2H of the one double — the reference's end-to-end scan query (alp_func +
 * aggr_plus under TBB workers, publication/source_code/bench_end_to_end/src/benchmarks/alp/queries/q1.cpp:63-102,650-679)
 * as one call.  Only the compressed bytes cross PCIe.  A padded tail (n_values < n_vectors * 1024) is excluded. */
int alpb200_sum_host_f64(alpb200_ctx* ctx, const alpb200_column* h_col, double* h_sum);
int alpb200_sum_host_f32(alpb200_ctx* ctx, const alpb200_column* h_col, double* h_sum);
/* Checks a HOST column container the way alpb200_decompress_host_* / alpb200_sum_host_* do before they launch anything:
 * every record well formed (scheme ALP or ALP_RD; bw <= lane width, resp. lane width - 16 <= right width < lane width;
 * exponent / factor / left width / dictionary size in range; at most 1024 exceptions, every position < 1024) and pointing
 * inside the container's arrays.  ALPB200_EINVAL otherwise — a damaged or truncated stored column is rejected, not decoded.
 * (The device entry points cannot afford that pass; their kernels instead clamp whatever a record says so that nothing
 * outside the column's arrays or the output vector is ever touched.) */
int alpb200_column_validate_host(const alpb200_column* h_col, int value_bytes);
/* Page-locked host memory for the buffers handed to the *_host entry points (pageable buffers also work, slower). */
void* alpb200_host_alloc(size_t bytes);
void  alpb200_host_free(void* p);

/* ------------------------------------------------------------------------------------------------
 * Single-vector primitives with HOST pointers: the reference's primitive API (PRIMITIVES.md) one call at a
 * time.  Each call is a 1-vector batch (H2D, one kernel, D2H) and exists for drop-in / parity use, not speed.
 * ---------------------------------------------------------------------------------------------- */

/* alp::encoder<PT>::encode (encoder.hpp:402-418).  e/f receive stt.exp / stt.fac. */
int alpb200_prim_encode_f64(const double* h_in, const alpb200_rg_state* h_state, double* h_exc, uint16_t* h_pos,
                            uint16_t* h_cnt, int64_t* h_enc, uint8_t* e, uint8_t* f);
int alpb200_prim_encode_f32(const float* h_in, const alpb200_rg_state* h_state, float* h_exc, uint16_t* h_pos,
                            uint16_t* h_cnt, int32_t* h_enc, uint8_t* e, uint8_t* f);
/* alp::encoder<PT>::analyze_ffor (encoder.hpp:109-120) */
int alpb200_prim_analyze_ffor_i64(const int64_t* h_enc, uint8_t* bw, int64_t* base);
int alpb200_prim_analyze_ffor_i32(const int32_t* h_enc, uint8_t* bw, int32_t* base);
/* ffor::ffor (include/fastlanes/ffor.hpp:7-15) / unffor::unffor (include/fastlanes/unffor.hpp:7-15).
 * h_out of ffor receives 128*bw bytes; bw > lane width returns ALPB200_EINVAL (the reference silently no-ops). */
int alpb200_prim_ffor_u64(const uint64_t* h_in, uint64_t* h_out, uint8_t bw, uint64_t base);
int alpb200_prim_ffor_u32(const uint32_t* h_in, uint32_t* h_out, uint8_t bw, uint32_t base);
int alpb200_prim_ffor_u16(const uint16_t* h_in, uint16_t* h_out, uint8_t bw, uint16_t base);
int alpb200_prim_ffor_u8(const uint8_t* h_in, uint8_t* h_out, uint8_t bw, uint8_t base);
int alpb200_prim_unffor_u64(const uint64_t* h_in, uint64_t* h_out, uint8_t bw, uint64_t base);
int alpb200_prim_unffor_u32(const uint32_t* h_in, uint32_t* h_out, uint8_t bw, uint32_t base);
int alpb200_prim_unffor_u16(const uint16_t* h_in, uint16_t* h_out, uint8_t bw, uint16_t base);
int alpb200_prim_unffor_u8(const uint8_t* h_in, uint8_t* h_out, uint8_t bw, uint8_t base);
/* generated::falp::fallback::scalar::falp (include/alp/falp.hpp:10-44): fused unffor + decode.
 * At bw == lane width the reference's fused kernel is wrong (src/falp.cpp:33311-33319); this entry point has the
 * unfused semantics unffor + decoder::decode for every width. */
int alpb200_prim_falp_f64(const uint64_t* h_packed, double* h_out, uint8_t bw, uint64_t base, uint8_t f, uint8_t e);
int alpb200_prim_falp_f32(const uint32_t* h_packed, float* h_out, uint8_t bw, uint32_t base, uint8_t f, uint8_t e);
/* alp::decoder<PT>::decode (decoder.hpp:134-138) */
int alpb200_prim_decode_f64(const int64_t* h_enc, uint8_t f, uint8_t e, double* h_out);
int alpb200_prim_decode_f32(const int32_t* h_enc, uint8_t f, uint8_t e, float* h_out);
/* alp::decoder<PT>::patch_exceptions (decoder.hpp:141-149): h_out is read, patched on the device, written back */
int alpb200_prim_patch_f64(double* h_out, const double* h_exc, const uint16_t* h_pos, uint16_t cnt);
int alpb200_prim_patch_f32(float* h_out, const float* h_exc, const uint16_t* h_pos, uint16_t cnt);
/* alp::rd_encoder<PT>::encode (rd.hpp:109-147) / decode (rd.hpp:152-178) */
int alpb200_prim_rd_encode_f64(const double* h_in, const alpb200_rg_state* h_state, uint16_t* h_exc, uint16_t* h_pos,
                               uint16_t* h_cnt, uint64_t* h_right, uint16_t* h_left);
int alpb200_prim_rd_encode_f32(const float* h_in, const alpb200_rg_state* h_state, uint16_t* h_exc, uint16_t* h_pos,
                               uint16_t* h_cnt, uint32_t* h_right, uint16_t* h_left);
int alpb200_prim_rd_decode_f64(double* h_out, const uint64_t* h_right, const uint16_t* h_left, const uint16_t* h_exc,
                               const uint16_t* h_pos, uint16_t cnt, const alpb200_rg_state* h_state);
int alpb200_prim_rd_decode_f32(float* h_out, const uint32_t* h_right, const uint16_t* h_left, const uint16_t* h_exc,
                               const uint16_t* h_pos, uint16_t cnt, const alpb200_rg_state* h_state);
/* alp::encoder<PT>::init (+ rd_encoder<PT>::init when the row-group falls to ALP_RD) on a host column:
 * the values [offset, min(offset+102400, n_values)) form the row-group.  Whole vectors only: a partial last vector is not
 * sampled (the reference's sampler takes it when it is the row-group's first vector or holds at least 32 values,
 * sampler.hpp:30-47) and a row-group without one complete vector is ALPB200_EINVAL — pad the tail first
 * (alpb200_fill_invalid_*), as the batched path does. */
int alpb200_prim_init_f64(const double* h_col, uint64_t offset, uint64_t n_values, alpb200_rg_state* h_state);
int alpb200_prim_init_f32(const float* h_col, uint64_t offset, uint64_t n_values, alpb200_rg_state* h_state);
/* alp::rd_encoder<PT>::init (rd.hpp:180-185) on its own: the row-group is made an ALP_RD row-group whatever the ALP search
 * would have said — cut position, dictionary and exception indices from the same first-level sample. */
int alpb200_prim_rd_init_f64(const double* h_col, uint64_t offset, uint64_t n_values, alpb200_rg_state* h_state);
int alpb200_prim_rd_init_f32(const float* h_col, uint64_t offset, uint64_t n_values, alpb200_rg_state* h_state);

/* ------------------------------------------------------------------------------------------------
 * Synthetic column generators on the device (stateless splitmix64 per index so that
 * host and device produce identical columns).  kind: 2 = decimal-heavy f64, 3 = high-precision f64 (ALP_RD),
 * 4 = mixed f32 (f32 entry point only).  first_index lets a shard generate its slice of a global column.
 * ---------------------------------------------------------------------------------------------- */
int alpb200_generate_f64(double* d_out, uint64_t n_values, uint64_t first_index, uint64_t seed, int kind, void* stream);
int alpb200_generate_f32(float* d_out, uint64_t n_values, uint64_t first_index, uint64_t seed, int kind, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALP_B200_H */
