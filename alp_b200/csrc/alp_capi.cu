// alp_capi.cu — the C ABI of include/alp_b200.h: argument checks, launches, the host-buffer codec context and the
// single-vector primitive wrappers.  No CPU fallback anywhere: without a CUDA device every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "alp_b200.h"
#include "alp_device.cuh"
#include "alp_host.h"

static_assert(sizeof(alpb200_rg_state) == 1196, "alpb200_rg_state layout");
static_assert(sizeof(alpb200_vec_meta) == 32, "alpb200_vec_meta layout");
static_assert(sizeof(alpb200_column) == 80, "alpb200_column layout");

using namespace alpb200;

namespace {
thread_local std::string g_last_error;
}

#ifndef ALPB200_ENCODE_STREAM_DEFAULT
#define ALPB200_ENCODE_STREAM_DEFAULT 0
#endif

namespace alpb200 {

int fail(int code, const char* fmt, const char* a, const char* b) {
	char buf[512];
	snprintf(buf, sizeof(buf), fmt, a, b);
	g_last_error = buf;
	return code;
}

namespace {
constexpr uint32_t N_COUNTERS = 4096;
DeviceInfo g_dev[64];
std::mutex g_dev_mutex;
}  // namespace

int device_info(DeviceInfo& out) {
	int dev = 0;
	CUDA_TRY(cudaGetDevice(&dev));
	std::lock_guard<std::mutex> lock(g_dev_mutex);
	if (dev < 0 || dev >= 64) { return fail(ALPB200_EINVAL, "device index out of range"); }
	if (g_dev[dev].sms == 0) {
		CUDA_TRY(cudaDeviceGetAttribute(&g_dev[dev].sms, cudaDevAttrMultiProcessorCount, dev));
		CUDA_TRY(cudaDeviceGetAttribute(&g_dev[dev].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
		CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&g_dev[dev].counters), N_COUNTERS * sizeof(unsigned long long)));
	}
	g_dev[dev].next_counter = (g_dev[dev].next_counter + 2) % N_COUNTERS;  // launches take PAIRS of slots
	out = g_dev[dev];
	return ALPB200_OK;
}

bool encode_uses_stream() {
	static const int choice = [] {
		const char* e = getenv("ALPB200_ENCODE_KERNEL");
		if (e && strcmp(e, "block") == 0) { return 0; }
		if (e && strcmp(e, "stream") == 0) { return 1; }
		return ALPB200_ENCODE_STREAM_DEFAULT;
	}();
	return choice != 0;
}

size_t encode_workspace_bytes(uint64_t n_vectors) {
	const uint64_t blocks = (n_vectors + ENC_MIN_WARPS - 1) / ENC_MIN_WARPS;
	return (size_t)((2 + 2 * blocks) * sizeof(uint64_t) + 255) & ~(size_t)255;
}

}  // namespace alpb200

// =====================================================================================================================
// Host-buffer codec context
// =====================================================================================================================
constexpr int MAX_CHUNKS = 16;  // pipeline depth of compress_host

struct alpb200_ctx {
	int          device      = 0;
	int          value_bytes = 8;
	uint64_t     max_vectors = 0;
	cudaStream_t streams[3]  = {nullptr, nullptr, nullptr};
	cudaEvent_t  events[3]   = {nullptr, nullptr, nullptr};
	// device-side column + value buffers, sized for max_vectors
	void*             d_values = nullptr;
	alpb200_vec_meta* d_meta   = nullptr;
	uint8_t*          d_packed = nullptr;
	void*             d_exc_val = nullptr;
	uint16_t*         d_exc_pos = nullptr;
	uint64_t*         d_totals  = nullptr;
	alpb200_rg_state* d_states  = nullptr;
	void*             d_ws_enc  = nullptr;
	void*             d_ws_init = nullptr;
	uint64_t          packed_capacity = 0, exc_capacity = 0;
	uint64_t*         h_totals = nullptr;  // pinned: one 4-word snapshot of the column totals per compress chunk
	cudaEvent_t       chunk_in[MAX_CHUNKS]  = {};
	cudaEvent_t       chunk_enc[MAX_CHUNKS] = {};
	bool              unordered = false;   // ALPB200_OPT_UNORDERED: compress_host encodes with the completion-order layout
	uint64_t          chunks    = MAX_CHUNKS;  // ALPB200_OPT_CHUNKS: pipeline depth of the host entry points (1..16)
};

namespace {

// keeps the caller's current device across a host entry point
struct DeviceGuard {
	int  prev = -1;
	bool ok   = false;
	explicit DeviceGuard(int dev) {
		ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess;
		if (!ok) { cudaGetLastError(); }
	}
	~DeviceGuard() {
		if (prev >= 0) { cudaSetDevice(prev); }
	}
};

void ctx_release(alpb200_ctx* c) {
	if (!c) { return; }
	DeviceGuard guard(c->device);
	for (int i = 0; i < 3; i++) {
		if (c->streams[i]) { cudaStreamDestroy(c->streams[i]); }
		if (c->events[i]) { cudaEventDestroy(c->events[i]); }
	}
	for (int i = 0; i < MAX_CHUNKS; i++) {
		if (c->chunk_in[i]) { cudaEventDestroy(c->chunk_in[i]); }
		if (c->chunk_enc[i]) { cudaEventDestroy(c->chunk_enc[i]); }
	}
	cudaFree(c->d_values);
	cudaFree(c->d_meta);
	cudaFree(c->d_packed);
	cudaFree(c->d_exc_val);
	cudaFree(c->d_exc_pos);
	cudaFree(c->d_totals);
	cudaFree(c->d_states);
	cudaFree(c->d_ws_enc);
	cudaFree(c->d_ws_init);
	if (c->h_totals) { cudaFreeHost(c->h_totals); }
	delete c;
}

// Chunked pipeline over three streams: while chunk i is analysed and encoded, chunk i+1's values are already on their
// way to the device and chunk i-1's compressed bytes on their way back (PCIe is full duplex).  Chunks are whole
// row-groups; each chunk's encode APPENDS to the column (the kernel continues at the running totals, which never leave
// the device), so the result is the same column a single launch would produce.  The host learns a chunk's byte range
// from a 32-byte snapshot of the totals and drains it one chunk behind the launches.
template <typename PT>
int compress_host(alpb200_ctx* c, const PT* h_in, uint64_t n_values_in, alpb200_column* h_col) {
	if (!c || !h_in || !h_col || !h_col->meta || !h_col->packed || !h_col->exc_val || !h_col->exc_pos || !h_col->totals) {
		return fail(ALPB200_EINVAL, "compress_host: null argument");
	}
	if (c->value_bytes != (int)sizeof(PT)) { return fail(ALPB200_EINVAL, "compress_host: context was created for another value width"); }
	const uint64_t n_vec    = (n_values_in + VEC - 1) / VEC;
	const uint64_t n_values = n_vec * VEC;  // padded length
	if (n_vec > c->max_vectors || n_vec > h_col->n_vectors) { return fail(ALPB200_EINVAL, "compress_host: column larger than the context / container"); }
	DeviceGuard guard(c->device);
	if (!guard.ok) { return fail(ALPB200_ECUDA, "compress_host: cannot select the context's device"); }
	if (n_vec == 0) {
		h_col->n_vectors = 0;
		h_col->n_values  = 0;
		std::memset(h_col->totals, 0, 4 * sizeof(uint64_t));
		return ALPB200_OK;
	}
	cudaStream_t   s_in = c->streams[0], s_enc = c->streams[1], s_out = c->streams[2];
	const uint64_t n_rg      = (n_vec + ALPB200_ROWGROUP_VECTORS - 1) / ALPB200_ROWGROUP_VECTORS;
	const uint64_t chunk_rg  = std::max<uint64_t>(10, (n_rg + c->chunks - 1) / c->chunks);  // at least ~8 MiB of f64 per chunk
	const uint64_t chunk_vec = chunk_rg * ALPB200_ROWGROUP_VECTORS;
	const int      n_chunks  = (int)((n_vec + chunk_vec - 1) / chunk_vec);
	PT*            d_values  = static_cast<PT*>(c->d_values);
	uint64_t       prev_p = 0, prev_e = 0;
	int            rc = ALPB200_OK;

	// copy chunk k's part of the column back (called one chunk behind the launches, and once more at the end)
	auto drain = [&](int k) -> int {
		CUDA_TRY(cudaEventSynchronize(c->chunk_enc[k]));
		const uint64_t* tot = c->h_totals + 4 * k;
		if (tot[2] != 0) { return fail(ALPB200_ECAPACITY, "compress_host: device staging capacity exceeded"); }
		const uint64_t p1 = tot[0], e1 = tot[1];
		if (p1 > h_col->packed_capacity || e1 > h_col->exc_capacity) { return fail(ALPB200_ECAPACITY, "compress_host: the host column container is too small"); }
		const uint64_t v0 = (uint64_t)k * chunk_vec, v1 = std::min(n_vec, v0 + chunk_vec);
		CUDA_TRY(cudaMemcpyAsync(h_col->meta + v0, c->d_meta + v0, (v1 - v0) * sizeof(alpb200_vec_meta), cudaMemcpyDeviceToHost, s_out));
		if (p1 > prev_p) { CUDA_TRY(cudaMemcpyAsync(h_col->packed + prev_p, c->d_packed + prev_p, p1 - prev_p, cudaMemcpyDeviceToHost, s_out)); }
		if (e1 > prev_e) {
			CUDA_TRY(cudaMemcpyAsync(static_cast<PT*>(h_col->exc_val) + prev_e, static_cast<PT*>(c->d_exc_val) + prev_e, (e1 - prev_e) * sizeof(PT),
			                         cudaMemcpyDeviceToHost, s_out));
			CUDA_TRY(cudaMemcpyAsync(h_col->exc_pos + prev_e, c->d_exc_pos + prev_e, (e1 - prev_e) * sizeof(uint16_t), cudaMemcpyDeviceToHost, s_out));
		}
		prev_p = p1;
		prev_e = e1;
		return ALPB200_OK;
	};
	auto launch_chunk = [&](int k) -> int {
		const uint64_t v0 = (uint64_t)k * chunk_vec, v1 = std::min(n_vec, v0 + chunk_vec);
		const uint64_t x0 = v0 * VEC, x1 = std::min(n_values_in, v1 * VEC);
		CUDA_TRY(cudaMemcpyAsync(d_values + x0, h_in + x0, (x1 - x0) * sizeof(PT), cudaMemcpyHostToDevice, s_in));
		const bool ragged = v1 == n_vec && n_values != n_values_in;  // the column's last vector is partial
		if (ragged) { TRY(launch_fill_invalid<PT>(d_values, n_values_in, nullptr, nullptr, s_in)); }  // real values everywhere for the sampling
		CUDA_TRY(cudaEventRecord(c->chunk_in[k], s_in));
		CUDA_TRY(cudaStreamWaitEvent(s_enc, c->chunk_in[k], 0));
		alpb200_rg_state* states = c->d_states + v0 / ALPB200_ROWGROUP_VECTORS;
		TRY(launch_init<PT>(d_values + x0, (v1 - v0) * VEC, states, c->d_ws_init, s_enc));
		// the tail's slots take the vector's first non-exception value: no exceptions, no wider block (PRIMITIVES.md:141-144)
		if (ragged) { TRY(launch_fill_invalid<PT>(d_values, n_values_in, nullptr, c->d_states, s_enc)); }
		alpb200_column d_col {};
		d_col.n_vectors       = v1 - v0;
		d_col.meta            = c->d_meta + v0;
		d_col.packed          = c->d_packed;
		d_col.packed_capacity = c->packed_capacity;
		d_col.exc_val         = c->d_exc_val;
		d_col.exc_pos         = c->d_exc_pos;
		d_col.exc_capacity    = c->exc_capacity;
		d_col.totals          = c->d_totals;
		TRY(launch_encode<PT>(d_values + x0, v1 - v0, states, &d_col, c->d_ws_enc, s_enc, !c->unordered, /*append=*/k > 0));
		CUDA_TRY(cudaMemcpyAsync(c->h_totals + 4 * k, c->d_totals, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s_enc));
		CUDA_TRY(cudaEventRecord(c->chunk_enc[k], s_enc));
		return ALPB200_OK;
	};
	for (int k = 0; k < n_chunks && rc == ALPB200_OK; k++) {
		rc = launch_chunk(k);
		if (rc == ALPB200_OK && k > 0) { rc = drain(k - 1); }
	}
	if (rc == ALPB200_OK) { rc = drain(n_chunks - 1); }
	for (int i = 0; i < 3; i++) {  // also on failure: nothing may still be running on the caller's buffers
		cudaError_t err = cudaStreamSynchronize(c->streams[i]);
		if (err != cudaSuccess && rc == ALPB200_OK) { rc = fail(ALPB200_ECUDA, "compress_host: %s", cudaGetErrorString(err)); }
	}
	if (rc != ALPB200_OK) { return rc; }
	const uint64_t* tot    = c->h_totals + 4 * (n_chunks - 1);
	h_col->n_vectors       = n_vec;
	h_col->n_values        = n_values_in;
	h_col->max_block_bytes = tot[3];
	std::memcpy(h_col->totals, tot, 4 * sizeof(uint64_t));
	return ALPB200_OK;
}

// Byte / slot ranges of the compressed arrays that the vectors [v0, v1) touch.  For a vector-order column that is
// [offset of v0, offset of v1); a completion-order column (alpb200_encode_unordered_*) is only roughly sorted, so the
// range is the min / max over the records (neighbouring chunks then overlap by a few blocks, which are simply copied
// twice).
struct ChunkRange {
	uint64_t p0, p1, e0, e1;
	uint32_t widest;  // largest packed block of the range, bytes
	bool     valid;
};
inline uint64_t block_units(const alpb200_vec_meta& m) { return m.scheme == ALPB200_SCHEME_ALP_RD ? (uint64_t)m.bw + m.e : m.bw; }
// what a well-formed record can say (the encoders never write anything else)
inline bool record_ok(const alpb200_vec_meta& m, int value_bytes) {
	const uint32_t T = 8u * (uint32_t)value_bytes, max_exp = value_bytes == 8 ? 18u : 10u;
	if (m.exc_cnt > VEC) { return false; }
	if (m.scheme == ALPB200_SCHEME_ALP) { return m.bw <= T && m.e <= max_exp && m.f <= m.e; }
	if (m.scheme == ALPB200_SCHEME_ALP_RD) { return m.bw < T && m.bw + 16u >= T && m.e >= 1 && m.e <= 3 && m.f >= 1 && m.f <= ALPB200_RD_DICT_SIZE; }
	return false;
}
// Walks the records of [v0, v1): byte / slot ranges, the widest block, and whether every record is well formed.  With
// host_col the exception positions are checked too (each must lie inside the vector).
ChunkRange chunk_range(const alpb200_vec_meta* hm, uint64_t v0, uint64_t v1, int value_bytes, const alpb200_column* host_col = nullptr) {
	ChunkRange r {UINT64_MAX, 0, UINT64_MAX, 0, 0, true};
	for (uint64_t v = v0; v < v1; v++) {
		const alpb200_vec_meta& m = hm[v];
		if (!record_ok(m, value_bytes)) {
			r.valid = false;
			continue;
		}
		const uint64_t p = (uint64_t)m.packed_off * 128ull, e = m.exc_off;
		r.p0 = std::min(r.p0, p);
		r.p1 = std::max<uint64_t>(r.p1, p + block_units(m) * 128ull);
		r.e0 = std::min(r.e0, e);
		r.e1 = std::max<uint64_t>(r.e1, e + m.exc_cnt);
		r.widest = std::max<uint32_t>(r.widest, (uint32_t)(block_units(m) * 128ull));
		if (host_col && host_col->exc_pos && e + m.exc_cnt <= host_col->exc_capacity) {
			const uint16_t* ep = host_col->exc_pos + e;
			uint32_t        bad = 0;
			for (uint32_t i = 0; i < m.exc_cnt; i++) {
				bad |= ep[i] >= VEC;
			}
			if (bad) { r.valid = false; }
		}
	}
	if (v0 >= v1 || r.p0 == UINT64_MAX) { r.p0 = r.p1 = r.e0 = r.e1 = 0; }
	return r;
}
// the host column's own arrays must cover what its records point at
int check_host_column(const alpb200_column* h_col, const ChunkRange& cr, const char* who) {
	if (!cr.valid) { return fail(ALPB200_EINVAL, "%s: malformed vector record (scheme / bit width / exponent / exception count or position out of range)", who); }
	if (cr.p1 > h_col->packed_capacity || cr.e1 > h_col->exc_capacity) {
		return fail(ALPB200_EINVAL, "%s: a vector record points outside the column's packed / exception arrays", who);
	}
	return ALPB200_OK;
}
// Chunked pipeline: while chunk i decodes and drains to the host on one stream, chunk i+1's compressed bytes are
// already travelling on another (PCIe is full duplex).  Chunks are whole ranges of vectors; thanks to the vector-order
// layout each chunk's packed bytes and exceptions are one contiguous range.
template <typename PT>
int decompress_host(alpb200_ctx* c, const alpb200_column* h_col, PT* h_out) {
	if (!c || !h_col || !h_out || !h_col->meta) { return fail(ALPB200_EINVAL, "decompress_host: null argument"); }
	if (c->value_bytes != (int)sizeof(PT)) { return fail(ALPB200_EINVAL, "decompress_host: context was created for another value width"); }
	const uint64_t n_vec = h_col->n_vectors;
	if (n_vec > c->max_vectors) { return fail(ALPB200_EINVAL, "decompress_host: column larger than the context"); }
	if (n_vec == 0) { return ALPB200_OK; }
	if (!h_col->packed || !h_col->exc_val || !h_col->exc_pos) { return fail(ALPB200_EINVAL, "decompress_host: null argument"); }
	const uint64_t n_out = h_col->n_values ? h_col->n_values : n_vec * VEC;  // a padded tail is not returned
	if (n_out > n_vec * VEC || n_out + VEC <= n_vec * VEC) { return fail(ALPB200_EINVAL, "decompress_host: n_values does not match n_vectors"); }
	DeviceGuard guard(c->device);
	if (!guard.ok) { return fail(ALPB200_ECUDA, "decompress_host: cannot select the context's device"); }
	const alpb200_vec_meta* hm = h_col->meta;
	alpb200_column d_col {};
	d_col.n_vectors       = n_vec;
	d_col.meta            = c->d_meta;
	d_col.packed          = c->d_packed;
	d_col.packed_capacity = c->packed_capacity;
	d_col.exc_val         = c->d_exc_val;
	d_col.exc_pos         = c->d_exc_pos;
	d_col.exc_capacity    = c->exc_capacity;

	const uint64_t chunk = std::max<uint64_t>(1024, (n_vec + c->chunks - 1) / c->chunks);  // ~16 chunks, at least 8 MiB of f64 output each
	int            si = 0, rc = ALPB200_OK;
	auto           body = [&](uint64_t v0, uint64_t v1, cudaStream_t s) -> int {
        const ChunkRange cr = chunk_range(hm, v0, v1, sizeof(PT), h_col);  // scanned chunk by chunk: the first copy starts at once
        TRY(check_host_column(h_col, cr, "decompress_host"));
        if (cr.p1 > c->packed_capacity || cr.e1 > c->exc_capacity) { return fail(ALPB200_ECAPACITY, "decompress_host: column exceeds the context's staging capacity"); }
        const uint64_t p0 = cr.p0, p1 = cr.p1, e0 = cr.e0, e1 = cr.e1;
        CUDA_TRY(cudaMemcpyAsync(c->d_meta + v0, hm + v0, (v1 - v0) * sizeof(alpb200_vec_meta), cudaMemcpyHostToDevice, s));
        if (p1 > p0) { CUDA_TRY(cudaMemcpyAsync(c->d_packed + p0, h_col->packed + p0, p1 - p0, cudaMemcpyHostToDevice, s)); }
        if (e1 > e0) {
            CUDA_TRY(cudaMemcpyAsync(static_cast<PT*>(c->d_exc_val) + e0, static_cast<const PT*>(h_col->exc_val) + e0, (e1 - e0) * sizeof(PT),
                                     cudaMemcpyHostToDevice, s));
            CUDA_TRY(cudaMemcpyAsync(c->d_exc_pos + e0, h_col->exc_pos + e0, (e1 - e0) * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
        }
        PT* d_out            = static_cast<PT*>(c->d_values) + v0 * VEC;
        d_col.max_block_bytes = cr.widest;  // the chunk's true widest block, never the caller's hint
        TRY(launch_decode<PT>(&d_col, v0, v1 - v0, d_out, s));
        const uint64_t n_copy = std::min<uint64_t>((v1 - v0) * VEC, n_out - v0 * VEC);
        CUDA_TRY(cudaMemcpyAsync(h_out + v0 * VEC, d_out, n_copy * sizeof(PT), cudaMemcpyDeviceToHost, s));
        return ALPB200_OK;
	};
	for (uint64_t v0 = 0; v0 < n_vec && rc == ALPB200_OK; v0 += chunk, si = (si + 1) % 3) {
		rc = body(v0, std::min(n_vec, v0 + chunk), c->streams[si]);
	}
	for (int i = 0; i < 3; i++) {  // also on failure: nothing may still be running on the caller's buffers
		cudaError_t err = cudaStreamSynchronize(c->streams[i]);
		if (err != cudaSuccess && rc == ALPB200_OK) { rc = fail(ALPB200_ECUDA, "decompress_host: %s", cudaGetErrorString(err)); }
	}
	return rc;
}

// SUM of a host column: the chunk pipeline of decompress_host with the fused decode+SUM kernel and no values coming
// back.  Every chunk adds into the same device double (atomics), read back once at the end.
template <typename PT>
int sum_host(alpb200_ctx* c, const alpb200_column* h_col, double* h_sum) {
	if (!c || !h_col || !h_sum || !h_col->meta) { return fail(ALPB200_EINVAL, "sum_host: null argument"); }
	if (c->value_bytes != (int)sizeof(PT)) { return fail(ALPB200_EINVAL, "sum_host: context was created for another value width"); }
	const uint64_t n_vec = h_col->n_vectors;
	if (n_vec > c->max_vectors) { return fail(ALPB200_EINVAL, "sum_host: column larger than the context"); }
	*h_sum = 0.0;
	if (n_vec == 0) { return ALPB200_OK; }
	if (!h_col->packed || !h_col->exc_val || !h_col->exc_pos) { return fail(ALPB200_EINVAL, "sum_host: null argument"); }
	const uint64_t n_out = h_col->n_values ? h_col->n_values : n_vec * VEC;
	if (n_out > n_vec * VEC || n_out + VEC <= n_vec * VEC) { return fail(ALPB200_EINVAL, "sum_host: n_values does not match n_vectors"); }
	DeviceGuard guard(c->device);
	if (!guard.ok) { return fail(ALPB200_ECUDA, "sum_host: cannot select the context's device"); }
	const alpb200_vec_meta* hm = h_col->meta;
	alpb200_column d_col {};
	d_col.n_vectors       = n_vec;
	d_col.meta            = c->d_meta;
	d_col.packed          = c->d_packed;
	d_col.packed_capacity = c->packed_capacity;
	d_col.exc_val         = c->d_exc_val;
	d_col.exc_pos         = c->d_exc_pos;
	d_col.exc_capacity    = c->exc_capacity;
	double* d_sum         = reinterpret_cast<double*>(c->d_totals);  // 32 bytes of device scratch owned by the context
	// a padded tail vector is summed on its own so that the padding can be taken out again (it repeats the last value)
	const bool     ragged = n_out != n_vec * VEC;
	const uint64_t n_full = ragged ? n_vec - 1 : n_vec;
	const uint64_t chunk  = std::max<uint64_t>(1024, (n_vec + c->chunks - 1) / c->chunks);
	int            si = 0, rc = ALPB200_OK;
	uint32_t       tail_widest = 0;
	auto           prologue = [&]() -> int {
        CUDA_TRY(cudaMemsetAsync(d_sum, 0, sizeof(double), c->streams[0]));
        CUDA_TRY(cudaEventRecord(c->events[0], c->streams[0]));
        for (int i = 1; i < 3; i++) {
            CUDA_TRY(cudaStreamWaitEvent(c->streams[i], c->events[0], 0));
        }
        return ALPB200_OK;
	};
	auto body = [&](uint64_t v0, uint64_t v1, cudaStream_t s) -> int {
		const ChunkRange cr = chunk_range(hm, v0, v1, sizeof(PT), h_col);
		TRY(check_host_column(h_col, cr, "sum_host"));
		if (cr.p1 > c->packed_capacity || cr.e1 > c->exc_capacity) { return fail(ALPB200_ECAPACITY, "sum_host: column exceeds the context's staging capacity"); }
		const uint64_t p0 = cr.p0, p1 = cr.p1, e0 = cr.e0, e1 = cr.e1;
		CUDA_TRY(cudaMemcpyAsync(c->d_meta + v0, hm + v0, (v1 - v0) * sizeof(alpb200_vec_meta), cudaMemcpyHostToDevice, s));
		if (p1 > p0) { CUDA_TRY(cudaMemcpyAsync(c->d_packed + p0, h_col->packed + p0, p1 - p0, cudaMemcpyHostToDevice, s)); }
		if (e1 > e0) {
			CUDA_TRY(cudaMemcpyAsync(static_cast<PT*>(c->d_exc_val) + e0, static_cast<const PT*>(h_col->exc_val) + e0, (e1 - e0) * sizeof(PT),
			                         cudaMemcpyHostToDevice, s));
			CUDA_TRY(cudaMemcpyAsync(c->d_exc_pos + e0, h_col->exc_pos + e0, (e1 - e0) * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
		}
		const uint64_t stop   = std::min(v1, n_full);
		d_col.max_block_bytes = cr.widest;
		tail_widest           = cr.widest;
		if (stop > v0) { TRY(launch_decode_sum<PT>(&d_col, v0, stop - v0, d_sum, s)); }
		return ALPB200_OK;
	};
	rc = prologue();
	for (uint64_t v0 = 0; v0 < n_vec && rc == ALPB200_OK; v0 += chunk, si = (si + 1) % 3) {
		rc = body(v0, std::min(n_vec, v0 + chunk), c->streams[si]);
	}
	for (int i = 0; i < 3; i++) {  // also on failure: nothing may still be reading the caller's column
		cudaError_t err = cudaStreamSynchronize(c->streams[i]);
		if (err != cudaSuccess && rc == ALPB200_OK) { rc = fail(ALPB200_ECUDA, "sum_host: %s", cudaGetErrorString(err)); }
	}
	if (rc != ALPB200_OK) { return rc; }
	double total = 0.0;
	if (ragged) {  // decode the last vector (4-8 KiB) and add its real values on the host
		cudaStream_t s = c->streams[0];
		PT*          d_tail = static_cast<PT*>(c->d_values);
		d_col.max_block_bytes = tail_widest;
		TRY(launch_decode<PT>(&d_col, n_vec - 1, 1, d_tail, s));
		std::vector<PT> tail(VEC);
		CUDA_TRY(cudaMemcpyAsync(tail.data(), d_tail, VEC * sizeof(PT), cudaMemcpyDeviceToHost, s));
		CUDA_TRY(cudaStreamSynchronize(s));
		for (uint64_t i = 0; i < n_out - n_full * VEC; i++) {
			total += (double)tail[i];
		}
	}
	double dev_sum = 0.0;
	CUDA_TRY(cudaMemcpy(&dev_sum, d_sum, sizeof(double), cudaMemcpyDeviceToHost));
	*h_sum = dev_sum + total;
	return ALPB200_OK;
}

}  // namespace

namespace {
// decode exactly n_values values: whole vectors straight into d_out, a partial last vector through d_scratch
template <typename PT>
int decode_values(const alpb200_column* col, uint64_t first, uint64_t n_values, PT* d_out, PT* d_scratch, void* stream) {
	const uint64_t n_full = n_values / VEC, n_tail = n_values % VEC;
	if (!col) { return fail(ALPB200_EINVAL, "decode_values: null argument"); }
	if (first + n_full + (n_tail ? 1 : 0) > col->n_vectors) { return fail(ALPB200_EINVAL, "decode_values: value range outside the column"); }
	if (n_tail && !d_scratch) { return fail(ALPB200_EINVAL, "decode_values: a partial last vector needs d_scratch (1024 values)"); }
	TRY(launch_decode<PT>(col, first, n_full, d_out, stream));
	if (n_tail) {
		TRY(launch_decode<PT>(col, first + n_full, 1, d_scratch, stream));
		CUDA_TRY(cudaMemcpyAsync(d_out + n_full * VEC, d_scratch, n_tail * sizeof(PT), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
	}
	return ALPB200_OK;
}
}  // namespace

// =====================================================================================================================
// extern "C"
// =====================================================================================================================
extern "C" {

int alpb200_version(void) { return 100; }

void alpb200_abi_sizes(uint32_t out[3]) {
	out[0] = sizeof(alpb200_rg_state);
	out[1] = sizeof(alpb200_vec_meta);
	out[2] = sizeof(alpb200_column);
}

const char* alpb200_last_error(void) { return g_last_error.c_str(); }

int alpb200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return fail(ALPB200_ENODEVICE, "no CUDA device is visible; alp_b200 has no CPU fallback");
	}
	return n;
}

size_t alpb200_init_workspace_bytes(uint64_t n_values) { return init_workspace_bytes(n_values); }
int alpb200_rowgroup_init_f64(const double* d_in, uint64_t n_values, alpb200_rg_state* d_states, void* ws, void* stream) {
	return launch_init<double>(d_in, n_values, d_states, ws, stream);
}
int alpb200_rowgroup_init_f32(const float* d_in, uint64_t n_values, alpb200_rg_state* d_states, void* ws, void* stream) {
	return launch_init<float>(d_in, n_values, d_states, ws, stream);
}

size_t alpb200_encode_workspace_bytes(uint64_t n_vectors) { return encode_workspace_bytes(n_vectors); }
int alpb200_encode_f64(const double* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws,
                       void* stream) {
	return launch_encode<double>(d_in, n_vectors, d_states, col, ws, stream);
}
int alpb200_encode_f32(const float* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws,
                       void* stream) {
	return launch_encode<float>(d_in, n_vectors, d_states, col, ws, stream);
}

int alpb200_encode_unordered_f64(const double* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws,
                                 void* stream) {
	return launch_encode<double>(d_in, n_vectors, d_states, col, ws, stream, false);
}
int alpb200_encode_unordered_f32(const float* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws,
                                 void* stream) {
	return launch_encode<float>(d_in, n_vectors, d_states, col, ws, stream, false);
}

int alpb200_encode_ex_f64(const double* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws,
                          void* stream, uint32_t flags) {
	if (flags & ~(ALPB200_ENCODE_UNORDERED | ALPB200_ENCODE_APPEND)) { return fail(ALPB200_EINVAL, "encode_ex: unknown flag"); }
	return launch_encode<double>(d_in, n_vectors, d_states, col, ws, stream, !(flags & ALPB200_ENCODE_UNORDERED), (flags & ALPB200_ENCODE_APPEND) != 0);
}
int alpb200_encode_ex_f32(const float* d_in, uint64_t n_vectors, const alpb200_rg_state* d_states, const alpb200_column* col, void* ws,
                          void* stream, uint32_t flags) {
	if (flags & ~(ALPB200_ENCODE_UNORDERED | ALPB200_ENCODE_APPEND)) { return fail(ALPB200_EINVAL, "encode_ex: unknown flag"); }
	return launch_encode<float>(d_in, n_vectors, d_states, col, ws, stream, !(flags & ALPB200_ENCODE_UNORDERED), (flags & ALPB200_ENCODE_APPEND) != 0);
}

int alpb200_decode_f64(const alpb200_column* col, uint64_t first, uint64_t n, double* d_out, void* stream) {
	// d_out receives vector `first + i` at d_out[i * 1024]
	return launch_decode<double>(col, first, n, d_out, stream);
}
int alpb200_decode_f32(const alpb200_column* col, uint64_t first, uint64_t n, float* d_out, void* stream) {
	return launch_decode<float>(col, first, n, d_out, stream);
}

int alpb200_decode_sum_f64(const alpb200_column* col, uint64_t first, uint64_t n, double* d_sum, void* stream) {
	return launch_decode_sum<double>(col, first, n, d_sum, stream);
}
int alpb200_decode_sum_f32(const alpb200_column* col, uint64_t first, uint64_t n, double* d_sum, void* stream) {
	return launch_decode_sum<float>(col, first, n, d_sum, stream);
}
int alpb200_decode_minmax_f64(const alpb200_column* col, uint64_t first, uint64_t n, alpb200_minmax* d_out, void* stream) {
	return launch_decode_minmax<double>(col, first, n, d_out, stream);
}
int alpb200_decode_minmax_f32(const alpb200_column* col, uint64_t first, uint64_t n, alpb200_minmax* d_out, void* stream) {
	return launch_decode_minmax<float>(col, first, n, d_out, stream);
}
int alpb200_decode_filter_f64(const alpb200_column* col, uint64_t first, uint64_t n, uint32_t op, double constant, uint32_t* d_bitmap,
                              uint64_t* d_selected, void* stream) {
	return launch_decode_filter<double>(col, first, n, op, constant, d_bitmap, d_selected, stream);
}
int alpb200_decode_filter_f32(const alpb200_column* col, uint64_t first, uint64_t n, uint32_t op, double constant, uint32_t* d_bitmap,
                              uint64_t* d_selected, void* stream) {
	return launch_decode_filter<float>(col, first, n, op, constant, d_bitmap, d_selected, stream);
}
int alpb200_decode_sum_ex_f64(const alpb200_column* col, uint64_t first, uint64_t n, double* d_sum, uint32_t flags, void* stream) {
	if (flags & ~(uint32_t)ALPB200_SUM_DECIMAL) { return fail(ALPB200_EINVAL, "decode_sum_ex: unknown flag"); }
	return launch_decode_sum<double>(col, first, n, d_sum, stream, flags);
}
int alpb200_decode_sum_ex_f32(const alpb200_column* col, uint64_t first, uint64_t n, double* d_sum, uint32_t flags, void* stream) {
	if (flags & ~(uint32_t)ALPB200_SUM_DECIMAL) { return fail(ALPB200_EINVAL, "decode_sum_ex: unknown flag"); }
	return launch_decode_sum<float>(col, first, n, d_sum, stream, flags);
}

int alpb200_ctx_create_ex(alpb200_ctx** out, int device, uint64_t max_vectors, int value_bytes, uint64_t packed_capacity, uint64_t exc_capacity) {
	if (!out || max_vectors == 0 || max_vectors > (1ull << 22) || (value_bytes != 8 && value_bytes != 4)) {
		return fail(ALPB200_EINVAL, "ctx_create: bad argument (max_vectors in 1..2^22, value_bytes 8 or 4)");
	}
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
		cudaGetLastError();
		return fail(ALPB200_ENODEVICE, "no CUDA device is visible; alp_b200 has no CPU fallback");
	}
	DeviceGuard guard(device);
	if (!guard.ok) { return fail(ALPB200_ECUDA, "ctx_create: cannot select the device"); }
	alpb200_ctx* c = new (std::nothrow) alpb200_ctx();
	if (!c) { return fail(ALPB200_EINVAL, "ctx_create: out of host memory"); }
	c->device      = device;
	c->value_bytes = value_bytes;
	c->max_vectors = max_vectors;
	const uint64_t n_rg = (max_vectors + ALPB200_ROWGROUP_VECTORS - 1) / ALPB200_ROWGROUP_VECTORS;
	// staging for the compressed column: what the caller says its columns need (an overflow is reported as
	// ALPB200_ECAPACITY, nothing is written out of bounds), else the worst case — every vector ALP_RD at full width with
	// 1024 exceptions, ~27 bytes per f64 value, which only a caller that cannot bound its data should pay for
	const uint64_t worst_packed = max_vectors * ((value_bytes == 8 ? 66ull : 35ull) * 128ull), worst_exc = max_vectors * VEC;
	c->packed_capacity  = packed_capacity ? std::min<uint64_t>(worst_packed, (packed_capacity + 127ull) & ~127ull) : worst_packed;
	c->exc_capacity     = exc_capacity ? std::min<uint64_t>(worst_exc, exc_capacity) : worst_exc;
	cudaError_t err     = cudaSuccess;
	auto        alloc   = [&](void** p, size_t bytes) {
        if (err == cudaSuccess) { err = cudaMalloc(p, bytes ? bytes : 16); }
	};
	alloc(&c->d_values, max_vectors * VEC * value_bytes);
	alloc(reinterpret_cast<void**>(&c->d_meta), max_vectors * sizeof(alpb200_vec_meta));
	alloc(reinterpret_cast<void**>(&c->d_packed), c->packed_capacity);
	alloc(&c->d_exc_val, c->exc_capacity * value_bytes);
	alloc(reinterpret_cast<void**>(&c->d_exc_pos), c->exc_capacity * sizeof(uint16_t));
	alloc(reinterpret_cast<void**>(&c->d_totals), 4 * sizeof(uint64_t));
	alloc(reinterpret_cast<void**>(&c->d_states), n_rg * sizeof(alpb200_rg_state));
	alloc(&c->d_ws_enc, encode_workspace_bytes(max_vectors));
	alloc(&c->d_ws_init, init_workspace_bytes(max_vectors * VEC));
	if (err == cudaSuccess) { err = cudaHostAlloc(reinterpret_cast<void**>(&c->h_totals), MAX_CHUNKS * 4 * sizeof(uint64_t), cudaHostAllocDefault); }
	for (int i = 0; i < MAX_CHUNKS && err == cudaSuccess; i++) {
		err = cudaEventCreateWithFlags(&c->chunk_in[i], cudaEventDisableTiming);
		if (err == cudaSuccess) { err = cudaEventCreateWithFlags(&c->chunk_enc[i], cudaEventDisableTiming); }
	}
	for (int i = 0; i < 3 && err == cudaSuccess; i++) {
		err = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
		if (err == cudaSuccess) { err = cudaEventCreateWithFlags(&c->events[i], cudaEventDisableTiming); }
	}
	if (err != cudaSuccess) {
		ctx_release(c);
		cudaGetLastError();
		return fail(ALPB200_ECUDA, "ctx_create: %s", cudaGetErrorString(err));
	}
	*out = c;
	return ALPB200_OK;
}
int alpb200_ctx_create(alpb200_ctx** out, int device, uint64_t max_vectors, int value_bytes) {
	return alpb200_ctx_create_ex(out, device, max_vectors, value_bytes, 0, 0);
}
void alpb200_ctx_destroy(alpb200_ctx* ctx) { ctx_release(ctx); }
int  alpb200_ctx_set_option(alpb200_ctx* ctx, int option, int value) {
	if (!ctx) { return fail(ALPB200_EINVAL, "ctx_set_option: null context"); }
	switch (option) {
	case ALPB200_OPT_UNORDERED: ctx->unordered = value != 0; return ALPB200_OK;
	case ALPB200_OPT_CHUNKS:
		if (value < 1 || value > MAX_CHUNKS) { return fail(ALPB200_EINVAL, "ctx_set_option: ALPB200_OPT_CHUNKS must be in 1..16"); }
		ctx->chunks = (uint64_t)value;
		return ALPB200_OK;
	default: return fail(ALPB200_EINVAL, "ctx_set_option: unknown option");
	}
}

int alpb200_compress_host_f64(alpb200_ctx* ctx, const double* h_in, uint64_t n_values, alpb200_column* h_col) {
	return compress_host<double>(ctx, h_in, n_values, h_col);
}
int alpb200_compress_host_f32(alpb200_ctx* ctx, const float* h_in, uint64_t n_values, alpb200_column* h_col) {
	return compress_host<float>(ctx, h_in, n_values, h_col);
}
int alpb200_decompress_host_f64(alpb200_ctx* ctx, const alpb200_column* h_col, double* h_out) { return decompress_host<double>(ctx, h_col, h_out); }
int alpb200_decompress_host_f32(alpb200_ctx* ctx, const alpb200_column* h_col, float* h_out) { return decompress_host<float>(ctx, h_col, h_out); }
int alpb200_sum_host_f64(alpb200_ctx* ctx, const alpb200_column* h_col, double* h_sum) { return sum_host<double>(ctx, h_col, h_sum); }
int alpb200_sum_host_f32(alpb200_ctx* ctx, const alpb200_column* h_col, double* h_sum) { return sum_host<float>(ctx, h_col, h_sum); }

int alpb200_column_validate_host(const alpb200_column* h_col, int value_bytes) {
	if (!h_col || (value_bytes != 8 && value_bytes != 4)) { return fail(ALPB200_EINVAL, "column_validate_host: bad argument"); }
	if (h_col->n_vectors == 0) { return ALPB200_OK; }
	if (!h_col->meta || !h_col->packed || !h_col->exc_val || !h_col->exc_pos) { return fail(ALPB200_EINVAL, "column_validate_host: null array"); }
	const uint64_t n_out = h_col->n_values ? h_col->n_values : h_col->n_vectors * VEC;
	if (n_out > h_col->n_vectors * VEC || n_out + VEC <= h_col->n_vectors * VEC) { return fail(ALPB200_EINVAL, "column_validate_host: n_values does not match n_vectors"); }
	const ChunkRange cr = chunk_range(h_col->meta, 0, h_col->n_vectors, value_bytes, h_col);
	return check_host_column(h_col, cr, "column_validate_host");
}

int alpb200_fill_invalid_f64(double* d_values, uint64_t n_values, const uint8_t* d_validity, const alpb200_rg_state* d_states, void* stream) {
	return launch_fill_invalid<double>(d_values, n_values, d_validity, d_states, stream);
}
int alpb200_fill_invalid_f32(float* d_values, uint64_t n_values, const uint8_t* d_validity, const alpb200_rg_state* d_states, void* stream) {
	return launch_fill_invalid<float>(d_values, n_values, d_validity, d_states, stream);
}

int alpb200_decode_values_f64(const alpb200_column* col, uint64_t first, uint64_t n_values, double* d_out, double* d_scratch, void* stream) {
	return decode_values<double>(col, first, n_values, d_out, d_scratch, stream);
}
int alpb200_decode_values_f32(const alpb200_column* col, uint64_t first, uint64_t n_values, float* d_out, float* d_scratch, void* stream) {
	return decode_values<float>(col, first, n_values, d_out, d_scratch, stream);
}

int alpb200_column_validate_device(const alpb200_column* col, int value_bytes, uint64_t* h_max_block_bytes, void* stream) {
	return validate_device(col, value_bytes, h_max_block_bytes, stream);
}

void* alpb200_host_alloc(size_t bytes) {
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
		cudaGetLastError();
		fail(ALPB200_ECUDA, "host_alloc: cudaHostAlloc failed");
		return nullptr;
	}
	return p;
}
void alpb200_host_free(void* p) {
	if (p) { cudaFreeHost(p); }
}

}  // extern "C"
