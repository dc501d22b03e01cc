/*
 * oracle/ref_bench_main.cpp — TEST / BENCH INFRASTRUCTURE, not product code.
 *
 * The CPU baseline as an EXECUTABLE: the reference's header-only templates (include/alp/*.hpp) and its generated
 * FFOR / UNFFOR / FALP kernels (src/*.cpp), compiled where they lie under /root/reference by oracle/Makefile into
 * oracle/_ref/ref_bench_v{3,4} (v4 = AVX-512 kernels).  It exists because alp::encoder<PT>::encode_simdized keeps its
 * scratch arrays in `static thread_local` storage (encoder.hpp:314-319): inside a dlopen'ed shared library g++ reaches them
 * through a __tls_get_addr call on every access in the 1024-value loops (11-17 ns per value), while an executable — which
 * is how the reference's users build these headers — uses plain %fs-relative addressing (4.8 ns per value).  bench.py's CPU
 * legs run this binary so that the reference is not handicapped by our packaging; the drivers are the very templates of
 * oracle/ref_shim.cpp (included below), which call nothing but the reference's own primitives.
 *
 *   ref_bench --kind 2|3|4 --values N [--first I] [--threads T] [--steps K --warmup W] [--seconds S]
 *
 * prints ONE JSON object: decode (K timed steps after W warm-up steps when --steps is given, else as many as fit S/4
 * seconds), encode with given states, row-group init, scan (alp_func + aggr_plus), each in seconds per pass over the column,
 * and whether decode(encode(x)) == x bit for bit.
 */
#include "ref_shim.cpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

namespace {

/* SURVEY.md §8d generators — twins of alpo_generate_* (oracle/alp_oracle.c) and alpb200_generate_* (the device) */
inline uint64_t splitmix64(uint64_t seed, uint64_t i) {
	uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
	z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z          = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
void generate(double* out, uint64_t n, uint64_t first, uint64_t seed, int kind) {
	static const double DIV[4] = {1.0, 10.0, 100.0, 1000.0};
	for (uint64_t j = 0; j < n; j++) {
		const uint64_t i = first + j, r = splitmix64(seed, i);
		if (kind == 3) {
			const double u = static_cast<double>(r >> 11) * 0x1.0p-53;
			const double t = u * 180.0;
			out[j]         = t - 90.0;
		} else {
			out[j] = static_cast<double>(r % 1000000ULL) / DIV[(i / ALPB200_ROWGROUP_SIZE) % 4];
		}
	}
}
void generate(float* out, uint64_t n, uint64_t first, uint64_t seed, int) {
	for (uint64_t j = 0; j < n; j++) {
		const uint64_t r = splitmix64(seed, first + j);
		if (r % 100 >= 5) {
			out[j] = static_cast<float>((r >> 8) % 10000ULL) / 100.0f;
		} else {
			uint32_t b = static_cast<uint32_t>(r >> 32);
			b          = (b & 0x807FFFFFu) | ((20u + ((b >> 23) % 200u)) << 23);
			std::memcpy(&out[j], &b, sizeof(b));
		}
	}
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* seconds per call: one warm-up call, then calls until `budget` seconds have passed (at least one) */
template <typename F>
double timed(F&& fn, double budget, int* reps_out = nullptr) {
	fn();
	int          reps = 0;
	const double t0   = now();
	double       dt   = 0;
	do {
		fn();
		reps++;
		dt = now() - t0;
	} while (dt < budget && reps < 200);
	if (reps_out) { *reps_out = reps; }
	return dt / reps;
}

template <typename PT>
int run(int kind, uint64_t n_values, uint64_t first, int threads, int steps, int warmup, double seconds) {
	n_values = n_values / 1024 * 1024;
	const uint64_t  n_vec = n_values / 1024, n_rg = (n_vec + 99) / 100;
	const uint64_t  seed  = kind == 2 ? 42 : (kind == 3 ? 43 : 44);
	// big buffers are left uninitialised (no page is touched before it is used)
	std::unique_ptr<PT[]> x_buf(new PT[n_values + 1]), out_buf(new PT[n_values + 1]);
	struct View {
		PT* p;
		PT* data() const { return p; }
	} x {x_buf.get()}, out {out_buf.get()};
	{
		std::vector<std::thread> pool;
		for (int t = 0; t < threads; t++) {
			const uint64_t lo = n_values * t / threads, hi = n_values * (t + 1) / threads;
			pool.emplace_back([&, lo, hi]() { generate(x.data() + lo, hi - lo, first + lo, seed, kind); });
		}
		for (auto& t : pool) {
			t.join();
		}
	}
	const uint64_t                units = sizeof(PT) == 8 ? 67 : 35;
	std::vector<alpb200_vec_meta> meta(n_vec + 1);
	const uint64_t                packed_cap = (n_vec + 100ull * threads) * units * 128, exc_cap = n_values / 2 + 102400ull * threads;
	std::unique_ptr<uint8_t[]>    packed_raw(new uint8_t[packed_cap + 256]);
	uint8_t*                      packed = packed_raw.get() + ((128 - reinterpret_cast<uintptr_t>(packed_raw.get()) % 128) % 128);
	std::unique_ptr<PT[]>         exc_val(new PT[exc_cap]);
	std::unique_ptr<uint16_t[]>   exc_pos(new uint16_t[exc_cap]);
	std::vector<alpb200_rg_state> states(n_rg + 1);
	uint64_t                      totals[4] = {0, 0, 0, 0};
	alpb200_column                col {};
	col.n_vectors       = n_vec;
	col.meta            = meta.data();
	col.packed          = packed;
	col.packed_capacity = packed_cap;
	col.exc_val         = exc_val.get();
	col.exc_pos         = exc_pos.get();
	col.exc_capacity    = exc_cap;
	col.totals          = totals;

	const double leg   = seconds / 4;
	const double t_init = timed([&]() { ref_bench_encode<PT>(x.data(), n_values, threads, nullptr, states.data(), 1); }, leg);
	int          rc     = 0;
	const double t_enc  = timed([&]() { rc |= ref_bench_encode<PT>(x.data(), n_values, threads, &col, states.data(), 2); }, leg);
	if (rc != 0) {
		std::fprintf(stderr, "ref_bench: encode failed (%d)\n", rc);
		return 2;
	}
	double t_dec = 0;
	int    dec_reps = 0;
	if (steps > 0) {
		for (int i = 0; i < warmup; i++) {
			ref_decode_column<PT>(&col, 0, n_vec, threads, out.data());
		}
		const double t0 = now();
		for (int i = 0; i < steps; i++) {
			ref_decode_column<PT>(&col, 0, n_vec, threads, out.data());
		}
		t_dec    = (now() - t0) / steps;
		dec_reps = steps;
	} else {
		t_dec = timed([&]() { ref_decode_column<PT>(&col, 0, n_vec, threads, out.data()); }, leg, &dec_reps);
	}
	const bool   ok = std::memcmp(out.data(), x.data(), n_values * sizeof(PT)) == 0;
	double       sum = 0;
	const double t_scan = timed([&]() { ref_sum_column<PT>(&col, 0, n_vec, threads, &sum); }, leg);
	const uint64_t n1    = std::min<uint64_t>(n_vec, 4096);
	const double   t_dec1 = timed([&]() { ref_decode_column<PT>(&col, 0, n1, 1, out.data()); }, 0.3);
	uint64_t       exc = 0, units_sum = 0;
	for (uint64_t v = 0; v < n_vec; v++) {
		exc += meta[v].exc_cnt;
		units_sum += meta[v].scheme == ALPB200_SCHEME_ALP_RD ? meta[v].bw + meta[v].e : meta[v].bw;
	}
	std::printf("{\"kind\": %d, \"values\": %llu, \"value_bytes\": %d, \"threads\": %d, \"round_trip_bit_exact\": %s, "
	            "\"decode_s\": %.9g, \"decode_reps\": %d, \"decode_single_thread_s_per_value\": %.9g, \"encode_s\": %.9g, \"init_s\": %.9g, "
	            "\"scan_s\": %.9g, \"scan_sum\": %.17g, \"packed_bytes\": %llu, \"exceptions\": %llu, \"build\": \"%s\"}\n",
	            kind, (unsigned long long)n_values, (int)sizeof(PT), threads, ok ? "true" : "false", t_dec, dec_reps, t_dec1 / (n1 * 1024.0), t_enc,
	            t_init, t_scan, sum, (unsigned long long)(units_sum * 128), (unsigned long long)exc, alpref_build_info());
	return ok ? 0 : 3;
}

} // namespace

int main(int argc, char** argv) {
	int      kind = 2, threads = 1, steps = 0, warmup = 1;
	uint64_t n = 1 << 22, first = 0;
	double   seconds = 4.0;
	for (int i = 1; i + 1 < argc; i += 2) {
		const std::string k = argv[i];
		const char*       v = argv[i + 1];
		if (k == "--kind") {
			kind = std::atoi(v);
		} else if (k == "--values") {
			n = std::strtoull(v, nullptr, 10);
		} else if (k == "--first") {
			first = std::strtoull(v, nullptr, 10);
		} else if (k == "--threads") {
			threads = std::max(1, std::atoi(v));
		} else if (k == "--steps") {
			steps = std::atoi(v);
		} else if (k == "--warmup") {
			warmup = std::atoi(v);
		} else if (k == "--seconds") {
			seconds = std::atof(v);
		} else {
			std::fprintf(stderr, "ref_bench: unknown option %s\n", argv[i]);
			return 1;
		}
	}
	if (kind == 4) { return run<float>(kind, n, first, threads, steps, warmup, seconds); }
	if (kind == 2 || kind == 3) { return run<double>(kind, n, first, threads, steps, warmup, seconds); }
	std::fprintf(stderr, "ref_bench: --kind must be 2, 3 or 4\n");
	return 1;
}
