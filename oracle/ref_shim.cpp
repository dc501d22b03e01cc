/*
 * oracle/ref_shim.cpp — TEST INFRASTRUCTURE, not product code.
 *
 * A C-ABI face on the UNMODIFIED reference (cwida/ALP), compiled from the sources where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/libalp_ref_*.so.  Nothing from the reference is copied
 * into this repository: this file only #includes the reference headers at build time and calls the
 * reference's own primitives.  It is used (a) to pin the C restatement in oracle/alp_oracle.c,
 * (b) to generate tests/golden fixtures, (c) by the GPU parity tests as the ground truth, and
 * (d) as the CPU baseline bench.py times (`cpu_baseline.kind = "reference"`).
 *
 * The TU must be compiled WITHOUT AVX-512F: include/alp/encoder.hpp:351-371 only compiles with clang when
 * __AVX512F__ is defined.  The generated FFOR/UNFFOR/FALP kernels (src/ files) are built with
 * the vector ISA in the library name.
 */
#include "alp.hpp"
#include "alp_b200.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace {

template <typename PT>
void to_ref_state(const alpb200_rg_state* s, alp::state<PT>& stt) {
	stt.scheme         = static_cast<alp::Scheme>(s->scheme);
	stt.k_combinations = static_cast<uint16_t>(s->k);
	stt.best_k_combinations.clear();
	for (int i = 0; i < s->k; i++) {
		stt.best_k_combinations.emplace_back(s->combos[i][0], s->combos[i][1]);
	}
	stt.right_bit_width        = s->right_bw;
	stt.left_bit_width         = s->left_bw;
	stt.actual_dictionary_size = s->dict_size;
	stt.left_parts_dict_map.clear();
	for (int i = 0; i < ALPB200_RD_DICT_SIZE; i++) {
		stt.left_parts_dict[i] = s->dict[i];
	}
	for (int i = 0; i < s->dict_size; i++) {
		stt.left_parts_dict_map.insert({s->dict[i], static_cast<uint16_t>(i)});
	}
	for (int i = 0; i < s->n_extra; i++) {
		stt.left_parts_dict_map.insert({s->extra_key[i], s->extra_idx[i]});
	}
}

template <typename PT>
void from_ref_state(const alp::state<PT>& stt, alpb200_rg_state* s) {
	std::memset(s, 0, sizeof(*s));
	s->scheme = static_cast<int32_t>(stt.scheme);
	if (stt.scheme == alp::Scheme::ALP) {
		s->k = stt.k_combinations;
		for (int i = 0; i < s->k && i < ALPB200_MAX_K; i++) {
			s->combos[i][0] = static_cast<uint8_t>(stt.best_k_combinations[i].first);
			s->combos[i][1] = static_cast<uint8_t>(stt.best_k_combinations[i].second);
		}
	} else {
		s->right_bw  = stt.right_bit_width;
		s->left_bw   = stt.left_bit_width;
		s->dict_size = stt.actual_dictionary_size;
		for (int i = 0; i < ALPB200_RD_DICT_SIZE; i++) {
			s->dict[i] = stt.left_parts_dict[i];
		}
		int n = 0;
		for (auto const& kv : stt.left_parts_dict_map) {
			if (kv.second >= stt.actual_dictionary_size && n < ALPB200_MAX_SAMPLES) {
				s->extra_key[n] = kv.first;
				s->extra_idx[n] = kv.second;
				n++;
			}
		}
		s->n_extra = static_cast<uint16_t>(n);
	}
}

template <typename PT>
void ref_init(const PT* col, size_t offset, size_t n, alpb200_rg_state* out) {
	// the call sequence of test/test_alp_sample.cpp:137-141 and benchmarks/benchmark.cpp:200-230
	std::vector<PT> sample(alp::config::VECTOR_SIZE);
	alp::state<PT>  stt;
	alp::encoder<PT>::init(col, offset, n, sample.data(), stt);
	if (stt.scheme == alp::Scheme::ALP_RD) { alp::rd_encoder<PT>::init(col, offset, n, sample.data(), stt); }
	from_ref_state(stt, out);
}

/* ---- column drivers: the loop a caller of the reference writes (benchmarks/benchmark.cpp:200-285) ---- */

template <typename PT>
struct VecOut {
	alpb200_vec_meta                 meta;
	std::vector<uint8_t>             packed;
	std::vector<PT>                  exc;   // ALP exceptions or zero-extended RD left parts (bit pattern)
	std::vector<uint16_t>            pos;
};

template <typename PT>
void encode_rowgroup(const PT* col, size_t n_values, size_t rg, std::vector<VecOut<PT>>& out) {
	using UT = typename alp::inner_t<PT>::ut;
	using ST = typename alp::inner_t<PT>::st;
	const size_t first  = rg * alp::config::ROWGROUP_SIZE;
	const size_t n_vec  = (std::min(n_values - first, alp::config::ROWGROUP_SIZE)) / alp::config::VECTOR_SIZE;
	std::vector<PT> sample(alp::config::VECTOR_SIZE);
	alp::state<PT>  stt;
	alp::encoder<PT>::init(col, first, n_values, sample.data(), stt);
	const bool rd = stt.scheme == alp::Scheme::ALP_RD;
	if (rd) { alp::rd_encoder<PT>::init(col, first, n_values, sample.data(), stt); }

	alignas(64) PT       exc[1024];
	alignas(64) uint16_t rdexc[1024], pos[1024], cnt[8], left[1024], fleft[1024];
	alignas(64) ST       enc[1024], base[8], ff[1024];
	alignas(64) UT       right[1024], fright[1024];

	for (size_t v = 0; v < n_vec; v++) {
		const PT*   in = col + first + v * alp::config::VECTOR_SIZE;
		VecOut<PT>& o  = out[rg * alp::config::N_VECTORS_PER_ROWGROUP + v];
		std::memset(&o.meta, 0, sizeof(o.meta));
		if (rd) {
			alp::rd_encoder<PT>::encode(in, rdexc, pos, cnt, right, left, stt);
			ffor::ffor(right, fright, stt.right_bit_width, &stt.right_for_base);
			ffor::ffor(left, fleft, stt.left_bit_width, &stt.left_for_base);
			o.meta.scheme  = ALPB200_SCHEME_ALP_RD;
			o.meta.bw      = stt.right_bit_width;
			o.meta.e       = stt.left_bit_width;
			o.meta.f       = stt.actual_dictionary_size;
			o.meta.exc_cnt = cnt[0];
			for (int i = 0; i < ALPB200_RD_DICT_SIZE; i++) {
				o.meta.u.rd_dict[i] = stt.left_parts_dict[i];
			}
			o.packed.resize(128u * (stt.right_bit_width + stt.left_bit_width));
			std::memcpy(o.packed.data(), fright, 128u * stt.right_bit_width);
			std::memcpy(o.packed.data() + 128u * stt.right_bit_width, fleft, 128u * stt.left_bit_width);
			o.exc.resize(cnt[0]);
			o.pos.assign(pos, pos + cnt[0]);
			for (int i = 0; i < cnt[0]; i++) {
				UT bits = rdexc[i];
				std::memcpy(&o.exc[i], &bits, sizeof(UT));
			}
		} else {
			uint8_t bw = 0;
			alp::encoder<PT>::encode(in, exc, pos, cnt, enc, stt);
			alp::encoder<PT>::analyze_ffor(enc, bw, base);
			ffor::ffor(enc, ff, bw, base);
			o.meta.scheme     = ALPB200_SCHEME_ALP;
			o.meta.bw         = bw;
			o.meta.e          = stt.exp;
			o.meta.f          = stt.fac;
			o.meta.exc_cnt    = cnt[0];
			o.meta.u.alp.base = static_cast<int64_t>(base[0]);
			o.packed.resize(128u * bw);
			std::memcpy(o.packed.data(), ff, 128u * bw);
			o.exc.assign(exc, exc + cnt[0]);
			o.pos.assign(pos, pos + cnt[0]);
		}
	}
}

template <typename PT>
int ref_encode_column(const PT* col, size_t n_values, int n_threads, alpb200_column* out) {
	const size_t n_vec = n_values / alp::config::VECTOR_SIZE;
	const size_t n_rg  = (n_vec + alp::config::N_VECTORS_PER_ROWGROUP - 1) / alp::config::N_VECTORS_PER_ROWGROUP;
	std::vector<VecOut<PT>> vecs(n_rg * alp::config::N_VECTORS_PER_ROWGROUP);
	std::atomic<size_t>     next {0};
	auto                    work = [&]() {
        for (size_t rg = next.fetch_add(1); rg < n_rg; rg = next.fetch_add(1)) {
            encode_rowgroup<PT>(col, n_vec * alp::config::VECTOR_SIZE, rg, vecs);
        }
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < n_threads; t++) {
		pool.emplace_back(work);
	}
	work();
	for (auto& t : pool) {
		t.join();
	}
	// serialise in vector order
	uint64_t poff = 0, eoff = 0;
	int      overflow = 0;
	out->n_vectors = n_vec;
	for (size_t v = 0; v < n_vec; v++) {
		VecOut<PT>& o = vecs[v];
		if (poff + o.packed.size() > out->packed_capacity || eoff + o.pos.size() > out->exc_capacity) {
			overflow = 1;
			break;
		}
		o.meta.packed_off = static_cast<uint32_t>(poff / 128);
		o.meta.exc_off    = static_cast<uint32_t>(eoff);
		out->meta[v]      = o.meta;
		if (!o.packed.empty()) { std::memcpy(out->packed + poff, o.packed.data(), o.packed.size()); }
		if (!o.pos.empty()) {
			std::memcpy(static_cast<PT*>(out->exc_val) + eoff, o.exc.data(), o.exc.size() * sizeof(PT));
			std::memcpy(out->exc_pos + eoff, o.pos.data(), o.pos.size() * sizeof(uint16_t));
		}
		poff += o.packed.size();
		eoff += o.pos.size();
	}
	if (out->totals) {
		out->totals[0] = poff;
		out->totals[1] = eoff;
		out->totals[2] = overflow;
	}
	return overflow ? ALPB200_ECAPACITY : ALPB200_OK;
}

template <typename PT>
void decode_vector(const alpb200_column* col, size_t v, PT* out) {
	using UT = typename alp::inner_t<PT>::ut;
	using ST = typename alp::inner_t<PT>::st;
	const alpb200_vec_meta& m      = col->meta[v];
	const uint8_t*          packed = col->packed + static_cast<uint64_t>(m.packed_off) * 128u;
	const PT*               excv   = static_cast<const PT*>(col->exc_val) + m.exc_off;
	const uint16_t*         excp   = col->exc_pos + m.exc_off;
	if (m.scheme == ALPB200_SCHEME_ALP) {
		// test/test_alp_sample.cpp:169-170: fused falp + patch_exceptions (unfused at bw == lane width, see
		// SURVEY.md §7 "falp at full width is wrong in the reference")
		UT base = static_cast<UT>(static_cast<ST>(m.u.alp.base));
		if (m.bw == sizeof(UT) * 8) {
			alignas(64) UT unpacked[1024];
			unffor::unffor(reinterpret_cast<const UT*>(packed), unpacked, m.bw, &base);
			alp::decoder<PT>::decode(reinterpret_cast<const ST*>(unpacked), m.f, m.e, out);
		} else {
			generated::falp::fallback::scalar::falp(reinterpret_cast<const UT*>(packed), out, m.bw, &base, m.f, m.e);
		}
		alp::decoder<PT>::patch_exceptions(out, excv, excp, &m.exc_cnt);
	} else {
		// test/test_alp_sample.cpp:148-151
		alignas(64) UT       right[1024];
		alignas(64) uint16_t left[1024], rdexc[1024], pos[1024];
		alp::state<PT>       stt;
		stt.right_bit_width        = m.bw;
		stt.left_bit_width         = m.e;
		stt.actual_dictionary_size = m.f;
		for (int i = 0; i < ALPB200_RD_DICT_SIZE; i++) {
			stt.left_parts_dict[i] = m.u.rd_dict[i];
		}
		UT       zero   = 0;
		uint16_t zero16 = 0;
		unffor::unffor(reinterpret_cast<const UT*>(packed), right, m.bw, &zero);
		unffor::unffor(reinterpret_cast<const uint16_t*>(packed + 128u * m.bw), left, m.e, &zero16);
		for (int i = 0; i < m.exc_cnt; i++) {
			UT bits;
			std::memcpy(&bits, &excv[i], sizeof(UT));
			rdexc[i] = static_cast<uint16_t>(bits);
			pos[i]   = excp[i];
		}
		uint16_t cnt = m.exc_cnt;
		alp::rd_encoder<PT>::decode(out, right, left, rdexc, pos, &cnt, stt);
	}
}

template <typename PT>
int ref_decode_column(const alpb200_column* col, size_t first, size_t n, int n_threads, PT* out) {
	std::atomic<size_t> next {0};
	const size_t        chunk = 64;
	auto                work  = [&]() {
        for (size_t c = next.fetch_add(chunk); c < n; c = next.fetch_add(chunk)) {
            for (size_t v = c; v < std::min(n, c + chunk); v++) {
                decode_vector<PT>(col, first + v, out + v * alp::config::VECTOR_SIZE);
            }
        }
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < n_threads; t++) {
		pool.emplace_back(work);
	}
	work();
	for (auto& t : pool) {
		t.join();
	}
	return ALPB200_OK;
}


/* ---- bench drivers (bench.py's CPU legs): the same call sequences, laid out for throughput ------------------------
 * Row-groups are dealt to the threads as contiguous ranges and every thread writes its blocks / exceptions into its own
 * slice of the container (slice k of n_threads of the packed and exception arrays), so there is no serial pass and no
 * per-vector allocation inside the timed region.  The result is a valid column container (records carry absolute
 * offsets; the slices leave gaps between them), which the caller decodes to check the timed run.
 * mode bit 0: run alp::encoder<PT>::init (+ rd_encoder<PT>::init) per row-group — test/test_alp_sample.cpp:137-141 —
 *             otherwise take the states from `states`;  bit 1: encode the vectors — :143-145 / :164-166
 *             (encode + analyze_ffor + ffor, or rd encode + 2x ffor); without it only the states are produced. */
template <typename PT>
int ref_bench_encode(const PT* col, size_t n_values, int n_threads, alpb200_column* out, alpb200_rg_state* states, int mode) {
	using UT = typename alp::inner_t<PT>::ut;
	using ST = typename alp::inner_t<PT>::st;
	const size_t n_vec = n_values / alp::config::VECTOR_SIZE;
	const size_t n_rg  = (n_vec + alp::config::N_VECTORS_PER_ROWGROUP - 1) / alp::config::N_VECTORS_PER_ROWGROUP;
	n_values           = n_vec * alp::config::VECTOR_SIZE;
	if (n_threads < 1) { n_threads = 1; }
	const bool do_init = (mode & 1) != 0, do_encode = (mode & 2) != 0;
	if ((!do_init && !states) || (do_encode && !out)) { return ALPB200_EINVAL; }
	std::atomic<int>      overflow {0};
	std::vector<uint64_t> used_p(n_threads, 0), used_e(n_threads, 0);
	auto                  work = [&](int k) {
        const size_t rg0 = n_rg * k / n_threads, rg1 = n_rg * (k + 1) / n_threads;
        uint64_t     p_lo = 0, p_hi = 0, e_lo = 0, e_hi = 0;
        if (do_encode) {
            p_lo = (out->packed_capacity / n_threads * k) & ~uint64_t(127);
            p_hi = (out->packed_capacity / n_threads * (k + 1)) & ~uint64_t(127);
            e_lo = out->exc_capacity / n_threads * k;
            e_hi = out->exc_capacity / n_threads * (k + 1);
        }
        uint64_t        poff = p_lo, eoff = e_lo;
        std::vector<PT> sample(alp::config::VECTOR_SIZE);
        alignas(64) PT       exc[1024];
        alignas(64) uint16_t rdexc[1024], pos[1024], cnt[8], left[1024];
        alignas(64) ST       enc[1024], base[8];
        alignas(64) UT       right[1024];
        for (size_t rg = rg0; rg < rg1; rg++) {
            const size_t   first = rg * alp::config::ROWGROUP_SIZE;
            const size_t   nv    = std::min(n_values - first, alp::config::ROWGROUP_SIZE) / alp::config::VECTOR_SIZE;
            alp::state<PT> stt;
            if (do_init) {
                alp::encoder<PT>::init(col, first, n_values, sample.data(), stt);
                if (stt.scheme == alp::Scheme::ALP_RD) { alp::rd_encoder<PT>::init(col, first, n_values, sample.data(), stt); }
                if (states) { from_ref_state(stt, &states[rg]); }
            } else {
                to_ref_state(&states[rg], stt);
            }
            if (!do_encode) { continue; }
            const bool rd = stt.scheme == alp::Scheme::ALP_RD;
            for (size_t v = 0; v < nv; v++) {
                const PT*         in = col + first + v * alp::config::VECTOR_SIZE;
                alpb200_vec_meta& m  = out->meta[rg * alp::config::N_VECTORS_PER_ROWGROUP + v];
                std::memset(&m, 0, sizeof(m));
                const uint64_t worst = 128u * (sizeof(PT) * 8 + 3);
                if (poff + worst > p_hi || eoff + 1024 > e_hi) {
                    overflow = 1;
                    return;
                }
                uint8_t* dst = out->packed + poff;
                if (rd) {
                    alp::rd_encoder<PT>::encode(in, rdexc, pos, cnt, right, left, stt);
                    ffor::ffor(right, reinterpret_cast<UT*>(dst), stt.right_bit_width, &stt.right_for_base);
                    ffor::ffor(left, reinterpret_cast<uint16_t*>(dst + 128u * stt.right_bit_width), stt.left_bit_width, &stt.left_for_base);
                    m.scheme = ALPB200_SCHEME_ALP_RD;
                    m.bw     = stt.right_bit_width;
                    m.e      = stt.left_bit_width;
                    m.f      = stt.actual_dictionary_size;
                    for (int i = 0; i < ALPB200_RD_DICT_SIZE; i++) {
                        m.u.rd_dict[i] = stt.left_parts_dict[i];
                    }
                    UT* ev = static_cast<UT*>(out->exc_val) + eoff;
                    for (int i = 0; i < cnt[0]; i++) {
                        ev[i] = rdexc[i];
                    }
                    m.packed_off = static_cast<uint32_t>(poff / 128);
                    poff += 128u * (stt.right_bit_width + stt.left_bit_width);
                } else {
                    uint8_t bw = 0;
                    alp::encoder<PT>::encode(in, exc, pos, cnt, enc, stt);
                    alp::encoder<PT>::analyze_ffor(enc, bw, base);
                    ffor::ffor(enc, reinterpret_cast<ST*>(dst), bw, base);
                    m.scheme     = ALPB200_SCHEME_ALP;
                    m.bw         = bw;
                    m.e          = stt.exp;
                    m.f          = stt.fac;
                    m.u.alp.base = static_cast<int64_t>(base[0]);
                    std::memcpy(static_cast<PT*>(out->exc_val) + eoff, exc, cnt[0] * sizeof(PT));
                    m.packed_off = static_cast<uint32_t>(poff / 128);
                    poff += 128u * bw;
                }
                std::memcpy(out->exc_pos + eoff, pos, cnt[0] * sizeof(uint16_t));
                m.exc_off = static_cast<uint32_t>(eoff);
                m.exc_cnt = cnt[0];
                eoff += cnt[0];
            }
        }
        used_p[k] = poff - p_lo;
        used_e[k] = eoff - e_lo;
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < n_threads; t++) {
		pool.emplace_back(work, t);
	}
	work(0);
	for (auto& t : pool) {
		t.join();
	}
	if (do_encode) {
		out->n_vectors = n_vec;
		if (out->totals) {  // bytes / slots actually written (the container's slices are not dense)
			uint64_t p = 0, e = 0;
			for (int k = 0; k < n_threads; k++) {
				p += used_p[k];
				e += used_e[k];
			}
			out->totals[0] = p;
			out->totals[1] = e;
			out->totals[2] = overflow.load();
		}
	}
	return overflow.load() ? ALPB200_ECAPACITY : ALPB200_OK;
}

/* The reference's scan query: per vector the scan primitive `alp_func` (fused falp + patch_exceptions into a
 * thread-private 1024-value buffer, publication/source_code/bench_end_to_end/src/benchmarks/alp/queries/q1.cpp:63-89)
 * followed by `aggr_plus` (out[0] += in[i], q1.cpp:91-100), on n_threads workers over disjoint morsels of vectors
 * (TBB workers in the reference, q1.cpp:650-679); the workers' aggregates are added at the end. */
template <typename PT>
int ref_sum_column(const alpb200_column* col, size_t first, size_t n, int n_threads, double* out) {
	if (n_threads < 1) { n_threads = 1; }
	std::atomic<size_t> next {0};
	const size_t        chunk = 64;
	std::vector<double> partial(n_threads, 0.0);
	auto                work = [&](int k) {
        alignas(64) PT buf[1024];
        double         aggr = 0.0;
        for (size_t c = next.fetch_add(chunk); c < n; c = next.fetch_add(chunk)) {
            for (size_t v = c; v < std::min(n, c + chunk); v++) {
                decode_vector<PT>(col, first + v, buf);
                for (size_t i = 0; i < 1024; ++i) {
                    aggr += buf[i];
                }
            }
        }
        partial[k] = aggr;
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < n_threads; t++) {
		pool.emplace_back(work, t);
	}
	work(0);
	for (auto& t : pool) {
		t.join();
	}
	double total = 0.0;
	for (double p : partial) {
		total += p;
	}
	*out = total;
	return ALPB200_OK;
}

} // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* alpref_build_info() {
#if defined(ALPREF_ISA)
	return "cwida/ALP reference, g++ " __VERSION__ ", isa " ALPREF_ISA;
#else
	return "cwida/ALP reference, g++ " __VERSION__;
#endif
}

/* struct sizes this library was compiled with — lets the Python driver reject a stale build */
void alpref_abi_sizes(uint32_t* out) {
	out[0] = sizeof(alpb200_rg_state);
	out[1] = sizeof(alpb200_vec_meta);
	out[2] = sizeof(alpb200_column);
}

/* alp::encoder<PT>::init (+ rd_encoder<PT>::init) */
void alpref_init_f64(const double* col, size_t offset, size_t n, alpb200_rg_state* st) { ref_init<double>(col, offset, n, st); }
void alpref_init_f32(const float* col, size_t offset, size_t n, alpb200_rg_state* st) { ref_init<float>(col, offset, n, st); }

/* alp::encoder<PT>::encode */
void alpref_encode_f64(const double* in, const alpb200_rg_state* st, double* exc, uint16_t* pos, uint16_t* cnt,
                       int64_t* enc, uint8_t* e, uint8_t* f) {
	alp::state<double> stt;
	to_ref_state(st, stt);
	alp::encoder<double>::encode(in, exc, pos, cnt, enc, stt);
	*e = stt.exp;
	*f = stt.fac;
}
void alpref_encode_f32(const float* in, const alpb200_rg_state* st, float* exc, uint16_t* pos, uint16_t* cnt,
                       int32_t* enc, uint8_t* e, uint8_t* f) {
	alp::state<float> stt;
	to_ref_state(st, stt);
	alp::encoder<float>::encode(in, exc, pos, cnt, enc, stt);
	*e = stt.exp;
	*f = stt.fac;
}
/* alp::encoder<PT>::encode_value<true> / decoder<PT>::decode_value — for sampling-path parity */
int64_t alpref_encode_value_f64(double v, uint8_t f, uint8_t e) { return alp::encoder<double>::encode_value(v, f, e); }
int32_t alpref_encode_value_f32(float v, uint8_t f, uint8_t e) { return alp::encoder<float>::encode_value(v, f, e); }
double  alpref_decode_value_f64(int64_t v, uint8_t f, uint8_t e) { return alp::decoder<double>::decode_value(v, f, e); }
float   alpref_decode_value_f32(int32_t v, uint8_t f, uint8_t e) { return alp::decoder<float>::decode_value(v, f, e); }

void alpref_analyze_ffor_i64(const int64_t* enc, uint8_t* bw, int64_t* base) {
	alp::encoder<double>::analyze_ffor(enc, *bw, base);
}
void alpref_analyze_ffor_i32(const int32_t* enc, uint8_t* bw, int32_t* base) {
	alp::encoder<float>::analyze_ffor(enc, *bw, base);
}

void alpref_ffor_u64(const uint64_t* in, uint64_t* out, uint8_t bw, uint64_t base) { ffor::ffor(in, out, bw, &base); }
void alpref_ffor_u32(const uint32_t* in, uint32_t* out, uint8_t bw, uint32_t base) { ffor::ffor(in, out, bw, &base); }
void alpref_ffor_u16(const uint16_t* in, uint16_t* out, uint8_t bw, uint16_t base) { ffor::ffor(in, out, bw, &base); }
void alpref_ffor_u8(const uint8_t* in, uint8_t* out, uint8_t bw, uint8_t base) { ffor::ffor(in, out, bw, &base); }
void alpref_unffor_u8(const uint8_t* in, uint8_t* out, uint8_t bw, uint8_t base) { unffor::unffor(in, out, bw, &base); }
void alpref_unffor_u64(const uint64_t* in, uint64_t* out, uint8_t bw, uint64_t base) { unffor::unffor(in, out, bw, &base); }
void alpref_unffor_u32(const uint32_t* in, uint32_t* out, uint8_t bw, uint32_t base) { unffor::unffor(in, out, bw, &base); }
void alpref_unffor_u16(const uint16_t* in, uint16_t* out, uint8_t bw, uint16_t base) { unffor::unffor(in, out, bw, &base); }

/* the reference's fused kernel exactly as shipped (including its bw == lane-width quirk) */
void alpref_falp_f64(const uint64_t* in, double* out, uint8_t bw, uint64_t base, uint8_t f, uint8_t e) {
	generated::falp::fallback::scalar::falp(in, out, bw, &base, f, e);
}
void alpref_falp_f32(const uint32_t* in, float* out, uint8_t bw, uint32_t base, uint8_t f, uint8_t e) {
	generated::falp::fallback::scalar::falp(in, out, bw, &base, f, e);
}
void alpref_decode_f64(const int64_t* enc, uint8_t f, uint8_t e, double* out) { alp::decoder<double>::decode(enc, f, e, out); }
void alpref_decode_f32(const int32_t* enc, uint8_t f, uint8_t e, float* out) { alp::decoder<float>::decode(enc, f, e, out); }
void alpref_patch_f64(double* out, const double* exc, const uint16_t* pos, uint16_t cnt) {
	alp::decoder<double>::patch_exceptions(out, exc, pos, &cnt);
}
void alpref_patch_f32(float* out, const float* exc, const uint16_t* pos, uint16_t cnt) {
	alp::decoder<float>::patch_exceptions(out, exc, pos, &cnt);
}

void alpref_rd_encode_f64(const double* in, const alpb200_rg_state* st, uint16_t* exc, uint16_t* pos, uint16_t* cnt,
                          uint64_t* right, uint16_t* left) {
	alp::state<double> stt;
	to_ref_state(st, stt);
	alp::rd_encoder<double>::encode(in, exc, pos, cnt, right, left, stt);
}
void alpref_rd_encode_f32(const float* in, const alpb200_rg_state* st, uint16_t* exc, uint16_t* pos, uint16_t* cnt,
                          uint32_t* right, uint16_t* left) {
	alp::state<float> stt;
	to_ref_state(st, stt);
	alp::rd_encoder<float>::encode(in, exc, pos, cnt, right, left, stt);
}
void alpref_rd_decode_f64(double* out, const uint64_t* right, const uint16_t* left, const uint16_t* exc,
                          const uint16_t* pos, uint16_t cnt, const alpb200_rg_state* st) {
	alp::state<double> stt;
	to_ref_state(st, stt);
	alp::rd_encoder<double>::decode(out, const_cast<uint64_t*>(right), const_cast<uint16_t*>(left),
	                                const_cast<uint16_t*>(exc), const_cast<uint16_t*>(pos), &cnt, stt);
}
void alpref_rd_decode_f32(float* out, const uint32_t* right, const uint16_t* left, const uint16_t* exc,
                          const uint16_t* pos, uint16_t cnt, const alpb200_rg_state* st) {
	alp::state<float> stt;
	to_ref_state(st, stt);
	alp::rd_encoder<float>::decode(out, const_cast<uint32_t*>(right), const_cast<uint16_t*>(left),
	                               const_cast<uint16_t*>(exc), const_cast<uint16_t*>(pos), &cnt, stt);
}

/* column drivers (host pointers in the container) */
int alpref_encode_column_f64(const double* col, size_t n_values, int n_threads, alpb200_column* out) {
	return ref_encode_column<double>(col, n_values, n_threads, out);
}
int alpref_encode_column_f32(const float* col, size_t n_values, int n_threads, alpb200_column* out) {
	return ref_encode_column<float>(col, n_values, n_threads, out);
}
int alpref_decode_column_f64(const alpb200_column* col, size_t first, size_t n, int n_threads, double* out) {
	return ref_decode_column<double>(col, first, n, n_threads, out);
}
int alpref_decode_column_f32(const alpb200_column* col, size_t first, size_t n, int n_threads, float* out) {
	return ref_decode_column<float>(col, first, n, n_threads, out);
}

/* bench drivers (see ref_bench_encode / ref_sum_column above) */
int alpref_bench_encode_f64(const double* col, size_t n_values, int n_threads, alpb200_column* out, alpb200_rg_state* states, int mode) {
	return ref_bench_encode<double>(col, n_values, n_threads, out, states, mode);
}
int alpref_bench_encode_f32(const float* col, size_t n_values, int n_threads, alpb200_column* out, alpb200_rg_state* states, int mode) {
	return ref_bench_encode<float>(col, n_values, n_threads, out, states, mode);
}
int alpref_sum_column_f64(const alpb200_column* col, size_t first, size_t n, int n_threads, double* out) {
	return ref_sum_column<double>(col, first, n, n_threads, out);
}
int alpref_sum_column_f32(const alpb200_column* col, size_t first, size_t n, int n_threads, double* out) {
	return ref_sum_column<float>(col, first, n, n_threads, out);
}


} // extern "C"
#pragma GCC visibility pop
