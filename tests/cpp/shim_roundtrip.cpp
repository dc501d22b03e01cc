// tests/cpp/shim_roundtrip.cpp — a host program written against the REFERENCE's primitive API (the call sequence of
// cwida/ALP test/test_alp_sample.cpp:137-179), compiled against include/alp_b200.hpp instead of the reference's alp.hpp
// and linked to libalp_b200.so.  It shows the drop-in: not one call below is specific to this repository.
//
// usage: shim_roundtrip <cases.bin>      cases.bin = repeated records { u32 is_float, u32 golden_bw, u32 golden_exceptions,
//                                                                       u32 scheme, 1024 values (f64, or f32 padded to 8 KiB) }
#include "alp_b200.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

template <typename T>
static bool same_value(T a, T b) {
	if (std::isnan(a)) { return std::isnan(b); }
	return std::memcmp(&a, &b, sizeof(T)) == 0;  // distinguishes -0.0 from 0.0
}

template <typename PT>
static int run_case(const PT* input_arr, uint32_t golden_bw, uint32_t golden_exceptions, uint32_t want_scheme) {
	using UT = typename alp::inner_t<PT>::ut;
	using ST = typename alp::inner_t<PT>::st;
	constexpr size_t N = alp::config::VECTOR_SIZE;
	std::vector<PT>       sample_arr(N), exc_arr(N), dec_arr(N), glue_arr(N);
	std::vector<ST>       encoded_arr(N), ffor_arr(N), base_arr(N);
	std::vector<UT>       right_arr(N), ffor_right_arr(N), unffor_right_arr(N);
	std::vector<uint16_t> rd_exc_arr(N), pos_arr(N), exc_c_arr(N), left_arr(N), ffor_left_arr(N), unffor_left_arr(N);
	alp::bw_t             bit_width = 0;
	alp::state<PT>        stt;
	int                   bad = 0;

	alp::encoder<PT>::init(input_arr, 0, N, sample_arr.data(), stt);
	if (static_cast<uint32_t>(stt.scheme) != want_scheme) { return 1; }
	// the caller's sample array is filled like the reference's (encoder.hpp:420-427, sampler.hpp:14-52): values 0, 32, ..., 992
	bad += stt.sampled_values_n != 32;
	for (size_t i = 0; i < 32; i++) {
		bad += !same_value(sample_arr[i], input_arr[32 * i]);
	}
	{
		// rd.hpp:180-185: rd_encoder<PT>::init builds a dictionary for ANY row-group (bench_alp_cutter_encode.cpp:110 calls it on
		// columns the ALP search would keep): forced ALP_RD must round-trip every vector, decimal or not
		alp::state<PT> forced;
		alp::rd_encoder<PT>::init(input_arr, 0, N, sample_arr.data(), forced);
		bad += forced.scheme != alp::Scheme::ALP_RD;
		alp::rd_encoder<PT>::encode(input_arr, rd_exc_arr.data(), pos_arr.data(), exc_c_arr.data(), right_arr.data(), left_arr.data(), forced);
		alp::rd_encoder<PT>::decode(glue_arr.data(), right_arr.data(), left_arr.data(), rd_exc_arr.data(), pos_arr.data(), exc_c_arr.data(), forced);
		for (size_t i = 0; i < N; i++) {
			bad += !same_value(input_arr[i], glue_arr[i]);
		}
	}
	switch (stt.scheme) {
	case alp::Scheme::ALP_RD: {
		alp::rd_encoder<PT>::init(input_arr, 0, N, sample_arr.data(), stt);
		alp::rd_encoder<PT>::encode(input_arr, rd_exc_arr.data(), pos_arr.data(), exc_c_arr.data(), right_arr.data(), left_arr.data(), stt);
		ffor::ffor(right_arr.data(), ffor_right_arr.data(), stt.right_bit_width, &stt.right_for_base);
		ffor::ffor(left_arr.data(), ffor_left_arr.data(), stt.left_bit_width, &stt.left_for_base);
		unffor::unffor(ffor_right_arr.data(), unffor_right_arr.data(), stt.right_bit_width, &stt.right_for_base);
		unffor::unffor(ffor_left_arr.data(), unffor_left_arr.data(), stt.left_bit_width, &stt.left_for_base);
		alp::rd_encoder<PT>::decode(glue_arr.data(), unffor_right_arr.data(), unffor_left_arr.data(), rd_exc_arr.data(), pos_arr.data(),
		                            exc_c_arr.data(), stt);
		for (size_t i = 0; i < N; i++) {
			bad += !same_value(input_arr[i], glue_arr[i]);
		}
		break;
	}
	case alp::Scheme::ALP: {
		alp::encoder<PT>::encode(input_arr, exc_arr.data(), pos_arr.data(), exc_c_arr.data(), encoded_arr.data(), stt);
		alp::encoder<PT>::analyze_ffor(encoded_arr.data(), bit_width, base_arr.data());
		ffor::ffor(encoded_arr.data(), ffor_arr.data(), bit_width, base_arr.data());
		generated::falp::fallback::scalar::falp(ffor_arr.data(), dec_arr.data(), bit_width, base_arr.data(), stt.fac, stt.exp);
		alp::decoder<PT>::patch_exceptions(dec_arr.data(), exc_arr.data(), pos_arr.data(), exc_c_arr.data());
		for (size_t i = 0; i < N; i++) {
			bad += !same_value(input_arr[i], dec_arr[i]);
		}
		bad += exc_c_arr[0] != golden_exceptions;  // the reference test's two golden asserts
		bad += bit_width != golden_bw;
		break;
	}
	default: bad++;
	}
	return bad;
}

// the 8-bit-lane overloads (fastlanes/ffor.hpp:10,15; unffor.hpp:10,15): every width, round trip
static int run_u8() {
	std::vector<uint8_t> in(1024), packed(1024), out(1024);
	std::vector<int8_t>  sin(1024), spacked(1024), sout(1024);
	int                  bad = 0;
	for (uint8_t bw = 0; bw <= 8; bw++) {
		const uint8_t base = static_cast<uint8_t>(17 * bw + 3);
		for (size_t i = 0; i < 1024; i++) {
			in[i]  = static_cast<uint8_t>(base + ((i * 7 + i / 128) & ((1u << bw) - 1u)));
			sin[i] = static_cast<int8_t>(in[i]);
		}
		ffor::ffor(in.data(), packed.data(), bw, &base);
		unffor::unffor(packed.data(), out.data(), bw, &base);
		const int8_t sbase = static_cast<int8_t>(base);
		ffor::ffor(sin.data(), spacked.data(), bw, &sbase);
		unffor::unffor(spacked.data(), sout.data(), bw, &sbase);
		for (size_t i = 0; i < 1024; i++) {
			bad += out[i] != in[i];
			bad += sout[i] != sin[i];
		}
	}
	return bad;
}

int main(int argc, char** argv) {
	if (argc < 2) { return 2; }
	FILE* f = std::fopen(argv[1], "rb");
	if (!f) { return 2; }
	uint32_t             hdr[4];
	std::vector<uint8_t> payload(8192);
	int                  n = 0, bad = 0;
	try {
		while (std::fread(hdr, 4, 4, f) == 4 && std::fread(payload.data(), 1, 8192, f) == 8192) {
			const int b = hdr[0] ? run_case<float>(reinterpret_cast<const float*>(payload.data()), hdr[1], hdr[2], hdr[3])
			                     : run_case<double>(reinterpret_cast<const double*>(payload.data()), hdr[1], hdr[2], hdr[3]);
			if (b) { std::printf("case %d: %d mismatches\n", n, b); }
			bad += b != 0;
			n++;
		}
	} catch (const alp::gpu_error& e) {
		std::printf("gpu_error %d: %s\n", e.code, e.what());
		return 3;
	}
	std::fclose(f);
	try {
		const int b8 = run_u8();
		if (b8) { std::printf("8-bit lanes: %d mismatches\n", b8); }
		bad += b8 != 0;
	} catch (const alp::gpu_error& e) {
		std::printf("gpu_error %d: %s\n", e.code, e.what());
		return 3;
	}
	std::printf("cases %d bad %d\n", n, bad);
	return bad ? 1 : 0;
}
