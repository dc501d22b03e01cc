// gtest/gtest.h — a minimal stand-in for GoogleTest, just large enough for the reference's own unit test
// (test/test_alp_sample.cpp) to compile UNCHANGED in a container without gtest or network: fixtures (TEST_F with
// ::testing::Test, SetUp/TearDown), ASSERT_EQ / ASSERT_TRUE with GoogleTest's "return from the current function on
// failure" behaviour, a registry and a main().  Test infrastructure only.
#pragma once

#include <cstdio>
#include <exception>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

class Test {
public:
	virtual ~Test() = default;
	virtual void SetUp() {}
	virtual void TearDown() {}
	virtual void TestBody() = 0;
};

namespace internal {
struct Case {
	std::string                name;
	std::function<Test*()>     make;
};
inline std::vector<Case>& registry() {
	static std::vector<Case> r;
	return r;
}
inline int& failures_of_current_test() {
	static int n = 0;
	return n;
}
struct Registrar {
	Registrar(const char* suite, const char* name, std::function<Test*()> make) {
		registry().push_back({std::string(suite) + "." + name, std::move(make)});
	}
};
template <typename A, typename B>
bool check_eq(const A& a, const B& b, const char* ea, const char* eb, const char* file, int line) {
	if (a == b) { return true; }
	std::ostringstream os;
	os << file << ":" << line << ": Failure\nExpected equality of these values:\n  " << ea << "\n    Which is: " << +a << "\n  " << eb
	   << "\n    Which is: " << +b << "\n";
	std::cout << os.str();
	failures_of_current_test()++;
	return false;
}
inline bool check_true(bool v, const char* e, const char* file, int line) {
	if (v) { return true; }
	std::cout << file << ":" << line << ": Failure\nValue of: " << e << "\n  Actual: false\nExpected: true\n";
	failures_of_current_test()++;
	return false;
}
inline int run_all() {
	int failed = 0;
	std::cout << "[==========] Running " << registry().size() << " tests.\n";
	for (auto& c : registry()) {
		std::cout << "[ RUN      ] " << c.name << "\n";
		failures_of_current_test() = 0;
		try {
			Test* t = c.make();
			t->SetUp();
			t->TestBody();
			t->TearDown();
			delete t;
		} catch (const std::exception& e) {
			std::cout << "unexpected exception: " << e.what() << "\n";
			failures_of_current_test()++;
		}
		if (failures_of_current_test()) {
			failed++;
			std::cout << "[  FAILED  ] " << c.name << "\n";
		} else {
			std::cout << "[       OK ] " << c.name << "\n";
		}
	}
	std::cout << "[==========] " << registry().size() << " tests ran.\n";
	std::cout << "[  PASSED  ] " << (registry().size() - failed) << " tests.\n";
	if (failed) { std::cout << "[  FAILED  ] " << failed << " tests.\n"; }
	return failed ? 1 : 0;
}
}  // namespace internal
}  // namespace testing

#define GTEST_STUB_CLASS_(suite, name) suite##_##name##_Test

#define TEST_F(suite, name)                                                                                   \
	class GTEST_STUB_CLASS_(suite, name) : public suite {                                                     \
	public:                                                                                                   \
		void TestBody() override;                                                                             \
	};                                                                                                        \
	static ::testing::internal::Registrar gtest_stub_registrar_##suite##_##name(                              \
	    #suite, #name, []() -> ::testing::Test* { return new GTEST_STUB_CLASS_(suite, name)(); });            \
	void GTEST_STUB_CLASS_(suite, name)::TestBody()

#define ASSERT_EQ(a, b)                                                                        \
	if (!::testing::internal::check_eq((a), (b), #a, #b, __FILE__, __LINE__)) return
#define ASSERT_TRUE(v)                                                                         \
	if (!::testing::internal::check_true(static_cast<bool>(v), #v, __FILE__, __LINE__)) return

#ifndef GTEST_STUB_NO_MAIN
int main() { return ::testing::internal::run_all(); }
#endif
