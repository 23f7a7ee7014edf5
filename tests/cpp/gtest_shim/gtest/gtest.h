// gtest/gtest.h -- a small single-header stand-in for the subset of GoogleTest the cuZK test sources use
// (TEST / TEST_F, EXPECT_* / ASSERT_* with streamed messages, EXPECT_THROW, GTEST_SKIP, SCOPED_TRACE,
// testing::internal::CaptureStdout, DISABLED_ prefixes, --gtest_filter).  TEST INFRASTRUCTURE: GoogleTest itself is
// fetched from the network by the reference's CMake (CMakeLists.txt:37-42) and is not available offline; this header
// lets the reference's test sources compile unmodified against the cuzk_b200 host layer.  Define GTEST_SHIM_MAIN in
// exactly one translation unit (or link tests/cpp/gtest_shim/gtest_main.cpp) to get main().
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <iomanip>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <tuple>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <unistd.h>
#include <utility>
#include <vector>

namespace testing {

class Message {
public:
  template <class T> Message &operator<<(const T &v) { ss_ << v; return *this; }
  Message &operator<<(std::ostream &(*manip)(std::ostream &)) { ss_ << manip; return *this; }
  std::string str() const { return ss_.str(); }
private:
  std::ostringstream ss_;
};

namespace internal {

template <class T, class = void> struct is_streamable : std::false_type {};
template <class T> struct is_streamable<T, std::void_t<decltype(std::declval<std::ostream &>() << std::declval<const T &>())>> : std::true_type {};

template <class T> std::string Print(const T &v) {
  if constexpr (std::is_same_v<T, bool>) return v ? "true" : "false";
  else if constexpr (std::is_same_v<T, std::nullptr_t>) return "nullptr";
  else if constexpr (is_streamable<T>::value) { std::ostringstream ss; ss << v; return ss.str(); }
  else {
    std::ostringstream ss;
    const unsigned char *p = reinterpret_cast<const unsigned char *>(&v);
    ss << sizeof(T) << "-byte object <";
    for (size_t i = 0; i < sizeof(T) && i < 32; ++i) { char b[4]; std::snprintf(b, sizeof b, "%02x", p[i]); ss << (i ? " " : "") << b; }
    ss << ">";
    return ss.str();
  }
}

struct TestInfo {
  std::string suite, name;
  std::function<void()> run;
};

struct State {
  std::vector<TestInfo> tests;
  std::vector<std::string> traces;
  bool current_failed = false, current_skipped = false;
  int captured_fd = -1;
  std::string capture_path;
  static State &get() { static State s; return s; }
};

inline bool Register(const char *suite, const char *name, std::function<void()> run) {
  State::get().tests.push_back({suite, name, std::move(run)});
  return true;
}

enum class Kind { NonFatal, Fatal, Skip };

class AssertHelper {
public:
  AssertHelper(Kind kind, const char *file, int line, std::string summary) : kind_(kind), file_(file), line_(line), summary_(std::move(summary)) {}
  void operator=(const Message &m) const {
    State &st = State::get();
    if (kind_ == Kind::Skip) {
      st.current_skipped = true;
      std::cout << file_ << ":" << line_ << ": Skipped" << std::endl;
      if (!m.str().empty()) std::cout << m.str() << std::endl;
      return;
    }
    st.current_failed = true;
    std::cout << file_ << ":" << line_ << ": Failure" << std::endl << summary_ << std::endl;
    if (!m.str().empty()) std::cout << m.str() << std::endl;
    for (auto it = st.traces.rbegin(); it != st.traces.rend(); ++it) std::cout << "Google Test trace: " << *it << std::endl;
  }
private:
  Kind kind_;
  const char *file_;
  int line_;
  std::string summary_;
};

class ScopedTrace {
public:
  template <class T> ScopedTrace(const char *file, int line, const T &msg) {
    std::ostringstream ss;
    ss << file << ":" << line << ": " << msg;
    State::get().traces.push_back(ss.str());
  }
  ~ScopedTrace() { State::get().traces.pop_back(); }
};

struct Result {
  bool ok;
  std::string text;
  explicit operator bool() const { return ok; }
};

template <class A, class B, class Op>
Result Compare(const char *ea, const char *eb, const char *opname, const A &a, const B &b, Op op) {
  if (op(a, b)) return {true, ""};
  std::ostringstream ss;
  if (std::strcmp(opname, "==") == 0)
    ss << "Expected equality of these values:\n  " << ea << "\n    Which is: " << Print(a) << "\n  " << eb << "\n    Which is: " << Print(b);
  else
    ss << "Expected: (" << ea << ") " << opname << " (" << eb << "), actual: " << Print(a) << " vs " << Print(b);
  return {false, ss.str()};
}

inline Result Boolean(const char *expr, bool value, bool expected) {
  if (value == expected) return {true, ""};
  return {false, std::string("Value of: ") + expr + "\n  Actual: " + (value ? "true" : "false") + "\nExpected: " + (expected ? "true" : "false")};
}

// stdout capture (testing::internal::CaptureStdout / GetCapturedStdout)
inline void CaptureStdout() {
  State &st = State::get();
  std::cout.flush();
  std::fflush(stdout);
  char path[] = "/tmp/gtest_shim_capture_XXXXXX";
  int fd = mkstemp(path);
  st.capture_path = path;
  st.captured_fd = dup(1);
  dup2(fd, 1);
  close(fd);
}
inline std::string GetCapturedStdout() {
  State &st = State::get();
  std::cout.flush();
  std::fflush(stdout);
  dup2(st.captured_fd, 1);
  close(st.captured_fd);
  st.captured_fd = -1;
  std::string out;
  if (FILE *f = std::fopen(st.capture_path.c_str(), "rb")) {
    char buf[4096];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    std::fclose(f);
  }
  unlink(st.capture_path.c_str());
  return out;
}

// --gtest_filter: ':'-separated glob patterns, optional '-' negative section
inline bool Glob(const char *p, const char *s) {
  if (*p == '\0') return *s == '\0';
  if (*p == '*') return Glob(p + 1, s) || (*s && Glob(p, s + 1));
  return *s && (*p == '?' || *p == *s) && Glob(p + 1, s + 1);
}
inline bool MatchesAny(const std::string &patterns, const std::string &name) {
  size_t at = 0;
  while (at <= patterns.size()) {
    size_t end = patterns.find(':', at);
    if (end == std::string::npos) end = patterns.size();
    if (end > at && Glob(patterns.substr(at, end - at).c_str(), name.c_str())) return true;
    at = end + 1;
  }
  return false;
}

}  // namespace internal

class Test {
public:
  virtual ~Test() = default;
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
  static bool HasFailure() { return internal::State::get().current_failed; }
  static bool IsSkipped() { return internal::State::get().current_skipped; }
};

namespace internal {
template <class T> void RunOne() {
  T t;
  Test &base = t;  // fixtures override SetUp/TearDown as protected members; dispatch through the public base
  base.SetUp();
  if (!State::get().current_skipped && !State::get().current_failed) base.TestBody();
  base.TearDown();
}
}  // namespace internal

inline void InitGoogleTest(int *, char **) {}

inline int RunAllTests(int argc, char **argv) {
  std::string filter = "*";
  bool also_disabled = false, list_only = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a.rfind("--gtest_filter=", 0) == 0) filter = a.substr(15);
    else if (a == "--gtest_also_run_disabled_tests") also_disabled = true;
    else if (a == "--gtest_list_tests") list_only = true;
  }
  std::string pos = filter, neg;
  const size_t dash = filter.find('-');
  if (dash != std::string::npos) { pos = filter.substr(0, dash); neg = filter.substr(dash + 1); }
  if (pos.empty()) pos = "*";
  internal::State &st = internal::State::get();
  int ran = 0, failed = 0, skipped = 0, disabled = 0;
  std::vector<std::string> failures;
  for (auto &t : st.tests) {
    const std::string full = t.suite + "." + t.name;
    if (!internal::MatchesAny(pos, full) || (!neg.empty() && internal::MatchesAny(neg, full))) continue;
    if (!also_disabled && (t.name.rfind("DISABLED_", 0) == 0 || t.suite.rfind("DISABLED_", 0) == 0)) { ++disabled; continue; }
    if (list_only) { std::cout << full << std::endl; continue; }
    std::cout << "[ RUN      ] " << full << std::endl;
    st.current_failed = st.current_skipped = false;
    try {
      t.run();
    } catch (const std::exception &e) {
      st.current_failed = true;
      std::cout << "unknown file: Failure\nC++ exception with description \"" << e.what() << "\" thrown in the test body." << std::endl;
    } catch (...) {
      st.current_failed = true;
      std::cout << "unknown file: Failure\nUnknown C++ exception thrown in the test body." << std::endl;
    }
    ++ran;
    if (st.current_failed) { ++failed; failures.push_back(full); std::cout << "[  FAILED  ] " << full << std::endl; }
    else if (st.current_skipped) { ++skipped; std::cout << "[  SKIPPED ] " << full << std::endl; }
    else std::cout << "[       OK ] " << full << std::endl;
  }
  if (list_only) return 0;
  std::cout << "[==========] " << ran << " tests ran." << std::endl;
  std::cout << "[  PASSED  ] " << (ran - failed - skipped) << " tests." << std::endl;
  if (skipped) std::cout << "[  SKIPPED ] " << skipped << " tests." << std::endl;
  if (disabled) std::cout << "  YOU HAVE " << disabled << " DISABLED TESTS" << std::endl;
  if (failed) {
    std::cout << "[  FAILED  ] " << failed << " tests, listed below:" << std::endl;
    for (auto &f : failures) std::cout << "[  FAILED  ] " << f << std::endl;
  }
  return failed ? 1 : 0;
}

}  // namespace testing

#define GTEST_SHIM_CAT_(a, b) a##b
#define GTEST_SHIM_CAT(a, b) GTEST_SHIM_CAT_(a, b)
#define GTEST_SHIM_CLASS(suite, name) suite##_##name##_Test

#define GTEST_SHIM_TEST_(suite, name, parent)                                                                              \
  class GTEST_SHIM_CLASS(suite, name) : public parent {                                                                    \
  public:                                                                                                                  \
    void TestBody() override;                                                                                              \
    static bool registered_;                                                                                               \
  };                                                                                                                       \
  bool GTEST_SHIM_CLASS(suite, name)::registered_ =                                                                        \
      ::testing::internal::Register(#suite, #name, [] { ::testing::internal::RunOne<GTEST_SHIM_CLASS(suite, name)>(); }); \
  void GTEST_SHIM_CLASS(suite, name)::TestBody()

#define TEST(suite, name) GTEST_SHIM_TEST_(suite, name, ::testing::Test)
#define TEST_F(fixture, name) GTEST_SHIM_TEST_(fixture, name, fixture)

#define GTEST_SHIM_AMBIGUOUS_ELSE_BLOCKER switch (0) case 0: default:
#define GTEST_SHIM_NONFATAL(result_expr)                                     \
  GTEST_SHIM_AMBIGUOUS_ELSE_BLOCKER                                          \
  if (const ::testing::internal::Result gtest_shim_r = (result_expr)) ;      \
  else ::testing::internal::AssertHelper(::testing::internal::Kind::NonFatal, __FILE__, __LINE__, gtest_shim_r.text) = ::testing::Message()
#define GTEST_SHIM_FATAL(result_expr)                                        \
  GTEST_SHIM_AMBIGUOUS_ELSE_BLOCKER                                          \
  if (const ::testing::internal::Result gtest_shim_r = (result_expr)) ;      \
  else return ::testing::internal::AssertHelper(::testing::internal::Kind::Fatal, __FILE__, __LINE__, gtest_shim_r.text) = ::testing::Message()

#define GTEST_SHIM_CMP(a, b, opname, op) \
  ::testing::internal::Compare(#a, #b, opname, (a), (b), [](const auto &x, const auto &y) { return x op y; })

#define EXPECT_TRUE(c) GTEST_SHIM_NONFATAL(::testing::internal::Boolean(#c, static_cast<bool>(c), true))
#define EXPECT_FALSE(c) GTEST_SHIM_NONFATAL(::testing::internal::Boolean(#c, static_cast<bool>(c), false))
#define ASSERT_TRUE(c) GTEST_SHIM_FATAL(::testing::internal::Boolean(#c, static_cast<bool>(c), true))
#define ASSERT_FALSE(c) GTEST_SHIM_FATAL(::testing::internal::Boolean(#c, static_cast<bool>(c), false))
#define EXPECT_EQ(a, b) GTEST_SHIM_NONFATAL(GTEST_SHIM_CMP(a, b, "==", ==))
#define EXPECT_NE(a, b) GTEST_SHIM_NONFATAL(GTEST_SHIM_CMP(a, b, "!=", !=))
#define EXPECT_LT(a, b) GTEST_SHIM_NONFATAL(GTEST_SHIM_CMP(a, b, "<", <))
#define EXPECT_LE(a, b) GTEST_SHIM_NONFATAL(GTEST_SHIM_CMP(a, b, "<=", <=))
#define EXPECT_GT(a, b) GTEST_SHIM_NONFATAL(GTEST_SHIM_CMP(a, b, ">", >))
#define EXPECT_GE(a, b) GTEST_SHIM_NONFATAL(GTEST_SHIM_CMP(a, b, ">=", >=))
#define ASSERT_EQ(a, b) GTEST_SHIM_FATAL(GTEST_SHIM_CMP(a, b, "==", ==))
#define ASSERT_NE(a, b) GTEST_SHIM_FATAL(GTEST_SHIM_CMP(a, b, "!=", !=))
#define ASSERT_LT(a, b) GTEST_SHIM_FATAL(GTEST_SHIM_CMP(a, b, "<", <))
#define ASSERT_LE(a, b) GTEST_SHIM_FATAL(GTEST_SHIM_CMP(a, b, "<=", <=))
#define ASSERT_GT(a, b) GTEST_SHIM_FATAL(GTEST_SHIM_CMP(a, b, ">", >))
#define ASSERT_GE(a, b) GTEST_SHIM_FATAL(GTEST_SHIM_CMP(a, b, ">=", >=))

#define GTEST_SHIM_THROW_RESULT(stmt, exc)                                                                      \
  [&]() -> ::testing::internal::Result {                                                                        \
    try { stmt; } catch (const exc &) { return {true, ""}; } catch (...) {                                       \
      return {false, "Expected: " #stmt " throws an exception of type " #exc ".\n  Actual: it throws a different type."}; \
    }                                                                                                           \
    return {false, "Expected: " #stmt " throws an exception of type " #exc ".\n  Actual: it throws nothing."};    \
  }()
#define EXPECT_THROW(stmt, exc) GTEST_SHIM_NONFATAL(GTEST_SHIM_THROW_RESULT(stmt, exc))
#define ASSERT_THROW(stmt, exc) GTEST_SHIM_FATAL(GTEST_SHIM_THROW_RESULT(stmt, exc))
#define EXPECT_NO_THROW(stmt)                                                                                   \
  GTEST_SHIM_NONFATAL(([&]() -> ::testing::internal::Result {                                                   \
    try { stmt; } catch (...) { return {false, "Expected: " #stmt " doesn't throw.\n  Actual: it throws."}; }    \
    return {true, ""};                                                                                          \
  }()))

#define GTEST_SKIP() return ::testing::internal::AssertHelper(::testing::internal::Kind::Skip, __FILE__, __LINE__, "") = ::testing::Message()
#define SUCCEED() GTEST_SHIM_NONFATAL((::testing::internal::Result{true, ""}))
#define ADD_FAILURE() GTEST_SHIM_NONFATAL((::testing::internal::Result{false, "Failed"}))
#define FAIL() GTEST_SHIM_FATAL((::testing::internal::Result{false, "Failed"}))
#define SCOPED_TRACE(msg) ::testing::internal::ScopedTrace GTEST_SHIM_CAT(gtest_shim_trace_, __LINE__)(__FILE__, __LINE__, (msg))

#define RUN_ALL_TESTS() ::testing::RunAllTests(gtest_shim_argc, gtest_shim_argv)

#ifdef GTEST_SHIM_MAIN
int main(int argc, char **argv) { return ::testing::RunAllTests(argc, argv); }
#endif
