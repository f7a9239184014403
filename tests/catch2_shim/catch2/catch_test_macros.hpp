// Minimal stand-in for the parts of Catch2 v3 that cuCollections' own test sources use, so that
// /root/reference/tests/{static_map,static_set,static_multiset,utility}/*.cu compile UNCHANGED against
// this repository's include/ tree (and against the reference's own, for a line-by-line diff of the
// two outputs). Catch2 itself is fetched over the network by the reference's CMake
// (tests/CMakeLists.txt:23-27) and is not available offline. Test infrastructure only.
//
// Supported: TEST_CASE, SECTION (re-runs the test body once per leaf section, nested sections
// included), REQUIRE / REQUIRE_FALSE / CHECK / CHECK_FALSE, STATIC_REQUIRE, INFO, SKIP, GENERATE (the
// body re-runs over the cartesian product), and through catch_template_test_macros.hpp
// TEMPLATE_TEST_CASE and TEMPLATE_TEST_CASE_SIG. Every translation unit is its own program: this
// header defines main().
#pragma once

#include <cstdio>
#include <cstdlib>
#include <exception>
#include <functional>
#include <initializer_list>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace catch2_shim {

struct test_case {
  std::string name;
  void (*body)();
};

inline std::vector<test_case>& registry()
{
  static std::vector<test_case> r;
  return r;
}

struct registrar {
  registrar(std::string name, void (*body)()) { registry().push_back({std::move(name), body}); }
};

struct require_failed : std::exception {};
struct skipped : std::exception {
  std::string why;
  explicit skipped(std::string w) : why{std::move(w)} {}
};

/// State of one run of a test body: which child to enter at every section depth, which value every
/// GENERATE yields, and what the run discovered (siblings per depth, values per generator).
struct run_state {
  std::vector<int> section_path;      // target child index per depth
  // Children discovered at each depth inside the entered parent, in first-visit order. A section is
  // identified by its source line and name, as in Catch2 (SectionTracker: name + location): a SECTION
  // macro that control flow reaches several times in one run (a helper called in a loop or thrice in a
  // row, e.g. tests/static_multiset/retrieve_test.cu:152-154) is ONE section, entered on its first visit
  // of the run in which it is the target and skipped on every other visit.
  std::vector<std::vector<std::string>> section_seen;
  std::vector<bool> section_entered;  // the target child of this depth has already run in this run
  int depth = 0;
  std::vector<int> generator_choice;  // chosen value index per GENERATE call (in call order)
  std::vector<int> generator_size;
  int generator_calls = 0;
  long assertions = 0, failures = 0;
};

inline run_state*& current()
{
  static run_state* s = nullptr;
  return s;
}

class section {
 public:
  section(char const* name, int line)
  {
    auto& s = *current();
    if (static_cast<int>(s.section_seen.size()) <= s.depth) {
      s.section_seen.resize(s.depth + 1);
      s.section_entered.resize(s.depth + 1, false);
    }
    std::string const id = std::to_string(line) + ":" + name;
    auto& seen           = s.section_seen[s.depth];
    int index            = -1;
    for (std::size_t i = 0; i < seen.size(); ++i) {
      if (seen[i] == id) { index = static_cast<int>(i); }
    }
    if (index < 0) {
      index = static_cast<int>(seen.size());
      seen.push_back(id);
    }
    if (static_cast<int>(s.section_path.size()) <= s.depth) {
      // first visit of this depth in this run: take the first child
      if (index == 0 && !s.section_entered[s.depth]) {
        s.section_path.push_back(0);
        entered_ = true;
      }
    } else {
      entered_ = s.section_path[s.depth] == index && !s.section_entered[s.depth];
    }
    if (entered_) {
      s.section_entered[s.depth] = true;
      ++s.depth;
      if (static_cast<int>(s.section_seen.size()) <= s.depth) {
        s.section_seen.resize(s.depth + 1);
        s.section_entered.resize(s.depth + 1, false);
      }
      s.section_seen[s.depth].clear();
      s.section_entered[s.depth] = false;
    }
  }
  ~section()
  {
    if (entered_) { --current()->depth; }
  }
  explicit operator bool() const { return entered_; }

 private:
  bool entered_ = false;
};

template <typename T>
T generate(std::initializer_list<T> values)
{
  auto& s       = *current();
  int const me  = s.generator_calls++;
  if (static_cast<int>(s.generator_choice.size()) <= me) {
    s.generator_choice.push_back(0);
    s.generator_size.push_back(static_cast<int>(values.size()));
  }
  return *(values.begin() + s.generator_choice[me]);
}

inline void report(bool ok, bool fatal, char const* file, int line, char const* text)
{
  auto& s = *current();
  ++s.assertions;
  if (ok) { return; }
  ++s.failures;
  std::printf("  FAILED %s:%d: %s\n", file, line, text);
  if (fatal) { throw require_failed{}; }
}

/// Runs one test case over all its section leaves and generator values; returns failures.
inline long run_test_case(test_case const& tc, long& assertions, bool& was_skipped)
{
  long failures = 0;
  run_state s;
  std::vector<int> next_path;
  std::vector<int> choices;
  while (true) {
    s.section_path     = next_path;
    s.section_seen.clear();
    s.section_entered.clear();
    s.depth            = 0;
    s.generator_choice = choices;
    s.generator_calls  = 0;
    s.assertions = s.failures = 0;
    current() = &s;
    try {
      tc.body();
    } catch (require_failed const&) {
    } catch (skipped const& e) {
      std::printf("  SKIPPED: %s\n", e.why.c_str());
      was_skipped = true;
    } catch (std::exception const& e) {
      ++s.failures;
      std::printf("  FAILED: unexpected exception: %s\n", e.what());
    }
    current() = nullptr;
    assertions += s.assertions;
    failures += s.failures;
    // next leaf of the section tree: advance the deepest index that still has a sibling
    next_path = s.section_path;
    bool more_sections = false;
    while (!next_path.empty()) {
      auto const d = next_path.size() - 1;
      int const siblings = d < s.section_seen.size() ? static_cast<int>(s.section_seen[d].size()) : 0;
      if (next_path[d] + 1 < siblings) {
        ++next_path[d];
        more_sections = true;
        break;
      }
      next_path.pop_back();
    }
    if (more_sections) { continue; }
    // all sections done for this generator tuple: odometer over the generators
    choices = s.generator_choice;
    bool more_values = false;
    for (int g = static_cast<int>(choices.size()) - 1; g >= 0; --g) {
      if (choices[g] + 1 < s.generator_size[g]) {
        ++choices[g];
        more_values = true;
        break;
      }
      choices[g] = 0;
    }
    if (!more_values) { break; }
    next_path.clear();
  }
  return failures;
}

inline int run_all()
{
  long total_assertions = 0, total_failures = 0, failed_cases = 0, skipped_cases = 0;
  for (auto const& tc : registry()) {
    long assertions = 0;
    bool was_skipped = false;
    long const failures = run_test_case(tc, assertions, was_skipped);
    total_assertions += assertions;
    total_failures += failures;
    failed_cases += failures != 0;
    skipped_cases += was_skipped;
    std::printf("%s %s (%ld assertions)\n", failures ? "FAIL" : (was_skipped ? "SKIP" : "PASS"), tc.name.c_str(),
                assertions);
    std::fflush(stdout);
  }
  std::printf("== %zu test cases, %ld failed, %ld skipped; %ld assertions, %ld failed\n", registry().size(),
              failed_cases, skipped_cases, total_assertions, total_failures);
  return total_failures == 0 ? 0 : 1;
}

}  // namespace catch2_shim

#define CATCH2_SHIM_CAT2(a, b) a##b
#define CATCH2_SHIM_CAT(a, b)  CATCH2_SHIM_CAT2(a, b)
#define CATCH2_SHIM_UNIQUE(prefix) CATCH2_SHIM_CAT(prefix, __COUNTER__)

#define CATCH2_SHIM_TEST_CASE(fn, ...)                                                                    \
  static void fn();                                                                                       \
  namespace {                                                                                             \
  ::catch2_shim::registrar CATCH2_SHIM_CAT(fn, _registrar){                                               \
    ::catch2_shim::first_string(__VA_ARGS__), &fn};                                                       \
  }                                                                                                       \
  static void fn()

namespace catch2_shim {
inline std::string first_string(char const* name, char const* = "") { return name; }
}  // namespace catch2_shim

#define TEST_CASE(...) CATCH2_SHIM_TEST_CASE(CATCH2_SHIM_UNIQUE(catch2_shim_test_), __VA_ARGS__)

#define SECTION(...) if (::catch2_shim::section CATCH2_SHIM_UNIQUE(catch2_shim_section_){#__VA_ARGS__, __LINE__})

#define REQUIRE(...)       ::catch2_shim::report(static_cast<bool>(__VA_ARGS__), true, __FILE__, __LINE__, "REQUIRE(" #__VA_ARGS__ ")")
#define REQUIRE_FALSE(...) ::catch2_shim::report(!static_cast<bool>(__VA_ARGS__), true, __FILE__, __LINE__, "REQUIRE_FALSE(" #__VA_ARGS__ ")")
#define CHECK(...)         ::catch2_shim::report(static_cast<bool>(__VA_ARGS__), false, __FILE__, __LINE__, "CHECK(" #__VA_ARGS__ ")")
#define CHECK_FALSE(...)   ::catch2_shim::report(!static_cast<bool>(__VA_ARGS__), false, __FILE__, __LINE__, "CHECK_FALSE(" #__VA_ARGS__ ")")
#define STATIC_REQUIRE(...) static_assert(__VA_ARGS__, #__VA_ARGS__)
#define INFO(...)          (void)0
#define CAPTURE(...)       (void)0
#define SUCCEED(...)       (void)0
#define SKIP(...)                                       \
  do {                                                  \
    std::ostringstream catch2_shim_why;                 \
    catch2_shim_why << "" __VA_ARGS__;                  \
    throw ::catch2_shim::skipped{catch2_shim_why.str()}; \
  } while (0)

#ifndef CATCH2_SHIM_NO_MAIN
int main() { return ::catch2_shim::run_all(); }
#endif
