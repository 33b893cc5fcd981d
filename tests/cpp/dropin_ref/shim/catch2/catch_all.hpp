// catch2/catch_all.hpp — a minimal stand-in for the Catch2 test framework (not installed in this image), written for
// this repository: TEST_CASE / REQUIRE / REQUIRE_THROWS_AS, the only macros the reference's tests use.  It exists so
// that /root/reference/tests/*.cpp compile UNMODIFIED against the product's headers (tests/cpp/dropin_ref/Makefile).
// One translation unit defines CATCH_CONFIG_MAIN and thereby gets main(): it runs every registered case, or those whose
// name or tags contain argv[1] ("~text" = those that do NOT contain it), prints one line per case and exits non-zero when
// any REQUIRE failed.
#pragma once

#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

namespace minicatch {

struct Case {
    const char* name;
    const char* tags;
    void (*fn)();
};
inline std::vector<Case>& cases() {
    static std::vector<Case> c;
    return c;
}
struct Registrar {
    Registrar(const char* name, const char* tags, void (*fn)()) { cases().push_back({name, tags, fn}); }
};
struct Failure : std::exception {
    std::string msg;
    explicit Failure(std::string m) : msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};
inline int& assertions() {
    static int n = 0;
    return n;
}

}  // namespace minicatch

#define MINICATCH_CAT2(a, b) a##b
#define MINICATCH_CAT(a, b) MINICATCH_CAT2(a, b)
#define MINICATCH_TEST(fn, ...)                                                   \
    static void fn();                                                             \
    static ::minicatch::Registrar MINICATCH_CAT(fn, _reg)(MINICATCH_FIRST(__VA_ARGS__, ""), MINICATCH_SECOND(__VA_ARGS__, "", ""), &fn); \
    static void fn()
#define MINICATCH_FIRST(a, ...) a
#define MINICATCH_SECOND(a, b, ...) b
#define TEST_CASE(...) MINICATCH_TEST(MINICATCH_CAT(minicatch_case_, __COUNTER__), __VA_ARGS__)

#define REQUIRE(expr)                                                                                                          \
    do {                                                                                                                       \
        ++::minicatch::assertions();                                                                                           \
        if (!(expr)) throw ::minicatch::Failure(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": REQUIRE(" #expr ") failed"); \
    } while (0)

#define REQUIRE_THROWS_AS(expr, type)                                                                                          \
    do {                                                                                                                       \
        ++::minicatch::assertions();                                                                                           \
        bool minicatch_threw = false;                                                                                          \
        try {                                                                                                                  \
            (void)(expr);                                                                                                      \
        } catch (const type&) {                                                                                                \
            minicatch_threw = true;                                                                                            \
        } catch (...) {                                                                                                        \
        }                                                                                                                      \
        if (!minicatch_threw)                                                                                                  \
            throw ::minicatch::Failure(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": REQUIRE_THROWS_AS(" #expr ", " #type ") failed"); \
    } while (0)

#ifdef CATCH_CONFIG_MAIN
int main(int argc, char** argv) {
    int failed = 0, ran = 0;
    for (const auto& c : ::minicatch::cases()) {
        if (argc > 1) {
            const bool negate = argv[1][0] == '~';
            const char* want = argv[1] + (negate ? 1 : 0);
            const bool has = std::strstr(c.name, want) || std::strstr(c.tags, want);
            if (has == negate) continue;
        }
        ++ran;
        try {
            c.fn();
            std::printf("ok      %s %s\n", c.name, c.tags);
        } catch (const std::exception& e) {
            ++failed;
            std::printf("FAILED  %s %s\n        %s\n", c.name, c.tags, e.what());
        }
    }
    std::printf("%d test cases, %d failed, %d assertions\n", ran, failed, ::minicatch::assertions());
    return failed ? 1 : (ran ? 0 : 2);
}
#endif
