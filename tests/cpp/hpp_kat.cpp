// Exercises the header-only C++ mirror (include/triple_accel.hpp) against known answers of the reference's tests
// (tests/basic_tests.rs); built and run by tests/test_gpu_parity.py::test_cpp_header_mirror on the GPU box.
#include <cstdio>
#include <cstring>

#include "triple_accel.hpp"

using namespace triple_accel;
static bytes B(const char *s) { return bytes((const uint8_t *)s, strlen(s)); }
#define CHECK(x)                                             \
    do {                                                     \
        if (!(x)) {                                          \
            printf("FAILED: %s (line %d)\n", #x, __LINE__);  \
            return 1;                                        \
        }                                                    \
    } while (0)

int main() {
    Engine eng(0);
    CHECK(eng.hamming(B("abc"), B("abd")) == 1);                               // basic_tests.rs:9
    CHECK(eng.levenshtein(B("abcde"), B(" ab cde")) == 2);                      // :201
    CHECK(eng.levenshtein_exp(B("abcde"), B("")) == 5);                         // :234
    CHECK(eng.rdamerau(B("abcde"), B(" ab dce")) == 3);                         // :257
    CHECK(eng.rdamerau_exp(B("abcde"), B("bacdee")) == 2);                      // :295
    CHECK(*eng.levenshtein_simd_k_with_opts(B("abc"), B("ac"), 5, EditCosts(1, 1, 2)) == 3);  // :487
    CHECK(!eng.levenshtein_simd_k_with_opts(B("abcde"), B("hello"), 1, RDAMERAU_COSTS));      // :541
    auto m = eng.levenshtein_search(B("tst"), B("testing 123 tasting!"));       // :711
    CHECK(m.size() == 2 && m[0] == (Match{0, 4, 1}) && m[1] == (Match{12, 16, 1}));
    auto m2 = eng.levenshtein_search_simd_with_opts(B("test"), B(" etsting 123 tasting"), 2, SearchType::All,
                                                    RDAMERAU_COSTS, true);     // :747
    CHECK(m2.size() == 3 && m2[2] == (Match{1, 5, 2}));
    bool threw = false;
    try {
        eng.hamming(B("abc"), B("ab"));
    } catch (const std::logic_error &) {
        threw = true;
    }
    CHECK(threw);  // the crate panics on a length mismatch (src/hamming.rs:38)
    printf("hpp ok\n");
    return 0;
}
