// triple_accel.hpp -- header-only C++17 mirror of the triple_accel crate's hot-path API over the C ABI
// (include/triple_accel_b200.h).  Same names and argument meaning as the reference's public functions
// (src/lib.rs:126-127); contract violations the crate turns into panics throw std::logic_error here, device
// failures throw std::runtime_error (there is no CPU fallback).
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "triple_accel_b200.h"

namespace triple_accel {

using bytes = std::basic_string_view<uint8_t>;

struct Match {  // src/lib.rs:134-142
    size_t start, end;
    uint32_t k;
    bool operator==(const Match &o) const { return start == o.start && end == o.end && k == o.k; }
};
enum class SearchType { All = TA_SEARCH_ALL, Best = TA_SEARCH_BEST };  // src/lib.rs:170-174

class EditCosts {  // src/levenshtein.rs:20-72
   public:
    EditCosts(uint8_t mismatch, uint8_t gap, uint8_t start_gap, std::optional<uint8_t> transpose = std::nullopt)
        : c_{mismatch, gap, start_gap, transpose.value_or(0)} {
        if ((transpose && *transpose == 0) || !ta_costs_valid(c_)) throw std::logic_error("EditCosts::new assertion");
    }
    ta_costs raw() const { return c_; }

   private:
    ta_costs c_;
};
inline const EditCosts LEVENSHTEIN_COSTS{1, 1, 0};      // src/levenshtein.rs:76-81
inline const EditCosts RDAMERAU_COSTS{1, 1, 0, 1};      // src/levenshtein.rs:84-89

class Engine {
   public:
    explicit Engine(int device = 0) {
        int rc = ta_init(device, &ctx_);
        if (rc != TA_OK) throw std::runtime_error(std::string("ta_init: ") + ta_strerror(rc));
    }
    ~Engine() { ta_shutdown(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    ta_ctx *raw() { return ctx_; }

    uint32_t hamming(bytes a, bytes b) {  // src/hamming.rs:390-392
        uint32_t out = 0;
        check(ta_hamming(ctx_, a.data(), a.size(), b.data(), b.size(), &out));
        return out;
    }
    std::optional<uint32_t> levenshtein_simd_k_with_opts(bytes a, bytes b, uint32_t k, const EditCosts &costs) {
        uint32_t out = 0;  // src/levenshtein.rs:714-720 with trace_on = false
        check(ta_levenshtein_simd_k_with_opts(ctx_, a.data(), a.size(), b.data(), b.size(), k, costs.raw(), &out));
        return out == TA_NONE ? std::nullopt : std::optional<uint32_t>(out);
    }
    std::optional<uint32_t> levenshtein_simd_k(bytes a, bytes b, uint32_t k) {
        return levenshtein_simd_k_with_opts(a, b, k, LEVENSHTEIN_COSTS);
    }
    uint32_t levenshtein(bytes a, bytes b) { return *levenshtein_simd_k(a, b, 0xFFFFFFFFu); }
    uint32_t rdamerau(bytes a, bytes b) { return *levenshtein_simd_k_with_opts(a, b, 0xFFFFFFFFu, RDAMERAU_COSTS); }
    uint32_t levenshtein_exp_with_opts(bytes a, bytes b, const EditCosts &costs) {  // :1480-1494
        uint32_t out = 0;
        check(ta_levenshtein_exp_with_opts(ctx_, a.data(), a.size(), b.data(), b.size(), costs.raw(), &out));
        return out;
    }
    uint32_t levenshtein_exp(bytes a, bytes b) { return levenshtein_exp_with_opts(a, b, LEVENSHTEIN_COSTS); }
    uint32_t rdamerau_exp(bytes a, bytes b) { return levenshtein_exp_with_opts(a, b, RDAMERAU_COSTS); }

    std::vector<Match> levenshtein_search_simd_with_opts(bytes needle, bytes haystack, uint32_t k, SearchType st,
                                                         const EditCosts &costs, bool anchored) {  // :1911-1918
        const uint64_t off[2] = {0, haystack.size()};
        ta_match *m = nullptr;
        uint64_t *mo = nullptr;
        check(ta_levenshtein_search_batch(ctx_, needle.data(), needle.size(), haystack.data(), off, 1, k, (int)st,
                                          costs.raw(), anchored, &m, &mo));
        std::vector<Match> v;
        for (uint64_t i = 0; i < mo[1]; i++) v.push_back(Match{(size_t)m[i].start, (size_t)m[i].end, m[i].k});
        ta_free(m);
        ta_free(mo);
        return v;
    }
    std::vector<Match> levenshtein_search(bytes needle, bytes haystack) {  // src/levenshtein.rs:2508-2513
        return levenshtein_search_simd_with_opts(needle, haystack, ta_search_default_k(needle.size()),
                                                 SearchType::Best, LEVENSHTEIN_COSTS, false);
    }

    // The crate's scalar / word-wise / movemask variants are CPU micro-variants of the same contracts: names for the
    // same entry points, so that every public name of the crate resolves.
    uint32_t hamming_naive(bytes a, bytes b) { return hamming(a, b); }          // src/hamming.rs:36-47
    uint32_t hamming_words_64(bytes a, bytes b) { return hamming(a, b); }       // src/hamming.rs:176
    uint32_t hamming_words_128(bytes a, bytes b) { return hamming(a, b); }      // src/hamming.rs:249
    uint32_t hamming_simd_parallel(bytes a, bytes b) { return hamming(a, b); }  // src/hamming.rs:317-330
    uint32_t hamming_simd_movemask(bytes a, bytes b) { return hamming(a, b); }  // src/hamming.rs:354
    uint32_t levenshtein_naive(bytes a, bytes b) { return levenshtein(a, b); }  // src/levenshtein.rs:105
    std::optional<uint32_t> levenshtein_naive_k(bytes a, bytes b, uint32_t k) { return levenshtein_simd_k(a, b, k); }
    std::optional<uint32_t> levenshtein_naive_k_with_opts(bytes a, bytes b, uint32_t k, const EditCosts &costs) {
        return levenshtein_simd_k_with_opts(a, b, k, costs);  // src/levenshtein.rs:376-607: the contract itself
    }
    std::vector<Match> levenshtein_search_naive_with_opts(bytes needle, bytes haystack, uint32_t k, SearchType st,
                                                          const EditCosts &costs, bool anchored) {  // :1589-1838
        return levenshtein_search_simd_with_opts(needle, haystack, k, st, costs, anchored);
    }
    std::vector<Match> levenshtein_search_naive(bytes needle, bytes haystack) { return levenshtein_search(needle, haystack); }
    std::vector<Match> levenshtein_search_simd(bytes needle, bytes haystack) { return levenshtein_search(needle, haystack); }

   private:
    void check(int rc) {
        if (rc == TA_OK) return;
        if (rc == TA_ERR_LEN_MISMATCH || rc == TA_ERR_BAD_COSTS) throw std::logic_error(ta_strerror(rc));
        throw std::runtime_error(std::string(ta_strerror(rc)) + ": " + ta_last_error(ctx_));
    }
    ta_ctx *ctx_ = nullptr;
};

}  // namespace triple_accel
