// C++ host-mirror test: the reference's doc-tests (/root/reference/src/lib.rs:19-40) and the
// conversion / query properties of src/tests.rs on seeded inputs, through sab200_suffix_array.hpp.
// Needs a GPU at run time; compiled (only) by the CPU test-suite.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <sstream>
#include <vector>

#include "sab200_suffix_array.hpp"

static int fails = 0;
#define EXPECT(c)                                                         \
    do {                                                                  \
        if (!(c)) {                                                       \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c);      \
            ++fails;                                                      \
        }                                                                 \
    } while (0)

static const std::uint8_t* U(const char* s) { return reinterpret_cast<const std::uint8_t*>(s); }

int main() {
    using sab200::SuffixArray;
    {  // src/lib.rs:19-40
        const char* s = "splendid splendor";
        SuffixArray sa(U(s), std::strlen(s));
        EXPECT(sa.contains(U("splend"), 6));
        auto hits = sa.search_all(U("splend"), 6);
        EXPECT(hits.len == 2 && hits.data[0] == 0 && hits.data[1] == 9);
        auto r = sa.search_lcp(U("splash"), 6);
        EXPECT(std::string(s + r.start, s + r.end) == "spl");
        sa.enable_buckets();
        hits = sa.search_all(U("splend"), 6);
        EXPECT(hits.len == 2 && hits.data[0] == 0 && hits.data[1] == 9);
        EXPECT(sa.sa()[0] == 17);  // src/saca.rs:13
    }
    std::mt19937_64 rng(42);
    for (int trial = 0; trial < 40; ++trial) {
        const std::size_t n = rng() % 4096;
        const int sigma = (trial % 4 == 0) ? 2 : 256;
        std::vector<std::uint8_t> s(n);
        for (auto& c : s) c = (std::uint8_t)(rng() % sigma);
        SuffixArray sa(s.data(), n);
        // conversion_correctness (src/tests.rs:14-17)
        auto again = SuffixArray::from_parts(s.data(), n, sa.sa());
        EXPECT(again.has_value());
        // strict suffix order, literally src/sa.rs:76-82
        bool ordered = sa.sa().size() == n + 1;
        for (std::size_t i = 1; ordered && i <= n; ++i) {
            const std::size_t a = sa.sa()[i - 1], b = sa.sa()[i];
            ordered = std::lexicographical_compare(s.begin() + a, s.end(), s.begin() + b, s.end());
        }
        EXPECT(ordered);
        if (n >= 2) {
            std::vector<std::uint32_t> bad = sa.sa();
            std::swap(bad[1], bad[2]);
            EXPECT(!SuffixArray::from_parts(s.data(), n, bad).has_value());
        }
        {  // pack_correctness (src/tests.rs:63-76)
            const auto bytes = sa.dump_bytes();
            std::ostringstream os;
            sa.dump(os);
            const std::string o = os.str();
            EXPECT(o.size() == bytes.size() && std::equal(bytes.begin(), bytes.end(), (const std::uint8_t*)o.data()));
            auto back = SuffixArray::load_bytes(s.data(), n, bytes.data(), bytes.size());
            EXPECT(back.sa() == sa.sa());
            if (n >= 2 && s[0] != s[n - 1]) {
                std::vector<std::uint8_t> other(s.rbegin(), s.rend());
                bool threw = false;
                try {
                    SuffixArray::load_bytes(other.data(), n, bytes.data(), bytes.size());
                } catch (const std::runtime_error&) { threw = true; }
                EXPECT(threw || other == s);
            }
        }
        // search_all_correctness / contains_correctness vs the naive scan (src/tests.rs:104-121)
        const std::size_t m = n ? rng() % std::min<std::size_t>(n, 12) : 0;
        const std::size_t at = n > m ? rng() % (n - m) : 0;
        std::vector<std::uint8_t> pat(s.begin() + at, s.begin() + at + m);
        if (trial % 3 == 0 && m) pat[m - 1] ^= 1;
        std::vector<std::uint32_t> naive;
        for (std::size_t i = 0; i + m <= n; ++i)
            if (std::equal(pat.begin(), pat.end(), s.begin() + i)) naive.push_back((std::uint32_t)i);
        for (int with_bkt = 0; with_bkt < 2; ++with_bkt) {
            if (with_bkt) sa.enable_buckets();
            auto hits = sa.search_all(pat.data(), m);
            std::vector<std::uint32_t> got(hits.begin(), hits.end());
            std::sort(got.begin(), got.end());
            EXPECT(got == naive);
            EXPECT(sa.contains(pat.data(), m) == !naive.empty());
        }
    }
    sab200_shutdown();
    std::printf(fails ? "host mirror: %d failure(s)\n" : "host mirror: all checks passed\n", fails);
    return fails ? 1 : 0;
}
