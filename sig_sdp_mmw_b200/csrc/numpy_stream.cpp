// numpy's legacy normal stream, continued natively: np.random.randn on the global RandomState is MT19937 ->
// 53-bit doubles -> Marsaglia's polar method (numpy/random/src/mt19937/mt19937.c, legacy-distributions.c:
// legacy_gauss).  The reference draws one randn(K, D) per MMW iteration (mmw.py:226) and an unchanged driver must see
// the same numbers, but numpy produces them on one core at ~14 ns each: 45 ms per iteration at 100k nodes x 32 columns
// against 0.45 ms for the iteration itself.  The stream has structure that one core does not exploit: every candidate
// pair (x1, x2) occupies a FIXED four words of the Mersenne-Twister output, whether it is accepted or not, so the
// expensive part (log, sqrt, divide and the acceptance test) is independent per pair.  Here the twister runs
// sequentially on its own thread (~1.2 ns per word), one chunk ahead of the host cores that evaluate the pairs and
// compact the accepted ones in order.  Bit for bit the numbers numpy would have returned, and the state handed back is the state
// numpy would have been left in (position inside the twister block, cached second normal).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/sigsdp_mmw.h"
#include "plan_host.h"

namespace sigsdp {
namespace {

constexpr int MT_N = 624, MT_M = 397;

// (AVX-512 / AVX2 clones picked at load time: the recurrence reads only values that are 1, 397 or -227 places away, so
// it vectorises; 1.2 -> 0.6 ns per word)
__attribute__((target_clones("avx512f", "avx2", "default"))) void mt_regen(uint32_t* mt) {   // mt19937_gen
    constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
    int kk = 0;
    uint32_t y;
    for (; kk < MT_N - MT_M; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    for (; kk < MT_N - 1; ++kk) {
        y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    y = (mt[MT_N - 1] & UPPER) | (mt[0] & LOWER);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
}
inline __attribute__((always_inline)) uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
__attribute__((target_clones("avx512f", "avx2", "default"))) void temper_block(const uint32_t* k, uint32_t* o, size_t n) {
    for (size_t j = 0; j < n; ++j) o[j] = temper(k[j]);
}
struct Twister {
    uint32_t key[MT_N];
    int pos;
    // the next `count` output words (numpy: if (pos == 624) gen; y = key[pos++]; temper)
    void fill(uint32_t* out, size_t count) {
        size_t i = 0;
        while (i < count) {
            if (pos >= MT_N) {
                mt_regen(key);
                pos = 0;
            }
            const size_t take = std::min<size_t>((size_t)(MT_N - pos), count - i);
            temper_block(key + pos, out + i, take);
            pos += (int)take;
            i += take;
        }
    }
    void skip(size_t count) {
        while (count > 0) {
            if (pos >= MT_N) {
                mt_regen(key);
                pos = 0;
            }
            const size_t take = std::min<size_t>((size_t)(MT_N - pos), count);
            pos += (int)take;
            count -= take;
        }
    }
};
inline double to_double(uint32_t a, uint32_t b) {   // mt19937_next_double
    return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0;
}
// One compaction block of candidate pairs in four passes, so that everything except the libm logarithm vectorises
// (conversions, products, the divide and the square root are correctly rounded whether scalar or SIMD, so the numbers
// are the ones numpy's scalar loop produces):
//   1. x1, x2, r2 of every candidate (vector)            2. accepted candidates packed densely (branch-free scalar)
//   3. log(r2) of the accepted ones (scalar libm)        4. f = sqrt(-2 log(r2) / r2), (f x2, f x1) written out (vector)
__attribute__((target_clones("avx512f", "avx2", "default"))) void pairs_xy(const uint32_t* w, int64_t n, double* x1, double* x2, double* r2) {
    for (int64_t p = 0; p < n; ++p) {
        const double a = 2.0 * (((w[4 * p] >> 5) * 67108864.0 + (w[4 * p + 1] >> 6)) / 9007199254740992.0) - 1.0;
        const double b = 2.0 * (((w[4 * p + 2] >> 5) * 67108864.0 + (w[4 * p + 3] >> 6)) / 9007199254740992.0) - 1.0;
        x1[p] = a;
        x2[p] = b;
        r2[p] = a * a + b * b;
    }
}
inline int64_t pairs_pack(int64_t n, double* x1, double* x2, double* r2, uint8_t* ok) {   // in place: j <= p always
    int64_t j = 0;
    for (int64_t p = 0; p < n; ++p) {
        const double r = r2[p];
        const bool acc = !(r >= 1.0 || r == 0.0);
        ok[p] = acc;
        x1[j] = x1[p];
        x2[j] = x2[p];
        r2[j] = r;
        j += acc;
    }
    return j;
}
__attribute__((target_clones("avx512f", "avx2", "default"))) void pairs_out(int64_t n, const double* x1, const double* x2, const double* r2,
                                                                            const double* lg, double* dst) {
    for (int64_t j = 0; j < n; ++j) {
        const double f = std::sqrt(-2.0 * lg[j] / r2[j]);   // legacy_gauss: sqrt(-2.0 * log(r2) / r2)
        dst[2 * j] = f * x2[j];        // returned first
        dst[2 * j + 1] = f * x1[j];    // numpy caches it and returns it next
    }
}

}  // namespace

// count standard normals of numpy's legacy stream into out; key / pos / has_gauss / gauss: RandomState.get_state() in,
// the state after the draw out.
void numpy_legacy_normals(uint32_t* key, int32_t* pos, int32_t* has_gauss, double* gauss, int64_t count, double* out) {
    int64_t written = 0;
    if (count > 0 && *has_gauss) {
        out[written++] = *gauss;
        *has_gauss = 0;
        *gauss = 0.0;
    }
    if (written >= count) return;
    Twister tw;
    std::memcpy(tw.key, key, sizeof(tw.key));
    tw.pos = *pos;
    const int64_t pairs_needed = (count - written + 1) / 2;   // accepted pairs still to find
    const bool odd = ((count - written) & 1) != 0;            // the last pair's second number stays cached
    // A chunk = candidate pairs evaluated together (acceptance = pi / 4).  Two chunks: while the cores evaluate one, a
    // thread runs the twister for the next (the twister is the sequential part; the evaluation hides behind it).
    constexpr int64_t CH = 1 << 19, BLK = 4096;   // candidate pairs per chunk / per compaction block
    constexpr int SEG = 8;
    struct Chunk {
        std::vector<uint32_t> words;
        Twister snap[SEG];   // the twister in front of every eighth of the chunk (the draw ends inside one of them)
        int64_t seg_pairs = 0;
        int64_t cand = 0;    // candidate pairs in the chunk
    };
    static std::mutex mu;   // the scratch below is kept between calls (one draw at a time: numpy's global stream is one)
    std::lock_guard<std::mutex> lock(mu);
    static Chunk ch[2];
    static hvec<double> packed;          // per compaction block, dense: (f * x2, f * x1) of its accepted pairs, in order
    static hvec<double> scratch;         // per compaction block: x1 | x2 | r2 | log r2
    static std::vector<uint8_t> ok;
    static std::vector<int64_t> blk_cnt;
    if (ok.empty()) {
        ch[0].words.resize((size_t)CH * 4);
        ch[1].words.resize((size_t)CH * 4);
        packed.resize((size_t)CH * 2);
        scratch.resize((size_t)CH * 4);
        ok.resize((size_t)CH);
        blk_cnt.resize((size_t)(CH / BLK) + 2);
    }
    auto produce = [&](Chunk& c, int64_t want) {   // enough candidates for `want` pairs at the expected acceptance
        c.cand = std::min<int64_t>(CH, std::max<int64_t>(BLK, (int64_t)(want * 1.2739) + 64));   // 4 / pi = 1.2732
        c.seg_pairs = (c.cand + SEG - 1) / SEG;
        for (int g = 0; g < SEG; ++g) {
            const int64_t p0 = std::min(c.cand, g * c.seg_pairs), p1 = std::min(c.cand, p0 + c.seg_pairs);
            c.snap[g] = tw;
            tw.fill(c.words.data() + 4 * p0, (size_t)(p1 - p0) * 4);
        }
    };
    const bool dbg = getenv("SIGSDP_RANDN_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    auto t_start = now();
    produce(ch[0], pairs_needed);
    if (dbg) fprintf(stderr, "[randn] first produce %.2f ms (cand %lld)\n", ms(t_start, now()), (long long)ch[0].cand);
    int cur = 0;
    int64_t found = 0;
    for (;;) {
        Chunk& c = ch[cur];
        const int64_t cand = c.cand, want = pairs_needed - found;
        // the chunk after this one, if this one is not expected to finish the draw (a wrong guess costs nothing but the
        // words: the twister is rewound to the end of the last used pair anyway)
        const int64_t left = want - (int64_t)(cand * 0.7853981634);
        std::thread producer;
        const bool ahead = left > -(cand / 64) - 64;
        if (ahead) {
            side_thread_begin();
            producer = std::thread([&] {
                produce(ch[cur ^ 1], std::max<int64_t>(left, BLK));
                side_thread_end();
            });
        }
        const uint32_t* words = c.words.data();
        const int64_t nblk = (cand + BLK - 1) / BLK;
        auto t0 = now();
        parallel_for(nblk, [&](int64_t b0, int64_t b1) {
            for (int64_t b = b0; b < b1; ++b) {
                const int64_t p0 = b * BLK, p1 = std::min(cand, p0 + BLK);
                double* dst = packed.data() + 2 * p0;
                double* x1 = scratch.data() + 4 * p0;   // the block's own slice: x1 | x2 | r2 | log r2
                double* x2 = x1 + BLK;
                double* r2 = x2 + BLK;
                double* lg = r2 + BLK;
                pairs_xy(words + 4 * p0, p1 - p0, x1, x2, r2);
                const int64_t n_acc = pairs_pack(p1 - p0, x1, x2, r2, ok.data() + p0);
                for (int64_t j = 0; j < n_acc; ++j) lg[j] = std::log(r2[j]);
                pairs_out(n_acc, x1, x2, r2, lg, dst);
                blk_cnt[b] = n_acc;
            }
        }, 2);
        auto t1 = now();
        // exclusive scan of the accepted counts; where the last needed pair sits
        int64_t run = 0;
        for (int64_t b = 0; b < nblk; ++b) {
            const int64_t n_acc = blk_cnt[b];
            blk_cnt[b] = run;
            run += n_acc;
        }
        blk_cnt[nblk] = run;
        const int64_t take = std::min<int64_t>(run, want);    // accepted pairs of this chunk that are used
        int64_t last_pair = cand - 1;                          // candidate index of the last used pair
        double last_second = 0.0;                              // its second number
        if (run >= want) {
            int64_t b = 0;
            while (blk_cnt[b + 1] < want) ++b;
            int64_t seen = blk_cnt[b];
            for (int64_t p = b * BLK;; ++p)
                if (ok[p] && ++seen == want) {
                    last_pair = p;
                    last_second = packed[(size_t)(2 * b * BLK + 2 * (want - blk_cnt[b] - 1) + 1)];
                    break;
                }
        }
        const int64_t base = written + 2 * found;
        parallel_for(nblk, [&](int64_t b0, int64_t b1) {   // the blocks' dense runs, end to end
            for (int64_t b = b0; b < b1; ++b) {
                const int64_t q = blk_cnt[b];
                if (q >= take) break;
                const int64_t pairs = std::min(blk_cnt[b + 1], take) - q;
                const int64_t o = base + 2 * q;
                const int64_t doubles = std::min<int64_t>(2 * pairs, count - o);   // (an odd count drops the very last second number)
                std::memcpy(out + o, packed.data() + 2 * b * BLK, (size_t)doubles * sizeof(double));
            }
        }, 2);
        auto t2 = now();
        if (ahead) producer.join();
        if (dbg) fprintf(stderr, "[randn] chunk cand %lld: eval %.2f ms, compact %.2f ms, wait for producer %.2f ms\n", (long long)cand, ms(t0, t1), ms(t1, t2), ms(t2, now()));
        found += take;
        if (found >= pairs_needed) {
            // the stream stops right behind the last used pair: rewind to the start of its eighth and walk there
            const int64_t g = (last_pair + 1) / c.seg_pairs >= SEG ? SEG - 1 : (last_pair + 1) / c.seg_pairs;
            tw = c.snap[g];
            tw.skip((size_t)(last_pair + 1 - g * c.seg_pairs) * 4);
            if (odd) {
                *has_gauss = 1;
                *gauss = last_second;
            }
            break;
        }
        if (!ahead) produce(ch[cur ^ 1], pairs_needed - found);
        cur ^= 1;
    }
    std::memcpy(key, tw.key, sizeof(tw.key));
    *pos = tw.pos;
}

}  // namespace sigsdp
