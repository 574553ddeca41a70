// host_rng.cpp -- host-side integer bookkeeping of the minibatch step, bit-exact with the
// reference's compiled helpers:
//   * Mt19937Stream     <-> modl/utils/randomkit/random_fast.pyx:49-150 over randomkit.c
//   * FeatureSampler    <-> modl/utils/randomkit/sampler.pyx:10-69
//   * modl_batch_weight <-> modl/decomposition/dict_fact_fast.pyx:115-122
// The subset / permutation streams decide WHICH columns and atoms every kernel touches, so
// they must reproduce the reference integer for integer; tests/test_host_bookkeeping.py
// checks the reference's known-answer vectors through this C ABI.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/modl_b200.h"

namespace {

class Mt19937Stream {
public:
    explicit Mt19937Stream(uint64_t seed) { reseed(seed); }

    // Knuth multiplicative fill of the 624-word state from the low 32 seed bits.
    void reseed(uint64_t seed) {
        uint32_t x = static_cast<uint32_t>(seed);
        for (uint32_t i = 0; i < kWords; ++i) {
            state_[i] = x;
            x = 1812433253u * (x ^ (x >> 30)) + i + 1u;
        }
        cursor_ = kWords;
        binom_.valid = false;
    }

    uint32_t next32() {
        if (cursor_ >= kWords) refill();
        uint32_t y = state_[cursor_++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        return y ^ (y >> 18);
    }

    // uniform integer in the closed range [0, top]: smallest covering bit mask + rejection;
    // one 32-bit word per try when top fits 32 bits, else two words (high first).
    uint64_t bounded(uint64_t top) {
        if (top == 0) return 0;
        uint64_t mask = top;
        for (int sh = 1; sh < 64; sh <<= 1) mask |= mask >> sh;
        const bool narrow = top <= 0xffffffffull;
        for (;;) {
            uint64_t draw;
            if (narrow) {
                draw = next32();
            } else {
                const uint64_t hi = next32();
                draw = (hi << 32) | next32();
            }
            draw &= mask;
            if (draw <= top) return draw;
        }
    }

    // 53-bit uniform double in [0, 1) from a 27-bit and a 26-bit draw
    double unit() {
        const int64_t a = next32() >> 5;
        const int64_t b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }

    int64_t binomial(int64_t n, double p) {
        if (p <= 0.5) return (p * n <= 30.0) ? binomial_small(n, p) : binomial_btpe(n, p);
        const double q = 1.0 - p;
        return n - ((q * n <= 30.0) ? binomial_small(n, q) : binomial_btpe(n, q));
    }

    // Fisher-Yates from the last slot down, j ~ U[0, i]
    template <typename V>
    void shuffle(V *x, int64_t n) {
        if (n - 1 > 0xffffffffll) {
            for (int64_t i = n - 1; i > 0; --i) {
                const int64_t j = static_cast<int64_t>(bounded(static_cast<uint64_t>(i)));
                const V t = x[i];
                x[i] = x[j];
                x[j] = t;
            }
            return;
        }
        // same draws as bounded(i) for every i, with the covering mask carried along instead of
        // being rebuilt per element (it only changes when i crosses a power of two)
        uint32_t mask = 0;
        if (n > 1) {
            mask = static_cast<uint32_t>(n - 1);
            for (int sh = 1; sh < 32; sh <<= 1) mask |= mask >> sh;
        }
        for (int64_t i = n - 1; i > 0; --i) {
            const uint32_t top = static_cast<uint32_t>(i);
            while ((mask >> 1) >= top) mask >>= 1;
            uint32_t j;
            do { j = next32() & mask; } while (j > top);
            const V t = x[i];
            x[i] = x[j];
            x[j] = t;
        }
    }

private:
    static constexpr uint32_t kWords = 624, kShift = 397;

    static uint32_t twist(uint32_t cur, uint32_t nxt, uint32_t far) {
        const uint32_t mix = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
        return far ^ (mix >> 1) ^ ((0u - (mix & 1u)) & 0x9908b0dfu);
    }

    // one full state transition, split at the wrap points so that the loops carry no modulo
    void refill() {
        uint32_t i = 0;
        for (; i < kWords - kShift; ++i) state_[i] = twist(state_[i], state_[i + 1], state_[i + kShift]);
        for (; i < kWords - 1; ++i) state_[i] = twist(state_[i], state_[i + 1], state_[i + kShift - kWords]);
        state_[kWords - 1] = twist(state_[kWords - 1], state_[0], state_[kShift - 1]);
        cursor_ = 0;
    }

    // The reference caches the set-up constants keyed on (n, p) in the generator state and
    // shares ONE cache between the two algorithms; so do we (the cache decides nothing about
    // the stream, but keeping one key mirrors its behaviour when (n, p) alternate).
    struct BinomCache {
        bool valid = false;
        int64_t n = 0;
        double p = 0;
        // inversion
        double q = 0, qn = 0;
        int64_t bound = 0;
        // BTPE
        double r = 0, nrq = 0, xm = 0, xl = 0, xr = 0, c = 0, laml = 0, lamr = 0;
        double p1 = 0, p2 = 0, p3 = 0, p4 = 0;
        int64_t m = 0;
    } binom_;

    bool cache_hit(int64_t n, double p) const { return binom_.valid && binom_.n == n && binom_.p == p; }

    // sequential search from 0 with restart beyond a 10-sigma bound (small n*p)
    int64_t binomial_small(int64_t n, double p) {
        BinomCache &c = binom_;
        if (!cache_hit(n, p)) {
            c.valid = true; c.n = n; c.p = p;
            c.q = 1.0 - p;
            c.qn = std::exp(n * std::log(c.q));
            const double np = n * p;
            const double lim = np + 10.0 * std::sqrt(np * c.q + 1);
            c.bound = static_cast<int64_t>(static_cast<double>(n) < lim ? static_cast<double>(n) : lim);
        }
        int64_t x = 0;
        double px = c.qn;
        double u = unit();
        while (u > px) {
            ++x;
            if (x > c.bound) {
                x = 0; px = c.qn; u = unit();
            } else {
                u -= px;
                px = ((n - x + 1) * p * px) / (x * c.q);
            }
        }
        return x;
    }

    // BTPE: triangle / parallelogram / two exponential tails, with squeeze acceptance
    int64_t binomial_btpe(int64_t n, double p) {
        BinomCache &c = binom_;
        if (!cache_hit(n, p)) {
            c.valid = true; c.n = n; c.p = p;
            c.r = (p < 1.0 - p) ? p : 1.0 - p;
            c.q = 1.0 - c.r;
            const double fm = n * c.r + c.r;
            c.m = static_cast<int64_t>(std::floor(fm));
            c.p1 = std::floor(2.195 * std::sqrt(n * c.r * c.q) - 4.6 * c.q) + 0.5;
            c.xm = c.m + 0.5;
            c.xl = c.xm - c.p1;
            c.xr = c.xm + c.p1;
            c.c = 0.134 + 20.5 / (15.3 + c.m);
            double a = (fm - c.xl) / (fm - c.xl * c.r);
            c.laml = a * (1.0 + a / 2.0);
            a = (c.xr - fm) / (c.xr * c.q);
            c.lamr = a * (1.0 + a / 2.0);
            c.p2 = c.p1 * (1.0 + 2.0 * c.c);
            c.p3 = c.p2 + c.c / c.laml;
            c.p4 = c.p3 + c.c / c.lamr;
        }
        const double r = c.r, q = c.q;
        const int64_t m = c.m;
        const double nrq = n * r * q;
        int64_t y = 0;
        bool accepted = false;
        while (!accepted) {
            const double u = unit() * c.p4;
            double v = unit();
            if (u <= c.p1) {
                y = static_cast<int64_t>(std::floor(c.xm - c.p1 * v + u));
                break;
            }
            if (u <= c.p2) {
                const double x = c.xl + (u - c.p1) / c.c;
                v = v * c.c + 1.0 - std::fabs(m - x + 0.5) / c.p1;
                if (v > 1.0) continue;
                y = static_cast<int64_t>(std::floor(x));
            } else if (u <= c.p3) {
                y = static_cast<int64_t>(std::floor(c.xl + std::log(v) / c.laml));
                if (y < 0) continue;
                v = v * (u - c.p2) * c.laml;
            } else {
                y = static_cast<int64_t>(std::floor(c.xr - std::log(v) / c.lamr));
                if (y > n) continue;
                v = v * (u - c.p3) * c.lamr;
            }
            const int64_t dist = std::llabs(y - m);
            if (dist > 20 && dist < nrq / 2.0 - 1) {
                // squeeze on log f(y)/f(m)
                const double k = static_cast<double>(dist);
                const double rho = (k / nrq) * ((k * (k / 3.0 + 0.625) + 0.16666666666666666) / nrq + 0.5);
                const double t = -k * k / (2 * nrq);
                const double A = std::log(v);
                if (A < t - rho) { accepted = true; continue; }
                if (A > t + rho) continue;
                const double x1 = y + 1, f1 = m + 1, z = n + 1 - m, w = n - y + 1;
                const double x2 = x1 * x1, f2 = f1 * f1, z2 = z * z, w2 = w * w;
                auto stirling = [](double s2, double s1) {
                    return (13680. - (462. - (132. - (99. - 140. / s2) / s2) / s2) / s2) / s1 / 166320.;
                };
                const double bound = c.xm * std::log(f1 / x1) + (n - m + 0.5) * std::log(z / w)
                                     + (y - m) * std::log(w * r / (x1 * q))
                                     + stirling(f2, f1) + stirling(z2, z) + stirling(x2, x1) + stirling(w2, w);
                accepted = !(A > bound);
            } else {
                // explicit ratio f(y)/f(m) by the recurrence
                const double s = r / q;
                const double a = s * (n + 1);
                double F = 1.0;
                if (m < y) {
                    for (int64_t i = m + 1; i <= y; ++i) F *= (a / i - s);
                } else if (m > y) {
                    for (int64_t i = y + 1; i <= m; ++i) F /= (a / i - s);
                }
                accepted = !(v > F);
            }
        }
        return (p > 0.5) ? n - y : y;
    }

    uint32_t state_[kWords];
    uint32_t cursor_ = kWords;
};

// Feature-subset generator ("reduction"): keeps a shuffled box of all feature ids and
// hands out either a freshly reshuffled prefix (with replacement across calls) or
// consecutive windows that tile the box (without replacement).
class FeatureSampler {
public:
    FeatureSampler(int64_t range, bool rand_size, bool replacement, uint64_t seed)
        : range_(range), rand_size_(rand_size), replacement_(replacement), rng_(seed),
          box_(static_cast<size_t>(range > 0 ? range : 0)) {
        for (int64_t i = 0; i < range_; ++i) box_[static_cast<size_t>(i)] = i;
        rng_.shuffle(box_.data(), range_);   // permutation(range)   [sampler.pyx:34]
        rng_.shuffle(box_.data(), range_);   // shuffle(box)         [sampler.pyx:39]
    }

    int64_t draw(double reduction, int64_t *out) {
        int64_t len;
        if (rand_size_) {
            // the Cython wrapper narrows n and the result to C int [random_fast.pyx:146-147]
            len = static_cast<int32_t>(rng_.binomial(static_cast<int32_t>(range_), 1. / reduction));
        } else {
            len = static_cast<int64_t>(static_cast<double>(range_) / reduction);
        }
        if (replacement_) {
            rng_.shuffle(box_.data(), range_);
            lo_ = 0;
            hi_ = len;
        } else if (range_ != len) {
            lo_ = hi_;
            const int64_t left = range_ - lo_;
            if (left == 0) {
                rng_.shuffle(box_.data(), range_);
                lo_ = 0;
            } else if (left < len) {
                // bring the unused tail to the front, swap the displaced head into its place,
                // reshuffle everything behind the kept tail  [sampler.pyx:59-64]
                std::vector<int64_t> head(box_.begin(), box_.begin() + left);
                std::memmove(box_.data(), box_.data() + lo_, sizeof(int64_t) * static_cast<size_t>(left));
                std::memcpy(box_.data() + lo_, head.data(), sizeof(int64_t) * static_cast<size_t>(left));
                rng_.shuffle(box_.data() + left, range_ - left);
                lo_ = 0;
            }
            hi_ = lo_ + len;
        } else {
            lo_ = 0;
            hi_ = range_;
        }
        int64_t cnt = hi_ - lo_;
        if (cnt < 0) cnt = 0;
        if (hi_ > range_) cnt = range_ - lo_;   // defensive: never read past the box
        std::memcpy(out, box_.data() + lo_, sizeof(int64_t) * static_cast<size_t>(cnt));
        return cnt;
    }

private:
    int64_t range_;
    bool rand_size_, replacement_;
    Mt19937Stream rng_;
    std::vector<int64_t> box_;
    int64_t lo_ = 0, hi_ = 0;
};

}  // namespace

struct modl_rng { Mt19937Stream impl; explicit modl_rng(uint64_t s) : impl(s) {} };
// The sampler stream does not depend on the data, so the NEXT subset (10^4 Mersenne-twister draws at
// the benchmark shape, ~0.1 ms) is drawn by a helper thread on a COPY of the state while the caller
// launches the kernels of the current step.  yield_subset() adopts the copy when it was drawn with
// the same `reduction`, otherwise drops it and draws from the untouched original: the integer
// stream handed out is exactly the sequential one.
struct modl_sampler {
    FeatureSampler impl;
    int64_t range;
    // look-ahead state (guarded by mu)
    FeatureSampler ahead;
    std::vector<int64_t> ahead_out;
    int64_t ahead_len = 0;
    double ahead_reduction = 0;
    bool job = false, ready = false, stop = false, started = false;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;

    modl_sampler(int64_t r, bool rs, bool rp, uint64_t s)
        : impl(r, rs, rp, s), range(r), ahead(impl), ahead_out(static_cast<size_t>(r > 0 ? r : 1)) {}

    ~modl_sampler() {
        if (started) {
            { std::lock_guard<std::mutex> lk(mu); stop = true; }
            cv.notify_all();
            worker.join();
        }
    }

    void run() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return job || stop; });
            if (stop) return;
            job = false;
            const double red = ahead_reduction;
            lk.unlock();
            ahead = impl;                                   // the caller does not touch impl until it has waited for us
            const int64_t len = ahead.draw(red, ahead_out.data());
            lk.lock();
            ahead_len = len;
            ready = true;
            cv.notify_all();
        }
    }

    void start_ahead(double reduction) {
        if (!started) {
            started = true;
            worker = std::thread([this] { run(); });
        }
        { std::lock_guard<std::mutex> lk(mu); ahead_reduction = reduction; ready = false; job = true; }
        cv.notify_all();
    }

    int64_t yield(double reduction, int64_t *out) {
        int64_t len = -1;
        if (started) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return ready; });            // a look-ahead is always pending once started
            if (ahead_reduction == reduction) {
                len = ahead_len;
                std::memcpy(out, ahead_out.data(), sizeof(int64_t) * static_cast<size_t>(len));
                std::swap(impl, ahead);                     // commit: the copy becomes the state
            }
        }
        if (len < 0) len = impl.draw(reduction, out);
        if (lookahead) start_ahead(reduction);
        return len;
    }

    bool lookahead = true;
};

extern "C" {

modl_rng *modl_rs_create(uint64_t seed) { return new (std::nothrow) modl_rng(seed); }
void modl_rs_destroy(modl_rng *rs) { delete rs; }
void modl_rs_seed(modl_rng *rs, uint64_t seed) { rs->impl.reseed(seed); }
int64_t modl_rs_randint(modl_rng *rs, uint64_t high) { return static_cast<int64_t>(rs->impl.bounded(high)); }
int64_t modl_rs_binomial(modl_rng *rs, int64_t n, double p) {
    // int narrowing of the reference wrapper [random_fast.pyx:146-147]
    return static_cast<int32_t>(rs->impl.binomial(static_cast<int32_t>(n), p));
}
void modl_rs_permutation(modl_rng *rs, int64_t *h_out, int64_t size) {
    for (int64_t i = 0; i < size; ++i) h_out[i] = i;
    rs->impl.shuffle(h_out, size);
}
void modl_rs_shuffle(modl_rng *rs, int64_t *h_x, int64_t n) { rs->impl.shuffle(h_x, n); }
void modl_rs_shuffle_with_trace(modl_rng *rs, int64_t n, int64_t *h_swap, int64_t *h_trace) {
    for (int64_t i = 0; i < n; ++i) h_trace[i] = i;
    if (h_swap && n > 0) h_swap[0] = 0;
    for (int64_t i = n - 1; i > 0; --i) {
        const int64_t j = static_cast<int64_t>(rs->impl.bounded(static_cast<uint64_t>(i)));
        if (h_swap) h_swap[i] = j;
        const int64_t t = h_trace[i];
        h_trace[i] = h_trace[j];
        h_trace[j] = t;
    }
}

modl_sampler *modl_sampler_create(int64_t range, int rand_size, int replacement, uint64_t seed) {
    if (range < 0) return nullptr;
    modl_sampler *s = new (std::nothrow) modl_sampler(range, rand_size != 0, replacement != 0, seed);
    if (s) {
        const char *e = std::getenv("MODL_SAMPLER_LOOKAHEAD");
        if (e && std::atoi(e) == 0) s->lookahead = false;
    }
    return s;
}
void modl_sampler_destroy(modl_sampler *s) { delete s; }
int64_t modl_sampler_yield_subset(modl_sampler *s, double reduction, int64_t *h_out) {
    return s->yield(reduction, h_out);
}

double modl_batch_weight(int64_t count, int64_t batch_size, double learning_rate, double offset) {
    double keep = 1;
    for (int64_t i = count + 1 - batch_size; i <= count; ++i)
        keep *= (1 - std::pow((1 + offset) / (offset + i), learning_rate));
    return 1 - keep;
}

}  // extern "C"
