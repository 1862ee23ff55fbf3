// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU restatement (C++17, scalar, single transform at a time) of the algorithm family RustFFT's
// *scalar* backend runs behind FftPlanner / Fft::process.  It is the parity oracle for the CUDA
// path in rustfft_b200/csrc and the "port" CPU baseline of bench.py.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// The reference (Rust) cannot be compiled in this environment (no rustc/cargo), so this file is
// a restatement written from the reference's maths; every routine cites the reference file:line
// it follows (paths relative to /root/reference).  It is pinned by the reference's own
// known-answer tests (tests/test_oracle.py): src/algorithm/dft.rs:283-398,
// src/math_utils.rs:495-540,617-631, src/plan.rs:712-830, src/twiddles.rs:77-98, and against an
// f64 numpy.fft ground truth.
//
// Build for parity with `-ffp-contract=off` (Rust never fuses mul+add; num-complex's Complex
// product is the 4-mul/2-add textbook form) -- see oracle/Makefile.
//
// Leaf butterflies (src/algorithm/butterflies.rs): 1,2,3,4,5,6,7,8,16,32 and the composite ones 9 (3x3 mixed radix),
// 12 (4x3 Good-Thomas), 24 (6x4 mixed radix), 27 (9x3 mixed radix) are restated operation for operation; the prime
// ones 11,13,17,19,23,29,31 -- generated code in the reference (tools/genbutterflies.py, :842-6242) -- are restated
// as the loop that generator unrolls (same pairing, same summation order, same sign placement).

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

typedef std::size_t usize;

// ----------------------------------------------------------------------------------------
// Complex<T>: repr(C) {re, im}; + - * as in num-complex 0.4 (no FMA).
// ----------------------------------------------------------------------------------------
template <class T>
struct Cx {
    T re, im;
};
template <class T> inline Cx<T> operator+(Cx<T> a, Cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <class T> inline Cx<T> operator-(Cx<T> a, Cx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <class T> inline Cx<T> operator*(Cx<T> a, Cx<T> b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <class T> inline Cx<T> scale(Cx<T> a, T s) { return {a.re * s, a.im * s}; }
template <class T> inline Cx<T> conj(Cx<T> a) { return {a.re, -a.im}; }

// src/twiddles.rs:6-23 -- angle and trig in f64, one rounding to T, inverse = conjugate.
template <class T>
Cx<T> twiddle(usize index, usize fft_len, bool inverse) {
    const double constant = -2.0 * 3.14159265358979323846264338327950288 / (double)fft_len;
    const double angle = constant * (double)index;
    Cx<T> r{(T)std::cos(angle), (T)std::sin(angle)};
    return inverse ? conj(r) : r;
}

// src/twiddles.rs:59-70
template <class T>
inline Cx<T> rot90(Cx<T> v, bool inverse) {
    return inverse ? Cx<T>{-v.im, v.re} : Cx<T>{v.im, -v.re};
}

// ----------------------------------------------------------------------------------------
// Integer helpers: src/math_utils.rs
// ----------------------------------------------------------------------------------------
// src/math_utils.rs:23-37
uint64_t modular_exponent(uint64_t base, uint64_t exponent, uint64_t modulo) {
    uint64_t result = 1;
    while (exponent > 0) {
        if (exponent & 1) result = result * base % modulo;
        exponent >>= 1;
        base = (base * base) % modulo;
    }
    return result;
}

// src/math_utils.rs:40-74
std::vector<uint64_t> distinct_prime_factors(uint64_t n) {
    std::vector<uint64_t> out;
    if (n % 2 == 0) {
        while (n % 2 == 0) n /= 2;
        out.push_back(2);
    }
    if (n > 1) {
        uint64_t divisor = 3;
        uint64_t limit = (uint64_t)std::sqrt((float)n) + 1;
        while (divisor < limit) {
            if (n % divisor == 0) {
                while (n % divisor == 0) n /= divisor;
                out.push_back(divisor);
                limit = (uint64_t)std::sqrt((float)n) + 1;
            }
            divisor += 2;
        }
        if (n > 1) out.push_back(n);
    }
    return out;
}

// src/math_utils.rs:3-20; returns 0 when no root exists
uint64_t primitive_root(uint64_t prime) {
    std::vector<uint64_t> exps;
    for (uint64_t f : distinct_prime_factors(prime - 1)) exps.push_back((prime - 1) / f);
    for (uint64_t cand = 2; cand < prime; ++cand) {
        bool ok = true;
        for (uint64_t e : exps)
            if (modular_exponent(cand, e, prime) == 1) {
                ok = false;
                break;
            }
        if (ok) return cand;
    }
    return 0;
}

// multiplicative inverse via the extended Euclid recurrence (num-integer extended_gcd call sites:
// src/algorithm/raders_algorithm.rs:79-84, src/algorithm/good_thomas_algorithm.rs:377-393)
int64_t mod_inverse(int64_t a, int64_t m) {
    int64_t old_r = a, r = m, old_s = 1, s = 0;
    while (r != 0) {
        int64_t q = old_r / r;
        int64_t t = old_r - q * r;
        old_r = r;
        r = t;
        t = old_s - q * s;
        old_s = s;
        s = t;
    }
    int64_t x = old_s % m;
    if (x < 0) x += m;
    return x;
}

usize gcd_usize(usize a, usize b) {
    while (b) {
        usize t = a % b;
        a = b;
        b = t;
    }
    return a;
}

bool is_prime_u64(uint64_t n) {
    if (n < 2) return false;
    for (uint64_t d = 2; d * d <= n; ++d)
        if (n % d == 0) return false;
    return true;
}

// src/math_utils.rs:83-369
struct PrimeFactors {
    std::vector<std::pair<usize, uint32_t>> other;  // (value, count), ascending
    usize n = 1;
    uint32_t p2 = 0, p3 = 0, total = 0, distinct = 0;

    static usize ipow(usize b, uint32_t e) {
        usize r = 1;
        while (e--) r *= b;
        return r;
    }
    static PrimeFactors compute(usize n) {  // :92-160
        PrimeFactors r;
        r.n = n;
        while (n % 2 == 0 && n > 0) {
            r.p2++;
            n /= 2;
        }
        r.total += r.p2;
        if (r.p2) r.distinct++;
        while (n % 3 == 0 && n > 0) {
            r.p3++;
            n /= 3;
        }
        r.total += r.p3;
        if (r.p3) r.distinct++;
        if (n > 1) {
            usize divisor = 5;
            usize limit = (usize)std::sqrt((float)n) + 1;
            while (divisor < limit) {
                uint32_t count = 0;
                while (n % divisor == 0) {
                    n /= divisor;
                    count++;
                }
                if (count) {
                    r.other.push_back({divisor, count});
                    r.total += count;
                    r.distinct++;
                    limit = (usize)std::sqrt((float)n) + 1;
                }
                divisor += 2;
            }
            if (n > 1) {
                r.other.push_back({n, 1});
                r.total++;
                r.distinct++;
            }
        }
        return r;
    }
    bool is_prime() const { return total == 1; }
    bool has_factors_leq(usize f) const {  // :239-247
        return p2 > 0 || p3 > 0 || (!other.empty() && other.front().first <= f);
    }
    bool has_factors_gt(usize f) const {  // :250-258
        return (f < 2 && p2 > 0) || (f < 3 && p3 > 0) || (!other.empty() && other.back().first > f);
    }
    usize product_above(usize min_factor) const {  // :261-267
        usize p = 1;
        bool skipping = true;
        for (auto& f : other) {
            if (skipping && f.first <= min_factor) continue;
            skipping = false;
            p *= ipow(f.first, f.second);
        }
        return p;
    }
    // :269-368 -- split into two sets with products as close as possible
    std::pair<PrimeFactors, PrimeFactors> partition() const {
        PrimeFactors self = *this;
        bool all_even = (p2 % 2 == 0) && (p3 % 2 == 0);
        for (auto& f : other) all_even = all_even && (f.second % 2 == 0);
        if (all_even) {
            usize prod = 1;
            self.p2 /= 2;
            prod <<= self.p2;
            self.p3 /= 2;
            prod *= ipow(3, self.p3);
            for (auto& f : self.other) {
                f.second /= 2;
                prod *= ipow(f.first, f.second);
            }
            self.total /= 2;
            self.n = prod;
            return {self, self};
        } else if (distinct == 1) {
            PrimeFactors half;
            half.n = n;
            half.p2 = p2 / 2;
            half.p3 = p3 / 2;
            half.total = total / 2;
            half.distinct = 1;
            self.p2 -= half.p2;
            self.p3 -= half.p3;
            self.total -= half.total;
            if (!self.other.empty()) {
                auto& first = self.other.front();
                std::pair<usize, uint32_t> hf{first.first, first.second / 2};
                first.second -= hf.second;
                half.other.push_back(hf);
                self.n = ipow(first.first, first.second);
                half.n = ipow(hf.first, hf.second);
            } else if (half.p2 > 0) {
                half.n = (usize)1 << half.p2;
                self.n = (usize)1 << self.p2;
            } else if (half.p3 > 0) {
                half.n = ipow(3, half.p3);
                self.n = ipow(3, self.p3);
            }
            return {self, half};
        } else {
            usize left = 1, right = 1;
            for (auto& f : other) {
                usize fp = ipow(f.first, f.second);
                if (left <= right)
                    left *= fp;
                else
                    right *= fp;
            }
            if (left <= right)
                left <<= p2;
            else
                right <<= p2;
            if (p3 > 0 && left <= right)
                left *= ipow(3, p3);
            else
                right *= ipow(3, p3);
            return {compute(left), compute(right)};
        }
    }
};

// ----------------------------------------------------------------------------------------
// Algorithm nodes.  One node == one `impl Fft<T>` object of the reference; run() transforms one
// chunk of len() elements in place (the reference's three process_* variants perform the same
// arithmetic and differ only in which buffer holds the result; batching is the serial chunk loop
// of src/array_utils.rs:151-177).
// ----------------------------------------------------------------------------------------
template <class T>
struct Node {
    usize len;
    bool inv;
    Node(usize l, bool i) : len(l), inv(i) {}
    virtual ~Node() {}
    virtual void run(Cx<T>* x) = 0;
    virtual std::string describe() const = 0;
};
template <class T>
using NodeP = std::shared_ptr<Node<T>>;

// src/algorithm/dft.rs:28-71
template <class T>
struct DftNode : Node<T> {
    std::vector<Cx<T>> tw;
    std::vector<Cx<T>> tmp;
    bool stands_for_butterfly = false;  // see header: leaf sizes evaluated through the naive Dft
    DftNode(usize len, bool inv) : Node<T>(len, inv), tw(len), tmp(len) {
        for (usize i = 0; i < len; ++i) tw[i] = twiddle<T>(i, len, inv);
    }
    void run(Cx<T>* x) override {
        const usize n = this->len;
        for (usize k = 0; k < n; ++k) {
            Cx<T> acc{0, 0};
            usize ti = 0;
            for (usize j = 0; j < n; ++j) {
                acc = acc + tw[ti] * x[j];
                ti += k;
                if (ti >= n) ti -= n;
            }
            tmp[k] = acc;
        }
        for (usize k = 0; k < n; ++k) x[k] = tmp[k];
    }
    std::string describe() const override {
        return (stands_for_butterfly ? "Butterfly" : "Dft(") + std::to_string(this->len) + (stands_for_butterfly ? "" : ")");
    }
};

// --- leaf butterflies, src/algorithm/butterflies.rs ---------------------------------------
template <class T> inline void bf2(Cx<T>& a, Cx<T>& b) {  // :175-180
    Cx<T> t = a + b;
    b = a - b;
    a = t;
}
template <class T> inline void bf3(Cx<T>* v, Cx<T> tw) {  // :230-248
    Cx<T> xp = v[1] + v[2], xn = v[1] - v[2], sum = v[0] + xp;
    Cx<T> ta = v[0] + Cx<T>{tw.re * xp.re, tw.re * xp.im};
    Cx<T> tb{-tw.im * xn.im, tw.im * xn.re};
    v[0] = sum;
    v[1] = ta + tb;
    v[2] = ta - tb;
}
template <class T> inline void bf4(Cx<T>* v, bool inv) {  // :265-293
    bf2(v[0], v[2]);
    bf2(v[1], v[3]);
    v[3] = rot90(v[3], inv);
    bf2(v[0], v[1]);
    bf2(v[2], v[3]);
    Cx<T> t = v[1];
    v[1] = v[2];
    v[2] = t;
}
template <class T> inline void bf5(Cx<T>* v, Cx<T> t1, Cx<T> t2) {  // :339-471
    Cx<T> x14p = v[1] + v[4], x14n = v[1] - v[4], x23p = v[2] + v[3], x23n = v[2] - v[3];
    Cx<T> sum = v[0] + x14p + x23p;
    T b14re_a = v[0].re + t1.re * x14p.re + t2.re * x23p.re;
    T b14re_b = t1.im * x14n.im + t2.im * x23n.im;
    T b23re_a = v[0].re + t2.re * x14p.re + t1.re * x23p.re;
    T b23re_b = t2.im * x14n.im + -t1.im * x23n.im;
    T b14im_a = v[0].im + t1.re * x14p.im + t2.re * x23p.im;
    T b14im_b = t1.im * x14n.re + t2.im * x23n.re;
    T b23im_a = v[0].im + t2.re * x14p.im + t1.re * x23p.im;
    T b23im_b = t2.im * x14n.re + -t1.im * x23n.re;
    v[0] = sum;
    v[1] = {b14re_a - b14re_b, b14im_a + b14im_b};
    v[2] = {b23re_a - b23re_b, b23im_a + b23im_b};
    v[3] = {b23re_a + b23re_b, b23im_a - b23im_b};
    v[4] = {b14re_a + b14re_b, b14im_a - b14im_b};
}
template <class T> inline void bf6(Cx<T>* v, Cx<T> tw3) {  // :494-525 (Good-Thomas 2x3)
    Cx<T> a[3] = {v[0], v[2], v[4]};
    Cx<T> b[3] = {v[3], v[5], v[1]};
    bf3(a, tw3);
    bf3(b, tw3);
    bf2(a[0], b[0]);
    bf2(a[1], b[1]);
    bf2(a[2], b[2]);
    v[0] = a[0];
    v[1] = b[1];
    v[2] = a[2];
    v[3] = b[0];
    v[4] = a[1];
    v[5] = b[2];
}
template <class T> inline void bf7(Cx<T>* v, Cx<T> t1, Cx<T> t2, Cx<T> t3) {  // :545-716
    Cx<T> x16p = v[1] + v[6], x16n = v[1] - v[6], x25p = v[2] + v[5], x25n = v[2] - v[5];
    Cx<T> x34p = v[3] + v[4], x34n = v[3] - v[4];
    Cx<T> sum = v[0] + x16p + x25p + x34p;
    T x16re_a = v[0].re + t1.re * x16p.re + t2.re * x25p.re + t3.re * x34p.re;
    T x16re_b = t1.im * x16n.im + t2.im * x25n.im + t3.im * x34n.im;
    T x25re_a = v[0].re + t1.re * x34p.re + t2.re * x16p.re + t3.re * x25p.re;
    T x25re_b = -t1.im * x34n.im + t2.im * x16n.im - t3.im * x25n.im;
    T x34re_a = v[0].re + t1.re * x25p.re + t2.re * x34p.re + t3.re * x16p.re;
    T x34re_b = -t1.im * x25n.im + t2.im * x34n.im + t3.im * x16n.im;
    T x16im_a = v[0].im + t1.re * x16p.im + t2.re * x25p.im + t3.re * x34p.im;
    T x16im_b = t1.im * x16n.re + t2.im * x25n.re + t3.im * x34n.re;
    T x25im_a = v[0].im + t1.re * x34p.im + t2.re * x16p.im + t3.re * x25p.im;
    T x25im_b = -t1.im * x34n.re + t2.im * x16n.re - t3.im * x25n.re;
    T x34im_a = v[0].im + t1.re * x25p.im + t2.re * x34p.im + t3.re * x16p.im;
    T x34im_b = t1.im * x25n.re - t2.im * x34n.re - t3.im * x16n.re;
    v[0] = sum;
    v[1] = {x16re_a - x16re_b, x16im_a + x16im_b};
    v[2] = {x25re_a - x25re_b, x25im_a + x25im_b};
    v[3] = {x34re_a - x34re_b, x34im_a - x34im_b};
    v[4] = {x34re_a + x34re_b, x34im_a + x34im_b};
    v[5] = {x25re_a + x25re_b, x25im_a - x25im_b};
    v[6] = {x16re_a + x16re_b, x16im_a - x16im_b};
}
template <class T> inline void bf8(Cx<T>* v, bool inv) {  // :734-777
    const T root2 = (T)std::sqrt(0.5);
    Cx<T> s0[4] = {v[0], v[2], v[4], v[6]};
    Cx<T> s1[4] = {v[1], v[3], v[5], v[7]};
    bf4(s0, inv);
    bf4(s1, inv);
    s1[1] = scale(rot90(s1[1], inv) + s1[1], root2);
    s1[2] = rot90(s1[2], inv);
    s1[3] = scale(rot90(s1[3], inv) - s1[3], root2);
    for (int i = 0; i < 4; ++i) bf2(s0[i], s1[i]);
    for (int i = 0; i < 4; ++i) v[i] = s0[i];
    for (int i = 0; i < 4; ++i) v[i + 4] = s1[i];
}
template <class T> inline void bf16(Cx<T>* v, bool inv) {  // :1506-1579 (split radix)
    const Cx<T> t1 = twiddle<T>(1, 16, inv), t2 = twiddle<T>(2, 16, inv), t3 = twiddle<T>(3, 16, inv);
    Cx<T> ev[8], n1[4], n3[4];
    for (int i = 0; i < 8; ++i) ev[i] = v[2 * i];
    for (int i = 0; i < 4; ++i) n1[i] = v[4 * i + 1];
    n3[0] = v[15];
    for (int i = 1; i < 4; ++i) n3[i] = v[4 * i - 1];
    bf8(ev, inv);
    bf4(n1, inv);
    bf4(n3, inv);
    n1[1] = n1[1] * t1;
    n3[1] = n3[1] * conj(t1);
    n1[2] = n1[2] * t2;
    n3[2] = n3[2] * conj(t2);
    n1[3] = n1[3] * t3;
    n3[3] = n3[3] * conj(t3);
    for (int i = 0; i < 4; ++i) bf2(n1[i], n3[i]);
    for (int i = 0; i < 4; ++i) n3[i] = rot90(n3[i], inv);
    for (int i = 0; i < 4; ++i) {
        v[i] = ev[i] + n1[i];
        v[i + 4] = ev[i + 4] + n3[i];
        v[i + 8] = ev[i] - n1[i];
        v[i + 12] = ev[i + 4] - n3[i];
    }
}
template <class T> inline void bf32(Cx<T>* v, bool inv) {  // :6269-6392 (split radix)
    Cx<T> tw[7];
    for (int i = 0; i < 7; ++i) tw[i] = twiddle<T>(i + 1, 32, inv);
    Cx<T> ev[16], n1[8], n3[8];
    for (int i = 0; i < 16; ++i) ev[i] = v[2 * i];
    for (int i = 0; i < 8; ++i) n1[i] = v[4 * i + 1];
    n3[0] = v[31];
    for (int i = 1; i < 8; ++i) n3[i] = v[4 * i - 1];
    bf16(ev, inv);
    bf8(n1, inv);
    bf8(n3, inv);
    for (int i = 1; i < 8; ++i) {
        n1[i] = n1[i] * tw[i - 1];
        n3[i] = n3[i] * conj(tw[i - 1]);
    }
    for (int i = 0; i < 8; ++i) bf2(n1[i], n3[i]);
    for (int i = 0; i < 8; ++i) n3[i] = rot90(n3[i], inv);
    for (int i = 0; i < 8; ++i) {
        v[i] = ev[i] + n1[i];
        v[i + 8] = ev[i + 8] + n3[i];
        v[i + 16] = ev[i] - n1[i];
        v[i + 24] = ev[i + 8] - n3[i];
    }
}

// Butterfly9, :780-841: 3x3 mixed radix (rows r hold inputs r + 3 i)
template <class T> inline void bf9(Cx<T>* v, bool inv) {
    const Cx<T> t3 = twiddle<T>(1, 3, inv), w1 = twiddle<T>(1, 9, inv), w2 = twiddle<T>(2, 9, inv), w4 = twiddle<T>(4, 9, inv);
    Cx<T> s[3][3];
    for (int r = 0; r < 3; ++r)
        for (int i = 0; i < 3; ++i) s[r][i] = v[r + 3 * i];
    for (int r = 0; r < 3; ++r) bf3(s[r], t3);
    s[1][1] = s[1][1] * w1;
    s[1][2] = s[1][2] * w2;
    s[2][1] = s[2][1] * w2;
    s[2][2] = s[2][2] * w4;
    for (int c = 0; c < 3; ++c) {
        Cx<T> col[3] = {s[0][c], s[1][c], s[2][c]};
        bf3(col, t3);
        for (int r = 0; r < 3; ++r) v[3 * r + c] = col[r];
    }
}
// Butterfly12, :1091-1166: 4x3 Good-Thomas with hard-coded input / output orders, no twiddles
template <class T> inline void bf12(Cx<T>* v, bool inv) {
    static const int in_idx[3][4] = {{0, 3, 6, 9}, {4, 7, 10, 1}, {8, 11, 2, 5}};
    static const int out_row[12] = {0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2};
    static const int out_col[12] = {0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3};
    const Cx<T> t3 = twiddle<T>(1, 3, inv);
    Cx<T> s[3][4];
    for (int r = 0; r < 3; ++r)
        for (int i = 0; i < 4; ++i) s[r][i] = v[in_idx[r][i]];
    for (int r = 0; r < 3; ++r) bf4(s[r], inv);
    for (int c = 0; c < 4; ++c) {
        Cx<T> col[3] = {s[0][c], s[1][c], s[2][c]};
        bf3(col, t3);
        for (int r = 0; r < 3; ++r) s[r][c] = col[r];
    }
    for (int k = 0; k < 12; ++k) v[k] = s[out_row[k]][out_col[k]];
}
// Butterfly24, :3427-3587: 6x4 mixed radix (rows r hold inputs r + 4 i), the multiples of 1/8 turn done with rotate_90 / root2
template <class T> inline void bf24(Cx<T>* v, bool inv) {
    const Cx<T> t3 = twiddle<T>(1, 3, inv);
    const Cx<T> w1 = twiddle<T>(1, 24, inv), w2 = twiddle<T>(2, 24, inv), w4 = twiddle<T>(4, 24, inv), w5 = twiddle<T>(5, 24, inv),
                w8 = twiddle<T>(8, 24, inv), w10 = twiddle<T>(10, 24, inv);
    const T root2 = (T)std::sqrt(0.5);
    Cx<T> s[4][6];
    for (int r = 0; r < 4; ++r)
        for (int i = 0; i < 6; ++i) s[r][i] = v[r + 4 * i];
    for (int r = 0; r < 4; ++r) bf6(s[r], t3);
    s[1][1] = s[1][1] * w1;
    s[1][2] = s[1][2] * w2;
    s[1][3] = scale(rot90(s[1][3], inv) + s[1][3], root2);
    s[1][4] = s[1][4] * w4;
    s[1][5] = s[1][5] * w5;
    s[2][1] = s[2][1] * w2;
    s[2][2] = s[2][2] * w4;
    s[2][3] = rot90(s[2][3], inv);
    s[2][4] = s[2][4] * w8;
    s[2][5] = s[2][5] * w10;
    s[3][1] = scale(rot90(s[3][1], inv) + s[3][1], root2);
    s[3][2] = rot90(s[3][2], inv);
    s[3][3] = scale(rot90(s[3][3], inv) - s[3][3], root2);
    s[3][4] = Cx<T>{-s[3][4].re, -s[3][4].im};
    s[3][5] = scale(rot90(s[3][5], inv) + s[3][5], -root2);
    for (int c = 0; c < 6; ++c) {
        Cx<T> col[4] = {s[0][c], s[1][c], s[2][c], s[3][c]};
        bf4(col, inv);
        for (int r = 0; r < 4; ++r) v[6 * r + c] = col[r];
    }
}
// Butterfly27, :3588-3760: 9x3 mixed radix (rows r hold inputs r + 3 i)
template <class T> inline void bf27(Cx<T>* v, bool inv) {
    const Cx<T> t3 = twiddle<T>(1, 3, inv);
    Cx<T> s[3][9];
    for (int r = 0; r < 3; ++r)
        for (int i = 0; i < 9; ++i) s[r][i] = v[r + 3 * i];
    for (int r = 0; r < 3; ++r) bf9(s[r], inv);
    for (int c = 1; c < 9; ++c) {
        s[1][c] = s[1][c] * twiddle<T>((usize)c, 27, inv);
        s[2][c] = s[2][c] * twiddle<T>((usize)(2 * c), 27, inv);
    }
    for (int c = 0; c < 9; ++c) {
        Cx<T> col[3] = {s[0][c], s[1][c], s[2][c]};
        bf3(col, t3);
        for (int r = 0; r < 3; ++r) v[9 * r + c] = col[r];
    }
}
// Butterfly11/13/17/19/23/29/31 (:842-6242, generated): with h = (p-1)/2, x_jp = x[j] + x[p-j], x_jn = x[j] - x[p-j],
//   out[0] = x0 + x_1p + x_2p + ...            (left to right)
//   a_re = x0.re + sum_j tw(kj).re * x_jp.re,  b_re = sum_j +-tw(kj).im * x_jn.im     (j = 1..h, left to right)
//   a_im = x0.im + sum_j tw(kj).re * x_jp.im,  b_im = sum_j +-tw(kj).im * x_jn.re
//   out[k] = (a_re - b_re, a_im + b_im),  out[p-k] = (a_re + b_re, a_im - b_im)
// where tw(m) for m = kj mod p > h is the stored twiddle(p - m) with its imaginary part negated.
template <class T> inline void bf_prime(Cx<T>* v, usize p, bool inv) {
    const usize h = (p - 1) / 2;
    Cx<T> tw[16], xp[16], xn[16];
    for (usize j = 1; j <= h; ++j) {
        tw[j] = twiddle<T>(j, p, inv);
        xp[j] = v[j] + v[p - j];
        xn[j] = v[j] - v[p - j];
    }
    Cx<T> sum = v[0];
    for (usize j = 1; j <= h; ++j) sum = sum + xp[j];
    Cx<T> out[32];
    out[0] = sum;
    for (usize k = 1; k <= h; ++k) {
        T a_re = v[0].re, a_im = v[0].im, b_re = 0, b_im = 0;
        for (usize j = 1; j <= h; ++j) {
            usize m = (k * j) % p;
            const bool neg = m > h;
            if (neg) m = p - m;
            const T c = tw[m].re, sn = neg ? -tw[m].im : tw[m].im;
            a_re = a_re + c * xp[j].re;
            a_im = a_im + c * xp[j].im;
            if (j == 1) {
                b_re = sn * xn[j].im;
                b_im = sn * xn[j].re;
            } else {
                b_re = b_re + sn * xn[j].im;
                b_im = b_im + sn * xn[j].re;
            }
        }
        out[k] = Cx<T>{a_re - b_re, a_im + b_im};
        out[p - k] = Cx<T>{a_re + b_re, a_im - b_im};
    }
    for (usize k = 0; k < p; ++k) v[k] = out[k];
}

template <class T>
struct LeafNode : Node<T> {  // Butterfly1/2/3/4/5/6/7/8/16/32 + 9/12/24/27 + the prime ones 11..31
    Cx<T> t[3];
    LeafNode(usize len, bool inv) : Node<T>(len, inv) {
        if (len == 3 || len == 6) t[0] = twiddle<T>(1, 3, inv);
        if (len == 5) {
            t[0] = twiddle<T>(1, 5, inv);
            t[1] = twiddle<T>(2, 5, inv);
        }
        if (len == 7)
            for (int i = 0; i < 3; ++i) t[i] = twiddle<T>(i + 1, 7, inv);
    }
    static bool supported(usize n) {
        return n == 1 || n == 2 || n == 3 || n == 4 || n == 5 || n == 6 || n == 7 || n == 8 || n == 16 || n == 32 || n == 9 ||
               n == 12 || n == 24 || n == 27 || n == 11 || n == 13 || n == 17 || n == 19 || n == 23 || n == 29 || n == 31;
    }
    void run(Cx<T>* x) override {
        switch (this->len) {
            case 1: break;
            case 2: bf2(x[0], x[1]); break;
            case 3: bf3(x, t[0]); break;
            case 4: bf4(x, this->inv); break;
            case 5: bf5(x, t[0], t[1]); break;
            case 6: bf6(x, t[0]); break;
            case 7: bf7(x, t[0], t[1], t[2]); break;
            case 8: bf8(x, this->inv); break;
            case 16: bf16(x, this->inv); break;
            case 32: bf32(x, this->inv); break;
            case 9: bf9(x, this->inv); break;
            case 12: bf12(x, this->inv); break;
            case 24: bf24(x, this->inv); break;
            case 27: bf27(x, this->inv); break;
            case 11: case 13: case 17: case 19: case 23: case 29: case 31: bf_prime(x, this->len, this->inv); break;
        }
    }
    std::string describe() const override { return "Butterfly" + std::to_string(this->len); }
};

// src/algorithm/radixn.rs:54-155,250-333 (+ butterfly_2..7 :337-490) and, with all factors == 4,
// src/algorithm/radix4.rs:69-119,167-203.  Input permutation: src/array_utils.rs:372-437 (bit
// reversed transpose) / :469-558 (factor_transpose + reverse_remainders) -- identical for radix 4.
template <class T>
struct RadixNNode : Node<T> {
    std::vector<int> factors;  // in butterfly (application) order
    NodeP<T> base;
    std::vector<Cx<T>> tw;
    std::vector<Cx<T>> tmp;
    std::vector<usize> perm;  // output slot chunk index for each input column x
    bool is_radix4;
    Cx<T> t3, t5a, t5b, t7a, t7b, t7c;

    RadixNNode(const std::vector<int>& f, NodeP<T> b, bool radix4)
        : Node<T>(b->len, b->inv), factors(f), base(b), is_radix4(radix4) {
        usize cross = base->len;
        for (int r : factors) {
            usize cols = cross;
            cross *= (usize)r;
            for (usize i = 0; i < cols; ++i)
                for (int k = 1; k < r; ++k) tw.push_back(twiddle<T>(i * (usize)k, cross, this->inv));
        }
        this->len = cross;
        tmp.resize(cross);
        const usize width = cross / base->len;
        perm.resize(width);
        // reverse_remainders over the factor list reversed (array_utils.rs:514-558, radixn.rs:85-104)
        for (usize x = 0; x < width; ++x) {
            usize v = x, res = 0;
            for (usize i = factors.size(); i-- > 0;) {
                res = res * (usize)factors[i] + v % (usize)factors[i];
                v /= (usize)factors[i];
            }
            perm[x] = res;
        }
        t3 = twiddle<T>(1, 3, this->inv);
        t5a = twiddle<T>(1, 5, this->inv);
        t5b = twiddle<T>(2, 5, this->inv);
        t7a = twiddle<T>(1, 7, this->inv);
        t7b = twiddle<T>(2, 7, this->inv);
        t7c = twiddle<T>(3, 7, this->inv);
    }
    void run(Cx<T>* x) override {
        const usize n = this->len, h = base->len, w = n / h;
        if (w > 1) {
            for (usize xx = 0; xx < w; ++xx)
                for (usize y = 0; y < h; ++y) tmp[y + perm[xx] * h] = x[xx + y * w];
        } else {
            for (usize i = 0; i < n; ++i) tmp[i] = x[i];
        }
        for (usize c = 0; c < w; ++c) base->run(&tmp[c * h]);
        usize cross = h;
        const Cx<T>* lt = tw.data();
        for (int r : factors) {
            const usize cols = cross;
            cross *= (usize)r;
            for (usize blk = 0; blk < n; blk += cross) {
                Cx<T>* d = &tmp[blk];
                for (usize idx = 0; idx < cols; ++idx) {
                    Cx<T> s[7];
                    s[0] = d[idx];
                    for (int k = 1; k < r; ++k) s[k] = d[idx + (usize)k * cols] * lt[idx * (usize)(r - 1) + (usize)(k - 1)];
                    switch (r) {
                        case 2: bf2(s[0], s[1]); break;
                        case 3: bf3(s, t3); break;
                        case 4: bf4(s, this->inv); break;
                        case 5: bf5(s, t5a, t5b); break;
                        case 6: bf6(s, t3); break;
                        case 7: bf7(s, t7a, t7b, t7c); break;
                    }
                    for (int k = 0; k < r; ++k) d[idx + (usize)k * cols] = s[k];
                }
            }
            lt += cols * (usize)(r - 1);
        }
        for (usize i = 0; i < n; ++i) x[i] = tmp[i];
    }
    std::string describe() const override {
        if (is_radix4) return "Radix4{k=" + std::to_string(factors.size()) + ",base=" + base->describe() + "}";
        std::string s = "RadixN{[";
        for (usize i = 0; i < factors.size(); ++i) s += (i ? "," : "") + std::to_string(factors[i]);
        return s + "],base=" + base->describe() + "}";
    }
};

// src/algorithm/mixed_radix.rs:53-126,128-158 (and MixedRadixSmall :266-398: same arithmetic)
template <class T>
struct MixedRadixNode : Node<T> {
    NodeP<T> wfft, hfft;
    usize w, h;
    bool small;
    std::vector<Cx<T>> tw, tmp;
    MixedRadixNode(NodeP<T> width_fft, NodeP<T> height_fft, bool small_)
        : Node<T>(width_fft->len * height_fft->len, width_fft->inv), wfft(width_fft), hfft(height_fft),
          w(width_fft->len), h(height_fft->len), small(small_), tw(w * h), tmp(w * h) {
        for (usize x = 0; x < w; ++x)
            for (usize y = 0; y < h; ++y) tw[x * h + y] = twiddle<T>(x * y, w * h, this->inv);
    }
    void run(Cx<T>* b) override {
        const usize n = this->len;
        for (usize x = 0; x < w; ++x)  // transpose(buffer -> scratch, width, height)
            for (usize y = 0; y < h; ++y) tmp[x * h + y] = b[y * w + x];
        for (usize r = 0; r < w; ++r) hfft->run(&tmp[r * h]);
        for (usize i = 0; i < n; ++i) tmp[i] = tmp[i] * tw[i];
        for (usize y = 0; y < h; ++y)  // transpose(scratch -> buffer, height, width)
            for (usize x = 0; x < w; ++x) b[y * w + x] = tmp[x * h + y];
        for (usize r = 0; r < h; ++r) wfft->run(&b[r * w]);
        for (usize x = 0; x < w; ++x)  // transpose(-> output, width, height)
            for (usize y = 0; y < h; ++y) tmp[x * h + y] = b[y * w + x];
        for (usize i = 0; i < n; ++i) b[i] = tmp[i];
    }
    std::string describe() const override {
        return std::string(small ? "MixedRadixSmall{" : "MixedRadix{") + wfft->describe() + "," + hfft->describe() + "}";
    }
};

// src/algorithm/good_thomas_algorithm.rs:360-470 (GoodThomasAlgorithmSmall; the only PFA variant the
// scalar planner emits, src/plan.rs:427-506)
template <class T>
struct GoodThomasSmallNode : Node<T> {
    NodeP<T> wfft, hfft;
    usize w, h;
    std::vector<usize> in_map, out_map;
    std::vector<Cx<T>> a, b;
    GoodThomasSmallNode(NodeP<T> width_fft, NodeP<T> height_fft)
        : Node<T>(width_fft->len * height_fft->len, width_fft->inv), wfft(width_fft), hfft(height_fft),
          w(width_fft->len), h(height_fft->len) {
        const usize n = w * h;
        const usize winv = (usize)mod_inverse((int64_t)(w % h), (int64_t)h);  // w^-1 mod h
        const usize hinv = (usize)mod_inverse((int64_t)(h % w), (int64_t)w);  // h^-1 mod w
        in_map.resize(n);
        out_map.resize(n);
        for (usize i = 0; i < n; ++i) {
            usize x = i % w, y = i / w;
            in_map[i] = (x * h + y * w) % n;
        }
        for (usize i = 0; i < n; ++i) {
            usize y = i % h, x = i / h;
            out_map[i] = (x * h * hinv + y * w * winv) % n;
        }
        a.resize(n);
        b.resize(n);
    }
    void run(Cx<T>* buf) override {
        const usize n = this->len;
        for (usize i = 0; i < n; ++i) a[i] = buf[in_map[i]];
        for (usize r = 0; r < h; ++r) wfft->run(&a[r * w]);
        for (usize x = 0; x < w; ++x)
            for (usize y = 0; y < h; ++y) b[x * h + y] = a[y * w + x];
        for (usize r = 0; r < w; ++r) hfft->run(&b[r * h]);
        for (usize i = 0; i < n; ++i) buf[out_map[i]] = b[i];
    }
    std::string describe() const override {
        return "GoodThomasAlgorithmSmall{" + wfft->describe() + "," + hfft->describe() + "}";
    }
};

// src/algorithm/raders_algorithm.rs:65-124 (setup), :235-283 (transform)
template <class T>
struct RaderNode : Node<T> {
    NodeP<T> inner;
    uint64_t g, ginv;
    std::vector<Cx<T>> spectrum, s;
    RaderNode(NodeP<T> inner_fft) : Node<T>(inner_fft->len + 1, inner_fft->inv), inner(inner_fft) {
        const uint64_t p = this->len;
        const usize m = inner->len;
        g = primitive_root(p);
        ginv = (uint64_t)mod_inverse((int64_t)g, (int64_t)p);
        const T sc = (T)1 / (T)m;
        spectrum.resize(m);
        uint64_t ti = 1;
        for (usize i = 0; i < m; ++i) {
            spectrum[i] = scale(twiddle<T>((usize)ti, (usize)p, this->inv), sc);
            ti = (ti * ginv) % p;
        }
        inner->run(spectrum.data());
        s.resize(m);
    }
    void run(Cx<T>* buf) override {
        const uint64_t p = this->len;
        const usize m = inner->len;
        const Cx<T> first = buf[0];
        uint64_t idx = 1;
        for (usize i = 0; i < m; ++i) {
            idx = (idx * g) % p;
            s[i] = buf[idx];
        }
        inner->run(s.data());
        buf[0] = buf[0] + s[0];
        for (usize i = 0; i < m; ++i) s[i] = conj(s[i] * spectrum[i]);
        s[0] = s[0] + conj(first);
        inner->run(s.data());
        idx = 1;
        for (usize i = 0; i < m; ++i) {
            idx = (idx * ginv) % p;
            buf[idx] = conj(s[i]);
        }
    }
    std::string describe() const override { return "RadersAlgorithm{" + inner->describe() + "}"; }
};

// src/twiddles.rs:25-57
template <class T>
void fill_bluestein_twiddles(Cx<T>* dst, usize len, bool inv) {
    const unsigned __int128 twice = (unsigned __int128)len * 2;
    for (usize i = 0; i < len; ++i) {
        unsigned __int128 sq = (unsigned __int128)i * (unsigned __int128)i;
        dst[i] = twiddle<T>((usize)(sq % twice), len * 2, inv);
    }
}

// src/algorithm/bluesteins_algorithm.rs:58-98 (setup), :100-136 (transform)
template <class T>
struct BluesteinNode : Node<T> {
    NodeP<T> inner;
    std::vector<Cx<T>> mult, tw, a;
    BluesteinNode(usize len, NodeP<T> inner_fft) : Node<T>(len, inner_fft->inv), inner(inner_fft) {
        const usize m = inner->len;
        const T sc = (T)1 / (T)m;
        mult.assign(m, Cx<T>{0, 0});
        fill_bluestein_twiddles(mult.data(), len, !this->inv);
        mult[0] = scale(mult[0], sc);
        for (usize i = 1; i < len; ++i) {
            Cx<T> t = scale(mult[i], sc);
            mult[i] = t;
            mult[m - i] = t;
        }
        inner->run(mult.data());
        tw.resize(len);
        fill_bluestein_twiddles(tw.data(), len, this->inv);
        a.resize(m);
    }
    void run(Cx<T>* x) override {
        const usize n = this->len, m = inner->len;
        for (usize i = 0; i < n; ++i) a[i] = x[i] * tw[i];
        for (usize i = n; i < m; ++i) a[i] = Cx<T>{0, 0};
        inner->run(a.data());
        for (usize i = 0; i < m; ++i) a[i] = conj(a[i] * mult[i]);
        inner->run(a.data());
        for (usize i = 0; i < n; ++i) x[i] = conj(a[i]) * tw[i];
    }
    std::string describe() const override {
        return "BluesteinsAlgorithm{len=" + std::to_string(this->len) + ",inner=" + inner->describe() + "}";
    }
};

// ----------------------------------------------------------------------------------------
// FftPlannerScalar: src/plan.rs:270-666 (+ FftCache, src/fft_cache.rs: one instance per
// (len, direction)).
// ----------------------------------------------------------------------------------------
template <class T>
struct ScalarPlanner {
    std::map<usize, NodeP<T>> cache[2];

    static bool is_butterfly_len(usize n) {  // plan.rs:609-634
        static const usize b[] = {2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 16, 17, 19, 23, 24, 27, 29, 31, 32};
        for (usize v : b)
            if (v == n) return true;
        return false;
    }
    NodeP<T> leaf(usize n, bool inv) {
        if (LeafNode<T>::supported(n)) return std::make_shared<LeafNode<T>>(n, inv);
        auto d = std::make_shared<DftNode<T>>(n, inv);  // stated deviation, see header
        d->stands_for_butterfly = true;
        return d;
    }
    NodeP<T> plan(usize len, bool inv) {  // design_fft_for_len + build_fft, plan.rs:312-335
        if (len < 2) return std::make_shared<DftNode<T>>(len, inv);
        auto it = cache[inv].find(len);
        if (it != cache[inv].end()) return it->second;
        NodeP<T> n = with_factors(len, PrimeFactors::compute(len), inv);
        cache[inv][len] = n;
        return n;
    }
    NodeP<T> cached(usize len, bool inv, NodeP<T> fresh) {  // FftCache semantics for inner nodes
        auto it = cache[inv].find(len);
        if (it != cache[inv].end()) return it->second;
        cache[inv][len] = fresh;
        return fresh;
    }
    NodeP<T> with_factors(usize len, const PrimeFactors& f, bool inv) {  // plan.rs:412-425
        if (is_butterfly_len(len)) return leaf(len, inv);
        if (f.is_prime()) return prime(len, inv);
        if (NodeP<T> bp = butterfly_product(len, inv)) return bp;
        if (f.has_factors_leq(7)) return radixn(f, inv);
        auto lr = f.partition();
        return mixed_radix(lr.first, lr.second, inv);
    }
    NodeP<T> butterfly_product(usize len, bool inv) {  // plan.rs:427-472
        if (len > 992 || (len & (len - 1)) == 0) return nullptr;
        const usize limit = (usize)std::ceil(std::sqrt((double)len)) + 1;
        static const usize bs[] = {2, 3, 4, 5, 6, 7, 8, 9, 11, 13, 16, 17, 19, 23, 24, 27, 29, 31, 32};
        usize min_sum = (usize)-1, bl = 0, br = 0;
        for (usize left : bs) {
            if (!(left < limit)) break;
            usize right = len / left;
            bool right_ok = false;
            for (usize v : bs) right_ok = right_ok || v == right;
            if (left * right == len && right_ok) {
                usize sum = left + right;
                if (sum < min_sum) {
                    min_sum = sum;
                    bl = left;
                    br = right;
                }
            }
        }
        if (!bl) return nullptr;
        NodeP<T> l = plan(bl, inv), r = plan(br, inv);
        if (gcd_usize(bl, br) == 1) return std::make_shared<GoodThomasSmallNode<T>>(l, r);
        return std::make_shared<MixedRadixNode<T>>(l, r, true);
    }
    NodeP<T> mixed_radix(const PrimeFactors& lf, const PrimeFactors& rf, bool inv) {  // plan.rs:474-506
        const usize ll = lf.n, rl = rf.n;
        NodeP<T> l = cached(ll, inv, with_factors(ll, lf, inv));
        NodeP<T> r = cached(rl, inv, with_factors(rl, rf, inv));
        if (ll < 31 && rl < 31) {
            if (gcd_usize(ll, rl) == 1) return std::make_shared<GoodThomasSmallNode<T>>(l, r);
            return std::make_shared<MixedRadixNode<T>>(l, r, true);
        }
        return std::make_shared<MixedRadixNode<T>>(l, r, false);
    }
    NodeP<T> radixn(const PrimeFactors& f, bool inv) {  // plan.rs:508-607
        const uint32_t p2 = f.p2, p3 = f.p3;
        uint32_t p5 = 0, p7 = 0;
        for (auto& o : f.other) {
            if (o.first == 5) p5 = o.second;
            if (o.first == 7) p7 = o.second;
        }
        usize base_len;
        if (f.has_factors_gt(7))
            base_len = f.product_above(7);
        else if (p7 == 0 && p5 == 0 && p3 < 2) {
            if (p3 == 0)
                base_len = (p2 % 2 == 1) ? 8 : 16;
            else
                base_len = (p2 % 2 == 1) ? 24 : 12;
        } else if (p2 > 0 && p3 > 0) {
            uint32_t excess = p2 > p3 ? p2 - p3 : 0;
            base_len = excess == 0 ? 6 : (excess == 1 ? 12 : 24);
        } else if (p3 > 2)
            base_len = 27;
        else if (p3 > 1)
            base_len = 9;
        else if (p7 > 0)
            base_len = 7;
        else
            base_len = 5;
        NodeP<T> base = plan(base_len, inv);
        usize cross = f.n / base_len;
        uint32_t cross_bits = 0;
        while (cross_bits < 63 && ((cross >> cross_bits) & 1) == 0 && (cross >> cross_bits) != 0) cross_bits++;
        const bool pow2 = cross != 0 && (cross & (cross - 1)) == 0;
        if (pow2 && cross_bits % 2 == 0)
            return std::make_shared<RadixNNode<T>>(std::vector<int>(cross_bits / 2, 4), base, true);
        std::vector<int> fac;
        while (cross % 7 == 0) { cross /= 7; fac.push_back(7); }
        while (cross % 6 == 0) { cross /= 6; fac.push_back(6); }
        while (cross % 5 == 0) { cross /= 5; fac.push_back(5); }
        while (cross % 3 == 0) { cross /= 3; fac.push_back(3); }
        uint32_t bits = 0;
        while ((cross >> bits) > 1) bits++;
        if (bits % 2 == 1) fac.push_back(2);
        for (uint32_t i = 0; i < bits / 2; ++i) fac.push_back(4);
        return std::make_shared<RadixNNode<T>>(fac, base, false);
    }
    NodeP<T> prime(usize len, bool inv) {  // plan.rs:636-665
        const usize inner_len = len - 1;
        PrimeFactors rf = PrimeFactors::compute(inner_len);
        bool big = false;
        for (auto& o : rf.other) big = big || o.first > 23;
        if (big) {
            const usize min_inner = 2 * len - 1;
            usize pow2 = 1;
            while (pow2 < min_inner) pow2 <<= 1;
            const usize f3 = pow2 / 4 * 3;
            const usize m = f3 >= min_inner ? f3 : pow2;
            return std::make_shared<BluesteinNode<T>>(len, plan(m, inv));
        }
        NodeP<T> inner = cached(inner_len, inv, with_factors(inner_len, rf, inv));
        return std::make_shared<RaderNode<T>>(inner);
    }
};

// Radix4::new(len, direction): src/algorithm/radix4.rs:42-66
template <class T>
NodeP<T> make_radix4(usize len, bool inv) {
    uint32_t e = 0;
    while (((usize)1 << e) < len) e++;
    uint32_t be = e <= 3 ? e : (e % 2 == 1 ? 5 : 4);
    NodeP<T> base = std::make_shared<LeafNode<T>>((usize)1 << be, inv);
    return std::make_shared<RadixNNode<T>>(std::vector<int>((e - be) / 2, 4), base, true);
}

// tests/accuracy.rs:98-122 -- the acceptance test's control
template <class T>
NodeP<T> make_control(usize len, bool inv) {
    usize m = 1;
    while (m < 2 * len - 1) m <<= 1;
    return std::make_shared<BluesteinNode<T>>(len, make_radix4<T>(m, inv));
}

enum Kind { K_PLANNER = 0, K_CONTROL = 1, K_DFT = 2, K_RADIX4 = 3 };

template <class T>
NodeP<T> build(int kind, usize len, bool inv) {
    switch (kind) {
        case K_PLANNER: {
            ScalarPlanner<T> p;
            return p.plan(len, inv);
        }
        case K_CONTROL: return make_control<T>(len, inv);
        case K_DFT: return std::make_shared<DftNode<T>>(len, inv);
        case K_RADIX4: return make_radix4<T>(len, inv);
    }
    return nullptr;
}

// batch = serial loop over contiguous chunks (src/array_utils.rs:151-177); with nthreads > 1, one
// contiguous slice of the batch per thread, each with its own plan instance (the reference's Fft
// objects are Sync and shared, examples/concurrency.rs:17-29; separate instances here only because
// these nodes carry their scratch).
template <class T>
int run_batch(int kind, usize len, int inverse, T* data, usize batch, int nthreads) {
    if (len == 0) return 0;  // src/fft_helper.rs:16-18
    if (kind == K_RADIX4 && (len & (len - 1)) != 0) return -2;
    if (nthreads < 1) nthreads = 1;
    if ((usize)nthreads > batch) nthreads = batch ? (int)batch : 1;
    Cx<T>* x = reinterpret_cast<Cx<T>*>(data);
    if (nthreads == 1) {
        NodeP<T> n = build<T>(kind, len, inverse != 0);
        if (!n) return -1;
        for (usize b = 0; b < batch; ++b) n->run(x + b * len);
        return 0;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        usize b0 = batch * (usize)t / (usize)nthreads, b1 = batch * (usize)(t + 1) / (usize)nthreads;
        th.emplace_back([=]() {
            NodeP<T> n = build<T>(kind, len, inverse != 0);
            for (usize b = b0; b < b1; ++b) n->run(x + b * len);
        });
    }
    for (auto& t : th) t.join();
    return 0;
}

}  // namespace

extern "C" {

// kind: 0 = FftPlannerScalar plan, 1 = tests/accuracy.rs control (Bluestein over Radix4::new),
//       2 = naive Dft, 3 = Radix4::new(len).  data = interleaved {re,im}, batch*len complex, in place.
int oracle_fft_f32(int kind, uint64_t len, int inverse, float* data, uint64_t batch, int nthreads) {
    return run_batch<float>(kind, (usize)len, inverse, data, (usize)batch, nthreads);
}
int oracle_fft_f64(int kind, uint64_t len, int inverse, double* data, uint64_t batch, int nthreads) {
    return run_batch<double>(kind, (usize)len, inverse, data, (usize)batch, nthreads);
}

// Plan + time only the transform loop (plan built outside the timed region, as
// benches/bench_rustfft.rs:43-54 does).  Returns seconds, <0 on error.
double oracle_time_f32(int kind, uint64_t len, int inverse, float* data, uint64_t batch, int nthreads, int reps);

// Recipe string of the scalar planner for `len` (plan.rs unit tests :700-858 check these shapes).
int oracle_describe_plan(uint64_t len, char* out, uint64_t cap) {
    ScalarPlanner<double> p;
    std::string s = p.plan((usize)len, false)->describe();
    if (s.size() + 1 > cap) return -1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

uint64_t oracle_modular_exponent(uint64_t b, uint64_t e, uint64_t m) { return modular_exponent(b, e, m); }
uint64_t oracle_primitive_root(uint64_t p) { return primitive_root(p); }
int oracle_distinct_prime_factors(uint64_t n, uint64_t* out, int cap) {
    auto v = distinct_prime_factors(n);
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
// PrimeFactors::compute summary: out = {p2, p3, total, distinct, n_other, (value,count)...}
int oracle_prime_factors(uint64_t n, uint64_t* out, int cap) {
    PrimeFactors f = PrimeFactors::compute((usize)n);
    std::vector<uint64_t> v = {f.p2, f.p3, f.total, f.distinct, (uint64_t)f.other.size()};
    for (auto& o : f.other) {
        v.push_back(o.first);
        v.push_back(o.second);
    }
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
int oracle_partition_factors(uint64_t n, uint64_t* left, uint64_t* right) {
    PrimeFactors f = PrimeFactors::compute((usize)n);
    if (f.is_prime() || n < 2) return -1;
    auto lr = f.partition();
    *left = lr.first.n;
    *right = lr.second.n;
    return 0;
}
void oracle_twiddle_f64(uint64_t idx, uint64_t len, int inverse, double* out) {
    Cx<double> t = twiddle<double>((usize)idx, (usize)len, inverse != 0);
    out[0] = t.re;
    out[1] = t.im;
}
void oracle_twiddle_f32(uint64_t idx, uint64_t len, int inverse, float* out) {
    Cx<float> t = twiddle<float>((usize)idx, (usize)len, inverse != 0);
    out[0] = t.re;
    out[1] = t.im;
}
int oracle_is_prime(uint64_t n) { return is_prime_u64(n) ? 1 : 0; }

}  // extern "C"

#include <chrono>
extern "C" double oracle_time_f32(int kind, uint64_t len, int inverse, float* data, uint64_t batch, int nthreads,
                                  int reps) {
    if (len == 0 || batch == 0) return -1.0;
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > batch) nthreads = (int)batch;
    std::vector<NodeP<float>> plans;
    for (int t = 0; t < nthreads; ++t) plans.push_back(build<float>(kind, (usize)len, inverse != 0));
    Cx<float>* x = reinterpret_cast<Cx<float>*>(data);
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) {
            usize b0 = (usize)batch * (usize)t / (usize)nthreads, b1 = (usize)batch * (usize)(t + 1) / (usize)nthreads;
            NodeP<float> n = plans[t];
            th.emplace_back([=]() {
                for (usize b = b0; b < b1; ++b) n->run(x + b * (usize)len);
            });
        }
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- honest multi-core timing (bench.py cpu_baseline / --impl reference) --------------------------------------------
// benches/bench_rustfft.rs:43-54 times `fft.process_with_scratch` on a prepared plan; examples/concurrency.rs:17-29 shares one plan between
// threads, each with its own buffer.  Here: `nthreads` workers are created ONCE per call, pinned to distinct CPUs of the allowed set,
// first-touch and fill their own contiguous slice of the batch (so pages live on the worker's NUMA node), and then run `reps` timed
// passes; every pass is bracketed by a spin barrier and timed separately (times_out[r], seconds).  Between passes (untimed) the slices
// are rescaled so repeated in-place unnormalised transforms stay finite.
#include <atomic>
#include <sched.h>
namespace {
struct SpinBarrier {
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    int n;
    explicit SpinBarrier(int n_) : n(n_) {}
    void wait(int& local) {
        local ^= 1;
        if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            count.store(0, std::memory_order_relaxed);
            sense.store(local, std::memory_order_release);
        } else {
            while (sense.load(std::memory_order_acquire) != local) __builtin_ia32_pause();
        }
    }
};
}  // namespace
extern "C" int oracle_bench_f32(int kind, uint64_t len, uint64_t batch, int nthreads, int reps, float* data, double* times_out) {
    if (len == 0 || batch == 0 || reps < 1 || !data || !times_out) return -1;
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > batch) nthreads = (int)batch;
    std::vector<int> cpus;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
    Cx<float>* x = reinterpret_cast<Cx<float>*>(data);
    // only the workers take part in the barriers (a spinning main thread would steal a pinned worker's CPU); every worker stamps
    // its own start and end of a pass, pass time = last end - first start
    SpinBarrier bar(nthreads);
    typedef std::chrono::steady_clock clk;
    std::vector<clk::time_point> t_start((size_t)nthreads * reps), t_end((size_t)nthreads * reps);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t]() {
            if (!cpus.empty()) {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpus[(size_t)t % cpus.size()], &one);
                sched_setaffinity(0, sizeof(one), &one);
            }
            const usize b0 = (usize)batch * (usize)t / (usize)nthreads, b1 = (usize)batch * (usize)(t + 1) / (usize)nthreads;
            NodeP<float> plan = build<float>(kind, (usize)len, false);
            uint32_t s = 0x9E3779B9u * (uint32_t)(t + 1);
            for (usize i = b0 * (usize)len; i < b1 * (usize)len; ++i) {  // first touch + fill: U[0, 10)
                s = s * 1664525u + 1013904223u;
                const float re = (float)(s >> 8) * (10.0f / 16777216.0f);
                s = s * 1664525u + 1013904223u;
                x[i] = Cx<float>{re, (float)(s >> 8) * (10.0f / 16777216.0f)};
            }
            const float scale = 1.0f / std::sqrt((float)len);
            int local = 0;
            for (int r = 0; r < reps; ++r) {
                if (r > 0)
                    for (usize i = b0 * (usize)len; i < b1 * (usize)len; ++i) x[i] = Cx<float>{x[i].re * scale, x[i].im * scale};
                bar.wait(local);  // start of the timed pass
                t_start[(size_t)r * nthreads + t] = clk::now();
                for (usize b = b0; b < b1; ++b) plan->run(x + b * (usize)len);
                t_end[(size_t)r * nthreads + t] = clk::now();
                bar.wait(local);  // end of the timed pass
            }
        });
    }
    for (auto& t : th) t.join();
    for (int r = 0; r < reps; ++r) {
        clk::time_point a = t_start[(size_t)r * nthreads], b = t_end[(size_t)r * nthreads];
        for (int t = 1; t < nthreads; ++t) {
            if (t_start[(size_t)r * nthreads + t] < a) a = t_start[(size_t)r * nthreads + t];
            if (t_end[(size_t)r * nthreads + t] > b) b = t_end[(size_t)r * nthreads + t];
        }
        times_out[r] = std::chrono::duration<double>(b - a).count();
    }
    return 0;
}
