// In-register radix-R DFTs (forward sign, natural order in, natural order out).
// These are the per-thread butterflies of the Stockham stages; the algorithm family is the
// reference's (radix 2/4/8/16 for powers of two -- cf. src/algorithm/butterflies.rs Butterfly2/4/8/16
// and radixn.rs butterfly_2..7 for 3/5/7) but the formulation is a plain DIT factorisation written
// for FMA hardware, not a restatement of the reference's split-radix code.
#pragma once
#include "common.h"
#include "prime_constants.h"

namespace b2 {

template <typename T> struct K {
    static constexpr T rsqrt2 = (T)0.70710678118654752440084436210484903928L;
    static constexpr T c16_1 = (T)0.92387953251128675612818318939678828682L;  // cos(pi/8)
    static constexpr T s16_1 = (T)0.38268343236508977172845998403039886676L;  // sin(pi/8)
    static constexpr T c32_1 = (T)0.98078528040323044912618223613423903697L;  // cos(pi/16)
    static constexpr T s32_1 = (T)0.19509032201612826784828486847702224093L;  // sin(pi/16)
    static constexpr T c32_3 = (T)0.83146961230254523707878837761790575673L;  // cos(3pi/16)
    static constexpr T s32_3 = (T)0.55557023301960222474283081394853287438L;  // sin(3pi/16)
    // radix 3
    static constexpr T c3 = (T)-0.5L;
    static constexpr T s3 = (T)0.86602540378443864676372317075293618347L;  // sin(2pi/3)
    // radix 5
    static constexpr T c5_1 = (T)0.30901699437494742410229341718281905886L;   // cos(2pi/5)
    static constexpr T c5_2 = (T)-0.80901699437494742410229341718281905886L;  // cos(4pi/5)
    static constexpr T s5_1 = (T)0.95105651629515357211643933337938214340L;   // sin(2pi/5)
    static constexpr T s5_2 = (T)0.58778525229247312916870595463907276860L;   // sin(4pi/5)
    // radix 7
    static constexpr T c7_1 = (T)0.62348980185873353052500488400423981063L;
    static constexpr T c7_2 = (T)-0.22252093395631440428890256449679475947L;
    static constexpr T c7_3 = (T)-0.90096886790241912623610231950744505117L;
    static constexpr T s7_1 = (T)0.78183148246802980870844452667405775023L;
    static constexpr T s7_2 = (T)0.97492791218182360701813168299393121723L;
    static constexpr T s7_3 = (T)0.43388373911755812047576833284835875461L;
};

template <typename T> B2_HD void bf2(cx<T>& a, cx<T>& b) {
    cx<T> t = a + b;
    b = a - b;
    a = t;
}

// v0..v3 natural in/out
template <typename T> B2_HD void bf4(cx<T>& v0, cx<T>& v1, cx<T>& v2, cx<T>& v3) {
    cx<T> t0 = v0 + v2, t1 = v0 - v2, t2 = v1 + v3, d = v1 - v3;
    v0 = t0 + t2;
    v1 = add_mi(t1, d);  // t1 + (-i) d
    v2 = t0 - t2;
    v3 = sub_mi(t1, d);
}

template <typename T> B2_HD void bf3(cx<T>& v0, cx<T>& v1, cx<T>& v2) {
    cx<T> s = v1 + v2, d = scale(v1 - v2, K<T>::s3);
    cx<T> m = axpy(v0, K<T>::c3, s);
    v0 = v0 + s;
    v1 = add_mi(m, d);  // m + (-i) s3 d
    v2 = sub_mi(m, d);
}

template <typename T> B2_HD void bf5(cx<T>& v0, cx<T>& v1, cx<T>& v2, cx<T>& v3, cx<T>& v4) {
    cx<T> a = v1 + v4, b = v1 - v4, c = v2 + v3, d = v2 - v3;
    cx<T> m1 = mk<T>(v0.x + K<T>::c5_1 * a.x + K<T>::c5_2 * c.x, v0.y + K<T>::c5_1 * a.y + K<T>::c5_2 * c.y);
    cx<T> m2 = mk<T>(v0.x + K<T>::c5_2 * a.x + K<T>::c5_1 * c.x, v0.y + K<T>::c5_2 * a.y + K<T>::c5_1 * c.y);
    // -i * (s1*b + s2*d) and -i * (s2*b - s1*d)
    cx<T> q1 = mk<T>(K<T>::s5_1 * b.x + K<T>::s5_2 * d.x, K<T>::s5_1 * b.y + K<T>::s5_2 * d.y);
    cx<T> q2 = mk<T>(K<T>::s5_2 * b.x - K<T>::s5_1 * d.x, K<T>::s5_2 * b.y - K<T>::s5_1 * d.y);
    cx<T> r1 = mul_mi(q1), r2 = mul_mi(q2);
    v0 = v0 + a + c;
    v1 = m1 + r1;
    v4 = m1 - r1;
    v2 = m2 + r2;
    v3 = m2 - r2;
}

template <typename T>
B2_HD void bf7(cx<T>& v0, cx<T>& v1, cx<T>& v2, cx<T>& v3, cx<T>& v4, cx<T>& v5, cx<T>& v6) {
    cx<T> a1 = v1 + v6, b1 = v1 - v6, a2 = v2 + v5, b2 = v2 - v5, a3 = v3 + v4, b3 = v3 - v4;
    const T c1 = K<T>::c7_1, c2 = K<T>::c7_2, c3 = K<T>::c7_3, s1 = K<T>::s7_1, s2 = K<T>::s7_2, s3 = K<T>::s7_3;
    cx<T> m1 = mk<T>(v0.x + c1 * a1.x + c2 * a2.x + c3 * a3.x, v0.y + c1 * a1.y + c2 * a2.y + c3 * a3.y);
    cx<T> m2 = mk<T>(v0.x + c2 * a1.x + c3 * a2.x + c1 * a3.x, v0.y + c2 * a1.y + c3 * a2.y + c1 * a3.y);
    cx<T> m3 = mk<T>(v0.x + c3 * a1.x + c1 * a2.x + c2 * a3.x, v0.y + c3 * a1.y + c1 * a2.y + c2 * a3.y);
    cx<T> q1 = mk<T>(s1 * b1.x + s2 * b2.x + s3 * b3.x, s1 * b1.y + s2 * b2.y + s3 * b3.y);
    cx<T> q2 = mk<T>(s2 * b1.x - s3 * b2.x - s1 * b3.x, s2 * b1.y - s3 * b2.y - s1 * b3.y);
    cx<T> q3 = mk<T>(s3 * b1.x - s1 * b2.x + s2 * b3.x, s3 * b1.y - s1 * b2.y + s2 * b3.y);
    cx<T> r1 = mul_mi(q1), r2 = mul_mi(q2), r3 = mul_mi(q3);
    v0 = v0 + a1 + a2 + a3;
    v1 = m1 + r1;
    v6 = m1 - r1;
    v2 = m2 + r2;
    v5 = m2 - r2;
    v3 = m3 + r3;
    v4 = m3 - r3;
}

// multiply by W8^1 = (1-i)/sqrt2 = (a + (-i)a)/sqrt2   and   W8^3 = (-1-i)/sqrt2 = -(a - (-i)a)/sqrt2
template <typename T> B2_HD cx<T> mul_w8_1(cx<T> a) { return scale(add_mi(a, a), K<T>::rsqrt2); }
template <typename T> B2_HD cx<T> mul_w8_3(cx<T> a) { return scale(sub_mi(a, a), -K<T>::rsqrt2); }

template <typename T> B2_HD void bf8(cx<T> (&v)[8]) {
    // DIT: evens / odds 4-point, odd outputs twiddled by W8^k
    bf4(v[0], v[2], v[4], v[6]);
    bf4(v[1], v[3], v[5], v[7]);
    cx<T> o1 = mul_w8_1(v[3]), o2 = v[5], o3 = mul_w8_3(v[7]);
    cx<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = e0 + o0;
    v[4] = e0 - o0;
    v[1] = e1 + o1;
    v[5] = e1 - o1;
    v[2] = add_mi(e2, o2);  // W8^2 = -i
    v[6] = sub_mi(e2, o2);
    v[3] = e3 + o3;
    v[7] = e3 - o3;
}

template <typename T> B2_HD void bf16(cx<T> (&v)[16]) {
    // n = 4*n1 + n2, k = k1 + 4*k2:  4-point over n1 for each n2, twiddle W16^(n2*k1), 4-point over n2
    B2_UNROLL
    for (int n2 = 0; n2 < 4; ++n2) bf4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);  // v[4*k1 + n2]
    const cx<T> w1 = mk<T>(K<T>::c16_1, -K<T>::s16_1);  // W16^1
    const cx<T> w3 = mk<T>(K<T>::s16_1, -K<T>::c16_1);  // W16^3
    // k1 = 1: n2 = 1,2,3 -> W16^1, W16^2, W16^3
    v[5] = cmul(v[5], w1);
    v[6] = mul_w8_1(v[6]);
    v[7] = cmul(v[7], w3);
    // k1 = 2: W16^2, W16^4, W16^6
    v[9] = mul_w8_1(v[9]);
    v[10] = mul_mi(v[10]);
    v[11] = mul_w8_3(v[11]);
    // k1 = 3: W16^3, W16^6, W16^9 = -W16^1
    v[13] = cmul(v[13], w3);
    v[14] = mul_w8_3(v[14]);
    v[15] = cmul(v[15], mk<T>(-K<T>::c16_1, K<T>::s16_1));
    B2_UNROLL
    for (int k1 = 0; k1 < 4; ++k1) bf4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);  // -> k2
    // now v[4*k1 + k2] holds X[k1 + 4*k2]: transpose the 4x4 register tile to natural order
    B2_UNROLL
    for (int a = 0; a < 4; ++a) {
        B2_UNROLL
        for (int b = a + 1; b < 4; ++b) {
            cx<T> t = v[4 * a + b];
            v[4 * a + b] = v[4 * b + a];
            v[4 * b + a] = t;
        }
    }
}

// multiply by W32^m = exp(-2 pi i m / 32), m compile-time (0..31): trivial rotations and the eighth
// roots use the modifier forms, the rest one complex multiply with a literal constant
template <int M, typename T> B2_HD cx<T> mul_w32(cx<T> a) {
    constexpr int m = M & 31;
    if constexpr (m == 0) return a;
    else if constexpr (m == 8) return mul_mi(a);
    else if constexpr (m == 16) return mk<T>(-a.x, -a.y);
    else if constexpr (m == 24) return mk<T>(-a.y, a.x);
    else if constexpr (m == 4) return mul_w8_1(a);
    else if constexpr (m == 12) return mul_w8_3(a);
    else if constexpr (m == 20) return scale(add_mi(a, a), -K<T>::rsqrt2);
    else if constexpr (m == 28) return scale(sub_mi(a, a), K<T>::rsqrt2);
    else {
        // cos/sin(2 pi m / 32) from the first-octant constants
        constexpr int o = m & 7;  // position inside the octant pair
        constexpr int q = m >> 3; // quadrant
        // angle = q*90deg + o*11.25deg ; cos/sin of o*11.25: o=1 (c32_1,s32_1) 2 (c16_1,s16_1) 3 (c32_3,s32_3)
        // 5 -> (s32_3,c32_3) 6 -> (s16_1,c16_1) 7 -> (s32_1,c32_1)
        constexpr T c0 = (o == 1) ? K<T>::c32_1 : (o == 2) ? K<T>::c16_1 : (o == 3) ? K<T>::c32_3
                       : (o == 5) ? K<T>::s32_3 : (o == 6) ? K<T>::s16_1 : K<T>::s32_1;
        constexpr T s0 = (o == 1) ? K<T>::s32_1 : (o == 2) ? K<T>::s16_1 : (o == 3) ? K<T>::s32_3
                       : (o == 5) ? K<T>::c32_3 : (o == 6) ? K<T>::c16_1 : K<T>::c32_1;
        // rotate by quadrant: exp(-i(theta0 + q pi/2)) -> (cos, -sin)
        constexpr T c = (q == 0) ? c0 : (q == 1) ? -s0 : (q == 2) ? -c0 : s0;
        constexpr T sn = (q == 0) ? s0 : (q == 1) ? c0 : (q == 2) ? -s0 : -c0;
        return cmul(a, mk<T>(c, -sn));
    }
}

// radix 32 = 4 x 8:  n = 8*n1 + n2, k = k1 + 4*k2.  4-point over n1 for each n2, twiddle W32^(n2*k1),
// 8-point over n2 for each k1.
template <typename T> B2_HD void bf32(cx<T> (&v)[32]) {
    B2_UNROLL
    for (int n2 = 0; n2 < 8; ++n2) bf4(v[n2], v[8 + n2], v[16 + n2], v[24 + n2]);  // -> v[8*k1 + n2]
#define B2_TW32(k1, n2) v[8 * k1 + n2] = mul_w32<k1 * n2>(v[8 * k1 + n2]);
    B2_TW32(1, 1) B2_TW32(1, 2) B2_TW32(1, 3) B2_TW32(1, 4) B2_TW32(1, 5) B2_TW32(1, 6) B2_TW32(1, 7)
    B2_TW32(2, 1) B2_TW32(2, 2) B2_TW32(2, 3) B2_TW32(2, 4) B2_TW32(2, 5) B2_TW32(2, 6) B2_TW32(2, 7)
    B2_TW32(3, 1) B2_TW32(3, 2) B2_TW32(3, 3) B2_TW32(3, 4) B2_TW32(3, 5) B2_TW32(3, 6) B2_TW32(3, 7)
#undef B2_TW32
    cx<T> o[32];
    B2_UNROLL
    for (int k1 = 0; k1 < 4; ++k1) {
        cx<T> t[8];
        B2_UNROLL
        for (int n2 = 0; n2 < 8; ++n2) t[n2] = v[8 * k1 + n2];
        bf8(t);  // -> k2
        B2_UNROLL
        for (int k2 = 0; k2 < 8; ++k2) o[k1 + 4 * k2] = t[k2];
    }
    B2_UNROLL
    for (int i = 0; i < 32; ++i) v[i] = o[i];
}

// Generic odd-prime butterfly (P = 11, 13, 17, 19, 23, 29, 31): the symmetric form the reference's
// hard-coded prime butterflies use (src/algorithm/butterflies.rs:339-471 for 5, :545-716 for 7, generated code
// for 11..31): pair x_j with x_{P-j}, then
//   X_k, X_{P-k} = (x0 + sum_j cos(2 pi jk/P) (x_j + x_{P-j}))  -/+  i (sum_j sin(2 pi jk/P) (x_j - x_{P-j}))
// i.e. (P-1)^2/2 real-by-complex multiply-adds, each ONE packed FFMA2.
template <int P, typename T> B2_HD void bf_prime(cx<T> (&v)[P]) {
    constexpr int H = (P - 1) / 2;
    cx<T> a[H], b[H];
    B2_UNROLL
    for (int j = 1; j <= H; ++j) {
        a[j - 1] = v[j] + v[P - j];
        b[j - 1] = v[j] - v[P - j];
    }
    const cx<T> x0 = v[0];
    cx<T> sum = x0;
    B2_UNROLL
    for (int j = 0; j < H; ++j) sum = sum + a[j];
    B2_UNROLL
    for (int k = 1; k <= H; ++k) {
        cx<T> c = x0, s = mk<T>(0, 0);
        B2_UNROLL
        for (int j = 1; j <= H; ++j) {
            const int m = (j * k) % P;
            const int mm = m <= H ? m : P - m;
            const T cv = PrimeTw<P, T>::c(mm);
            const T sv = m <= H ? PrimeTw<P, T>::s(mm) : -PrimeTw<P, T>::s(mm);
            c = axpy(c, cv, a[j - 1]);
            s = axpy(s, sv, b[j - 1]);
        }
        v[k] = add_mi(c, s);      // c - i s
        v[P - k] = sub_mi(c, s);  // c + i s
    }
    v[0] = sum;
}

// dispatch on a register array
template <int R, typename T> struct Bfly;
template <typename T> struct Bfly<1, T> { static B2_HD void run(cx<T> (&)[1]) {} };
template <typename T> struct Bfly<2, T> { static B2_HD void run(cx<T> (&v)[2]) { bf2(v[0], v[1]); } };
template <typename T> struct Bfly<3, T> { static B2_HD void run(cx<T> (&v)[3]) { bf3(v[0], v[1], v[2]); } };
template <typename T> struct Bfly<4, T> { static B2_HD void run(cx<T> (&v)[4]) { bf4(v[0], v[1], v[2], v[3]); } };
template <typename T> struct Bfly<5, T> { static B2_HD void run(cx<T> (&v)[5]) { bf5(v[0], v[1], v[2], v[3], v[4]); } };
template <typename T> struct Bfly<7, T> {
    static B2_HD void run(cx<T> (&v)[7]) { bf7(v[0], v[1], v[2], v[3], v[4], v[5], v[6]); }
};
template <typename T> struct Bfly<8, T> { static B2_HD void run(cx<T> (&v)[8]) { bf8(v); } };
template <typename T> struct Bfly<16, T> { static B2_HD void run(cx<T> (&v)[16]) { bf16(v); } };
template <typename T> struct Bfly<32, T> { static B2_HD void run(cx<T> (&v)[32]) { bf32(v); } };
template <typename T> struct Bfly<11, T> { static B2_HD void run(cx<T> (&v)[11]) { bf_prime<11>(v); } };
template <typename T> struct Bfly<13, T> { static B2_HD void run(cx<T> (&v)[13]) { bf_prime<13>(v); } };
template <typename T> struct Bfly<17, T> { static B2_HD void run(cx<T> (&v)[17]) { bf_prime<17>(v); } };
template <typename T> struct Bfly<19, T> { static B2_HD void run(cx<T> (&v)[19]) { bf_prime<19>(v); } };
template <typename T> struct Bfly<23, T> { static B2_HD void run(cx<T> (&v)[23]) { bf_prime<23>(v); } };
template <typename T> struct Bfly<29, T> { static B2_HD void run(cx<T> (&v)[29]) { bf_prime<29>(v); } };
template <typename T> struct Bfly<31, T> { static B2_HD void run(cx<T> (&v)[31]) { bf_prime<31>(v); } };

}  // namespace b2
