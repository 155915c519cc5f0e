// Radix butterflies for the Stockham passes (registers only, fully unrolled).
//
// All trigonometric constants are produced by constexpr polynomial evaluation in double
// and folded by the compiler after unrolling (check: no DFMA/MUFU.SIN in the SASS of the
// FFT kernels).  DIR = -1 is the forward transform exp(-2 pi i jk/N), DIR = +1 the
// (unscaled) inverse.  T is float or double.
#pragma once
#include <cuda_runtime.h>

namespace p3d {

template <typename T> struct Cx { T x, y; };
template <> struct __align__(8)  Cx<float>  { float x, y; };
template <> struct __align__(16) Cx<double> { double x, y; };

template <typename T> __host__ __device__ __forceinline__ Cx<T> cmake(T a, T b) { Cx<T> r; r.x = a; r.y = b; return r; }
template <typename T> __host__ __device__ __forceinline__ Cx<T> cadd(Cx<T> a, Cx<T> b) { return cmake<T>(a.x + b.x, a.y + b.y); }
template <typename T> __host__ __device__ __forceinline__ Cx<T> csub(Cx<T> a, Cx<T> b) { return cmake<T>(a.x - b.x, a.y - b.y); }
template <typename T> __host__ __device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
    return cmake<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
template <typename T> __host__ __device__ __forceinline__ Cx<T> cmulc(Cx<T> a, Cx<T> b) {
    return cmake<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
template <typename T> __host__ __device__ __forceinline__ Cx<T> cscale(Cx<T> a, T s) { return cmake<T>(a.x * s, a.y * s); }
// multiply by DIR * i  (DIR=-1: -i ; DIR=+1: +i)
template <int DIR, typename T> __host__ __device__ __forceinline__ Cx<T> cmul_i(Cx<T> a) {
    return DIR < 0 ? cmake<T>(a.y, -a.x) : cmake<T>(-a.y, a.x);
}
// c + s * a  (real scalar s times complex a)
template <typename T> __host__ __device__ __forceinline__ Cx<T> cfma_real(T s, Cx<T> a, Cx<T> c) {
    return cmake<T>(c.x + s * a.x, c.y + s * a.y);
}
// s * a
template <typename T> __host__ __device__ __forceinline__ Cx<T> cmul_real(T s, Cx<T> a) { return cmake<T>(s * a.x, s * a.y); }

// ---------------------------------------------------------------------------------------
// Blackwell packed FP32: add/mul/fma.f32x2 execute both halves of a 64-bit register pair in one
// instruction (SASS FADD2 / FMUL2 / FFMA2, with free per-operand swap, broadcast and sign
// modifiers).  A complex64 value IS such a pair, so complex add/sub, real-scalar FMA and the
// complex multiply map to one or two packed instructions instead of two or four scalar ones;
// on sm_100 a scalar 3-register FFMA issues at half the FP32 rate, the packed forms at full rate.
// These overloads are picked for Cx<float> in device code; double and host code use the
// templates above.
// ---------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(P3D_NO_F32X2)
__device__ __forceinline__ unsigned long long cx_bits(Cx<float> a) { return *reinterpret_cast<unsigned long long*>(&a); }
__device__ __forceinline__ Cx<float> cx_from(unsigned long long u) { return *reinterpret_cast<Cx<float>*>(&u); }
__device__ __forceinline__ Cx<float> add2(Cx<float> a, Cx<float> b) {
    unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cx_bits(a)), "l"(cx_bits(b))); return cx_from(r);
}
__device__ __forceinline__ Cx<float> sub2(Cx<float> a, Cx<float> b) {
    unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cx_bits(a)), "l"(cx_bits(b))); return cx_from(r);
}
__device__ __forceinline__ Cx<float> mul2(Cx<float> a, Cx<float> b) {
    unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cx_bits(a)), "l"(cx_bits(b))); return cx_from(r);
}
__device__ __forceinline__ Cx<float> fma2(Cx<float> a, Cx<float> b, Cx<float> c) {
    unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(cx_bits(a)), "l"(cx_bits(b)), "l"(cx_bits(c))); return cx_from(r);
}
__device__ __forceinline__ Cx<float> cadd(Cx<float> a, Cx<float> b) { return add2(a, b); }
__device__ __forceinline__ Cx<float> csub(Cx<float> a, Cx<float> b) { return sub2(a, b); }
__device__ __forceinline__ Cx<float> cmul(Cx<float> a, Cx<float> w) {
    return fma2(cmake<float>(a.x, a.x), w, mul2(cmake<float>(a.y, a.y), cmake<float>(-w.y, w.x)));
}
__device__ __forceinline__ Cx<float> cmulc(Cx<float> a, Cx<float> w) {
    return fma2(cmake<float>(a.x, a.x), cmake<float>(w.x, -w.y), mul2(cmake<float>(a.y, a.y), cmake<float>(w.y, w.x)));
}
__device__ __forceinline__ Cx<float> cscale(Cx<float> a, float s) { return mul2(a, cmake<float>(s, s)); }
__device__ __forceinline__ Cx<float> cfma_real(float s, Cx<float> a, Cx<float> c) { return fma2(cmake<float>(s, s), a, c); }
__device__ __forceinline__ Cx<float> cmul_real(float s, Cx<float> a) { return mul2(cmake<float>(s, s), a); }
#endif

// ---------------------------------------------------------------------------------------
// constexpr trigonometry: cos(2 pi m / n), sin(2 pi m / n) for integer m, n
// ---------------------------------------------------------------------------------------
__host__ __device__ constexpr double cx_sin_poly(double x) {   // |x| <= pi/4
    double x2 = x * x;
    return x * (1.0 + x2 * (-1.0 / 6 + x2 * (1.0 / 120 + x2 * (-1.0 / 5040 + x2 * (1.0 / 362880 + x2 * (-1.0 / 39916800 +
           x2 * (1.0 / 6227020800.0 + x2 * (-1.0 / 1307674368000.0))))))));
}
__host__ __device__ constexpr double cx_cos_poly(double x) {   // |x| <= pi/4
    double x2 = x * x;
    return 1.0 + x2 * (-0.5 + x2 * (1.0 / 24 + x2 * (-1.0 / 720 + x2 * (1.0 / 40320 + x2 * (-1.0 / 3628800 +
           x2 * (1.0 / 479001600.0 + x2 * (-1.0 / 87178291200.0 + x2 * (1.0 / 20922789888000.0))))))));
}
constexpr double CX_PI = 3.141592653589793238462643383279502884;

// angle = 2 pi m / n reduced with exact integer arithmetic to an octant
__host__ __device__ constexpr double cx_cos2pi(int m, int n) {
    m %= n; if (m < 0) m += n;
    // use cos symmetry about pi: m -> n - m
    if (2 * m > n) m = n - m;                 // now angle in [0, pi]
    bool neg = false;
    // angle > pi/2  <=> 4m > n : cos(a) = -cos(pi - a), pi - a = 2 pi (n/2 - m)/n = 2 pi (n - 2m)/(2n)
    int num = m, den = n;                     // angle = 2 pi num/den
    if (4 * m > n) { neg = true; num = n - 2 * m; den = 2 * n; }   // now angle in [0, pi/2]
    // angle > pi/4 <=> 8 num > den : cos(a) = sin(pi/2 - a) = sin(2 pi (den - 4 num)/(4 den))
    double r = 0.0;
    if (8 * num > den) r = cx_sin_poly(2.0 * CX_PI * (double)(den - 4 * num) / (4.0 * (double)den));
    else               r = cx_cos_poly(2.0 * CX_PI * (double)num / (double)den);
    return neg ? -r : r;
}
__host__ __device__ constexpr double cx_sin2pi(int m, int n) {
    // sin(2 pi m/n) = cos(2 pi m/n - pi/2) = cos(2 pi (4m - n)/(4n))
    return cx_cos2pi(4 * m - n, 4 * n);
}

// W_n^m for direction DIR: exp(DIR * 2 pi i m / n)
template <int DIR, typename T> __host__ __device__ __forceinline__ Cx<T> cx_w(int m, int n) {
    return cmake<T>((T)cx_cos2pi(m, n), (T)((double)DIR * cx_sin2pi(m, n)));
}

// ---------------------------------------------------------------------------------------
// butterflies: in-place DFT of v[0..R-1]
// ---------------------------------------------------------------------------------------
template <int R, int DIR, typename T> struct Bfly;

template <int DIR, typename T> struct Bfly<1, DIR, T> {
    __device__ __forceinline__ static void run(Cx<T>* v) {}
};

template <int DIR, typename T> struct Bfly<2, DIR, T> {
    __device__ __forceinline__ static void run(Cx<T>* v) {
        Cx<T> a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    }
};

template <int DIR, typename T> struct Bfly<4, DIR, T> {
    __device__ __forceinline__ static void run(Cx<T>* v) {
        Cx<T> a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        Cx<T> c = cadd(v[1], v[3]), d = cmul_i<DIR>(csub(v[1], v[3]));
        v[0] = cadd(a, c); v[2] = csub(a, c);
        v[1] = cadd(b, d); v[3] = csub(b, d);
    }
};

// odd prime radix, symmetric form: pairs (j, P-j)
template <int P, int DIR, typename T> struct BflyPrime {
    __device__ __forceinline__ static void run(Cx<T>* v) {
        constexpr int H = (P - 1) / 2;
        Cx<T> s[H], d[H];
#pragma unroll
        for (int j = 0; j < H; ++j) { s[j] = cadd(v[j + 1], v[P - 1 - j]); d[j] = csub(v[j + 1], v[P - 1 - j]); }
        Cx<T> x0 = v[0];
        Cx<T> sum = x0;
#pragma unroll
        for (int j = 0; j < H; ++j) sum = cadd(sum, s[j]);
        v[0] = sum;
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            Cx<T> A = x0, B;
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const T c = (T)cx_cos2pi(j * k, P);
                const T sn = (T)cx_sin2pi(j * k, P);
                A = cfma_real(c, s[j - 1], A);
                B = (j == 1) ? cmul_real(sn, d[0]) : cfma_real(sn, d[j - 1], B);
            }
            // forward: X[k] = A - i B, X[P-k] = A + i B ; inverse swaps
            Cx<T> iB = cmul_i<DIR>(B);         // DIR*i*B
            v[k] = cadd(A, iB);
            v[P - k] = csub(A, iB);
        }
    }
};
template <int DIR, typename T> struct Bfly<3, DIR, T>  { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPrime<3, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<5, DIR, T>  { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPrime<5, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<7, DIR, T>  { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPrime<7, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<11, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPrime<11, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<13, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPrime<13, DIR, T>::run(v); } };

// composite radix R = R1 * R2 (Cooley-Tukey inside registers, constant twiddles)
//   X[k1 + R1 k2] = sum_{n2} W_R^{n2 k1} [ sum_{n1} x[R2 n1 + n2] W_R1^{n1 k1} ] W_R2^{n2 k2}
template <int R1, int R2, int DIR, typename T> struct BflyCT {
    __device__ __forceinline__ static void run(Cx<T>* v) {
        constexpr int R = R1 * R2;
        Cx<T> y[R2][R1];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) y[n2][n1] = v[R2 * n1 + n2];
            Bfly<R1, DIR, T>::run(y[n2]);
#pragma unroll
            for (int k1 = 1; k1 < R1; ++k1) {
                if (n2 == 0) continue;
                const int m = (n2 * k1) % R;
                // exploit trivial twiddles
                if (4 * m == R)            y[n2][k1] = cmul_i<DIR>(y[n2][k1]);
                else if (2 * m == R)       y[n2][k1] = cmake<T>(-y[n2][k1].x, -y[n2][k1].y);
                else if (4 * m == 3 * R)   y[n2][k1] = cmul_i<-DIR>(y[n2][k1]);
                else                       y[n2][k1] = cmul(y[n2][k1], cx_w<DIR, T>(m, R));
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            Cx<T> z[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) z[n2] = y[n2][k1];
            Bfly<R2, DIR, T>::run(z);
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = z[k2];
        }
    }
};
// composite radix R = R1 * R2 with gcd(R1, R2) = 1: Good-Thomas prime-factor mapping, NO twiddles.
//   input  n = (R2 n1 + R1 n2) mod R,   output k with k = k1 (mod R1), k = k2 (mod R2)
//   X[k] = sum_{n2} W_R2^{n2 k2} sum_{n1} x[n] W_R1^{n1 k1}
__host__ __device__ constexpr int pfa_out_index(int R1, int R2, int k1, int k2) {
    for (int k = 0; k < R1 * R2; ++k) if (k % R1 == k1 && k % R2 == k2) return k;
    return -1;
}
template <int R1, int R2, int DIR, typename T> struct BflyPFA {
    __device__ __forceinline__ static void run(Cx<T>* v) {
        constexpr int R = R1 * R2;
        Cx<T> y[R2][R1];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) y[n2][n1] = v[(R2 * n1 + R1 * n2) % R];
            Bfly<R1, DIR, T>::run(y[n2]);
        }
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            Cx<T> z[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) z[n2] = y[n2][k1];
            Bfly<R2, DIR, T>::run(z);
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) v[pfa_out_index(R1, R2, k1, k2)] = z[k2];
        }
    }
};
template <int DIR, typename T> struct Bfly<6, DIR, T>  { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<2, 3, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<8, DIR, T>  { __device__ __forceinline__ static void run(Cx<T>* v) { BflyCT<2, 4, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<9, DIR, T>  { __device__ __forceinline__ static void run(Cx<T>* v) { BflyCT<3, 3, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<10, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<2, 5, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<12, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<4, 3, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<16, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyCT<4, 4, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<14, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<2, 7, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<15, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<3, 5, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<18, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<2, 9, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<21, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<3, 7, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<22, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<2, 11, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<30, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<5, 6, DIR, T>::run(v); } };
template <int DIR, typename T> struct Bfly<20, DIR, T> { __device__ __forceinline__ static void run(Cx<T>* v) { BflyPFA<4, 5, DIR, T>::run(v); } };

}  // namespace p3d
