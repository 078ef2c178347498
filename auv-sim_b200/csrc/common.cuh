// common.cuh -- arithmetic policy, counter-based uniform stream, lane-group helpers.
//
// Two builds of every kernel live in the same library (include/auvrrt.h `precision`):
//   R = float,  VERIFY = false : fast build.  FMA contraction allowed, numerically stable
//                                restatements of the reference's formulas, MUFU where accurate.
//   R = double, VERIFY = true  : verification build.  Every operation the reference (CPython
//                                float arithmetic) performs is a separately rounded IEEE operation
//                                here too (__dadd_rn / __dmul_rn / __ddiv_rn are never contracted),
//                                in the reference's order, so booleans and indices match it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define AUV_SMS 148

namespace auv {

// ------------------------------------------------------------------ strict / relaxed arithmetic
template <typename R, bool VERIFY> struct Ar;

template <> struct Ar<double, true> {
    typedef double R;
    static __device__ __forceinline__ R add(R a, R b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ R sub(R a, R b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ R mul(R a, R b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ R div(R a, R b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ R sqrt(R a) { return __dsqrt_rn(a); }
    // dx**2 + dy**2 : two rounded squares, one rounded sum (rrt_dubins.py:562)
    static __device__ __forceinline__ R sq2(R dx, R dy) { return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)); }
    static __device__ __forceinline__ void sincos(R a, R *s, R *c) { ::sincos(a, s, c); }
    static __device__ __forceinline__ R floor(R a) { return ::floor(a); }
    static __device__ __forceinline__ R fabs(R a) { return ::fabs(a); }
    static __device__ __forceinline__ R fma(R a, R b, R c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ R inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};

template <> struct Ar<float, false> {
    typedef float R;
    static __device__ __forceinline__ R add(R a, R b) { return a + b; }
    static __device__ __forceinline__ R sub(R a, R b) { return a - b; }
    static __device__ __forceinline__ R mul(R a, R b) { return a * b; }
    static __device__ __forceinline__ R div(R a, R b) { return __fdividef(a, b); }
    static __device__ __forceinline__ R sqrt(R a) { return sqrtf(a); }
    static __device__ __forceinline__ R sq2(R dx, R dy) { return fmaf(dy, dy, dx * dx); }
    // fast build: two-term Cody-Waite reduction to [-pi, pi] (exact product for |a| < 2^10 * 2pi),
    // then the MUFU sine / cosine (abs. error ~4e-7 on the reduced range): headings here are a few
    // radians, so the error per arc primitive is < 1e-6 m, far inside the 1e-5 relative bar
    static __device__ __forceinline__ void sincos(R a, R *s, R *c) {
        const float k = rintf(a * 0.15915494309189535f);
        float r = fmaf(k, -6.28318548202514648f, a);         // 2pi rounded to fp32 ...
        r = fmaf(k, 1.74845553146951715e-7f, r);             // ... plus its rounding error
        *s = __sinf(r); *c = __cosf(r);
    }
    static __device__ __forceinline__ R floor(R a) { return floorf(a); }
    static __device__ __forceinline__ R fabs(R a) { return fabsf(a); }
    static __device__ __forceinline__ R fma(R a, R b, R c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ R inf() { return __int_as_float(0x7f800000); }
};

template <typename R> struct Policy;
template <> struct Policy<float> { static const bool VERIFY = false; typedef Ar<float, false> A; };
template <> struct Policy<double> { static const bool VERIFY = true; typedef Ar<double, true> A; };

// ------------------------------------------------------------------ the uniform stream
// u_k(seed) = top bits of SplitMix64's finaliser applied to key(seed) + (k+1) * golden.  Pure
// function of (seed, k): any lane can evaluate any position (what makes the warp-parallel steer
// possible) and the host can pre-generate the identical sequence for the reference
// (auvrrt_stream_u, oracle/harness.py stream_block).
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t stream_key(uint64_t seed) {
    return mix64((seed + 1) * 0x9E3779B97F4A7C15ULL);
}
__host__ __device__ __forceinline__ uint64_t stream_bits(uint64_t key, uint64_t k) {
    return mix64(key + (k + 1) * 0x9E3779B97F4A7C15ULL);
}
template <typename R> __device__ __forceinline__ R bits_to_u(uint64_t z);
template <> __device__ __forceinline__ double bits_to_u<double>(uint64_t z) {
    return (double)(z >> 11) * 0x1.0p-53;
}
template <> __device__ __forceinline__ float bits_to_u<float>(uint64_t z) {
    return (float)(uint32_t)(z >> 40) * 0x1.0p-24f;   // 24 bits: exact in fp32, in [0, 1)
}

// Where a stream comes from: the counter hash, or an explicit pre-generated array (replaying a
// recorded CPython Mersenne-Twister sequence through the same kernels).
template <typename R> struct Stream {
    uint64_t key;
    const double *ext;    // explicit doubles (host-provided), or nullptr
    int64_t n_ext;
    // Positions past the end of an explicit stream read as 0.5: lanes prefetch a window of positions
    // speculatively, so only a CONSUMED position past the end is an error (see consumed_ok()).
    __device__ __forceinline__ R u(uint32_t k) const {
        if (ext) return ((int64_t)k < n_ext) ? (R)ext[k] : (R)0.5;
        return bits_to_u<R>(stream_bits(key, k));
    }
    __device__ __forceinline__ bool consumed_ok(uint32_t ctr_end) const {
        return !ext || (int64_t)ctr_end <= n_ext;
    }
};

// random.uniform(a, b) = a + (b - a) * random()      (CPython Lib/random.py)
template <typename R> __device__ __forceinline__ R uniform_ab(R a, R b, R u) {
    typedef typename Policy<R>::A A;
    return A::add(a, A::mul(A::sub(b, a), u));
}

// Python float floor division  t // w  for w > 0 (Objects/floatobject.c float_floor_div): the
// floor of the EXACT quotient, not of the rounded one.
template <typename R> __device__ __forceinline__ R floordiv_pos(R t, R w) {
    typedef typename Policy<R>::A A;
    R q = A::floor(t / w);
    // exact residual via FMA: r = q*w - t
    if (A::fma(q, w, -t) > (R)0) q -= (R)1;            // rounded quotient overshot
    else if (A::fma(q + (R)1, w, -t) <= (R)0) q += (R)1;  // or undershot
    return q;
}

// ------------------------------------------------------------------ lane groups
// G lanes (8, 16 or 32) cooperate on one tree / edge; a warp holds 32/G groups.
template <int G> struct Grp {
    int lane, gl, gbase;
    unsigned gmask;
    __device__ __forceinline__ Grp() {
        lane = threadIdx.x & 31;
        gl = lane & (G - 1);
        gbase = lane - gl;
        gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << gbase);
    }
    __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
    __device__ __forceinline__ unsigned ballot(bool p) const {
        unsigned b = __ballot_sync(gmask, p);
        return (G == 32) ? b : ((b >> gbase) & ((1u << (G & 31)) - 1u));
    }
    template <typename T> __device__ __forceinline__ T bcast(T v, int src) const {
        return __shfl_sync(gmask, v, gbase + src);
    }
    template <typename T> __device__ __forceinline__ T up(T v, int d) const {   // lane gl-d (gl >= d)
        return __shfl_up_sync(gmask, v, d, G);
    }
    template <typename T> __device__ __forceinline__ T xorv(T v, int m) const {
        return __shfl_xor_sync(gmask, v, m, G);
    }
};

template <int G, typename T> __device__ __forceinline__ T grp_sum(const Grp<G> &g, T v) {
#pragma unroll
    for (int m = G / 2; m > 0; m >>= 1) v += g.xorv(v, m);
    return v;
}
template <int G, typename T> __device__ __forceinline__ T grp_scan_incl(const Grp<G> &g, T v) {
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        T o = g.up(v, d);
        if (g.gl >= d) v += o;
    }
    return v;
}
// inclusive running sum in the reference's serial order: ((carry + v0) + v1) + ...   (VERIFY)
template <int G> __device__ __forceinline__ double grp_scan_serial(const Grp<G> &g, double v, double carry) {
    double acc = carry, mine = carry;
#pragma unroll 1
    for (int j = 0; j < G; j++) {
        double vj = g.bcast(v, j);
        acc = __dadd_rn(acc, vj);
        if (j == g.gl) mine = acc;
    }
    return mine;
}

// ------------------------------------------------------------------ host-side helpers
extern thread_local char g_err[512];
extern long long g_launches;
int set_err(int code, const char *fmt, ...);

#define AUV_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return auv::set_err(AUVRRT_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,    \
                                cudaGetErrorString(e_));                                       \
    } while (0)

}  // namespace auv
