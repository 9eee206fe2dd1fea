// fpt_portable.cuh — the handful of intrinsics the warp-autonomous scoring kernel (fpt_warp_core.cuh) uses,
// each with a plain C++ equivalent. The device build maps them onto the sm_100a instructions; the host build
// (FPT_HOST_EMU, used only by tests/emu to run the kernel's per-lane steps lane by lane against the CPU oracle)
// takes the portable bodies. The product never runs the host bodies: libfpt_b200.so is compiled by nvcc and
// launches the kernel on the GPU only.
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstring>

#if defined(FPT_HOST_EMU)
#include <vector_types.h>
#define FPT_HD inline
#define FPT_NOINLINE_HD inline
#define FPT_EMU_ASSERT(c)                                                                            \
    do {                                                                                             \
        if (!(c)) { fprintf(stderr, "emu assert failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); abort(); } \
    } while (0)
#include <cstdio>
#include <cstdlib>
#else
#include <cuda_runtime.h>
#define FPT_HD __host__ __device__ __forceinline__
#define FPT_NOINLINE_HD __host__ __device__ __noinline__
#define FPT_EMU_ASSERT(c) ((void)0)
#endif

namespace fpt {
namespace pt {

FPT_HD unsigned vadd2(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __vadd2(a, b);
#else
    return ((a + b) & 0xFFFFu) | (((a >> 16) + (b >> 16)) << 16);
#endif
}
FPT_HD unsigned vminu2(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b);
#else
    const unsigned al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}
FPT_HD unsigned vmaxu2(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __vmaxu2(a, b);
#else
    const unsigned al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16;
    return (al > bl ? al : bl) | ((ah > bh ? ah : bh) << 16);
#endif
}
// lo16(a) | lo16(b) << 16
FPT_HD unsigned pack_lo16(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, 0x5410);
#else
    return (a & 0xFFFFu) | (b << 16);
#endif
}
// bits [sh, sh + 32) of the 64-bit value hi:lo, 0 <= sh < 32
FPT_HD unsigned funnel_r(unsigned lo, unsigned hi, int sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}
FPT_HD unsigned long long brevll(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __brevll(v);
#else
    unsigned long long r = 0;
    for (int i = 0; i < 64; ++i) r |= ((v >> i) & 1ull) << (63 - i);
    return r;
#endif
}
FPT_HD int ffs32(unsigned v) {
#if defined(__CUDA_ARCH__)
    return __ffs(v);
#else
    return __builtin_ffs((int)v);
#endif
}
FPT_HD float fdividef(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdividef(a, b);
#else
    return a / b;
#endif
}
FPT_HD float uint_as_float(unsigned v) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}
FPT_HD int float_as_int(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int v;
    memcpy(&v, &f, 4);
    return v;
#endif
}
FPT_HD float fmaf_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return fmaf(a, b, c);
#else
    return std::fmaf(a, b, c);
#endif
}
// IEEE double operations that must not be contracted into FMAs (the reference is compiled without FMA)
FPT_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
FPT_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
FPT_HD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b;
    return r;
#endif
}
// 1/x for normal positive x, relative error 2^-23 on the device (MUFU.RCP)
FPT_HD float rcp_approx(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

// Two single-precision values in one 64-bit register pair: Blackwell issues add / mul / fma on both halves as ONE
// instruction (FADD2 / FMUL2 / FFMA2, PTX .f32x2). The scoring step keeps the plus strand in the low half and the
// minus strand in the high half. Each half is rounded exactly as the scalar operation would be.
struct f32x2 {
#if defined(__CUDA_ARCH__)
    unsigned long long v;
#else
    float lo_, hi_;
#endif
};
FPT_HD f32x2 pack2(float lo, float hi) {
    f32x2 r;
#if defined(__CUDA_ARCH__)
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
#else
    r.lo_ = lo; r.hi_ = hi;
#endif
    return r;
}
FPT_HD float lo2(f32x2 a) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float((unsigned)a.v);
#else
    return a.lo_;
#endif
}
FPT_HD float hi2(f32x2 a) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float((unsigned)(a.v >> 32));
#else
    return a.hi_;
#endif
}
// the pair stored at p (8-byte aligned): low half first
FPT_HD f32x2 ld_pair(const float *p) {
    f32x2 r;
#if defined(__CUDA_ARCH__)
    r.v = *reinterpret_cast<const unsigned long long *>(p);
#else
    r.lo_ = p[0]; r.hi_ = p[1];
#endif
    return r;
}
FPT_HD f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
#if defined(__CUDA_ARCH__)
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
#else
    r.lo_ = a.lo_ + b.lo_; r.hi_ = a.hi_ + b.hi_;
#endif
    return r;
}
FPT_HD f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
#if defined(__CUDA_ARCH__)
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
#else
    r.lo_ = a.lo_ * b.lo_; r.hi_ = a.hi_ * b.hi_;
#endif
    return r;
}
FPT_HD f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
#if defined(__CUDA_ARCH__)
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
#else
    r.lo_ = std::fmaf(a.lo_, b.lo_, c.lo_); r.hi_ = std::fmaf(a.hi_, b.hi_, c.hi_);
#endif
    return r;
}

// v with its high word replaced by `hi` when cond holds (one select on the high word; the low word stays)
FPT_HD double set_hi_if(double v, unsigned hi, bool cond) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(cond ? (int)hi : __double2hiint(v), __double2loint(v));
#else
    if (!cond) return v;
    unsigned long long b;
    std::memcpy(&b, &v, 8);
    b = (b & 0xFFFFFFFFull) | ((unsigned long long)hi << 32);
    std::memcpy(&v, &b, 8);
    return v;
#endif
}

template <class T>
FPT_HD T ldg(const T *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

}  // namespace pt
}  // namespace fpt
