// atan2f with the results of the C library the reference links against (glibc 2.39, x86-64: the fdlibm-derived float
// kernels).  vectorToColor (ripcurrents_module.cpp:1031) truncates atan2(y, x)*180/pi/2 to an 8-bit hue, so a 1-ulp
// difference in atan2f moves pixels across hue boundaries; CUDA's own atan2f is a different approximation.  Written from
// the published algorithm (argument reduction to |t| < 7/16 around 0.5, 1, 1.5, inf; odd/even split 11-term polynomial);
// every operation is a single IEEE fp32 operation (compile without FMA contraction).  tests/test_atan2f_port.py compiles
// this header for the host and compares it with libm on ~10^8 arguments.
#ifndef RC_ATAN2F_REF_H
#define RC_ATAN2F_REF_H
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define RC_HD __host__ __device__ __forceinline__
#else
#define RC_HD static inline
#endif

RC_HD int32_t rc_fbits(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
RC_HD float rc_bitsf(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

#if defined(__CUDA_ARCH__)
#define RC_MUL(a, b) __fmul_rn((a), (b))
#define RC_ADD(a, b) __fadd_rn((a), (b))
#define RC_SUB(a, b) __fsub_rn((a), (b))
#define RC_DIV(a, b) __fdiv_rn((a), (b))
#else
#define RC_MUL(a, b) ((a) * (b))
#define RC_ADD(a, b) ((a) + (b))
#define RC_SUB(a, b) ((a) - (b))
#define RC_DIV(a, b) ((a) / (b))
#endif

#define RC_ATAN_BIG 0x4c000000      /* |x| >= 2^25: +-pi/2 */

RC_HD float rc_atanf_ref(float x)
{
    const float hi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float lo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f, a3 = -1.1111110449e-01f,
                a4 = 9.0908870101e-02f, a5 = -7.6918758452e-02f, a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f,
                a8 = 4.9768779427e-02f, a9 = -3.6531571299e-02f, a10 = 1.6285819933e-02f;
    const int32_t hx = rc_fbits(x), ix = hx & 0x7fffffff;
    int id;
    if (ix >= RC_ATAN_BIG) {
        if (ix > 0x7f800000) return RC_ADD(x, x);
        const float r = RC_ADD(hi[3], lo[3]);
        return hx > 0 ? r : -r;
    }
    if (ix < 0x3ee00000) {                 // |x| < 7/16
        if (ix < 0x31000000) return x;     // |x| < 2^-29
        id = -1;
    } else {
        x = rc_bitsf(ix);                  // |x|
        if (ix < 0x3f980000) {             // |x| < 19/16
            if (ix < 0x3f300000) { id = 0; x = RC_DIV(RC_SUB(RC_MUL(2.0f, x), 1.0f), RC_ADD(2.0f, x)); }
            else { id = 1; x = RC_DIV(RC_SUB(x, 1.0f), RC_ADD(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = RC_DIV(RC_SUB(x, 1.5f), RC_ADD(1.0f, RC_MUL(1.5f, x))); }
            else { id = 3; x = RC_DIV(-1.0f, x); }
        }
    }
    const float z = RC_MUL(x, x), w = RC_MUL(z, z);
    float s1 = RC_ADD(a8, RC_MUL(w, a10));
    s1 = RC_ADD(a6, RC_MUL(w, s1)); s1 = RC_ADD(a4, RC_MUL(w, s1)); s1 = RC_ADD(a2, RC_MUL(w, s1)); s1 = RC_ADD(a0, RC_MUL(w, s1));
    s1 = RC_MUL(z, s1);
    float s2 = RC_ADD(a7, RC_MUL(w, a9));
    s2 = RC_ADD(a5, RC_MUL(w, s2)); s2 = RC_ADD(a3, RC_MUL(w, s2)); s2 = RC_ADD(a1, RC_MUL(w, s2));
    s2 = RC_MUL(w, s2);
    const float p = RC_MUL(x, RC_ADD(s1, s2));
    if (id < 0) return RC_SUB(x, p);
    const float r = RC_SUB(hi[id], RC_SUB(RC_SUB(p, lo[id]), x));
    return hx < 0 ? -r : r;
}

RC_HD float rc_atan2f_ref(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    const int32_t hx = rc_fbits(x), hy = rc_fbits(y), ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return RC_ADD(x, y);
    if (hx == 0x3f800000) return rc_atanf_ref(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) { case 0: case 1: return y; case 2: return RC_ADD(pi, tiny); default: return RC_SUB(-pi, tiny); }
    }
    if (ix == 0) return hy < 0 ? RC_SUB(-pi_o_2, tiny) : RC_ADD(pi_o_2, tiny);
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
                case 0: return RC_ADD(pi_o_4, tiny);
                case 1: return RC_SUB(-pi_o_4, tiny);
                case 2: return RC_ADD(RC_MUL(3.0f, pi_o_4), tiny);
                default: return RC_SUB(RC_MUL(-3.0f, pi_o_4), tiny);
            }
        }
        switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return RC_ADD(pi, tiny); default: return RC_SUB(-pi, tiny); }
    }
    if (iy == 0x7f800000) return hy < 0 ? RC_SUB(-pi_o_2, tiny) : RC_ADD(pi_o_2, tiny);
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = RC_ADD(pi_o_2, RC_MUL(0.5f, pi_lo));
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = rc_atanf_ref(rc_bitsf(rc_fbits(RC_DIV(y, x)) & 0x7fffffff));
    switch (m) {
        case 0: return z;
        case 1: return rc_bitsf(rc_fbits(z) ^ (int32_t)0x80000000);
        case 2: return RC_SUB(pi, RC_SUB(z, pi_lo));
        default: return RC_SUB(RC_SUB(z, pi_lo), pi);
    }
}
#endif
