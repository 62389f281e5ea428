/* Host check of ripcurrents_b200/csrc/atan2f_ref.h against libm (tests/test_atan2f_port.py). Prints mismatch counts. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "atan2f_ref.h"
static uint64_t st = 88172645463325252ull;
static uint32_t rnd(void) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); }
static int same(float a, float b) { return rc_fbits(a) == rc_fbits(b) || (a != a && b != b); }
int main(int argc, char** argv)
{
    const long npairs = argc > 1 ? atol(argv[1]) : 20000000;
    long bad1 = 0, bad2 = 0;
    for (uint64_t b = 0; b < 0x100000000ull; b += 211) {                       /* every 211th float */
        const float x = rc_bitsf((int32_t)b);
        if (!same(atanf(x), rc_atanf_ref(x))) bad1++;
    }
    const float sp[] = {0.f, -0.f, 1.f, -1.f, INFINITY, -INFINITY, NAN, 1e-45f, -1e-45f, 3e38f, -3e38f, 1e-30f, 0.4375f, 0.6875f, 1.1875f, 2.4375f};
    for (unsigned i = 0; i < sizeof sp / 4; i++)
        for (unsigned j = 0; j < sizeof sp / 4; j++)
            if (!same(atan2f(sp[i], sp[j]), rc_atan2f_ref(sp[i], sp[j]))) bad2++;
    for (long i = 0; i < npairs; i++) {
        float y, x;
        if (i & 1) { y = rc_bitsf((int32_t)rnd()); x = rc_bitsf((int32_t)rnd()); }
        else {                                                                  /* flow-like magnitudes */
            y = ((int)(rnd() % 2000001) - 1000000) * 1e-5f; x = ((int)(rnd() % 2000001) - 1000000) * 1e-5f;
            if (i % 7 == 0) x *= 1e-6f;
            if (i % 11 == 0) y *= 1e-7f;
        }
        if (!same(atan2f(y, x), rc_atan2f_ref(y, x))) bad2++;
    }
    printf("%ld %ld\n", bad1, bad2);
    return 0;
}
