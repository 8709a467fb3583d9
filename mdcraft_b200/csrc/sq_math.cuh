// sq_math.cuh — phase -> (sin, cos) pipeline shared by the S(q) kernels and the
// on-box accuracy probe, so that what is probed is exactly what is run.
#pragma once
#include <stdint.h>

namespace mdc {

// ptxas lowers sin/cos.approx.f32(x) to  FMUL.RZ t, x, RRO ; MUFU.SIN/COS t
// with RRO = float(1/(2 pi)) = 0x3E22F983, i.e. the SFU consumes TURNS.  PTX
// offers no way to hand it turns directly, so the kernels pre-multiply by
// 1/RRO split into two floats (hi + lo): x = RN(t / RRO) has only a random
// rounding error, whereas a single-constant  t * float(2 pi)  carries a
// coherent relative phase-scale error of -1.25e-8 that a sum over N particles
// multiplies by N (SURVEY.md §7 hard part 2).
//   RRO        = 0.159154936671257019043   (exact value of the float)
//   1/RRO      = 6.2831855606562375
#define MDC_INV_RRO_HI 6.28318548202514648438f
#define MDC_INV_RRO_LO 7.863108919536899e-08f

// Phase is a 32-bit binary fraction of a turn (value = ph / 2^32).  The top 23
// bits are placed in the mantissa of a float in [1, 2): every phase lives in
// ONE binade, so the .RZ rounding of ptxas' range-reduction multiply and the
// truncation of the low 9 bits are constant phase shifts (a global rotation of
// rho, which |rho|^2 and Re(rho_j conj rho_k) do not see) instead of
// magnitude-dependent ones.
__device__ __forceinline__ void sincos_phase(uint32_t ph, float &s, float &c)
{
    const float t = __uint_as_float(0x3f800000u | (ph >> 9));
    const float x = fmaf(t, MDC_INV_RRO_HI, t * MDC_INV_RRO_LO);
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(x));
}

}  // namespace mdc
