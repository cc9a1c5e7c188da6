// eos_device.cuh -- the reference's equation of state on the device, operation for operation (never contracted):
// sigmai_dep (src/eos.f90:842-882) and sigmantr (src/eos.f90:661-684).  Shared by K2 (cdfmocsig), cdfmoc -decomp and K6.
#pragma once
#include "common.cuh"
#include "../../include/cdf_eos_coeffs.h"

namespace cdfgpu {

struct EosConst {
    double c[CDF_EOS_NCOEF];
    double r0[6];
    double rdeltaS, r1_S0;
};
__constant__ EosConst c_eos;

// ---- equation of state ----------------------------------------------------------------------------------------
#define CE(i, j, k) c_eos.c[I_EOS##i##j##k]
#define DM(a, b) __dmul_rn((a), (b))
#define DA(a, b) __dadd_rn((a), (b))

// dlr0 of eos.f90:871-877 -- the only polynomial needed when pref == 0.
__device__ __forceinline__ double eos_dlr0(double t, double s)
{
    double a = DA(DA(DM(CE(0, 6, 0), t), DM(CE(1, 5, 0), s)), CE(0, 5, 0));
    a = DA(DA(DM(a, t), DM(DA(DM(CE(2, 4, 0), s), CE(1, 4, 0)), s)), CE(0, 4, 0));
    a = DA(DA(DM(a, t), DM(DA(DM(DA(DM(CE(3, 3, 0), s), CE(2, 3, 0)), s), CE(1, 3, 0)), s)), CE(0, 3, 0));
    a = DA(DA(DM(a, t), DM(DA(DM(DA(DM(DA(DM(CE(4, 2, 0), s), CE(3, 2, 0)), s), CE(2, 2, 0)), s), CE(1, 2, 0)), s)),
           CE(0, 2, 0));
    a = DA(DA(DM(a, t),
              DM(DA(DM(DA(DM(DA(DM(DA(DM(CE(5, 1, 0), s), CE(4, 1, 0)), s), CE(3, 1, 0)), s), CE(2, 1, 0)), s), CE(1, 1, 0)), s)),
           CE(0, 1, 0));
    a = DA(DA(DM(a, t),
              DM(DA(DM(DA(DM(DA(DM(DA(DM(DA(DM(CE(6, 0, 0), s), CE(5, 0, 0)), s), CE(4, 0, 0)), s), CE(3, 0, 0)), s), CE(2, 0, 0)), s),
                    CE(1, 0, 0)),
                 s)),
           CE(0, 0, 0));
    return a;
}
__device__ __forceinline__ double eos_dlr1(double t, double s)
{
    double a = DA(DA(DM(CE(0, 4, 1), t), DM(CE(1, 3, 1), s)), CE(0, 3, 1));
    a = DA(DA(DM(a, t), DM(DA(DM(CE(2, 2, 1), s), CE(1, 2, 1)), s)), CE(0, 2, 1));
    a = DA(DA(DM(a, t), DM(DA(DM(DA(DM(CE(3, 1, 1), s), CE(2, 1, 1)), s), CE(1, 1, 1)), s)), CE(0, 1, 1));
    a = DA(DA(DM(a, t), DM(DA(DM(DA(DM(DA(DM(CE(4, 0, 1), s), CE(3, 0, 1)), s), CE(2, 0, 1)), s), CE(1, 0, 1)), s)),
           CE(0, 0, 1));
    return a;
}
__device__ __forceinline__ double eos_dlr2(double t, double s)
{
    double a = DA(DA(DM(CE(0, 2, 2), t), DM(CE(1, 1, 2), s)), CE(0, 1, 2));
    a = DA(DA(DM(a, t), DM(DA(DM(CE(2, 0, 2), s), CE(1, 0, 2)), s)), CE(0, 0, 2));
    return a;
}
__device__ __forceinline__ double eos_dlr3(double t, double s)
{
    return DA(DA(DM(CE(0, 1, 3), t), DM(CE(1, 0, 3), s)), CE(0, 0, 3));
}

// sigmai_dep for one cell (eos.f90:848-882).  SIGMA0: pref == 0, then dlh == 0 and dlr == dlr0 exactly.
template <bool SIGMA0>
__device__ __forceinline__ double eos_sigma_exact(float tem, float sal, double dlh, double dlref)
{
    const double t = DM((double)tem, 1.0 / 40.0);
    const double s = __dsqrt_rn(DM(fabs(DA((double)sal, c_eos.rdeltaS)), c_eos.r1_S0));
    double dlr = eos_dlr0(t, s);
    if (!SIGMA0) {
        const double r1 = eos_dlr1(t, s), r2 = eos_dlr2(t, s), r3 = eos_dlr3(t, s);
        dlr = DA(DM(DA(DM(DA(DM(r3, dlh), r2), dlh), r1), dlh), dlr);
    }
    const double dltm = (sal == 0.0f) ? 0.0 : 1.0;
    return DM(DA(DA(dlr, dlref), -1000.0), dltm);
}

// sigmantr for one cell (eos.f90:663-682).
__device__ __forceinline__ double eos_sigma_neutral(float tem, float sal)
{
    const double t = (double)tem, s = (double)sal;
    const double sr = __dsqrt_rn(fabs(s));
    const double r1 = DA(DM(DA(DM(DA(DM(-4.3159255086706703e-4, t), 8.1157118782170051e-2), t), 2.2280832068441331e-1), t),
                         1002.3063688892480e0);
    // -a*s - b*t - c  ==  ((-a*s) - (b*t)) - c
    const double r2 = DM(DA(DA(DM(-1.7052298331414675e-7, s), -DM(3.1710675488863952e-3, t)), -1.0304537539692924e-4), s);
    const double r3 = DA(DM(DA(DM(DA(DM(DA(DM(-2.3850178558212048e-9, t), -1.6212552470310961e-7), t), 7.8717799560577725e-5), t),
                               4.3907692647825900e-5), t), 1.0);
    const double r4 = DM(DA(DM(DA(DM(DM(-2.2744455733317707e-9, t), t), 6.0399864718597388e-6), t), -5.1268124398160734e-4), s);
    const double r5 = DM(DM(DA(DM(DM(-1.3409379420216683e-9, t), t), -3.6138532339703262e-5), s), sr);
    return DA(__ddiv_rn(DA(r1, r2), DA(DA(r3, r4), r5)), -1000.0);
}

#undef CE
#undef DM
#undef DA

}  // namespace cdfgpu
