// phmm_log10.h -- glibc 2.39's log10f, restated operation for operation, for host AND device.
//
// Why: the reference's final value on the FP32 path is (double)(log10f(raw) - log10f(2^120)) evaluated in
// float by the host's libm (pairhmm/intel_pairhmm.hpp:142).  glibc's log10f is NOT correctly rounded (1.4% of
// all inputs differ from the correctly rounded value in the last ulp), CUDA's log10f differs from it as
// well, so a device-side final step -- and with it the device-side cap / filter / genotype reduction of
// SURVEY.md section 8(f)-3 -- has to reproduce glibc's algorithm, not just its accuracy.  Every operation
// below is an IEEE-754 float or double operation that x86-64 and sm_100 round identically, so equality is by
// construction; tools/check_log10f.cpp verifies it EXHAUSTIVELY on the host (all 2^31 non-negative float bit
// patterns against the libm of this image), tests/test_log10_gpu.py on the device.
//
// Source restated (third-party, absent from /root/reference): GNU libc 2.39 (Ubuntu GLIBC 2.39-0ubuntu8.5),
//   sysdeps/ieee754/flt-32/e_log10f.c   __ieee754_log10f: k = exponent, x' = mantissa scaled into [0.5, 2),
//                                        z = y*log10_2lo + ivln10*logf(x');  return z + y*log10_2hi   (float, unfused)
//   sysdeps/ieee754/flt-32/e_logf.c     __logf (ARM optimized-routines): 16-entry table, degree-3 polynomial in
//                                        double; the x86-64 multiarch build selects the FMA variant on every CPU
//                                        with AVX2+FMA (sysdeps/x86_64/fpu/multiarch/e_logf.c), whose contraction
//                                        pattern -- read off the disassembly of libm.so.6 -- is the one below.
// Hosts without FMA would run the unfused variant; phmm_log10_selftest() (phmm_tables.cpp) compares this
// restatement with the running libm at engine creation and the engine falls back to the host log10f pass
// when they disagree, so the device path is never silently different from the reference's.
#pragma once
#include <cstdint>
#include <cstring>
#ifdef __CUDACC__
#define PHMM_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define PHMM_HD inline
#endif

namespace phmm {

#ifdef __CUDA_ARCH__
__device__ __forceinline__ double l10_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double l10_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double l10_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float l10_fmulf(float a, float b) { return __fmul_rn(a, b); }      // never contracted, never flushed
__device__ __forceinline__ float l10_faddf(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ uint32_t l10_bits(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float l10_float(uint32_t u) { return __uint_as_float(u); }
#else
inline double l10_fma(double a, double b, double c) { return std::fma(a, b, c); }
inline double l10_mul(double a, double b) { volatile double r = a * b; return r; }              // volatile: no contraction
inline double l10_add(double a, double b) { volatile double r = a + b; return r; }
inline float l10_fmulf(float a, float b) { volatile float r = a * b; return r; }
inline float l10_faddf(float a, float b) { volatile float r = a + b; return r; }
inline uint32_t l10_bits(float x) { uint32_t u; std::memcpy(&u, &x, 4); return u; }
inline float l10_float(uint32_t u) { float x; std::memcpy(&x, &u, 4); return x; }
#endif

// __logf_data (sysdeps/ieee754/flt-32/e_logf_data.c): {1/c, log(c)} for 16 subintervals of [OFF, 2 OFF)
#ifdef __CUDA_ARCH__
__device__
#endif
static const double kLogfTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2}, {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

// __logf, FMA variant, for normal positive finite x (the callers below guarantee it)
PHMM_HD float glibc_logf_normal(float x)
{
    const double Ln2 = 0x1.62e42fefa39efp-1;
    const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    const uint32_t ix = l10_bits(x);
    if (ix == 0x3f800000u) return 0.0f;
    const uint32_t tmp = ix - 0x3f330000u;                 // OFF
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = kLogfTab[i][0], logc = kLogfTab[i][1];
    const double z = (double)l10_float(iz);
    const double y0 = l10_fma((double)k, Ln2, logc);      // vfmadd132sd: k*Ln2 + logc
    const double r = l10_fma(z, invc, -1.0);              // z*invc - 1
    double y = l10_fma(r, A1, A2);                        // A1*r + A2
    const double r2 = l10_mul(r, r);
    const double y0r = l10_add(r, y0);                    // r + y0
    y = l10_fma(r2, A0, y);                               // A0*r2 + y
    return (float)l10_fma(r2, y, y0r);                    // y*r2 + (y0 + r), rounded to float
}

// __ieee754_log10f for x >= 0 (zero, subnormals, inf and NaN included; negative arguments never occur here)
PHMM_HD float glibc_log10f(float x)
{
    const float two25 = 3.3554432000e+07f, ivln10 = 0x1.bcb7b2p-2f, log10_2hi = 0x1.3441p-2f, log10_2lo = 0x1.a84fb6p-21f;
    int32_t hx = (int32_t)l10_bits(x), k = 0;
    if (hx < 0x00800000) {                                 // x < 2^-126
        if ((hx & 0x7fffffff) == 0) return -two25 / 0.0f;  // log(+-0) = -inf
        k -= 25; x = l10_fmulf(x, two25);                  // subnormal: scale up
        hx = (int32_t)l10_bits(x);
    }
    if (hx >= 0x7f800000) return x + x;
    k += (hx >> 23) - 127;
    const int32_t i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
    hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
    const float y = (float)(k + i);
    const float lf = glibc_logf_normal(l10_float((uint32_t)hx));
    const float z = l10_faddf(l10_fmulf(ivln10, lf), l10_fmulf(y, log10_2lo));
    return l10_faddf(z, l10_fmulf(y, log10_2hi));
}

// ---- double precision: __ieee754_log10 (sysdeps/ieee754/dbl-64/e_log10.c) over __log (ARM optimized-routines,
// sysdeps/ieee754/dbl-64/e_log.c, 128-entry table), FMA variant of the x86-64 multiarch build: the final value of
// an FP64-rescued pair is log10(raw64) - log10(2^1020) in double on the host (intel_pairhmm.hpp:139), and the
// device-side genotype reduction needs that very double.  Contraction pattern read off the disassembly of
// libm.so.6 (__log_fma); tools/check_log10f.cpp compares with the running libm on > 10^9 sampled doubles (all
// binades, dense around 1.0); exhaustive testing is not possible for doubles, equality is by construction.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ double l10_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ uint64_t l10_bits64(double x) { return (uint64_t)__double_as_longlong(x); }
__device__ __forceinline__ double l10_double(uint64_t u) { return __longlong_as_double((long long)u); }
#else
inline double l10_sub(double a, double b) { volatile double r = a - b; return r; }
inline uint64_t l10_bits64(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
inline double l10_double(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
#endif

// __log_data.tab: {1/c, log(c)} for 128 subintervals of [OFF, 2 OFF), OFF = 0x3fe6000000000000
#ifdef __CUDA_ARCH__
__device__
#endif
static const double kLogTab[128][2] = {
    {0x1.734f0c3e0de9fp+0, -0x1.7cc7f79e69000p-2},
    {0x1.713786a2ce91fp+0, -0x1.76feec20d0000p-2},
    {0x1.6f26008fab5a0p+0, -0x1.713e31351e000p-2},
    {0x1.6d1a61f138c7dp+0, -0x1.6b85b38287800p-2},
    {0x1.6b1490bc5b4d1p+0, -0x1.65d5590807800p-2},
    {0x1.69147332f0cbap+0, -0x1.602d076180000p-2},
    {0x1.6719f18224223p+0, -0x1.5a8ca86909000p-2},
    {0x1.6524f99a51ed9p+0, -0x1.54f4356035000p-2},
    {0x1.63356aa8f24c4p+0, -0x1.4f637c36b4000p-2},
    {0x1.614b36b9ddc14p+0, -0x1.49da7fda85000p-2},
    {0x1.5f66452c65c4cp+0, -0x1.445923989a800p-2},
    {0x1.5d867b5912c4fp+0, -0x1.3edf439b0b800p-2},
    {0x1.5babccb5b90dep+0, -0x1.396ce448f7000p-2},
    {0x1.59d61f2d91a78p+0, -0x1.3401e17bda000p-2},
    {0x1.5805612465687p+0, -0x1.2e9e2ef468000p-2},
    {0x1.56397cee76bd3p+0, -0x1.2941b3830e000p-2},
    {0x1.54725e2a77f93p+0, -0x1.23ec58cda8800p-2},
    {0x1.52aff42064583p+0, -0x1.1e9e129279000p-2},
    {0x1.50f22dbb2bddfp+0, -0x1.1956d2b48f800p-2},
    {0x1.4f38f4734ded7p+0, -0x1.141679ab9f800p-2},
    {0x1.4d843cfde2840p+0, -0x1.0edd094ef9800p-2},
    {0x1.4bd3ec078a3c8p+0, -0x1.09aa518db1000p-2},
    {0x1.4a27fc3e0258ap+0, -0x1.047e65263b800p-2},
    {0x1.4880524d48434p+0, -0x1.feb224586f000p-3},
    {0x1.46dce1b192d0bp+0, -0x1.f474a7517b000p-3},
    {0x1.453d9d3391854p+0, -0x1.ea4443d103000p-3},
    {0x1.43a2744b4845ap+0, -0x1.e020d44e9b000p-3},
    {0x1.420b54115f8fbp+0, -0x1.d60a22977f000p-3},
    {0x1.40782da3ef4b1p+0, -0x1.cc00104959000p-3},
    {0x1.3ee8f5d57fe8fp+0, -0x1.c202956891000p-3},
    {0x1.3d5d9a00b4ce9p+0, -0x1.b81178d811000p-3},
    {0x1.3bd60c010c12bp+0, -0x1.ae2c9ccd3d000p-3},
    {0x1.3a5242b75dab8p+0, -0x1.a45402e129000p-3},
    {0x1.38d22cd9fd002p+0, -0x1.9a877681df000p-3},
    {0x1.3755bc5847a1cp+0, -0x1.90c6d69483000p-3},
    {0x1.35dce49ad36e2p+0, -0x1.87120a645c000p-3},
    {0x1.34679984dd440p+0, -0x1.7d68fb4143000p-3},
    {0x1.32f5cceffcb24p+0, -0x1.73cb83c627000p-3},
    {0x1.3187775a10d49p+0, -0x1.6a39a9b376000p-3},
    {0x1.301c8373e3990p+0, -0x1.60b3154b7a000p-3},
    {0x1.2eb4ebb95f841p+0, -0x1.5737d76243000p-3},
    {0x1.2d50a0219a9d1p+0, -0x1.4dc7b8fc23000p-3},
    {0x1.2bef9a8b7fd2ap+0, -0x1.4462c51d20000p-3},
    {0x1.2a91c7a0c1babp+0, -0x1.3b08abc830000p-3},
    {0x1.293726014b530p+0, -0x1.31b996b490000p-3},
    {0x1.27dfa5757a1f5p+0, -0x1.2875490a44000p-3},
    {0x1.268b39b1d3bbfp+0, -0x1.1f3b9f879a000p-3},
    {0x1.2539d838ff5bdp+0, -0x1.160c8252ca000p-3},
    {0x1.23eb7aac9083bp+0, -0x1.0ce7f57f72000p-3},
    {0x1.22a012ba940b6p+0, -0x1.03cdc49fea000p-3},
    {0x1.2157996cc4132p+0, -0x1.f57bdbc4b8000p-4},
    {0x1.201201dd2fc9bp+0, -0x1.e370896404000p-4},
    {0x1.1ecf4494d480bp+0, -0x1.d17983ef94000p-4},
    {0x1.1d8f5528f6569p+0, -0x1.bf9674ed8a000p-4},
    {0x1.1c52311577e7cp+0, -0x1.adc79202f6000p-4},
    {0x1.1b17c74cb26e9p+0, -0x1.9c0c3e7288000p-4},
    {0x1.19e010c2c1ab6p+0, -0x1.8a646b372c000p-4},
    {0x1.18ab07bb670bdp+0, -0x1.78d01b3ac0000p-4},
    {0x1.1778a25efbcb6p+0, -0x1.674f145380000p-4},
    {0x1.1648d354c31dap+0, -0x1.55e0e6d878000p-4},
    {0x1.151b990275fddp+0, -0x1.4485cdea1e000p-4},
    {0x1.13f0ea432d24cp+0, -0x1.333d94d6aa000p-4},
    {0x1.12c8b7210f9dap+0, -0x1.22079f8c56000p-4},
    {0x1.11a3028ecb531p+0, -0x1.10e4698622000p-4},
    {0x1.107fbda8434afp+0, -0x1.ffa6c6ad20000p-5},
    {0x1.0f5ee0f4e6bb3p+0, -0x1.dda8d4a774000p-5},
    {0x1.0e4065d2a9fcep+0, -0x1.bbcece4850000p-5},
    {0x1.0d244632ca521p+0, -0x1.9a1894012c000p-5},
    {0x1.0c0a77ce2981ap+0, -0x1.788583302c000p-5},
    {0x1.0af2f83c636d1p+0, -0x1.5715e67d68000p-5},
    {0x1.09ddb98a01339p+0, -0x1.35c8a49658000p-5},
    {0x1.08cabaf52e7dfp+0, -0x1.149e364154000p-5},
    {0x1.07b9f2f4e28fbp+0, -0x1.e72c082eb8000p-6},
    {0x1.06ab58c358f19p+0, -0x1.a55f152528000p-6},
    {0x1.059eea5ecf92cp+0, -0x1.63d62cf818000p-6},
    {0x1.04949cdd12c90p+0, -0x1.228fb8caa0000p-6},
    {0x1.038c6c6f0ada9p+0, -0x1.c317b20f90000p-7},
    {0x1.02865137932a9p+0, -0x1.419355daa0000p-7},
    {0x1.0182427ea7348p+0, -0x1.81203c2ec0000p-8},
    {0x1.008040614b195p+0, -0x1.0040979240000p-9},
    {0x1.fe01ff726fa1ap-1, 0x1.feff384900000p-9},
    {0x1.fa11cc261ea74p-1, 0x1.7dc41353d0000p-7},
    {0x1.f6310b081992ep-1, 0x1.3cea3c4c28000p-6},
    {0x1.f25f63ceeadcdp-1, 0x1.b9fc114890000p-6},
    {0x1.ee9c8039113e7p-1, 0x1.1b0d8ce110000p-5},
    {0x1.eae8078cbb1abp-1, 0x1.58a5bd001c000p-5},
    {0x1.e741aa29d0c9bp-1, 0x1.95c8340d88000p-5},
    {0x1.e3a91830a99b5p-1, 0x1.d276aef578000p-5},
    {0x1.e01e009609a56p-1, 0x1.07598e598c000p-4},
    {0x1.dca01e577bb98p-1, 0x1.253f5e30d2000p-4},
    {0x1.d92f20b7c9103p-1, 0x1.42edd8b380000p-4},
    {0x1.d5cac66fb5ccep-1, 0x1.606598757c000p-4},
    {0x1.d272caa5ede9dp-1, 0x1.7da76356a0000p-4},
    {0x1.cf26e3e6b2ccdp-1, 0x1.9ab434e1c6000p-4},
    {0x1.cbe6da2a77902p-1, 0x1.b78c7bb0d6000p-4},
    {0x1.c8b266d37086dp-1, 0x1.d431332e72000p-4},
    {0x1.c5894bd5d5804p-1, 0x1.f0a3171de6000p-4},
    {0x1.c26b533bb9f8cp-1, 0x1.067152b914000p-3},
    {0x1.bf583eeece73fp-1, 0x1.147858292b000p-3},
    {0x1.bc4fd75db96c1p-1, 0x1.2266ecdca3000p-3},
    {0x1.b951e0c864a28p-1, 0x1.303d7a6c55000p-3},
    {0x1.b65e2c5ef3e2cp-1, 0x1.3dfc33c331000p-3},
    {0x1.b374867c9888bp-1, 0x1.4ba366b7a8000p-3},
    {0x1.b094b211d304ap-1, 0x1.5933928d1f000p-3},
    {0x1.adbe885f2ef7ep-1, 0x1.66acd2418f000p-3},
    {0x1.aaf1d31603da2p-1, 0x1.740f8ec669000p-3},
    {0x1.a82e63fd358a7p-1, 0x1.815c0f51af000p-3},
    {0x1.a5740ef09738bp-1, 0x1.8e92954f68000p-3},
    {0x1.a2c2a90ab4b27p-1, 0x1.9bb3602f84000p-3},
    {0x1.a01a01393f2d1p-1, 0x1.a8bed1c2c0000p-3},
    {0x1.9d79f24db3c1bp-1, 0x1.b5b515c01d000p-3},
    {0x1.9ae2505c7b190p-1, 0x1.c2967ccbcc000p-3},
    {0x1.9852ef297ce2fp-1, 0x1.cf635d5486000p-3},
    {0x1.95cbaeea44b75p-1, 0x1.dc1bd3446c000p-3},
    {0x1.934c69de74838p-1, 0x1.e8c01b8cfe000p-3},
    {0x1.90d4f2f6752e6p-1, 0x1.f5509c0179000p-3},
    {0x1.8e6528effd79dp-1, 0x1.00e6c121fb800p-2},
    {0x1.8bfce9fcc007cp-1, 0x1.071b80e93d000p-2},
    {0x1.899c0dabec30ep-1, 0x1.0d46b9e867000p-2},
    {0x1.87427aa2317fbp-1, 0x1.13687334bd000p-2},
    {0x1.84f00acb39a08p-1, 0x1.1980d67234800p-2},
    {0x1.82a49e8653e55p-1, 0x1.1f8ffe0cc8000p-2},
    {0x1.8060195f40260p-1, 0x1.2595fd7636800p-2},
    {0x1.7e22563e0a329p-1, 0x1.2b9300914a800p-2},
    {0x1.7beb377dcb5adp-1, 0x1.3187210436000p-2},
    {0x1.79baa679725c2p-1, 0x1.377266dec1800p-2},
    {0x1.77907f2170657p-1, 0x1.3d54ffbaf3000p-2},
    {0x1.756cadbd6130cp-1, 0x1.432eee32fe000p-2},
};

// __log, FMA variant, for normal positive finite x
PHMM_HD double glibc_log_normal(double x)
{
    const double Ln2hi = 0x1.62e42fefa3800p-1, Ln2lo = 0x1.ef35793c76730p-45;
    const double A0 = -0x1.0000000000001p-1, A1 = 0x1.555555551305bp-2, A2 = -0x1.fffffffeb4590p-3, A3 = 0x1.999b324f10111p-3,
                 A4 = -0x1.55575e506c89fp-3;
    const uint64_t ix = l10_bits64(x);
    if (ix - 0x3fee000000000000ull < 0x3090000000000ull) {        // 1 - 2^-4 <= x < 1 + 0x1.09p-4: separate polynomial
        const double B0 = -0x1.0000000000000p-1, B1 = 0x1.5555555555577p-2, B2 = -0x1.ffffffffffdcbp-3, B3 = 0x1.999999995dd0cp-3,
                     B4 = -0x1.55555556745a7p-3, B5 = 0x1.24924a344de30p-3, B6 = -0x1.fffffa4423d65p-4, B7 = 0x1.c7184282ad6cap-4,
                     B8 = -0x1.999eb43b068ffp-4, B9 = 0x1.78182f7afd085p-4, B10 = -0x1.5521375d145cdp-4;
        if (ix == 0x3ff0000000000000ull) return 0.0;
        const double r = l10_sub(x, 1.0);
        double p1 = l10_fma(r, B2, B1);
        double p2 = l10_fma(r, B5, B4);
        const double r2 = l10_mul(r, r);
        double p3 = l10_fma(r, B8, B7);
        p1 = l10_fma(r2, B3, p1);                                  // B1 + r B2 + r2 B3
        p2 = l10_fma(r2, B6, p2);                                  // B4 + r B5 + r2 B6
        const double r3 = l10_mul(r, r2);
        p3 = l10_fma(r2, B9, p3);                                  // B7 + r B8 + r2 B9
        p3 = l10_fma(r3, B10, p3);
        p3 = l10_fma(p3, r3, p2);
        p3 = l10_fma(p3, r3, p1);                                  // the polynomial; y = r3 * p3 below
        const double t = l10_fma(r, 0x1p27, r);                    // r + r 2^27, one rounding
        const double rhi = l10_fma(-0x1p27, r, t);                 // t - r 2^27
        const double rhi2 = l10_mul(rhi, rhi);
        const double rlo = l10_sub(r, rhi);
        const double hi = l10_fma(rhi2, B0, r);                    // r + rhi^2 B0, one rounding
        const double rmh = l10_sub(r, hi);
        const double rsum = l10_add(r, rhi);
        double lo = l10_fma(rhi2, B0, rmh);                        // (r - hi) + rhi^2 B0
        lo = l10_fma(l10_mul(B0, rlo), rsum, lo);                  // + B0 rlo (rhi + r)
        return l10_add(hi, l10_fma(p3, r3, lo));
    }
    const uint64_t tmp = ix - 0x3fe6000000000000ull;               // OFF
    const int i = (int)((tmp >> 45) & 127u);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
    const double invc = kLogTab[i][0], logc = kLogTab[i][1];
    const double z = l10_double(iz), kd = (double)k;
    const double w = l10_fma(kd, Ln2hi, logc);
    const double r = l10_fma(z, invc, -1.0);
    const double q12 = l10_fma(r, A2, A1);                         // A1 + r A2
    const double hi = l10_add(r, w);
    const double r2 = l10_mul(r, r);
    double lo = l10_add(l10_sub(w, hi), r);
    lo = l10_fma(kd, Ln2lo, lo);
    const double r3 = l10_mul(r, r2);
    const double q34 = l10_fma(r, A4, A3);                         // A3 + r A4
    lo = l10_fma(r2, A0, lo);
    const double q = l10_fma(q34, r2, q12);
    return l10_add(l10_fma(r3, q, lo), hi);
}

// __ieee754_log10 for x >= 0 (zero, subnormals, inf, NaN included)
PHMM_HD double glibc_log10(double x)
{
    const double two54 = 1.80143985094819840000e+16, ivln10 = 0x1.bcb7b1526e50ep-2, log10_2hi = 0x1.34413509f6000p-2,
                 log10_2lo = 0x1.9fef311f12b36p-42;
    int64_t hx = (int64_t)l10_bits64(x);
    int32_t k = -1023;
    if (hx <= 0x000fffffffffffffll) {                               // x < 2^-1022
        if ((hx & 0x7fffffffffffffffll) == 0) return -two54 / 0.0;  // log(+-0) = -inf
        k = -1077; x = l10_mul(x, two54);                           // subnormal: scale up
        hx = (int64_t)l10_bits64(x);
    }
    if ((uint64_t)hx > 0x7fefffffffffffffull) return x + x;
    k += (int32_t)(hx >> 52);
    const int32_t i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
    const uint64_t mx = ((uint64_t)hx & 0x000fffffffffffffull) | ((uint64_t)(0x3ff - i) << 52);
    const double y = (double)(k + i);
    const double lg = glibc_log_normal(l10_double(mx));
    const double z = l10_add(l10_mul(lg, ivln10), l10_mul(y, log10_2lo));
    return l10_add(z, l10_mul(y, log10_2hi));
}

}  // namespace phmm
