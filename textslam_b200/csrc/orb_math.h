// Deterministic sin/cos shared by the CUDA ORB kernels and (as its checker) the CPU oracle.
// The reference evaluates cos/sin of the keypoint angle with the host libm's float routines
// (src/ORBextractor.cc:112-113), whose last-ulp behaviour differs between libm builds and from CUDA's
// math library. Both sides therefore evaluate the same fixed sequence of IEEE double operations
// (no FMA contraction: explicit _rn intrinsics on the device, -ffp-contract=off on the host) and round
// the result to float; the double value is within ~2e-16 of the true cos/sin.
#pragma once
#if defined(__CUDA_ARCH__)
#define TSL_HD __host__ __device__ __forceinline__
#define TSL_MUL(a, b) __dmul_rn((a), (b))
#define TSL_ADD(a, b) __dadd_rn((a), (b))
#elif defined(__CUDACC__)
#define TSL_HD __host__ __device__ __forceinline__
#define TSL_MUL(a, b) ((a) * (b))
#define TSL_ADD(a, b) ((a) + (b))
#else
#define TSL_HD inline
#define TSL_MUL(a, b) ((a) * (b))
#define TSL_ADD(a, b) ((a) + (b))
#endif

// x in radians, |x| <= ~8. Quadrant reduction with k = round-half-even(x * 2/pi), then Taylor series.
TSL_HD void tsl_det_sincos(double x, double* s_out, double* c_out) {
  const double two_over_pi = 0.63661977236758134308, pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
  const double kd = rint(TSL_MUL(x, two_over_pi));
  const int k = (int)kd;
  double r = TSL_ADD(x, -TSL_MUL(kd, pio2_hi));
  r = TSL_ADD(r, -TSL_MUL(kd, pio2_lo));
  const double r2 = TSL_MUL(r, r);
  // sin r = r (1 - r2/6 (1 - r2/20 (1 - r2/42 ( ... ))))  Horner in the nested form, 9 terms
  double ps = 1.0;
  ps = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 272.0), ps));  // 16*17
  ps = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 210.0), ps));  // 14*15
  ps = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 156.0), ps));  // 12*13
  ps = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 110.0), ps));  // 10*11
  ps = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 72.0), ps));   // 8*9
  ps = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 42.0), ps));   // 6*7
  ps = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 20.0), ps));   // 4*5
  ps = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 6.0), ps));    // 2*3
  const double sr = TSL_MUL(r, ps);
  double pc = 1.0;
  pc = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 306.0), pc));  // 17*18
  pc = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 240.0), pc));  // 15*16
  pc = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 182.0), pc));  // 13*14
  pc = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 132.0), pc));  // 11*12
  pc = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 90.0), pc));   // 9*10
  pc = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 56.0), pc));   // 7*8
  pc = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 30.0), pc));   // 5*6
  pc = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 12.0), pc));   // 3*4
  pc = TSL_ADD(1.0, -TSL_MUL(TSL_MUL(r2, 1.0 / 2.0), pc));    // 1*2
  const double cr = pc;
  switch (k & 3) {
    case 0: *s_out = sr; *c_out = cr; break;
    case 1: *s_out = cr; *c_out = -sr; break;
    case 2: *s_out = -sr; *c_out = -cr; break;
    default: *s_out = -cr; *c_out = sr; break;
  }
}
