// ikb_common.cuh -- shared device-side definitions of libikb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ikb200.h"

#define IKB_FULL_MASK 0xffffffffu
#define IKB_I64_MAX 0x7fffffffffffffffLL

// Robot / solver constants handed to every kernel by value (kernel parameter space, broadcast
// through the constant bank).  Derived once on the host from ikb_config.
// d / sqrt(x) constants of ikb_rsqrt_times (below): d, d/2, 3d/8 -- precomputed on the host for every link length so
// that the kernels use them straight from the parameter block
struct IkbScaledRsqrt {
    double d, d_half, d_38;
};

struct IkbRobot {
    // seed chain of FabrikInverseKinematics.ikine (reference inverse.py:123-130): FK of
    // [theta_1, dh[0][1], dh[0][2], dh[0][3]].  Rz(theta_1) is the left-most factor of the DH
    // product, so the chain is the theta_1 = 0 chain rotated about z by theta_1 = atan2(y, x);
    // when that chain lies in the x-z plane (`planar`), chain and target share the vertical plane
    // through the z axis and the whole solve is 2-D in (r, z).
    double seed_r[4], seed_z[4];
    double seed_xyz[12];  // the theta_1 = 0 chain in 3-D (generic path)
    double seed_ab;       // distance origin -> first joint (|AB| of inverse.py:68)
    double seed_ab2, half_inv_ab;  // |AB|^2 and 1 / (2 |AB|) for the cosine of inverse.py:77-81
    // the three cosines of inverse.py:77-100 when the segments have their link lengths (any chain after >= 1 pass):
    // cos_k = (cos_sum[k] - opposite^2) * cos_inv[k], sums |AB|^2 + d1^2, d1^2 + d2^2, d2^2 + d3^2 and
    // 1 / (2 |AB| d1), 1 / (2 d1 d2), 1 / (2 d2 d3)
    double cos_sum[3], cos_inv[3];
    double links[4];      // joints_distances
    double limits[6];     // xlo, xhi, ylo, yhi, zlo, zhi
    double tol;
    // convergence bands on squared lengths (start end: links[0], goal end: links[3]), see fabrik.cu
    double band_lo2[2], band_hi2[2];
    float band_lo2_f[2], band_hi2_f[2];  // the same edges rounded to fp32 (fp32 iterate mode)
    int max_iter;
    int planar;
    int zero_iter;        // tol >= 1 or max_iter <= 0: the reference's while loop never runs
    // forward kinematics, reference forward.py:62-70: T_i = Rz(th_i) Tz(eps_i) Tx(a_i) Rx(alpha_i)
    double eps[4], a[4], cos_alpha[4], sin_alpha[4], alpha[4];
    int fk_planar_tail;   // alpha[1..3] == 0: joints 2..4 rotate about parallel axes (closed-form FK)
    // the closed form's constants: a[0..3], eps[0], eps[1]+eps[2]+eps[3], cos/sin alpha[0] -- in both precisions, so
    // that the fp32 kernels read them from the constant bank instead of converting doubles per row
    double fkc[8];
    float fkc_f[8];
    IkbScaledRsqrt link_k[4];  // per link length: see ikb_rsqrt_times
};

// Device-side statistics block; host mirror is ikb_stats.  first_* start at IKB_I64_MAX.
struct IkbDeviceStats {
    unsigned long long n_solved;
    unsigned long long sum_iterations;
    unsigned long long n_iter_capped;
    long long first_out_of_limits;
    long long first_zero_division;
    long long first_domain_error;
    long long first_fk_angle_range;
    double sum_fk_error;
    unsigned long long n_fk_error;
};

__device__ __forceinline__ unsigned ikb_lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// AoS target row i as doubles from an fp32 or fp64 buffer.
__device__ __forceinline__ void ikb_load_xyz(const void *xyz, int f64, long long i, double &x,
                                             double &y, double &z)
{
    if (f64) {
        const double *p = reinterpret_cast<const double *>(xyz) + 3 * i;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
    } else {
        const float *p = reinterpret_cast<const float *>(xyz) + 3 * i;
        x = (double)__ldg(p); y = (double)__ldg(p + 1); z = (double)__ldg(p + 2);
    }
}

// reference inverse.py:26-35: any axis < lo or > hi (NaN compares False and passes)
__device__ __forceinline__ bool ikb_out_of_limits(const IkbRobot &rc, double x, double y, double z)
{
    return (x < rc.limits[0]) | (x > rc.limits[1]) | (y < rc.limits[2]) | (y > rc.limits[3]) |
           (z < rc.limits[4]) | (z > rc.limits[5]);
}

__device__ __forceinline__ void ikb_store_angles(void *out, int f64, long long i, const double th[4])
{
    if (f64) {
        double2 *p = reinterpret_cast<double2 *>(out) + 2 * i;
        p[0] = make_double2(th[0], th[1]);
        p[1] = make_double2(th[2], th[3]);
    } else {
        reinterpret_cast<float4 *>(out)[i] =
            make_float4((float)th[0], (float)th[1], (float)th[2], (float)th[3]);
    }
}

// 1/sqrt(x): one MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22 relative) followed by one
// third-order correction y(1 + e/2 + 3e^2/8), e = 1 - x y^2 -- 5 DP instructions, result good to
// ~1 ulp for normal x.  x == 0 yields inf * 0 -> NaN downstream, which is how a zero-length segment
// (ZeroDivisionError in reference point.py:40) is detected; no slow-path branches.
// (A second-order step would save one instruction of five and was measured 7 % faster, but its 4e-14
// relative error flips the reference's round(cos, 8) on ~2e-5 of the angles, which for nearly straight
// joints -- every unreachable target -- can move an angle by more than the 1e-4 rad bar: not taken.)
__device__ __forceinline__ double ikb_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = y * y;
    double e = fma(-x, t, 1.0);
    double p = fma(e, 0.375, 0.5);
    double q = y * e;
    return fma(q, p, y);
}

__device__ __forceinline__ float ikb_rsqrt(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// d / sqrt(x) with the same seed and the same third-order correction, the factor folded into the correction's
// constants:  d y (1 + e/2 + 3 e^2 / 8) = y (d + e (d/2 + (3d/8) e)) -- 5 DP instructions for the scaled result
// instead of 6 (rsqrt, then times d).  The pass loop of K1 is bound by the fp64 pipe, so this is 1 of 12 per update.
__host__ __device__ inline IkbScaledRsqrt ikb_scaled_rsqrt_constants(double d) { return IkbScaledRsqrt{d, 0.5 * d, 0.375 * d}; }
__device__ __forceinline__ double ikb_rsqrt_times(double x, const IkbScaledRsqrt &k)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = y * y;
    const double e = fma(-x, t, 1.0);
    const double p = fma(e, k.d_38, k.d_half);
    const double w = fma(e, p, k.d);
    return y * w;
}

template <typename T>
__device__ __forceinline__ T ikb_warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(IKB_FULL_MASK, v, o);
    return v;
}
