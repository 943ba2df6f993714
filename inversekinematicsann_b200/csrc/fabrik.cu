// fabrik.cu -- K1: batched FABRIK inverse kinematics + fp64 angle extraction (sm_100a).
//
// Replaces, per target, reference FabrikInverseKinematics.ikine (inverse.py:115-139):
//   check_limits (inverse.py:26-35) -> Fabrik.calculate (fabrik.py:44-67, built from
//   get_point_between / get_distance_between, point.py:25-45) -> __get_angles (inverse.py:54-112).
//
// Design (see DESIGN.md "K1"):
//  * One chain per lane, chain state in registers.  The seed chain and the target lie in one
//    vertical plane through the z axis, so the iteration runs in 2-D (r, z) coordinates.
//  * P0 is pinned to the start joint and P3 is only needed at the end, so one iteration is
//    4 point-at-distance updates + 5 reciprocal square roots; the two convergence errors are
//    | |S - b1| - d0 | and | |T - f2| - d3 | (algebraically the reference's |b0 - S|, |f3 - T|).
//  * Iteration counts are bimodal (3..8 inside the workspace, exactly max_iter outside), so lanes
//    are REFILLED: a lane that converges parks its chain in a per-warp shared-memory queue and
//    takes the next target from a pre-staged input queue; the expensive fp64 epilogue (8 sqrt,
//    3 acos, atan2) runs only when 32 parked chains are available, i.e. always at full warp width.
//  * Work is handed out in chunks of IKB_FABRIK_CHUNK consecutive targets from a global counter
//    (persistent warps, no tail imbalance); row i of the output always belongs to row i of the input.
#include <cstdlib>

#include "fk_device.cuh"

#define IKB_FABRIK_CHUNK 256
#define IKB_FABRIK_WARPS 8
#define IKB_Q 64  // per-warp input queue capacity (ring), power of two
#ifndef IKB_FABRIK_MIN_CTAS
#define IKB_FABRIK_MIN_CTAS 3
#endif
#ifndef IKB_FABRIK_MIN_CTAS2
#define IKB_FABRIK_MIN_CTAS2 2
#endif
#ifndef IKB_FABRIK_CHAINS
#define IKB_FABRIK_CHAINS 1
#endif

namespace {

struct FabrikArgs {
    const void *xyz;
    int xyz_f64;
    long long n;
    long long index_base;  // global row of xyz[0] (host pipeline chunks)
    void *angles;
    int angles_f64;
    int *iters;  // nullable
    void *fk_err;  // nullable: ||FK(angles) - target|| per row in the angles' dtype (fused K3)
    int fk_stats;  // accumulate sum_fk_error / n_fk_error even without the per-row array
    IkbDeviceStats *stats;
    unsigned long long *work_counter;
    unsigned long long *work_counter_far;
    double far_thr2;  // > 0: targets with |T - S|^2 above it belong to fabrik_far_kernel (see is_far)
    IkbRobot rc;
};

template <typename Real>
struct PlanarChain {
    Real r1, z1, r2, z2;  // joints 1 and 2; joint 0 is pinned, joint 3 is derived
};

// Convergence band of one end of the chain: |sqrt(n2) - d| > tol  <=>  n2 outside [lo2, hi2] with
// lo2 = max(d - tol, 0)^2, hi2 = (d + tol)^2.  Testing the squared length saves the square root; the two
// forms differ only when n2 sits within an ulp of a band edge (probability ~1e-13 per test in fp64).
template <typename Real>
struct Band {
    Real lo2, hi2;  // precomputed on the host (IkbRobot::band_*): kernel-parameter constants
    __device__ __forceinline__ bool outside(Real n2) const { return (n2 < lo2) | (n2 > hi2); }
};

template <typename Real>
__device__ __forceinline__ Band<Real> make_band(const IkbRobot &rc, int which);
template <>
__device__ __forceinline__ Band<double> make_band<double>(const IkbRobot &rc, int which)
{
    return Band<double>{rc.band_lo2[which], rc.band_hi2[which]};
}
template <>
__device__ __forceinline__ Band<float> make_band<float>(const IkbRobot &rc, int which)
{
    return Band<float>{rc.band_lo2_f[which], rc.band_hi2_f[which]};
}

// One forward-and-backward-reaching pass (reference fabrik.py:60-63) in the (r, z) plane.
// Returns true when the reference's loop condition (start_error > tol or goal_error > tol) holds.
// start_error = |b0 - S| = | |S - b1| - d0 | and goal_error = |f3 - T| = | |T - f2| - d3 |, because b0 and
// f3 are the points at distance d0 / d3 from b1 / f2 towards S / T.
// d / sqrt(n2): fp64 folds the link length into the correction step (ikb_rsqrt_times), fp32 is seed times d
template <typename Real>
struct LinkScale;
template <>
struct LinkScale<double> {
    const IkbScaledRsqrt &k;  // in the kernel's parameter block: the DFMAs take the constants as c[][] operands
    __device__ __forceinline__ explicit LinkScale(const IkbScaledRsqrt &kk) : k(kk) {}
    __device__ __forceinline__ double over_sqrt(double n2) const { return ikb_rsqrt_times(n2, k); }
};
template <>
struct LinkScale<float> {
    float d;
    __device__ __forceinline__ explicit LinkScale(const IkbScaledRsqrt &kk) : d((float)kk.d) {}
    __device__ __forceinline__ float over_sqrt(float n2) const { return d * ikb_rsqrt(n2); }
};

template <typename Real>
__device__ __forceinline__ bool fabrik_pass(PlanarChain<Real> &c, Real Tr, Real Tz, Real R0, Real Z0,
                                            const LinkScale<Real> &d1, const LinkScale<Real> &d2,
                                            const Band<Real> &start_band, const Band<Real> &goal_band)
{
    // backward (fabrik.py:19-29): b3 = T, b2 = PB(b3, P2, d2), b1 = PB(b2, P1, d1), b0 = PB(b1, P0, d0)
    // (this file is compiled with --fmad=false: every fused multiply-add is written out, so that a row's result does
    //  not depend on which kernel, or which inlined copy of this function, happens to process it)
    Real dr = c.r2 - Tr, dz = c.z2 - Tz;
    Real s = d2.over_sqrt(fma(dz, dz, dr * dr));
    const Real b2r = fma(s, dr, Tr), b2z = fma(s, dz, Tz);
    dr = c.r1 - b2r; dz = c.z1 - b2z;
    s = d1.over_sqrt(fma(dz, dz, dr * dr));
    const Real b1r = fma(s, dr, b2r), b1z = fma(s, dz, b2z);
    dr = R0 - b1r; dz = Z0 - b1z;
    Real n2 = fma(dz, dz, dr * dr);
    const bool start_off = start_band.outside(n2);                       // fabrik.py:61
    // forward (fabrik.py:32-42): f0 = S, f1 = PB(f0, b1, d1), f2 = PB(f1, b2, d2), f3 = PB(f2, b3, d3)
    s = d1.over_sqrt(n2);
    c.r1 = fma(-s, dr, R0); c.z1 = fma(-s, dz, Z0);  // (b1 - S) = -(dr, dz)
    dr = b2r - c.r1; dz = b2z - c.z1;
    s = d2.over_sqrt(fma(dz, dz, dr * dr));
    c.r2 = fma(s, dr, c.r1); c.z2 = fma(s, dz, c.z1);
    dr = Tr - c.r2; dz = Tz - c.z2;
    n2 = fma(dz, dz, dr * dr);                // f3 itself is only needed after the last pass
    return start_off | goal_band.outside(n2);                            // fabrik.py:63
}

#ifndef IKB_FABRIK_IDLE_T
#define IKB_FABRIK_IDLE_T 8  // lanes allowed to sit idle before the warp leaves its inner loop to refill
#endif

// sqrt(x) for x >= 0 through the reciprocal square root (2 ulp); exact 0 for x == 0
__device__ __forceinline__ double fast_sqrt(double x)
{
    const double s = x * ikb_rsqrt(x);
    return x == 0.0 ? 0.0 : s;
}

__device__ __forceinline__ double dist2d_sq(double ar, double az, double br, double bz)
{
    const double dr = ar - br, dz = az - bz;
    return fma(dz, dz, dr * dr);
}

// round(x, 8) of reference inverse.py:81,92,100: q = rint(x * 1e8) is an integer, q / 1e8 is formed as
// q * 1e-8 plus one exact-residual correction (1e8 is exactly representable), which reproduces the
// correctly rounded quotient np.round / Python's round give, in particular exactly +-1.0 at q = +-1e8.
__device__ __forceinline__ double round8(double x)
{
    const double q = rint(x * 1e8);
    const double r = q * 1e-8;
    return fma(fma(-r, 1e8, q), 1e-8, r);
}

// asin(s) for 0 <= s <= 0.5 given t = s^2: s + s t P(t), P fitted by tools/fit_asin_poly.py
// (degree 9, max abs error 1.2e-15).  Shared core of the acos / atan2 replacements below: the CUDA
// library versions cost ~4x more instructions and the epilogue is the second largest consumer of
// issue slots in this kernel.
__constant__ double c_asin_poly[10] = {
    0.16666666666025234, 0.07500000093993356, 0.04464280537685713, 0.03038341968562146,
    0.022347394121891163, 0.017613143412093663, 0.012211296489012176, 0.01903265892950786,
    -0.009325464322296697, 0.03306665466887422};

__device__ __forceinline__ double asin_core(double s, double t)
{
    double p = c_asin_poly[9];  // constant-bank operands: no immediates to materialise
#pragma unroll
    for (int i = 8; i >= 0; --i)
        p = fma(p, t, c_asin_poly[i]);
    return fma(s * t, p, s);
}

// acos(c), |c| <= 1 (NaN outside, like math.acos raising ValueError upstream); abs error < 3e-15.
// |c| <= 1/2: pi/2 - asin(c);  otherwise 2 asin(sqrt((1 - |c|) / 2)), reflected for c < 0.
__device__ __forceinline__ double acos_fast(double c)
{
    const double PI = 3.141592653589793;
    const double a = fabs(c);
    const bool small = a <= 0.5;
    const double t = small ? c * c : fma(-0.5, a, 0.5);
    const double s = small ? a : fast_sqrt(t);
    const double r = asin_core(s, t);
    const double big = c > 0.0 ? 2.0 * r : PI - 2.0 * r;
    return small ? PI / 2 - copysign(r, c) : big;
}

// atan2(uy, ux) for a UNIT vector (ux, uy); abs error < 3e-15.  The angle to the x axis comes from
// asin(|uy|) when |uy| <= 1/2, from pi/2 - asin(|ux|) when |ux| <= 1/2, else from the half-angle form.
__device__ __forceinline__ double atan2_unit(double uy, double ux)
{
    const double PI = 3.141592653589793;
    const double ay = fabs(uy), ax = fabs(ux);
    const bool m0 = ay <= 0.5, m1 = ax <= 0.5;
    const double t = m0 ? ay * ay : (m1 ? ax * ax : fma(-0.5, ax, 0.5));
    const double s = m0 ? ay : (m1 ? ax : fast_sqrt(t));
    const double r = asin_core(s, t);
    double phi = m0 ? r : (m1 ? PI / 2 - r : 2.0 * r);
    phi = ux < 0.0 ? PI - phi : phi;
    return copysign(phi, uy);
}

// ---- the same two functions on the fp32 pipe, for float32 OUTPUT buffers ------------------------------------------
// K1 is bound by the fp64 pipe while the fp32 pipe idles (ncu: 5 % active).  When the caller's angle buffer is float32
// the result is rounded to 24 bits anyway, so everything after the last step that needs fp64 -- the cosine with its
// 8-decimal rounding (inverse.py:81,92,100: the 1e-8 quantum must not flip) -- runs in fp32: t = (1 - |c|) / 2 is
// still formed in fp64 (no cancellation near |c| = 1, where the reference itself quantises), then sqrt, a degree-4
// polynomial (tools/fit_asin_poly.py 4: 1.6e-9 abs) and the final combination with pi in fp32.  Measured against the
// fp64 tail: <= 4e-7 rad (1.7 ulp of pi in fp32), 250x inside the 1e-4 rad bar.  float64 buffers (the list API, the
// reference's own tests) keep the fp64 tail above (2e-15).
#ifndef IKB_FABRIK_TAIL32
#define IKB_FABRIK_TAIL32 1
#endif
#define IKB_PI_F 3.14159274101257324f

__device__ __forceinline__ float asin_core_f(float s, float t)
{
    float p = 0.0437449409671523f;
    p = fmaf(p, t, 0.023150225160843828f);
    p = fmaf(p, t, 0.045707160478796506f);
    p = fmaf(p, t, 0.07493066952010845f);
    p = fmaf(p, t, 0.1666682263762875f);
    return fmaf(s * t, p, s);
}

__device__ __forceinline__ float sqrt_approx_f(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// acos of an fp64 cosine (NaN outside [-1, 1], like acos_fast)
__device__ __forceinline__ float acos_f32(double c)
{
    const double a = fabs(c);
    const bool small = a <= 0.5;
    const float v = (float)(small ? a : fma(-0.5, a, 0.5));  // |c|, or t = (1 - |c|) / 2 formed in fp64
    const float t = small ? v * v : v;
    const float s = small ? v : sqrt_approx_f(v);
    const float r = asin_core_f(s, t);
    const float big = c > 0.0 ? 2.0f * r : IKB_PI_F - 2.0f * r;
    return small ? 0.5f * IKB_PI_F - (c < 0.0 ? -r : r) : big;
}

// atan2(uy, ux) of a unit vector given in fp64, evaluated in fp32
__device__ __forceinline__ float atan2_unit_f32(double uy64, double ux64)
{
    const float uy = (float)uy64, ux = (float)ux64;
    const float ay = fabsf(uy), ax = fabsf(ux);
    const bool m0 = ay <= 0.5f, m1 = ax <= 0.5f;
    const float h = fmaf(-0.5f, ax, 0.5f);                   // exact: ax >= 1/2 on this branch
    const float t = m0 ? ay * ay : (m1 ? ax * ax : h);
    const float s = m0 ? ay : (m1 ? ax : sqrt_approx_f(h));
    const float r = asin_core_f(s, t);
    float phi = m0 ? r : (m1 ? 0.5f * IKB_PI_F - r : 2.0f * r);
    phi = ux < 0.0f ? IKB_PI_F - phi : phi;
    return copysignf(phi, uy);
}

// In-plane radius of a target and the unit direction of its vertical plane.  Targets on the z axis
// have no direction (the reference's theta_1 is rounding noise there): +x is used.
__device__ __forceinline__ double planar_radius(double x, double y, double &ux, double &uy)
{
    const double n2 = fma(y, y, x * x);
    const double rs = ikb_rsqrt(n2);
    const bool on_axis = (n2 == 0.0);
    ux = on_axis ? 1.0 : x * rs;
    uy = on_axis ? 0.0 : y * rs;
    return on_axis ? 0.0 : n2 * rs;
}

// Out-of-reach split.  With |T - S| > d1 + d2 + d3 + tol the goal error |f3 - T| = ||T - f2| - d3| can never get
// below tol (f2 stays within d1 + d2 of S), so the reference runs exactly max_iter passes on such a target and its
// last verdict is "not converged".  Those targets -- 38 % of a uniform workspace sample, 87 % of its passes -- go to
// fabrik_far_kernel, which runs the passes in lockstep without the per-pass verdict, parking and refill; the
// lane-refill kernel skips them when it stages its input.  Both kernels evaluate this predicate on the raw target,
// with explicitly rounded operations, so that every row is claimed by exactly one of them.
__device__ __forceinline__ bool is_far(double x, double y, double z, double R0, double Z0, double thr2)
{
    const double q = __fma_rn(x, x, __dmul_rn(y, y)), dz = __dadd_rn(z, -Z0);
    double d2;
    if (R0 == 0.0) {  // base on the z axis (every arm of the reference's family): no square root needed
        d2 = __fma_rn(dz, dz, q);
    } else {
        const double dr = __dadd_rn(__dsqrt_rn(q), -R0);
        d2 = __fma_rn(dr, dr, __dmul_rn(dz, dz));
    }
    return (d2 > thr2) & (d2 < 1.0e300);  // false for NaN and for infinite targets (those stop after one pass)
}

// Finish one solved chain: derive the effector, lift to 3-D, extract the four angles in fp64 as
// reference inverse.py:54-112 does, write outputs, raise the per-row flags.
// Out of line on purpose: the fp32 FK needs ~40 registers of its own, and inlined into the epilogue it pushed the
// register allocation of the whole kernel (the solve got 10 % slower even with the error switched off).  Scalars
// only, so nothing of the kernel's parameter block has to be spilled for the call.
__device__ __noinline__ float fused_fk_error_f32(float t0, float t1, float t2, float t3, float tx, float ty, float tz,
                                                 float a0, float a1, float a2, float a3, float eps0, float w, float ca,
                                                 float sa)
{
    const float th[4] = {t0, t1, t2, t3};
    return ikb_fk_error_planar_tail<float>(th, tx, ty, tz, a0, a1, a2, a3, eps0, w, ca, sa);
}
__device__ __noinline__ double fused_fk_error_f64(double t0, double t1, double t2, double t3, double tx, double ty,
                                                  double tz, double a0, double a1, double a2, double a3, double eps0,
                                                  double w, double ca, double sa)
{
    const double th[4] = {t0, t1, t2, t3};
    return ikb_fk_error_planar_tail<double>(th, tx, ty, tz, a0, a1, a2, a3, eps0, w, ca, sa);
}

template <bool FUSE_FK, bool OUT32>
__device__ __forceinline__ void fabrik_epilogue(const FabrikArgs &a, int idx, int k, double r1,
                                                double z1, double r2, double z2, double &fk_sum, unsigned &fk_cnt)
{
    const IkbRobot &rc = a.rc;
    const double PI = 3.141592653589793;
    double x, y, z;
    ikb_load_xyz(a.xyz, a.xyz_f64, idx, x, y, z);
    const long long row = a.index_base + idx;
    if (ikb_out_of_limits(rc, x, y, z))
        atomicMin(&a.stats->first_out_of_limits, row);
    double ux, uy;
    const double Tr = planar_radius(x, y, ux, uy), Tz = z;
    const double R0 = rc.seed_r[0], Z0 = rc.seed_z[0];
    double r3, z3;
    bool zero_div = false;
    if (k == 0) {  // the reference's loop never ran: the chain is the seed chain
        r1 = rc.seed_r[1]; z1 = rc.seed_z[1]; r2 = rc.seed_r[2]; z2 = rc.seed_z[2];
        r3 = rc.seed_r[3]; z3 = rc.seed_z[3];
    } else {       // f3 = PB(f2, T, d3) (fabrik.py:40)
        const double dr = Tr - r2, dz = Tz - z2;
        const double n2 = fma(dz, dz, dr * dr);
        zero_div |= (n2 == 0.0);
        const double s = ikb_rsqrt_times(n2, rc.link_k[3]);
        r3 = fma(s, dr, r2); z3 = fma(s, dz, z2);
    }
    // a NaN chain from finite input can only come from 0 * inf, i.e. a zero-length segment
    const bool finite_in = isfinite(x) & isfinite(y) & isfinite(z);
    zero_div |= finite_in & !(isfinite(r1) & isfinite(z1) & isfinite(r2) & isfinite(z2));
    const bool flip = r3 < 0.0;  // theta_1 = atan2(E.y, E.x) with E = r3 (ux, uy) (inverse.py:60): only the sign of r3 matters
    const double ab = rc.seed_ab;                                        // |AB|, A = origin: a constant
    // Everything on squared lengths: the reference takes seven roots, squares six of them again (pow(ac, 2) etc.)
    // and divides by products of three; a squared rounded root differs from the squared length by at most 2 ulp --
    // the order of this epilogue's own rounding -- so no root is taken except inside the three cosines' rsqrt, and the
    // "bd > dista" test of inverse.py:103 compares squares.
    const double n_bc = dist2d_sq(R0, Z0, r1, z1), n_cd = dist2d_sq(r1, z1, r2, z2), n_de = dist2d_sq(r2, z2, r3, z3);
    const double n_ac = fma(z1, z1, r1 * r1), n_bd = dist2d_sq(R0, Z0, r2, z2), n_ce = dist2d_sq(r1, z1, r3, z3);
    // cos = numerator / (2 |.| |.|) = numerator * rsqrt(product of the squared lengths) / 2: one reciprocal square
    // root per angle instead of two roots and a division (|AB| is a constant of the arm)
    double den2 = n_bc;
    zero_div |= (den2 == 0.0) | (ab == 0.0);
    const double c2 = round8(((rc.seed_ab2 + n_bc) - n_ac) * ikb_rsqrt(den2) * rc.half_inv_ab);   // inverse.py:77-81
    den2 = n_bc * n_cd;
    zero_div |= (den2 == 0.0);
    const double c3 = round8(((n_bc + n_cd) - n_bd) * ikb_rsqrt(den2) * 0.5);            // :90-92
    den2 = n_cd * n_de;
    zero_div |= (den2 == 0.0) | (n_ce == 0.0);
    const double c4 = round8(((n_cd + n_de) - n_ce) * ikb_rsqrt(den2) * 0.5);            // :98-100
    // t4_point_bt = PB(C, E, |CE| / 2) (inverse.py:102): (|CE|/2)/|CE| is exactly 0.5
    const double mr = fma(0.5, r3 - r1, r1), mz = fma(0.5, z3 - z1, z1);
    const bool elbow_neg = (r1 * ux) * (r2 * ux) < 0;                       // :82
    const bool wrist_neg = n_bd > dist2d_sq(R0, Z0, mr, mz);                // :103
    zero_div &= finite_in;
    const bool domain = !zero_div & (fabs(c2) > 1.0 | fabs(c3) > 1.0 | fabs(c4) > 1.0);
    if (zero_div)
        atomicMin(&a.stats->first_zero_division, row);
    if (domain)
        atomicMin(&a.stats->first_domain_error, row);
    if (a.iters)
        a.iters[idx] = k;
    if (OUT32 && IKB_FABRIK_TAIL32) {
        // float32 buffer: the trigonometric tail on the fp32 pipe (see acos_f32)
        float t0 = atan2_unit_f32(flip ? -uy : uy, flip ? -ux : ux);
        t0 = r3 == 0.0 ? 0.0f : (r3 != r3 ? __int_as_float(0x7fc00000) : t0);
        const float acos2 = acos_f32(c2), acos3 = acos_f32(c3), acos4 = acos_f32(c4);
        float t1 = elbow_neg ? 1.5f * IKB_PI_F - acos2 : -(0.5f * IKB_PI_F - acos2);     // :82-85
        float t2 = -(IKB_PI_F - acos3);                                                  // :93
        float t3 = wrist_neg ? -(IKB_PI_F - acos4) : (IKB_PI_F - acos4);                 // :103-108
        if (zero_div)
            t0 = t1 = t2 = t3 = __int_as_float(0x7fc00000);
        reinterpret_cast<float4 *>(a.angles)[idx] = make_float4(t0, t1, t2, t3);
        if (FUSE_FK) {  // fused K3: FK of the angles as stored
            const float e = fused_fk_error_f32(t0, t1, t2, t3, (float)x, (float)y, (float)z, rc.fkc_f[0], rc.fkc_f[1],
                                               rc.fkc_f[2], rc.fkc_f[3], rc.fkc_f[4], rc.fkc_f[5], rc.fkc_f[6], rc.fkc_f[7]);
            if (a.fk_err)
                reinterpret_cast<float *>(a.fk_err)[idx] = e;
            if (isfinite(e)) {
                fk_sum += (double)e;
                ++fk_cnt;
            }
        }
        return;
    }
    double th[4];
    th[0] = atan2_unit(flip ? -uy : uy, flip ? -ux : ux);  // one evaluation: the compiler will not merge two calls under a select
    th[0] = r3 == 0.0 ? 0.0 : (r3 != r3 ? r3 : th[0]);
    const double acos2 = acos_fast(c2);
    th[1] = elbow_neg ? (3 * PI / 2) - acos2 : -(PI / 2 - acos2);  // :82-85
    th[2] = -(PI - acos_fast(c3));                                                        // :93
    const double acos4 = acos_fast(c4);
    th[3] = wrist_neg ? -(PI - acos4) : (PI - acos4);   // :103-108
    if (zero_div)
        th[0] = th[1] = th[2] = th[3] = __longlong_as_double(0x7ff8000000000000LL);
    ikb_store_angles(a.angles, OUT32 ? 0 : 1, idx, th);
    if (FUSE_FK) {  // fused K3: FK of the angles as stored, in the stored precision
        double err;
        if (!OUT32) {
            err = fused_fk_error_f64(th[0], th[1], th[2], th[3], x, y, z, rc.fkc[0], rc.fkc[1], rc.fkc[2], rc.fkc[3],
                                     rc.fkc[4], rc.fkc[5], rc.fkc[6], rc.fkc[7]);
            if (a.fk_err)
                reinterpret_cast<double *>(a.fk_err)[idx] = err;
        } else {
            const float e = fused_fk_error_f32((float)th[0], (float)th[1], (float)th[2], (float)th[3], (float)x, (float)y,
                                               (float)z, rc.fkc_f[0], rc.fkc_f[1], rc.fkc_f[2], rc.fkc_f[3], rc.fkc_f[4],
                                               rc.fkc_f[5], rc.fkc_f[6], rc.fkc_f[7]);
            if (a.fk_err)
                reinterpret_cast<float *>(a.fk_err)[idx] = e;
            err = (double)e;
        }
        if (isfinite(err)) {
            fk_sum += err;
            ++fk_cnt;
        }
    }
}

// OUT_Q: parked-chain ring; < 32 entries wait when a pass starts and a pass parks at most 32 per chain slot
template <typename Real, int OUT_Q>
struct WarpQueues {
    int in_idx[IKB_Q];
    Real in_tr[IKB_Q], in_tz[IKB_Q];
    int out_idx[OUT_Q];
    int out_k[OUT_Q];  // iterations; bit 30 set when the chain stopped on max_iter, not on the tolerance
    Real out_c[4][OUT_Q];
    double fk_sum[32];  // per-lane sums of the fused FK error (kept out of the register file)
    unsigned fk_cnt[32];
};

#define IKB_CAPPED_BIT 0x40000000

// CHAINS = independent chains per lane.  The pass is one long dependency chain (every instruction
// needs the previous result), so a second chain per lane doubles the instruction-level parallelism a
// warp offers the FP64 pipe; its cost is registers (fewer resident warps).
template <typename Real, int CHAINS, bool FUSE_FK, bool OUT32>
__device__ __forceinline__ void fabrik_refill_loop(const FabrikArgs &a, WarpQueues<Real, 32 + 32 * CHAINS> *s_queues)
{
    // per-warp rings: input queue (pre-staged targets) and output queue (parked solved chains);
    // one struct per warp so every access is "warp base + constant + slot"
    constexpr int OUT_Q = 32 + 32 * CHAINS;

    const int lane = threadIdx.x & 31;
    WarpQueues<Real, OUT_Q> &q = s_queues[threadIdx.x >> 5];
    const unsigned lt = ikb_lanemask_lt();
    if (FUSE_FK) {
        q.fk_sum[lane] = 0.0;
        q.fk_cnt[lane] = 0u;
    }
    const IkbRobot &rc = a.rc;
    const Real R0 = (Real)rc.seed_r[0], Z0 = (Real)rc.seed_z[0];
    const LinkScale<Real> d1(rc.link_k[1]), d2(rc.link_k[2]);
    const Band<Real> start_band = make_band<Real>(rc, 0), goal_band = make_band<Real>(rc, 1);
    const int max_iter = rc.max_iter;
    const bool zero_iter = rc.zero_iter != 0;

    bool active[CHAINS], exhausted = false;
    unsigned active_mask[CHAINS];  // warp-uniform copies of `active`
    int idx[CHAINS], k[CHAINS];
    Real Tr[CHAINS], Tz[CHAINS];
    PlanarChain<Real> c[CHAINS];
#pragma unroll
    for (int u = 0; u < CHAINS; ++u) {
        active[u] = false; active_mask[u] = 0; idx[u] = 0; k[u] = 0; Tr[u] = 0; Tz[u] = 0;
        c[u] = PlanarChain<Real>{0, 0, 0, 0};
    }
    long long cur = 0, cur_end = 0;
    int in_head = 0, in_cnt = 0, out_head = 0, out_cnt = 0;
    unsigned long long iters_local = 0;
    unsigned solved_local = 0, capped_local = 0;

    for (;;) {
        // 1. top up the input queue: one coalesced full-warp load of up to 32 targets
        if (in_cnt <= 32 && !exhausted) {
            if (cur == cur_end) {
                unsigned long long base = 0;
                if (lane == 0)
                    base = atomicAdd(a.work_counter, (unsigned long long)IKB_FABRIK_CHUNK);
                base = __shfl_sync(IKB_FULL_MASK, base, 0);
                if ((long long)base >= a.n) {
                    exhausted = true;
                } else {
                    cur = (long long)base;
                    cur_end = min(cur + IKB_FABRIK_CHUNK, a.n);
                }
            }
            if (!exhausted) {
                const int m = (int)min((long long)32, cur_end - cur);
                double x = 0, y = 0, z = 0, ux, uy;
                bool mine = lane < m;
                if (mine) {
                    ikb_load_xyz(a.xyz, a.xyz_f64, cur + lane, x, y, z);
                    if (a.far_thr2 > 0.0)
                        mine = !is_far(x, y, z, rc.seed_r[0], rc.seed_z[0], a.far_thr2);
                }
                const unsigned keep = __ballot_sync(IKB_FULL_MASK, mine);
                if (mine) {
                    const int slot = (in_head + in_cnt + __popc(keep & lt)) & (IKB_Q - 1);
                    q.in_idx[slot] = (int)(cur + lane);
                    q.in_tr[slot] = (Real)planar_radius(x, y, ux, uy);
                    q.in_tz[slot] = (Real)z;
                }
                in_cnt += __popc(keep);
                cur += m;
                __syncwarp();
            }
        }
        // 2. idle chain slots take the next staged targets
        unsigned any_active = 0;
#pragma unroll
        for (int u = 0; u < CHAINS; ++u) {
            if (active_mask[u] != IKB_FULL_MASK && in_cnt > 0) {
                const unsigned need = ~active_mask[u];
                const int rank = __popc(need & lt);
                if (!active[u] && rank < in_cnt) {
                    const int slot = (in_head + rank) & (IKB_Q - 1);
                    idx[u] = q.in_idx[slot];
                    Tr[u] = q.in_tr[slot];
                    Tz[u] = q.in_tz[slot];
                    c[u].r1 = (Real)rc.seed_r[1]; c[u].z1 = (Real)rc.seed_z[1];
                    c[u].r2 = (Real)rc.seed_r[2]; c[u].z2 = (Real)rc.seed_z[2];
                    k[u] = 0;
                    active[u] = true;
                }
                const int take = min(__popc(need), in_cnt);
                in_head = (in_head + take) & (IKB_Q - 1);
                in_cnt -= take;
                active_mask[u] = __ballot_sync(IKB_FULL_MASK, active[u]);
            }
            any_active |= active_mask[u];
        }
        // 3. FABRIK passes (reference fabrik.py:57-65).  Every lane executes the pass -- chain slots without
        //    a live chain compute on stale registers and are ignored -- so the loop body is branch-free;
        //    a chain that finishes is parked at once and the warp only leaves the loop when enough slots
        //    idle to make a refill worthwhile or a full warp of parked chains is ready.
        if (any_active != 0) {
            const int idle_limit = (exhausted && in_cnt == 0) ? 32 * CHAINS + 1 : IKB_FABRIK_IDLE_T * CHAINS;
            int n_idle = 0;
#pragma unroll
            for (int u = 0; u < CHAINS; ++u)
                n_idle += 32 - __popc(active_mask[u]);
            bool leave = false;
            do {
                bool more[CHAINS], fin[CHAINS];
                unsigned m[CHAINS], many = 0;
#pragma unroll
                for (int u = 0; u < CHAINS; ++u) {
                    more[u] = false;
                    if (!zero_iter) {
                        more[u] = fabrik_pass(c[u], Tr[u], Tz[u], R0, Z0, d1, d2, start_band, goal_band);
                        ++k[u];
                    }
                }
#pragma unroll
                for (int u = 0; u < CHAINS; ++u) {
                    fin[u] = active[u] & (!more[u] | (k[u] >= max_iter));
                    m[u] = __ballot_sync(IKB_FULL_MASK, fin[u]);
                    many |= m[u];
                }
                if (many != 0) {
#pragma unroll
                    for (int u = 0; u < CHAINS; ++u) {
                        if (m[u] != 0) {
                            if (fin[u]) {
                                int slot = out_head + out_cnt + __popc(m[u] & lt);
                                slot -= slot >= OUT_Q ? OUT_Q : 0;
                                q.out_idx[slot] = idx[u];
                                q.out_k[slot] = more[u] ? (k[u] | IKB_CAPPED_BIT) : k[u];
                                q.out_c[0][slot] = c[u].r1; q.out_c[1][slot] = c[u].z1;
                                q.out_c[2][slot] = c[u].r2; q.out_c[3][slot] = c[u].z2;
                                active[u] = false;
                            }
                            const int nf = __popc(m[u]);
                            out_cnt += nf;
                            n_idle += nf;
                            active_mask[u] &= ~m[u];
                        }
                    }
                    any_active = 0;
#pragma unroll
                    for (int u = 0; u < CHAINS; ++u)
                        any_active |= active_mask[u];
                    // the parked-chain ring holds IKB_Q entries: leave as soon as one more pass could overflow it
                    leave = (n_idle >= idle_limit) | (out_cnt >= 32) | (any_active == 0);
                }
            } while (!leave);
            __syncwarp();
        }
        // 4. the fp64 epilogue runs on a full warp of parked chains (or on the remainder once all
        //    work is done).  ONE call site: every row goes through the same instruction sequence,
        //    so results do not depend on where in the batch a target sits.
        const bool drained = any_active == 0 && exhausted && in_cnt == 0;
        while (out_cnt >= 32 || (drained && out_cnt > 0)) {
            const int n_take = min(32, out_cnt);
            if (lane < n_take) {
                int slot = out_head + lane;
                slot -= slot >= OUT_Q ? OUT_Q : 0;
                const int k_raw = q.out_k[slot], k_done = k_raw & (IKB_CAPPED_BIT - 1);
                fabrik_epilogue<FUSE_FK, OUT32>(a, q.out_idx[slot], k_done, (double)q.out_c[0][slot], (double)q.out_c[1][slot],
                                (double)q.out_c[2][slot], (double)q.out_c[3][slot], q.fk_sum[lane], q.fk_cnt[lane]);
                iters_local += (unsigned)k_done;
                ++solved_local;
                capped_local += (k_raw & IKB_CAPPED_BIT) ? 1u : 0u;
            }
            out_head += n_take;
            out_head -= out_head >= OUT_Q ? OUT_Q : 0;
            out_cnt -= n_take;
            __syncwarp();
        }
        if (drained && out_cnt == 0)
            break;
    }
    // statistics: one atomic per warp and counter
    const unsigned long long it = ikb_warp_sum(iters_local);
    const unsigned sv = ikb_warp_sum(solved_local), cp = ikb_warp_sum(capped_local);
    if (lane == 0 && sv != 0) {
        atomicAdd(&a.stats->sum_iterations, it);
        atomicAdd(&a.stats->n_solved, (unsigned long long)sv);
        if (cp)
            atomicAdd(&a.stats->n_iter_capped, (unsigned long long)cp);
    }
    if (FUSE_FK) {
        const double es = ikb_warp_sum(q.fk_sum[lane]);
        const unsigned ec = ikb_warp_sum(q.fk_cnt[lane]);
        if (lane == 0 && ec != 0) {
            atomicAdd(&a.stats->sum_fk_error, es);
            atomicAdd(&a.stats->n_fk_error, (unsigned long long)ec);
        }
    }
}

template <typename Real, int CHAINS, bool FUSE_FK, bool OUT32>
__global__ void __launch_bounds__(IKB_FABRIK_WARPS * 32, (CHAINS == 1 ? IKB_FABRIK_MIN_CTAS : IKB_FABRIK_MIN_CTAS2))
    fabrik_planar_kernel(const FabrikArgs a)
{
    __shared__ WarpQueues<Real, 32 + 32 * CHAINS> s_queues[IKB_FABRIK_WARPS];
    fabrik_refill_loop<Real, CHAINS, FUSE_FK, OUT32>(a, s_queues);
}

// ---- out-of-reach targets: max_iter passes in lockstep (see is_far) -----------------------------------------
#ifndef IKB_FAR_CHAINS
#define IKB_FAR_CHAINS 2
#endif
#ifndef IKB_FAR_MIN_CTAS
#define IKB_FAR_MIN_CTAS 3
#endif

template <typename Real>
struct FarQueue {
    int idx[32 * IKB_FAR_CHAINS + 32];
    Real tr[32 * IKB_FAR_CHAINS + 32], tz[32 * IKB_FAR_CHAINS + 32];
    Real out_c[4][32 * IKB_FAR_CHAINS];
};

template <typename Real, bool OUT32>
__device__ __forceinline__ void fabrik_far_loop(const FabrikArgs &a, FarQueue<Real> *s_q)
{
    constexpr int CH = IKB_FAR_CHAINS, BATCH = 32 * CH;
    FarQueue<Real> &q = s_q[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt = ikb_lanemask_lt();
    const IkbRobot &rc = a.rc;
    const Real R0 = (Real)rc.seed_r[0], Z0 = (Real)rc.seed_z[0];
    const LinkScale<Real> d1(rc.link_k[1]), d2(rc.link_k[2]);
    const Band<Real> start_band = make_band<Real>(rc, 0), goal_band = make_band<Real>(rc, 1);
    const int max_iter = rc.max_iter;
    long long cur = 0, cur_end = 0;
    int cnt = 0;  // queued targets, always at the front of the arrays
    bool exhausted = false;
    unsigned solved_local = 0;
    double fk_sum = 0.0;
    unsigned fk_cnt = 0;
    for (;;) {
        // stage far targets until a full batch is queued (or the input is exhausted)
        while (cnt < BATCH && !exhausted) {
            if (cur == cur_end) {
                unsigned long long base = 0;
                if (lane == 0)
                    base = atomicAdd(a.work_counter_far, (unsigned long long)IKB_FABRIK_CHUNK);
                base = __shfl_sync(IKB_FULL_MASK, base, 0);
                if ((long long)base >= a.n) {
                    exhausted = true;
                    break;
                }
                cur = (long long)base;
                cur_end = min(cur + IKB_FABRIK_CHUNK, a.n);
            }
            const int m = (int)min((long long)32, cur_end - cur);
            double x = 0, y = 0, z = 0, ux, uy;
            bool mine = false;
            if (lane < m) {
                ikb_load_xyz(a.xyz, a.xyz_f64, cur + lane, x, y, z);
                mine = is_far(x, y, z, rc.seed_r[0], rc.seed_z[0], a.far_thr2);
            }
            const unsigned keep = __ballot_sync(IKB_FULL_MASK, mine);
            if (mine) {
                const int slot = cnt + __popc(keep & lt);
                q.idx[slot] = (int)(cur + lane);
                q.tr[slot] = (Real)planar_radius(x, y, ux, uy);
                q.tz[slot] = (Real)z;
            }
            cnt += __popc(keep);
            cur += m;
            __syncwarp();
        }
        if (cnt == 0)
            break;
        // one batch: chain slot u of lane l takes entry u * 32 + l; empty slots run on a harmless dummy target
        const int nb = min(cnt, BATCH);
        PlanarChain<Real> c[CH];
        Real Tr[CH], Tz[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int e = u * 32 + lane;
            Tr[u] = e < nb ? q.tr[e] : (Real)100;
            Tz[u] = e < nb ? q.tz[e] : (Real)100;
            c[u].r1 = (Real)rc.seed_r[1]; c[u].z1 = (Real)rc.seed_z[1];
            c[u].r2 = (Real)rc.seed_r[2]; c[u].z2 = (Real)rc.seed_z[2];
        }
#pragma unroll 1
        for (int it = 0; it < max_iter; ++it) {
#pragma unroll
            for (int u = 0; u < CH; ++u)
                (void)fabrik_pass(c[u], Tr[u], Tz[u], R0, Z0, d1, d2, start_band, goal_band);
        }
        // park the chains, then ONE epilogue call site for every entry (results must not depend on the chain slot)
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int e = u * 32 + lane;
            q.out_c[0][e] = c[u].r1; q.out_c[1][e] = c[u].z1; q.out_c[2][e] = c[u].r2; q.out_c[3][e] = c[u].z2;
        }
        __syncwarp();
#pragma unroll 1
        for (int e = lane; e < nb; e += 32) {
            fabrik_epilogue<false, OUT32>(a, q.idx[e], max_iter, (double)q.out_c[0][e], (double)q.out_c[1][e],
                                   (double)q.out_c[2][e], (double)q.out_c[3][e], fk_sum, fk_cnt);
            ++solved_local;
        }
        __syncwarp();
        // move the rest (< 32 entries) to the front
        const int rest = cnt - nb;
        int m_idx = 0;
        Real m_tr = 0, m_tz = 0;
        if (lane < rest) { m_idx = q.idx[nb + lane]; m_tr = q.tr[nb + lane]; m_tz = q.tz[nb + lane]; }
        __syncwarp();
        if (lane < rest) { q.idx[lane] = m_idx; q.tr[lane] = m_tr; q.tz[lane] = m_tz; }
        __syncwarp();
        cnt = rest;
    }
    const unsigned sv = ikb_warp_sum(solved_local);
    if (lane == 0 && sv != 0) {
        atomicAdd(&a.stats->sum_iterations, (unsigned long long)sv * (unsigned long long)max_iter);
        atomicAdd(&a.stats->n_solved, (unsigned long long)sv);
        atomicAdd(&a.stats->n_iter_capped, (unsigned long long)sv);
    }
}

// One launch for both populations: a CTA starts in one role and, when the rows of its role are used up, carries on
// in the other, so the two populations balance themselves whatever the mix.  Each role scans the whole input through
// its own chunk counter and claims its rows with is_far.  (Which role a CTA starts in hardly matters -- 17.1 to
// 17.7 ms per 1e8 rows across 0..100 % -- because a lockstep warp is bound by the latency of its own dependent fp64
// chain rather than by the pipe once fewer than six of them share a scheduler.)
template <typename Real, bool OUT32>
__global__ void __launch_bounds__(IKB_FABRIK_WARPS * 32, IKB_FABRIK_MIN_CTAS) fabrik_split_kernel(const FabrikArgs a)
{
    constexpr size_t kNear = sizeof(WarpQueues<Real, 64>) * IKB_FABRIK_WARPS, kFar = sizeof(FarQueue<Real>) * IKB_FABRIK_WARPS;
    __shared__ __align__(16) unsigned char s_raw[kNear > kFar ? kNear : kFar];
#ifndef IKB_SPLIT_FAR_PCT
#define IKB_SPLIT_FAR_PCT 34
#endif
    const bool far_first = blockIdx.x * 100u < gridDim.x * (unsigned)IKB_SPLIT_FAR_PCT;
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
        if ((phase == 0) == far_first)
            fabrik_far_loop<Real, OUT32>(a, reinterpret_cast<FarQueue<Real> *>(s_raw));
        else
            fabrik_refill_loop<Real, 1, false, OUT32>(a, reinterpret_cast<WarpQueues<Real, 64> *>(s_raw));
        __syncthreads();  // the two roles overlay the same shared memory
    }
}

// ---- generic 3-D path ---------------------------------------------------------------------------
// IEEE fp64 with the reference's operation order (point.py:25-45): correctly rounded sqrt and
// division, explicit _rn intrinsics so nothing is contracted into FMAs.  One chain per thread.
struct Vec3 {
    double x, y, z;
};

__device__ __forceinline__ double dist3(const Vec3 &a, const Vec3 &b)
{
    const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

// get_point_between(s, e, L) = s + (L / |s - e|) * (e - s); NaN when the distance is 0
__device__ __forceinline__ Vec3 point_between3(const Vec3 &s, const Vec3 &e, double L)
{
    const double t = __ddiv_rn(L, dist3(s, e));
    Vec3 o;
    o.x = __dadd_rn(s.x, __dmul_rn(t, __dsub_rn(e.x, s.x)));
    o.y = __dadd_rn(s.y, __dmul_rn(t, __dsub_rn(e.y, s.y)));
    o.z = __dadd_rn(s.z, __dmul_rn(t, __dsub_rn(e.z, s.z)));
    return o;
}

struct FabrikGenericArgs {
    const double *init;  // n_init x 4 x 3
    long long n_init;
    const double *goals;  // n x 3
    long long n;
    double *chain_out;  // n x 4 x 3
    int *iters;         // nullable
    IkbDeviceStats *stats;
    IkbRobot rc;
};

__global__ void __launch_bounds__(128) fabrik_generic_kernel(const FabrikGenericArgs a)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n)
        return;
    const double *ip = a.init + (a.n_init == 1 ? 0 : 12 * i);
    Vec3 P[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        P[j] = Vec3{ip[3 * j], ip[3 * j + 1], ip[3 * j + 2]};
    const Vec3 S = P[0];
    const Vec3 T{a.goals[3 * i], a.goals[3 * i + 1], a.goals[3 * i + 2]};
    const double *d = a.rc.links;
    const double tol = a.rc.tol;
    double se = 1.0, ge = 1.0;
    int step = 0;
    while ((se > tol || ge > tol) && a.rc.max_iter > step) {  // fabrik.py:57-59
        const Vec3 b2 = point_between3(T, P[2], d[2]);
        const Vec3 b1 = point_between3(b2, P[1], d[1]);
        const Vec3 b0 = point_between3(b1, P[0], d[0]);
        se = dist3(b0, S);
        P[0] = S;
        P[1] = point_between3(P[0], b1, d[1]);
        P[2] = point_between3(P[1], b2, d[2]);
        P[3] = point_between3(P[2], T, d[3]);
        ge = dist3(P[3], T);
        ++step;
    }
    bool nan_out = false;
    double *o = a.chain_out + 12 * i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        o[3 * j] = P[j].x; o[3 * j + 1] = P[j].y; o[3 * j + 2] = P[j].z;
        nan_out |= !(isfinite(P[j].x) & isfinite(P[j].y) & isfinite(P[j].z));
    }
    if (a.iters)
        a.iters[i] = step;
    if (nan_out && isfinite(T.x) && isfinite(T.y) && isfinite(T.z))
        atomicMin(&a.stats->first_zero_division, i);
    atomicAdd(&a.stats->sum_iterations, (unsigned long long)step);
    atomicAdd(&a.stats->n_solved, 1ULL);
}

// ---- generic ikine: robots whose seed chain leaves the vertical plane (general DH tables) ----------------------
// One target per thread, IEEE fp64 in 3-D with the reference's operation order: seed chain = the theta_1 = 0
// chain rotated about z by atan2(y, x) (inverse.py:123-130), Fabrik.calculate (fabrik.py:44-67), __get_angles
// (inverse.py:54-112).  Slower than the planar kernel (warp-vote exit, no lane refill) -- the reference's own
// robot never takes this path.
struct FabrikGenericIkineArgs {
    const void *xyz;
    int xyz_f64;
    long long n;
    long long index_base;
    void *angles;
    int angles_f64;
    int *iters;
    IkbDeviceStats *stats;
    IkbRobot rc;
};

__global__ void __launch_bounds__(128) fabrik_generic_ikine_kernel(const FabrikGenericIkineArgs a)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = tid < a.n;                 // lanes past the end redo the last row and write nothing,
    const long long i = valid ? tid : a.n - 1;    // so the warp-wide reductions below see full warps
    const IkbRobot &rc = a.rc;
    const double PI = 3.141592653589793;
    double x, y, z;
    ikb_load_xyz(a.xyz, a.xyz_f64, i, x, y, z);
    const long long row = a.index_base + i;
    if (valid && ikb_out_of_limits(rc, x, y, z))
        atomicMin(&a.stats->first_out_of_limits, row);
    double s1, c1;
    sincos(atan2(y, x), &s1, &c1);
    Vec3 P[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double sx = rc.seed_xyz[3 * j], sy = rc.seed_xyz[3 * j + 1];
        P[j] = Vec3{sx * c1 - sy * s1, sx * s1 + sy * c1, rc.seed_xyz[3 * j + 2]};
    }
    const Vec3 S = P[0], T{x, y, z};
    const double *d = rc.links;
    double se = 1.0, ge = 1.0;
    int step = 0;
    while ((se > rc.tol || ge > rc.tol) && rc.max_iter > step) {
        const Vec3 b2 = point_between3(T, P[2], d[2]);
        const Vec3 b1 = point_between3(b2, P[1], d[1]);
        const Vec3 b0 = point_between3(b1, P[0], d[0]);
        se = dist3(b0, S);
        P[0] = S;
        P[1] = point_between3(P[0], b1, d[1]);
        P[2] = point_between3(P[1], b2, d[2]);
        P[3] = point_between3(P[2], T, d[3]);
        ge = dist3(P[3], T);
        ++step;
    }
    const Vec3 A{0.0, 0.0, 0.0};
    const Vec3 &B = P[0], &C = P[1], &D = P[2], &E = P[3];
    double th[4];
    th[0] = atan2(E.y, E.x);
    const double ab = dist3(A, B), bc = dist3(B, C), cd = dist3(C, D), de = dist3(D, E);
    const double ac = dist3(A, C), bd = dist3(B, D), ce = dist3(C, E);
    bool zero_div = false;
    double den = 2 * ab * bc;
    zero_div |= (den == 0.0);
    const double c2 = round8((ab * ab + bc * bc - ac * ac) / den);
    const double acos2 = acos(c2);
    th[1] = (C.x * D.x < 0) ? (3 * PI / 2) - acos2 : -(PI / 2 - acos2);
    den = 2 * bc * cd;
    zero_div |= (den == 0.0);
    const double c3 = round8((bc * bc + cd * cd - bd * bd) / den);
    th[2] = -(PI - acos(c3));
    den = 2 * cd * de;
    zero_div |= (den == 0.0) | (ce == 0.0);
    const double c4 = round8((cd * cd + de * de - ce * ce) / den);
    const double acos4 = acos(c4);
    const Vec3 mid = point_between3(C, E, ce / 2);
    th[3] = (bd > dist3(B, mid)) ? -(PI - acos4) : (PI - acos4);
    const bool finite_in = isfinite(x) & isfinite(y) & isfinite(z);
    bool chain_nan = false;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        chain_nan |= !(isfinite(P[j].x) & isfinite(P[j].y) & isfinite(P[j].z));
    zero_div = (zero_div | chain_nan) & finite_in;
    const bool domain = !zero_div & (fabs(c2) > 1.0 | fabs(c3) > 1.0 | fabs(c4) > 1.0);
    if (zero_div) {
        th[0] = th[1] = th[2] = th[3] = __longlong_as_double(0x7ff8000000000000LL);
        if (valid)
            atomicMin(&a.stats->first_zero_division, row);
    }
    if (domain && valid)
        atomicMin(&a.stats->first_domain_error, row);
    if (valid) {
        ikb_store_angles(a.angles, a.angles_f64, i, th);
        if (a.iters)
            a.iters[i] = step;
    }
    const bool more = (se > rc.tol) | (ge > rc.tol);
    const unsigned long long it = ikb_warp_sum(valid ? (unsigned long long)step : 0ULL);
    const unsigned sv = ikb_warp_sum(valid ? 1u : 0u), cp = ikb_warp_sum((valid && more && step > 0) ? 1u : 0u);
    if ((threadIdx.x & 31) == 0 && sv != 0) {
        atomicAdd(&a.stats->sum_iterations, it);
        atomicAdd(&a.stats->n_solved, (unsigned long long)sv);
        if (cp)
            atomicAdd(&a.stats->n_iter_capped, (unsigned long long)cp);
    }
}

}  // namespace

// ---- launchers (called from capi.cu) --------------------------------------------------------------
// Rows below which the out-of-reach split is not worth a second launch.
#ifndef IKB_SPLIT_MIN_ROWS
#define IKB_SPLIT_MIN_ROWS 65536
#endif

// work_counter: TWO consecutive counters (lane-refill kernel, far kernel).
cudaError_t ikb_launch_fabrik_planar(const void *xyz, int xyz_f64, long long n, long long index_base,
                                     void *angles, int angles_f64, int *iters, void *fk_err, int fk_stats,
                                     int precision, IkbDeviceStats *stats, unsigned long long *work_counter,
                                     const IkbRobot &rc, int num_sms, cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    FabrikArgs a;
    a.xyz = xyz; a.xyz_f64 = xyz_f64; a.n = n; a.index_base = index_base;
    a.angles = angles; a.angles_f64 = angles_f64; a.iters = iters;
    a.fk_err = fk_err; a.fk_stats = fk_stats;
    a.stats = stats; a.work_counter = work_counter; a.rc = rc;
    a.far_thr2 = -1.0;
    a.work_counter_far = nullptr;
    cudaError_t err = cudaMemsetAsync(work_counter, 0, 2 * sizeof(unsigned long long), stream);
    if (err != cudaSuccess)
        return err;
    const bool fuse = fk_err != nullptr || fk_stats != 0;  // the caller checked rc.fk_planar_tail
    static const bool split_enabled = [] { const char *v = std::getenv("IKB_FABRIK_SPLIT"); return !v || v[0] != '0'; }();
    const bool split = split_enabled && !fuse && n >= IKB_SPLIT_MIN_ROWS && !rc.zero_iter && rc.max_iter >= 8;
    // persistent grid: resident CTAs per SM x SM count, trimmed for small batches
    const int per_cta = IKB_FABRIK_WARPS * 32;
    long long want = (n + per_cta - 1) / per_cta;
    long long grid = (long long)num_sms * (IKB_FABRIK_CHAINS == 1 ? IKB_FABRIK_MIN_CTAS : IKB_FABRIK_MIN_CTAS2);
    if (want < grid)
        grid = want;
    if (split) {
        const double reach = rc.links[1] + rc.links[2] + rc.links[3];
        const double thr = reach + rc.tol + 1e-6 * (1.0 + reach);  // margin >> the rounding of |f2 - S| <= d1 + d2
        a.far_thr2 = thr * thr;
        a.work_counter_far = work_counter + 1;
        const void *kargs[] = {&a};
        const void *fn;
        if (precision == IKB_FABRIK_F32)
            fn = angles_f64 ? (const void *)fabrik_split_kernel<float, false> : (const void *)fabrik_split_kernel<float, true>;
        else
            fn = angles_f64 ? (const void *)fabrik_split_kernel<double, false> : (const void *)fabrik_split_kernel<double, true>;
        return cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(per_cta), const_cast<void **>(kargs), 0, stream);
    }
    // kernel variant = (iterate precision, fused FK error, output precision); the trigonometric tail of the angle
    // extraction follows the output buffer's precision (see acos_f32)
    const void *fn;
#define IKB_PICK(REAL)                                                                                               \
    (fuse ? (angles_f64 ? (const void *)fabrik_planar_kernel<REAL, IKB_FABRIK_CHAINS, true, false>                   \
                        : (const void *)fabrik_planar_kernel<REAL, IKB_FABRIK_CHAINS, true, true>)                   \
          : (angles_f64 ? (const void *)fabrik_planar_kernel<REAL, IKB_FABRIK_CHAINS, false, false>                  \
                        : (const void *)fabrik_planar_kernel<REAL, IKB_FABRIK_CHAINS, false, true>))
    if (precision == IKB_FABRIK_F32)
        fn = IKB_PICK(float);
    else
        fn = IKB_PICK(double);
#undef IKB_PICK
    const void *kargs[] = {&a};
    return cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(per_cta), const_cast<void **>(kargs), 0, stream);
}

cudaError_t ikb_launch_fabrik_generic_ikine(const void *xyz, int xyz_f64, long long n, long long index_base,
                                            void *angles, int angles_f64, int *iters, IkbDeviceStats *stats,
                                            const IkbRobot &rc, cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    FabrikGenericIkineArgs a;
    a.xyz = xyz; a.xyz_f64 = xyz_f64; a.n = n; a.index_base = index_base; a.angles = angles;
    a.angles_f64 = angles_f64; a.iters = iters; a.stats = stats; a.rc = rc;
    fabrik_generic_ikine_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t ikb_launch_fabrik_generic(const double *init, long long n_init, const double *goals,
                                      long long n, double *chain_out, int *iters,
                                      IkbDeviceStats *stats, const IkbRobot &rc, cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    FabrikGenericArgs a;
    a.init = init; a.n_init = n_init; a.goals = goals; a.n = n; a.chain_out = chain_out;
    a.iters = iters; a.stats = stats; a.rc = rc;
    const unsigned grid = (unsigned)((n + 127) / 128);
    fabrik_generic_kernel<<<grid, 128, 0, stream>>>(a);
    return cudaGetLastError();
}
