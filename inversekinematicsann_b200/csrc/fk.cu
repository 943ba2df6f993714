// fk.cu -- K0 (workspace-limit check), K3 (batched DH forward kinematics / position error) and the
// FMA-pipe microbenchmark (sm_100a).
//
// K3 replaces reference ForwardKinematics.fkine (forward.py:73-94) for batches:
//   T = prod_i Rz(theta_i) Tz(eps_i) Tx(a_i) Rx(alpha_i)                      (forward.py:62-70)
// Only what callers read is produced in the batched form -- the end-effector position T[0:3, 3]
// (cli.py:60, inverse.py:130, tests/forward_unit.py:30) and ||pos - target||.  The chain kernel
// returns all four cumulative 4x4 matrices for the (T, [T1..T4]) return value of fkine.
// HBM-bound: 16 B angles + 12 B target in, 4 B error out per row (fp32 buffers).
#include <cstdlib>

#include "fk_device.cuh"

namespace {

struct FkArgs {
    const void *angles;
    int angles_f64;
    long long n;
    long long index_base;
    void *pos_out;        // nullable, n x 3 (angles dtype)
    const void *targets;  // nullable, n x 3 (xyz dtype)
    int xyz_f64;
    void *err_out;        // nullable, n (angles dtype)
    IkbDeviceStats *stats;
    IkbRobot rc;
};

// FK_ROWS rows per thread and loop trip (2 rows x 4 resident CTAs measured best), all of their loads issued
// before any arithmetic: at 28 B of
// input per row the kernel needs ~5 MB in flight chip-wide to cover HBM latency.
#ifndef FK_ROWS
#define FK_ROWS 2
#endif
#ifndef FK_MINB
#define FK_MINB 4
#endif

// Real = dtype of the angle / position / error buffers, XYZ_F64 = dtype of the target buffer: both compile-time, so the
// row loop carries no dtype branches.
template <typename Real, bool XYZ_F64>
__global__ void __launch_bounds__(256, FK_MINB) fk_kernel(const FkArgs a)
{
    constexpr bool ANG_F64 = sizeof(Real) == 8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    double err_sum = 0.0;
    unsigned err_cnt = 0;
    bool alpha_ok = true;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        alpha_ok &= !(a.rc.alpha[j] < -6.283185307179586) & !(a.rc.alpha[j] > 6.283185307179586);
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < a.n; base += stride * FK_ROWS) {
        Real err_part = 0;  // per-trip partial sum in the working precision, folded into the fp64 total once
        Real th[FK_ROWS][4];
        Real tg[FK_ROWS][3];
#pragma unroll
        for (int u = 0; u < FK_ROWS; ++u) {
            const long long i = base + u * stride;
            if (i < a.n) {
                if (ANG_F64) {
                    const double2 *p = reinterpret_cast<const double2 *>(a.angles) + 2 * i;
                    const double2 v0 = __ldg(p), v1 = __ldg(p + 1);
                    th[u][0] = (Real)v0.x; th[u][1] = (Real)v0.y; th[u][2] = (Real)v1.x; th[u][3] = (Real)v1.y;
                } else {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(a.angles) + i);
                    th[u][0] = (Real)v.x; th[u][1] = (Real)v.y; th[u][2] = (Real)v.z; th[u][3] = (Real)v.w;
                }
                if (a.targets) {
                    if (XYZ_F64) {
                        const double *p = reinterpret_cast<const double *>(a.targets) + 3 * i;
                        tg[u][0] = (Real)__ldg(p); tg[u][1] = (Real)__ldg(p + 1); tg[u][2] = (Real)__ldg(p + 2);
                    } else {  // no detour through double for fp32 buffers
                        const float *p = reinterpret_cast<const float *>(a.targets) + 3 * i;
                        tg[u][0] = (Real)__ldg(p); tg[u][1] = (Real)__ldg(p + 1); tg[u][2] = (Real)__ldg(p + 2);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < FK_ROWS; ++u) {
            const long long i = base + u * stride;
            if (i >= a.n)
                break;
            Real px, py, pz;
            const bool ok = fk_position<Real>(a.rc, th[u], px, py, pz) & alpha_ok;
            if (!ok) {
                px = py = pz = (Real)__int_as_float(0x7fc00000);
                atomicMin(&a.stats->first_fk_angle_range, a.index_base + i);
            }
            if (a.pos_out) {
                if (ANG_F64) {
                    double *o = reinterpret_cast<double *>(a.pos_out) + 3 * i;
                    o[0] = px; o[1] = py; o[2] = pz;
                } else {
                    float *o = reinterpret_cast<float *>(a.pos_out) + 3 * i;
                    o[0] = (float)px; o[1] = (float)py; o[2] = (float)pz;
                }
            }
            if (a.targets) {
                const Real dx = px - tg[u][0], dy = py - tg[u][1], dz = pz - tg[u][2];
                const Real err = sqrt(dx * dx + dy * dy + dz * dz);
                if (a.err_out) {
                    if (ANG_F64)
                        reinterpret_cast<double *>(a.err_out)[i] = err;
                    else
                        reinterpret_cast<float *>(a.err_out)[i] = (float)err;
                }
                if (isfinite(err)) {
                    err_part += err;
                    ++err_cnt;
                }
            }
        }
        err_sum += (double)err_part;
    }
    if (a.targets) {
        err_sum = ikb_warp_sum(err_sum);
        err_cnt = ikb_warp_sum(err_cnt);
        if ((threadIdx.x & 31) == 0 && err_cnt) {
            atomicAdd(&a.stats->sum_fk_error, err_sum);
            atomicAdd(&a.stats->n_fk_error, (unsigned long long)err_cnt);
        }
    }
}

// ---- the common case at HBM speed: fp32 buffers, closed-form arm, position error only ----------------------------------
// The generic kernel above is bound by instruction issue (194 instructions per row, 69 % of the HBM roofline).  Here a
// thread takes two ADJACENT rows per trip -- angles as two 16-byte loads, the six target floats as three 8-byte loads,
// the two errors as one 8-byte store -- and evaluates the eight sin/cos pairs two at a time on the packed fp32 pipe
// (ikb_sincos2: bit-identical to the scalar code, about half its instructions).  The chain arithmetic is the same inline
// function as everywhere else, so the errors are the ones the generic kernel and the solvers' fused epilogues produce.
#ifndef FK_PAIR_MINB
#define FK_PAIR_MINB 3  // 80 registers: room for the software-pipelined loads without spills (4 CTAs: 64 registers, spills)
#endif
__global__ void __launch_bounds__(256, FK_PAIR_MINB) fk_error_pairs_kernel(const FkArgs a)
{
    const long long n_pairs = a.n >> 1;  // the odd last row (if any) is handled by one thread after the loop
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float *k = a.rc.fkc_f;
    const float TWO_PI = 6.283185307179586f;
    double err_sum = 0.0;
    unsigned err_cnt = 0;
    const float4 *ang = reinterpret_cast<const float4 *>(a.angles);
    const float2 *tgt = reinterpret_cast<const float2 *>(a.targets);
    float2 *out = reinterpret_cast<float2 *>(a.err_out);
    auto row_error = [&](const float (&s)[4], const float (&c)[4], const float4 &th, float tx, float ty, float tz) {
        float px, py, pz;
        fk_planar_tail_position<float>(s, c, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], px, py, pz);
        const float dx = px - tx, dy = py - ty, dz = pz - tz;
        // sqrt.approx (one MUFU, 1 ulp) instead of the IEEE root with its fix-up branch: the error is a diagnostic
        // in metres, and the kernel is bound by instruction issue
        float err;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(err) : "f"(dx * dx + dy * dy + dz * dz));
        const bool ok = !(fabsf(th.x) > TWO_PI) & !(fabsf(th.y) > TWO_PI) & !(fabsf(th.z) > TWO_PI) & !(fabsf(th.w) > TWO_PI);
        return ok ? err : __int_as_float(0x7fc00000);
    };
#ifndef FK_PAIR_PREFETCH
#define FK_PAIR_PREFETCH 1
#endif
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float4 n_th0, n_th1;
    float2 n_t0, n_t1, n_t2;
    if (FK_PAIR_PREFETCH && p < n_pairs) {  // software pipeline: the next trip's loads are in flight during this trip's arithmetic
        n_th0 = __ldg(ang + 2 * p); n_th1 = __ldg(ang + 2 * p + 1);
        n_t0 = __ldg(tgt + 3 * p); n_t1 = __ldg(tgt + 3 * p + 1); n_t2 = __ldg(tgt + 3 * p + 2);
    }
    for (; p < n_pairs; p += stride) {
        float4 th0, th1;
        float2 t0, t1, t2;
        if (FK_PAIR_PREFETCH) {
            th0 = n_th0; th1 = n_th1; t0 = n_t0; t1 = n_t1; t2 = n_t2;
            const long long pn = p + stride;
            if (pn < n_pairs) {
                n_th0 = __ldg(ang + 2 * pn); n_th1 = __ldg(ang + 2 * pn + 1);
                n_t0 = __ldg(tgt + 3 * pn); n_t1 = __ldg(tgt + 3 * pn + 1); n_t2 = __ldg(tgt + 3 * pn + 2);
            }
        } else {
            th0 = __ldg(ang + 2 * p); th1 = __ldg(ang + 2 * p + 1);
            t0 = __ldg(tgt + 3 * p); t1 = __ldg(tgt + 3 * p + 1); t2 = __ldg(tgt + 3 * p + 2);
        }
        float s0[4], c0[4], s1[4], c1[4];
        ikb_sincos2(th0.x, th1.x, s0[0], c0[0], s1[0], c1[0]);
        ikb_sincos2(th0.y, th1.y, s0[1], c0[1], s1[1], c1[1]);
        ikb_sincos2(th0.z, th1.z, s0[2], c0[2], s1[2], c1[2]);
        ikb_sincos2(th0.w, th1.w, s0[3], c0[3], s1[3], c1[3]);
        const float e0 = row_error(s0, c0, th0, t0.x, t0.y, t1.x);
        const float e1 = row_error(s1, c1, th1, t1.y, t2.x, t2.y);
        if (e0 != e0)  // NaN from the angle guard of forward.py:23-25 (or from NaN input: then the row index is not recorded)
            if (fabsf(th0.x) > TWO_PI || fabsf(th0.y) > TWO_PI || fabsf(th0.z) > TWO_PI || fabsf(th0.w) > TWO_PI)
                atomicMin(&a.stats->first_fk_angle_range, a.index_base + 2 * p);
        if (e1 != e1)
            if (fabsf(th1.x) > TWO_PI || fabsf(th1.y) > TWO_PI || fabsf(th1.z) > TWO_PI || fabsf(th1.w) > TWO_PI)
                atomicMin(&a.stats->first_fk_angle_range, a.index_base + 2 * p + 1);
        if (out)
            out[p] = make_float2(e0, e1);
        float part = 0.f;
        if (isfinite(e0)) { part += e0; ++err_cnt; }
        if (isfinite(e1)) { part += e1; ++err_cnt; }
        err_sum += (double)part;
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // the odd last row: scalar code, same arithmetic
        const long long i = a.n - 1;
        const float4 v = __ldg(ang + i);
        const float th[4] = {v.x, v.y, v.z, v.w};
        const float *t = reinterpret_cast<const float *>(a.targets) + 3 * i;
        const float e = ikb_fk_error_planar_tail<float>(th, __ldg(t), __ldg(t + 1), __ldg(t + 2), k[0], k[1], k[2], k[3], k[4],
                                                        k[5], k[6], k[7]);
        if (e != e && (fabsf(v.x) > TWO_PI || fabsf(v.y) > TWO_PI || fabsf(v.z) > TWO_PI || fabsf(v.w) > TWO_PI))
            atomicMin(&a.stats->first_fk_angle_range, a.index_base + i);
        if (a.err_out)
            reinterpret_cast<float *>(a.err_out)[i] = e;
        if (isfinite(e)) { err_sum += (double)e; ++err_cnt; }
    }
    err_sum = ikb_warp_sum(err_sum);
    err_cnt = ikb_warp_sum(err_cnt);
    if ((threadIdx.x & 31) == 0 && err_cnt) {
        atomicAdd(&a.stats->sum_fk_error, err_sum);
        atomicAdd(&a.stats->n_fk_error, (unsigned long long)err_cnt);
    }
}

// ---- the same rows through an asynchronous-copy ring -------------------------------------------------------------------
// With the instruction count down, what keeps the pair kernel at ~0.8 of the HBM roofline is bytes in flight: a thread
// holds its next loads in registers (56 B) and 24 warps per SM make 43 KB.  Here a WARP owns 32 consecutive pairs per
// trip (1 KB of angles + 768 B of targets, both contiguous) and streams them global -> shared with cp.async through a
// ring of FK_STREAM_STAGES trips, no registers involved: up to 3.5 KB per warp in flight, 32 warps per SM.  Arithmetic
// and results are those of fk_error_pairs_kernel.  Needs 16-byte aligned buffers; the last partial trip goes to the
// pair kernel.
#ifndef FK_STREAM_STAGES
#define FK_STREAM_STAGES 3
#endif
#define FK_STREAM_WARPS 8
struct FkStage {
    float4 ang[64];   // 32 pairs x 2 rows
    float2 tgt[96];   // 32 pairs x 6 floats
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}

__global__ void __launch_bounds__(FK_STREAM_WARPS * 32, 4) fk_error_stream_kernel(const FkArgs a, long long n_trips)
{
    __shared__ __align__(16) FkStage s_ring[FK_STREAM_WARPS][FK_STREAM_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FkStage *ring = s_ring[warp];
    const long long w = (long long)blockIdx.x * FK_STREAM_WARPS + warp, W = (long long)gridDim.x * FK_STREAM_WARPS;
    const float *k = a.rc.fkc_f;
    const float TWO_PI = 6.283185307179586f;
    const char *ang = reinterpret_cast<const char *>(a.angles);
    const char *tgt = reinterpret_cast<const char *>(a.targets);
    float2 *out = reinterpret_cast<float2 *>(a.err_out);
    double err_sum = 0.0;
    unsigned err_cnt = 0;
    auto issue = [&](long long trip, int stage) {
        if (trip < n_trips) {
            const char *ga = ang + trip * 1024, *gt = tgt + trip * 768;
            char *sa = reinterpret_cast<char *>(ring[stage].ang), *st = reinterpret_cast<char *>(ring[stage].tgt);
            cp_async16(sa + lane * 16, ga + lane * 16);
            cp_async16(sa + 512 + lane * 16, ga + 512 + lane * 16);
            cp_async16(st + lane * 16, gt + lane * 16);
            if (lane < 16)
                cp_async16(st + 512 + lane * 16, gt + 512 + lane * 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // one group per trip, possibly empty: the wait below counts groups
    };
    auto row_error = [&](const float (&s)[4], const float (&c)[4], const float4 &th, float tx, float ty, float tz) {
        float px, py, pz;
        fk_planar_tail_position<float>(s, c, k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], px, py, pz);
        const float dx = px - tx, dy = py - ty, dz = pz - tz;
        float err;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(err) : "f"(dx * dx + dy * dy + dz * dz));
        const bool ok = !(fabsf(th.x) > TWO_PI) & !(fabsf(th.y) > TWO_PI) & !(fabsf(th.z) > TWO_PI) & !(fabsf(th.w) > TWO_PI);
        return ok ? err : __int_as_float(0x7fc00000);
    };
#pragma unroll
    for (int sidx = 0; sidx < FK_STREAM_STAGES - 1; ++sidx)
        issue(w + sidx * W, sidx);
    int stage = 0;
    for (long long trip = w; trip < n_trips; trip += W) {
        int ahead = stage + FK_STREAM_STAGES - 1;
        ahead -= ahead >= FK_STREAM_STAGES ? FK_STREAM_STAGES : 0;
        issue(trip + (FK_STREAM_STAGES - 1) * W, ahead);
        asm volatile("cp.async.wait_group %0;" ::"n"(FK_STREAM_STAGES - 1) : "memory");
        __syncwarp();
        const float4 th0 = ring[stage].ang[2 * lane], th1 = ring[stage].ang[2 * lane + 1];
        const float2 t0 = ring[stage].tgt[3 * lane], t1 = ring[stage].tgt[3 * lane + 1], t2 = ring[stage].tgt[3 * lane + 2];
        __syncwarp();  // every lane has read its rows: the stage may be refilled by the next issue()
        float s0[4], c0[4], s1[4], c1[4];
        ikb_sincos2(th0.x, th1.x, s0[0], c0[0], s1[0], c1[0]);
        ikb_sincos2(th0.y, th1.y, s0[1], c0[1], s1[1], c1[1]);
        ikb_sincos2(th0.z, th1.z, s0[2], c0[2], s1[2], c1[2]);
        ikb_sincos2(th0.w, th1.w, s0[3], c0[3], s1[3], c1[3]);
        const float e0 = row_error(s0, c0, th0, t0.x, t0.y, t1.x);
        const float e1 = row_error(s1, c1, th1, t1.y, t2.x, t2.y);
        const long long p = trip * 32 + lane;
        if (e0 != e0)
            if (fabsf(th0.x) > TWO_PI || fabsf(th0.y) > TWO_PI || fabsf(th0.z) > TWO_PI || fabsf(th0.w) > TWO_PI)
                atomicMin(&a.stats->first_fk_angle_range, a.index_base + 2 * p);
        if (e1 != e1)
            if (fabsf(th1.x) > TWO_PI || fabsf(th1.y) > TWO_PI || fabsf(th1.z) > TWO_PI || fabsf(th1.w) > TWO_PI)
                atomicMin(&a.stats->first_fk_angle_range, a.index_base + 2 * p + 1);
        if (out)
            out[p] = make_float2(e0, e1);
        float part = 0.f;
        if (isfinite(e0)) { part += e0; ++err_cnt; }
        if (isfinite(e1)) { part += e1; ++err_cnt; }
        err_sum += (double)part;
        stage = stage + 1 == FK_STREAM_STAGES ? 0 : stage + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    err_sum = ikb_warp_sum(err_sum);
    err_cnt = ikb_warp_sum(err_cnt);
    if (lane == 0 && err_cnt) {
        atomicAdd(&a.stats->sum_fk_error, err_sum);
        atomicAdd(&a.stats->n_fk_error, (unsigned long long)err_cnt);
    }
}

// all four cumulative homogeneous matrices (forward.py:79-94), fp64, one thread per angle set
__global__ void fk_chain_kernel(const double *angles, long long n, double *chain_out, int *status,
                                const IkbRobot rc)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double TWO_PI = 6.283185307179586;
    double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    int st = 0;
    for (int j = 0; j < 4; ++j) {
        const double t = angles[4 * i + j], al = rc.alpha[j];
        if (t < -TWO_PI || t > TWO_PI || al < -TWO_PI || al > TWO_PI)
            st = 1;
        double s, c;
        sincos(t, &s, &c);
        const double ca = rc.cos_alpha[j], sa = rc.sin_alpha[j], aj = rc.a[j], ej = rc.eps[j];
        // D = Rz(t) Tz(e) Tx(a) Rx(alpha)
        const double D[16] = {c, -s * ca, s * sa, aj * c,
                              s, c * ca, -c * sa, aj * s,
                              0, sa, ca, ej,
                              0, 0, 0, 1};
        double N[16];
        for (int r = 0; r < 4; ++r)
            for (int q = 0; q < 4; ++q) {
                double acc = 0;
                for (int m = 0; m < 4; ++m)
                    acc += T[4 * r + m] * D[4 * m + q];
                N[4 * r + q] = acc;
            }
        for (int m = 0; m < 16; ++m) {
            T[m] = N[m];
            chain_out[64 * i + 16 * j + m] = N[m];
        }
    }
    status[i] = st;
}

struct LimitsArgs {
    const void *xyz;
    int xyz_f64;
    long long n;
    long long index_base;
    IkbDeviceStats *stats;
    IkbRobot rc;
};

// K0: reference InverseKinematics.check_limits (inverse.py:26-35) as a min-index reduction.
__global__ void __launch_bounds__(256) check_limits_kernel(const LimitsArgs a)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long first = IKB_I64_MAX;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        double x, y, z;
        ikb_load_xyz(a.xyz, a.xyz_f64, i, x, y, z);
        if (ikb_out_of_limits(a.rc, x, y, z) && i < first)
            first = i;  // grid-stride order is increasing per thread
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(IKB_FULL_MASK, first, o);
        first = other < first ? other : first;
    }
    if ((threadIdx.x & 31) == 0 && first != IKB_I64_MAX)
        atomicMin(&a.stats->first_out_of_limits, a.index_base + first);
}

// Dependent-FMA-chain microbenchmark: 8 independent chains per thread keep the pipe full.
template <typename Real>
__global__ void __launch_bounds__(256) fma_peak_kernel(Real *sink, int iters, Real seed)
{
    Real v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
        v[j] = seed + (Real)(threadIdx.x + j);
    const Real m = (Real)0.999, c = (Real)0.001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[j] = fma(v[j], m, c);
    }
    Real s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        s += v[j];
    if (s == (Real)-1)
        sink[0] = s;
}

}  // namespace

static unsigned ikb_stream_grid(long long n, int num_sms, int threads, int ctas_per_sm)
{
    long long want = (n + threads - 1) / threads;
    long long cap = (long long)num_sms * ctas_per_sm;
    return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

cudaError_t ikb_launch_fk(const void *angles, int angles_f64, long long n, long long index_base,
                          void *pos_out, const void *targets, int xyz_f64, void *err_out,
                          IkbDeviceStats *stats, const IkbRobot &rc, int num_sms, cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    FkArgs a;
    a.angles = angles; a.angles_f64 = angles_f64; a.n = n; a.index_base = index_base;
    a.pos_out = pos_out; a.targets = targets; a.xyz_f64 = xyz_f64; a.err_out = err_out;
    a.stats = stats; a.rc = rc;
    // fp32 buffers, closed-form arm with alpha inside the guard, error only, 8-byte aligned rows: the pair kernel
    bool alpha_ok = rc.fk_planar_tail != 0;
    for (int j = 0; j < 4; ++j)
        alpha_ok = alpha_ok && !(rc.alpha[j] < -6.283185307179586) && !(rc.alpha[j] > 6.283185307179586);
    if (!angles_f64 && !xyz_f64 && targets && !pos_out && alpha_ok && ((uintptr_t)targets & 7) == 0 &&
        ((uintptr_t)err_out & 7) == 0 && n >= 2) {
        static const bool use_stream = [] { const char *v = std::getenv("IKB_FK_STREAM"); return !v || v[0] != '0'; }();
        const long long n_trips = n / 64;  // full trips of 32 pairs for the asynchronous-copy kernel
        if (use_stream && n_trips >= 4LL * num_sms * FK_STREAM_WARPS && ((uintptr_t)targets & 15) == 0 &&
            ((uintptr_t)angles & 15) == 0) {
            const long long want = (n_trips + FK_STREAM_WARPS - 1) / FK_STREAM_WARPS, cap = (long long)num_sms * 4;
            static const cudaError_t carve = cudaFuncSetAttribute(fk_error_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                                  (int)cudaSharedmemCarveoutMaxShared);
            if (carve != cudaSuccess)
                return carve;
            fk_error_stream_kernel<<<(unsigned)(want < cap ? want : cap), FK_STREAM_WARPS * 32, 0, stream>>>(a, n_trips);
            cudaError_t err = cudaGetLastError();
            const long long done = n_trips * 64;
            if (err != cudaSuccess || done == n)
                return err;
            // the last partial trip (< 64 rows): the pair kernel on the tail, same arithmetic
            FkArgs t = a;
            t.n = n - done;
            t.index_base = index_base + done;
            t.angles = reinterpret_cast<const float *>(angles) + 4 * done;
            t.targets = reinterpret_cast<const float *>(targets) + 3 * done;
            t.err_out = err_out ? reinterpret_cast<float *>(err_out) + done : nullptr;
            if (t.n >= 2) {
                fk_error_pairs_kernel<<<1, 256, 0, stream>>>(t);
                return cudaGetLastError();
            }
            a = t;  // a single odd row: the generic kernel below
            fk_kernel<float, false><<<1, 256, 0, stream>>>(a);
            return cudaGetLastError();
        }
        const unsigned pgrid = ikb_stream_grid(n / 2, num_sms, 256, FK_PAIR_MINB);
        fk_error_pairs_kernel<<<pgrid, 256, 0, stream>>>(a);
        return cudaGetLastError();
    }
    const unsigned grid = ikb_stream_grid(n, num_sms, 256, 8);
    if (angles_f64) {
        if (xyz_f64)
            fk_kernel<double, true><<<grid, 256, 0, stream>>>(a);
        else
            fk_kernel<double, false><<<grid, 256, 0, stream>>>(a);
    } else {
        if (xyz_f64)
            fk_kernel<float, true><<<grid, 256, 0, stream>>>(a);
        else
            fk_kernel<float, false><<<grid, 256, 0, stream>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t ikb_launch_fk_chain(const double *angles, long long n, double *chain_out, int *status,
                                const IkbRobot &rc, cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    fk_chain_kernel<<<(unsigned)((n + 63) / 64), 64, 0, stream>>>(angles, n, chain_out, status, rc);
    return cudaGetLastError();
}

cudaError_t ikb_launch_check_limits(const void *xyz, int xyz_f64, long long n, long long index_base,
                                    IkbDeviceStats *stats, const IkbRobot &rc, int num_sms,
                                    cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    LimitsArgs a;
    a.xyz = xyz; a.xyz_f64 = xyz_f64; a.n = n; a.index_base = index_base; a.stats = stats; a.rc = rc;
    check_limits_kernel<<<ikb_stream_grid(n, num_sms, 256, 8), 256, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t ikb_launch_fma_peak(int f64, void *sink, int iters, int num_sms, cudaStream_t stream)
{
    if (f64)
        fma_peak_kernel<double><<<num_sms * 8, 256, 0, stream>>>((double *)sink, iters, 1.0);
    else
        fma_peak_kernel<float><<<num_sms * 8, 256, 0, stream>>>((float *)sink, iters, 1.0f);
    return cudaGetLastError();
}
