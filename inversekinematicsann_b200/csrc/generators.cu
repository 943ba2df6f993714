// generators.cu -- on-device trajectory generators: the shapes of reference
// robot/position_generator.py:26-97 produced directly in HBM so that large synthetic workloads need no
// host generation and no H2D copy.  Deterministic shapes follow the reference formulas in fp64;
// the random shapes use a counter-based Philox4x32-10 stream (seed, row) -- same distributions as the
// reference's numpy/scipy calls, not the same numbers (documented in DESIGN.md).
#include "ikb_common.cuh"

#define IKB_GEN_CIRCLE 0
#define IKB_GEN_SPRING 1
#define IKB_GEN_CUBE 2
#define IKB_GEN_CUBE_RANDOM 3
#define IKB_GEN_NORMAL 4

namespace {

struct GenArgs {
    int kind;
    long long n;
    void *out;
    int out_f64;
    double p[12];  // kind-specific parameters, see ikb_generate_device
    unsigned long long seed;
    long long row_offset;  // first global row (shards of one trajectory generate disjoint ranges)
};

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], const uint32_t (&k)[2])
{
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Philox4x32-10: 128 random bits for (seed, counter)
__device__ __forceinline__ void philox4x32(unsigned long long seed, unsigned long long ctr, uint32_t (&r)[4])
{
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x1BD11BDAu, 0x5851F42Du};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    r[0] = c[0]; r[1] = c[1]; r[2] = c[2]; r[3] = c[3];
}

// uniform in [0, 1) with 53 random bits (numpy.random.rand resolution)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    return (double)((((unsigned long long)a >> 5) << 26) | ((unsigned long long)b >> 6)) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(256) generate_kernel(const GenArgs a)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const long long g = a.row_offset + i;
        double x, y, z;
        if (a.kind == IKB_GEN_CIRCLE) {          // position_generator.py:26-31, integer time stamps (radians)
            double s, c;
            sincos((double)g, &s, &c);
            x = a.p[1]; y = a.p[2] + a.p[0] * s; z = a.p[3] + a.p[0] * c;
        } else if (a.kind == IKB_GEN_SPRING) {   // :72-78, z = linspace(0, len_z, n_total)
            const double n_total = a.p[3];
            const double step = n_total > 1.0 ? a.p[2] / (n_total - 1.0) : 0.0;
            z = (double)g == n_total - 1.0 ? a.p[2] : (double)g * step;
            double s, c;
            sincos(z, &s, &c);
            x = ((s * a.p[0]) + a.p[0]) / 2;
            y = ((c * a.p[1]) + a.p[1]) / 2;
        } else if (a.kind == IKB_GEN_CUBE) {     // :39-46, x fastest, then y, then z
            const long long nx = (long long)a.p[7], ny = (long long)a.p[8];
            const long long ix = g % nx, iy = (g / nx) % ny, iz = g / (nx * ny);
            x = (double)ix * a.p[0] + a.p[4]; y = (double)iy * a.p[0] + a.p[5]; z = (double)iz * a.p[0] + a.p[6];
        } else {
            uint32_t r[4], r2[4];
            philox4x32(a.seed, 2ULL * (unsigned long long)g, r);
            philox4x32(a.seed, 2ULL * (unsigned long long)g + 1, r2);
            const double u0 = u53(r[0], r[1]), u1 = u53(r[2], r[3]), u2 = u53(r2[0], r2[1]);
            if (a.kind == IKB_GEN_CUBE_RANDOM) { // :48-55: start + len * U[0,1) per axis
                // separately rounded product and sum (no FMA contraction): the stream is restated bit for bit in NumPy
                // by the test infrastructure, which is how the CPU arm of bench.py gets the same rows
                x = __dadd_rn(__dmul_rn(u0, a.p[0]), a.p[3]);
                y = __dadd_rn(__dmul_rn(u1, a.p[1]), a.p[4]);
                z = __dadd_rn(__dmul_rn(u2, a.p[2]), a.p[5]);
            } else {                             // :80-97 'normal': truncnorm(mean 0, std, lo, hi) per axis
                const double sd = a.p[6];        // inverse-CDF sampling, as scipy's truncnorm.rvs does
                const double us[3] = {u0, u1, u2};
                double v[3];
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const double ca = normcdf(a.p[2 * ax] / sd), cb = normcdf(a.p[2 * ax + 1] / sd);
                    double q = sd * normcdfinv(ca + us[ax] * (cb - ca));
                    q = fmin(fmax(q, a.p[2 * ax]), a.p[2 * ax + 1]);
                    v[ax] = q;
                }
                x = v[0]; y = v[1]; z = v[2];
            }
        }
        if (a.out_f64) {
            double *o = reinterpret_cast<double *>(a.out) + 3 * i;
            o[0] = x; o[1] = y; o[2] = z;
        } else {
            float *o = reinterpret_cast<float *>(a.out) + 3 * i;
            o[0] = (float)x; o[1] = (float)y; o[2] = (float)z;
        }
    }
}

}  // namespace

cudaError_t ikb_launch_generate(int kind, const double *params, int n_params, long long n, long long row_offset,
                                void *out, int out_f64, unsigned long long seed, int num_sms, cudaStream_t stream)
{
    if (n <= 0)
        return cudaSuccess;
    GenArgs a;
    a.kind = kind; a.n = n; a.out = out; a.out_f64 = out_f64; a.seed = seed; a.row_offset = row_offset;
    for (int i = 0; i < 12; ++i)
        a.p[i] = i < n_params ? params[i] : 0.0;
    long long want = (n + 255) / 256, cap = (long long)num_sms * 8;
    generate_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(a);
    return cudaGetLastError();
}
