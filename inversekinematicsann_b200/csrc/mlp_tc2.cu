// mlp_tc2.cu -- K2, second tensor-core layout (IKB_MLP_FP16X3_TS): activations as the A operand in TMEM.
//
// mlp_tc.cu is bound by shared-memory bandwidth: with 64 rows per CTA, streaming 1 MB of weights per layer
// through smem (TMA write + MMA read) and re-reading the activation granules for every feature tile costs
// 2.9 MB of smem traffic per layer.  This layout halves the traffic per row:
//   * 128 REAL batch rows per CTA = UMMA M = 128 = TMEM lanes, so every (row, feature) sum lives in one lane;
//   * x_hi (fp16) lives in TMEM (256 columns, two k per 32-bit column) and feeds the MMA as a TMEM A operand
//     (TS mode, no smem reads at all); only x_lo (fp16) stays in shared memory (128 KB, K-major, 128B swizzle);
//   * weights are the B operand, N = 256 features per UMMA, tiles of 256 features x 64 k (32 KB) pre-swizzled
//     on the host and streamed by the TMA engine through a 3-stage ring;
//   * three products: x_hi w_hi (TS), x_lo w_hi (SS), x_hi w_lo (TS), all into ONE fp32 accumulator
//     D[128 lanes x 256 columns] per N half -- the two small products for every k chunk FIRST, the main product last
//     (IKB_TS_CORR_FIRST: the accumulator is truncated toward zero at every K = 16 step, and only steps taken at the
//     sum's full size cost accuracy).  x is stored times 2^6 so that x_lo stays a normal fp16 number while sharing the
//     accumulator with x_hi; the factor is folded into the per-layer output scale.
// TMEM: columns [0, 256) = x_hi, [256, 512) = D.  Per layer: MMA(half 0) -> epilogue drains D into registers ->
// MMA(half 1) runs while the epilogue turns half 0 into the next layer's activations (kept packed in registers
// until the issuer's tcgen05.commit says the old activations are dead) -> store (tcgen05.st for x_hi, st.shared
// for x_lo) -> the next layer starts on the k range that is already there while half 1 is still being converted.
// The 3-input first layer and the 4-output last layer run on the CUDA cores of the epilogue warps.
#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fk_device.cuh"
#include "mlp.cuh"

#ifdef IKB_TC_DEBUG
__device__ unsigned long long g_tc2_dbg[16];
#define DBG_T0() const long long _t0 = clock64()
#define DBG_ADD(i) do { if (blockIdx.x == 0) g_tc2_dbg[i] += (unsigned long long)(clock64() - _t0); } while (0)
#else
#define DBG_T0() do {} while (0)
#define DBG_ADD(i) do {} while (0)
#endif

namespace {

constexpr int ROWS = 128;             // targets per CTA tile = UMMA M
constexpr int GRAN_BYTES = 16384;     // x_lo granule: 128 rows x 64 k x fp16
// k extent of one weight tile (compile-time): 64 (default: rows of 128 B, 128-byte swizzle, three 32 KB ring stages) or 32
// (rows of 64 B, 64-byte swizzle, six 16 KB stages: the same 96 KB with more of the ring in flight at any time).
// Measured on B200, 1 M rows: 14.24 ms (64) against 14.45 ms (32), identical results -- the issuer's waits for weights
// drop by a fifth with the finer stages, but twice as many commits / waits on the single issuing thread cost as much,
// and the kernel runs at the power-capped clock (1.58 GHz, 88 % of what a cuBLAS bf16 GEMM sustains under the same cap)
// either way.
#ifndef IKB_TS_WK
#define IKB_TS_WK 64
#endif
constexpr int WK = IKB_TS_WK;                   // 64: rows of 128 B, 128-byte swizzle; 32: rows of 64 B, 64-byte swizzle
static_assert(WK == 64 || WK == 32, "weight tiles are 64 or 32 k wide");
constexpr int WROW_BYTES = WK * 2;
constexpr int WKSTEPS = WK / 16;                // UMMA K steps per weight tile
constexpr int WSUB = 64 / WK;                   // weight tiles per x_lo granule (64 k)
constexpr int WTILE_BYTES = 256 * WROW_BYTES;   // weight tile: up to 256 features x WK k x fp16
constexpr int W_RING_BYTES = 98304;             // weight ring: 96 KB = 3 / 6 stages of a whole tile (half tiles for CTA pairs: twice as many)
#ifndef IKB_TS_EPI_WARPS
#define IKB_TS_EPI_WARPS 16
#endif
// Accumulation order within one (layer, N half).  tcgen05.mma truncates the fp32 accumulator toward zero after every
// K = 16 step, by up to an ulp of the RUNNING SUM.  Interleaved (round 1: per k chunk x_hi w_hi, x_lo w_hi, x_hi w_lo)
// all 96 steps of an output truncate at the sum's full size.  Corrections first: the 64 steps of the two small
// products (2^-11 of the result) run while the accumulator is still tiny, and only the 32 steps of x_hi w_hi
// truncate at full size -- a third of the noise (tools/tc_trunc_model.py replays both orders on the CPU).  The price
// is that the w_hi tiles pass through shared memory twice (3 tiles per k chunk instead of 2: +50 % L2 -> smem
// traffic, same MMA count).
#ifndef IKB_TS_CORR_FIRST
#define IKB_TS_CORR_FIRST 1
#endif
constexpr int TILES_PER_KC = IKB_TS_CORR_FIRST ? 3 : 2;   // weight tiles streamed per k chunk
constexpr int N_EPI_WARPS = IKB_TS_EPI_WARPS;   // CW warps share the columns of one TMEM sub-partition
constexpr int CW = N_EPI_WARPS / 4;
static_assert(CW == 4, "the act_ready granularity (128 features) assumes 4 warps per TMEM sub-partition");
constexpr int QMAX = 256 / CW / 32;             // groups of 32 accumulator columns per thread and N half
constexpr int THREADS = (2 + N_EPI_WARPS) * 32;
constexpr float X_SCALE = 64.0f;      // activations are stored times 2^6
constexpr int TMEM_A_COL = 0, TMEM_D_COL = 256;

struct Tc2Net {
    int n_mma_layers;        // hidden layers 2..NH
    int hp;                  // common padded hidden width (multiple of 128, <= 512)
    const __half *w_tiles;   // [layer][n half][k chunk][hi|lo] tiles of WTILE_BYTES (rows beyond the half's width unused)
    const float *w_first;    // [3][hp]
    const float *b_hidden;   // [1 + n_mma_layers][hp]
    const float *out_scale;  // [n_mma_layers]  1 / (weight scale * X_SCALE)
    int bias_in_mma;         // hidden width < padded width: feature `hmax` is a constant 1 and the biases are a weight row
    const float *w_last;     // [hp][4]
    float b_last[4];
    double mean_x[3], scale_x[3];
    float mean_y[4], scale_y[4];
};

struct Tc2Args {
    const void *xyz;
    int xyz_f64;
    long long n;
    long long index_base;
    float *out;
    float *fk_err;  // nullable: ||FK(out) - target|| per row (fused K3)
    int fk_stats;   // accumulate sum_fk_error / n_fk_error even without the per-row array
    IkbDeviceStats *stats;
    IkbRobot rc;
    Tc2Net net;
};

// ---- PTX helpers (same conventions as mlp_tc.cu) --------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TS_DONE;\n"
        "bra TS_WAIT;\n"
        "TS_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// ---- CTA pairs (cta_group::2; opt-in with IKB_TS_CLUSTER=2): two CTAs on the SMs of one TPC run ONE M = 256 MMA: each
// holds 128 batch rows (its own x_hi in TMEM, x_lo in shared memory, accumulator in TMEM) and HALF of every weight tile
// (128 of the 256 features), which the tensor cores of both SMs read.  Per SM that halves the weight bytes TMA writes
// into shared memory and the L2 -> SM traffic.  Bit-identical results -- and SLOWER on B200: 17.4 ms against 13.9 ms per
// 1 M rows.  The issuer's time per layer grows from 37 k to 55 k cycles and stays there with a third or two thirds of
// the MMAs removed (timing experiments with -DIKB_DBG_SKIP_SS / _TS), i.e. it is not tensor throughput but the longer
// hand-off chain of every ring stage (multicast commit -> both producers -> TMA -> the peer's "landed" forwarded by a
// remote mbarrier arrive -> issuer) and of every accumulator / activation barrier, which now waits for two epilogues.
// Kept as a tested switch, not the default.
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t rank)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait that also acquires at cluster scope (the arrivals come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TSC_WAIT:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TSC_DONE;\n"
        "bra TSC_WAIT;\n"
        "TSC_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)  // K-major, 128B swizzle, 8-row groups 1024 B apart
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// weight tiles: K-major rows of WROW_BYTES, 8-row groups 8 * WROW_BYTES apart, layout type 2 (128B swizzle) / 4 (64B)
__device__ __forceinline__ uint64_t make_desc_w(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(8 * WROW_BYTES >> 4) << 32) | (1ull << 46) |
           ((WK == 64 ? 2ull : 4ull) << 61);
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// the same three for a CTA pair: issued by the leader CTA only, M = 256 over both CTAs; the commit arrives on the barrier
// at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma2_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t *bar)
{
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
        "h"((unsigned short)3)
        : "memory");
}
template <int NCTA>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
#ifndef IKB_DBG_SKIP_SS  // timing experiments only: results are wrong without this product
    if (NCTA == 2) umma2_ss(d, a, b, idesc, acc); else umma_ss(d, a, b, idesc, acc);
#endif
}
template <int NCTA>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
#ifndef IKB_DBG_SKIP_TS
    if (NCTA == 2) umma2_ts(d, a, b, idesc, acc); else umma_ts(d, a, b, idesc, acc);
#endif
}
template <int NCTA>
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    if (NCTA == 2) umma2_commit(bar); else umma_commit(bar);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __host__ __forceinline__ int swz_off(int row, int k)  // [rows x 64] fp16 K-major tile, 128B swizzle
{
    return row * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + ((k & 7) << 1);
}
// byte offset of element (row, k) in a weight tile [rows x WK] fp16, K-major: 128-byte swizzle for 128-byte rows (16-byte
// chunk index ^= row & 7), 64-byte swizzle for 64-byte rows (chunk index ^= (row >> 1) & 3: address bits [4,6) ^= [7,9))
__device__ __host__ __forceinline__ int wswz_off(int row, int k)
{
    if (WK == 64)
        return swz_off(row, k);
    return row * 64 + ((((k >> 3) ^ (row >> 1)) & 3) << 4) + ((k & 7) << 1);
}

// tanh with the accumulator scale folded into the exponent and X_SCALE folded into the result:
// returns X_SCALE * tanh(v) for v = acc * oscale given t = |acc| * (-2 log2(e) oscale); 2 MUFU + 5 ALU ops
__device__ __forceinline__ float tanh_scaled(float t, float sign_src)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return copysignf(fmaf(e, -X_SCALE, X_SCALE) * r, sign_src);
}

// (hi, lo) fp16 split of two values that already carry the X_SCALE factor
__device__ __forceinline__ void split_pair_scaled(float s0, float s1, uint32_t &hi, uint32_t &lo)
{
    const __half2 h = __floats2half2_rn(s0, s1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// y -> (hi, lo) fp16 split of X_SCALE * y, two consecutive features packed per 32-bit word (even feature low)
__device__ __forceinline__ void split_pair(float y0, float y1, uint32_t &hi, uint32_t &lo)
{
    // cvt.rn.f16x2.f32 converts and packs two floats in one full-rate instruction (the scalar F2F.F16.F32
    // shares the quarter-rate pipe with MUFU and made the epilogue, not the MMA, the bottleneck)
    const float s0 = y0 * X_SCALE, s1 = y1 * X_SCALE;
    const __half2 h = __floats2half2_rn(s0, s1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// Fused K3: ||FK(angles) - target|| for the row this thread just produced, same arithmetic as
// fk.cu on the same fp32 angles; one warp-level sum and one atomic pair per warp and tile.
__device__ __forceinline__ void fused_fk_error(const Tc2Args &a, long long i, bool live, const float (&th)[4])
{
    float err = 0.f;
    bool counted = false;
    if (live) {
        double x, y, z;
        ikb_load_xyz(a.xyz, a.xyz_f64, i, x, y, z);
        err = ikb_fk_error<float>(a.rc, th, (float)x, (float)y, (float)z);
        if (a.fk_err)
            a.fk_err[i] = err;
        counted = isfinite(err);
    }
    const float part = ikb_warp_sum(counted ? err : 0.f);
    const unsigned cnt = __popc(__ballot_sync(IKB_FULL_MASK, counted));
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(&a.stats->sum_fk_error, (double)part);
        atomicAdd(&a.stats->n_fk_error, (unsigned long long)cnt);
    }
}

// Everything an epilogue thread needs to turn pre-activations into the next layer's operands.
struct EpiCtx {
    const Tc2Net *net;
    unsigned char *xlo;
    uint64_t *a_free, *act_ready;
    uint32_t tmem_base, lane_base;
    int row, lane, HP;
    int pair_peer;         // this CTA is rank 1 of a CTA pair: epilogue -> MMA barriers live in rank 0
    float in0, in1, in2;   // scaled inputs of this thread's row (first layer)
};

// epilogue -> MMA issuer ("these activations are stored" / "the accumulator is drained"): the issuer lives in the
// leader CTA of a pair, so the peer's epilogue warps arrive there
__device__ __forceinline__ void epi_arrive(const EpiCtx &cx, uint64_t *bar)
{
    if (cx.pair_peer)
        mbar_arrive_cluster(bar, 0);
    else
        mbar_arrive(bar);
}

// Layer 1 (3 -> HP) for one thread: `ncols` features starting at f0, a rolled loop over chunks of 8 features
// stored at once (nothing reads the activation buffers while a tile's first layer runs), or fed to the output sums
// when the network has a single hidden layer.
__device__ __forceinline__ void first_layer_half(const EpiCtx &cx, int f0, int ncols, bool last_hidden, int ready_idx,
                                                 float (&out_acc)[4])
{
    const Tc2Net &net = *cx.net;
    const uint32_t a_addr = cx.tmem_base + cx.lane_base + TMEM_A_COL + (f0 >> 1);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < ncols; c += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float y[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int f = f0 + c + 2 * t + u;
                y[u] = tanhf(fmaf(cx.in2, __ldg(net.w_first + 2 * cx.HP + f),
                                  fmaf(cx.in1, __ldg(net.w_first + cx.HP + f),
                                       fmaf(cx.in0, __ldg(net.w_first + f), __ldg(net.b_hidden + f)))));
                if (last_hidden) {  // single hidden layer: same X_SCALE convention as the MMA layers' output sums
                    const float4 w = __ldg(reinterpret_cast<const float4 *>(net.w_last) + f);
                    const float ys = y[u] * X_SCALE;
                    out_acc[0] = fmaf(ys, w.x, out_acc[0]); out_acc[1] = fmaf(ys, w.y, out_acc[1]);
                    out_acc[2] = fmaf(ys, w.z, out_acc[2]); out_acc[3] = fmaf(ys, w.w, out_acc[3]);
                }
            }
            split_pair(y[0], y[1], hi[t], lo[t]);
        }
        if (!last_hidden) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_addr + (c >> 1)), "r"(hi[0]),
                         "r"(hi[1]), "r"(hi[2]), "r"(hi[3])
                         : "memory");
            const int f = f0 + c;
            unsigned char *dst = cx.xlo + (size_t)(f >> 6) * GRAN_BYTES + cx.row * 128 +
                                 (((((f & 63) >> 3) ^ (cx.row & 7)) & 7) << 4);
            *reinterpret_cast<uint4 *>(dst) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    if (last_hidden)
        return;
    tmem_st_wait();
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (cx.lane == 0)
        epi_arrive(cx, &cx.act_ready[ready_idx]);
}

// One N half of one hidden layer for one thread.  The thread owns QMAX groups of 32 accumulator columns; group q
// covers features fbase + q * 32 * CW .. + 31 (so that the CW warps of a sub-partition together complete 128
// consecutive features = two k chunks of the next layer at a time, reported through act_ready[2 nh + q]).
// Results are packed IN PLACE over `d` (word 2t = the x_hi pair of features 2t, 2t+1 of the group, word 2t+1 = the
// x_lo pair); they become the next layer's x_hi (TMEM) / x_lo (smem), or, for the last hidden layer, go straight
// into the 4 output sums (the output layer needs fp32 activations anyway).
// early_compute: this is not the layer's last N half, i.e. the MMAs still read the old activations -- convert every
// group first (overlapping the MMAs of the next half), then wait for a_free and store.  Otherwise a_free has fired
// together with d_full: convert, store and report group by group so the next layer's MMAs can start on the first
// 128 features while the second group is still being converted.
__device__ __forceinline__ void finish_half(const EpiCtx &cx, uint32_t (&d)[QMAX][32], int fbase, int ngroups,
                                            bool last_hidden, bool early_compute, uint32_t free_parity, int ready_base,
                                            float oscale, const float *bias, float (&out_acc)[4])
{
    const Tc2Net &net = *cx.net;
    const float cexp = -2.8853900817779268f * oscale;
    auto convert = [&](int q) {
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int f = fbase + q * 32 * CW + 2 * t;
            const float a0 = __uint_as_float(d[q][2 * t]), a1 = __uint_as_float(d[q][2 * t + 1]);
            float s0, s1;  // X_SCALE * tanh(pre-activation)
            if (net.bias_in_mma) {  // the bias row already sits in the accumulator
#ifdef IKB_TS_TANH_EXACT  // experiment: libdevice tanhf, to separate the tanh error from the accumulation error
                s0 = X_SCALE * tanhf(a0 * oscale);
                s1 = X_SCALE * tanhf(a1 * oscale);
#else
                s0 = tanh_scaled(fabsf(a0) * cexp, a0);
                s1 = tanh_scaled(fabsf(a1) * cexp, a1);
#endif
            } else {
                const float v0 = fmaf(a0, oscale, __ldg(bias + f)), v1 = fmaf(a1, oscale, __ldg(bias + f + 1));
                s0 = tanh_scaled(fabsf(v0) * -2.8853900817779268f, v0);
                s1 = tanh_scaled(fabsf(v1) * -2.8853900817779268f, v1);
            }
            if (last_hidden) {  // out_acc carries the X_SCALE factor, removed once at the end
                const float4 w0 = __ldg(reinterpret_cast<const float4 *>(net.w_last) + f);
                const float4 w1 = __ldg(reinterpret_cast<const float4 *>(net.w_last) + f + 1);
                out_acc[0] = fmaf(s1, w1.x, fmaf(s0, w0.x, out_acc[0]));
                out_acc[1] = fmaf(s1, w1.y, fmaf(s0, w0.y, out_acc[1]));
                out_acc[2] = fmaf(s1, w1.z, fmaf(s0, w0.z, out_acc[2]));
                out_acc[3] = fmaf(s1, w1.w, fmaf(s0, w0.w, out_acc[3]));
            } else {
                split_pair_scaled(s0, s1, d[q][2 * t], d[q][2 * t + 1]);
            }
        }
    };
    auto store = [&](int q) {
        const int f0 = fbase + q * 32 * CW;
        // x_hi pairs of this group -> 16 TMEM columns
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15, %16};" ::"r"(cx.tmem_base + cx.lane_base + TMEM_A_COL + (f0 >> 1)),
            "r"(d[q][0]), "r"(d[q][2]), "r"(d[q][4]), "r"(d[q][6]), "r"(d[q][8]), "r"(d[q][10]), "r"(d[q][12]),
            "r"(d[q][14]), "r"(d[q][16]), "r"(d[q][18]), "r"(d[q][20]), "r"(d[q][22]), "r"(d[q][24]), "r"(d[q][26]),
            "r"(d[q][28]), "r"(d[q][30])
            : "memory");
        // x_lo pairs -> shared memory, 16-byte chunks of 8 features, swizzled K-major rows
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int f = f0 + c * 8;
            unsigned char *dst = cx.xlo + (size_t)(f >> 6) * GRAN_BYTES + cx.row * 128 +
                                 (((((f & 63) >> 3) ^ (cx.row & 7)) & 7) << 4);
            *reinterpret_cast<uint4 *>(dst) = make_uint4(d[q][8 * c + 1], d[q][8 * c + 3], d[q][8 * c + 5], d[q][8 * c + 7]);
        }
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (cx.lane == 0)
            epi_arrive(cx, &cx.act_ready[ready_base + q]);
    };
    if (last_hidden) {
#pragma unroll
        for (int q = 0; q < QMAX; ++q)
            if (q < ngroups)
                convert(q);
        return;
    }
    if (early_compute) {
#pragma unroll
        for (int q = 0; q < QMAX; ++q)
            if (q < ngroups)
                convert(q);
        {
            DBG_T0();
            mbar_wait(cx.a_free, free_parity);  // the MMAs of this layer no longer read x_hi / x_lo
            if (threadIdx.x == 64) DBG_ADD(5);
        }
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < QMAX; ++q)
            if (q < ngroups)
                store(q);
    } else {
        mbar_wait(cx.a_free, free_parity);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < QMAX; ++q)
            if (q < ngroups) {
                convert(q);
                store(q);
            }
    }
}

// NCTA = 1: one CTA per SM on its own.  NCTA = 2: CTA pairs (clusters of two), see the cluster helpers above; rank 0
// is the leader (it issues every MMA), each CTA of a pair works on its own tile of 128 rows.
template <int NCTA>
__global__ void __launch_bounds__(THREADS, 1) mlp_tc2_kernel(const Tc2Args a)
{
    constexpr int STAGE_BYTES = WTILE_BYTES / NCTA;        // a CTA of a pair holds half of every weight tile
    constexpr int W_STAGES = W_RING_BYTES / STAGE_BYTES;   // 3 whole tiles or 6 half tiles
    extern __shared__ __align__(1024) unsigned char smem[];
    const Tc2Net &net = a.net;
    const int HP = net.hp, KG = HP >> 6, NHALF = (HP + 255) >> 8, NM = net.n_mma_layers;
    const int KT = HP / WK;                                        // weight k tiles per layer
    unsigned char *xlo = smem;                                     // KG granules of x_lo
    unsigned char *wring = smem + (size_t)KG * GRAN_BYTES;          // W_STAGES weight (half) tiles
    float *s_io = reinterpret_cast<float *>(wring + (size_t)W_RING_BYTES);  // [128][4] inputs / output partials
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_io + ROWS * 4);
    static_assert(W_STAGES <= 12, "barrier block holds 12 stages");
    uint64_t *w_full = bars, *w_empty = bars + 12, *w_peer = bars + 24;  // w_peer (leader): the peer's half of a stage has landed
    uint64_t *d_full = bars + 36, *d_empty = d_full + 1;            // the single accumulator: MMA <-> epilogue
    uint64_t *act_ready = d_empty + 1;                              // [4] epilogue -> MMA: 128 features (2 k chunks) of the next input stored
    uint64_t *a_free = act_ready + 4;                               // MMA -> epilogue: this layer's input is dead
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(a_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;       // 0 = leader of the pair
    if (threadIdx.x == 0) {
        for (int i = 0; i < W_STAGES; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
            mbar_init(&w_peer[i], 1);
        }
        mbar_init(d_full, 1);
        mbar_init(d_empty, N_EPI_WARPS * NCTA);      // a pair: the epilogue warps of both CTAs arrive in the leader
        for (int i = 0; i < 4; ++i)
            mbar_init(&act_ready[i], N_EPI_WARPS * NCTA);
        mbar_init(a_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (NCTA == 2)
        cluster_sync_all();   // both CTAs' barriers are initialised before anybody arrives on a remote one
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    const long long n_tiles = (a.n + ROWS - 1) / ROWS;
    // both CTAs of a pair run the same number of trips (the MMAs span both); a tile index past the end is an empty tile
    const long long n_groups = (n_tiles + NCTA - 1) / NCTA, group0 = blockIdx.x / NCTA, group_stride = gridDim.x / NCTA;

    if (warp == 0) {
        // ===== TMA producer: the weight tiles of every (layer, N half) in the order the issuer consumes them =====
        // arena layout: [layer][N half][k chunk][hi | lo] tiles of WTILE_BYTES.
        // interleaved order: kc0 hi, kc0 lo, kc1 hi, ...; corrections first: (kc hi, kc lo) for all kc, then kc hi again.
        // A CTA of a pair loads its half of the tile's features: rows [rank * nfeat / 2, ...) of the swizzled image, which
        // is a valid image of its own (the swizzle pattern repeats every 8 rows).
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 1;
            const unsigned char *base = reinterpret_cast<const unsigned char *>(net.w_tiles);
            for (long long g = group0; g < n_groups; g += group_stride) {
                for (int mh = 0; mh < NM * NHALF; ++mh) {
                    const unsigned char *half_base = base + (size_t)mh * KT * 2 * WTILE_BYTES;
                    const uint32_t bytes = (uint32_t)min(256, HP - 256 * (mh % NHALF)) * (uint32_t)WROW_BYTES / NCTA;
                    for (int t = 0; t < KT * TILES_PER_KC; ++t) {
                        // t < 2 KT: tile t of the half as stored; after that the hi tile of k tile t - 2 KT
                        const int src_tile = t < 2 * KT ? t : 2 * (t - 2 * KT);
                        mbar_wait(&w_empty[s], ph);
                        mbar_expect_tx(&w_full[s], bytes);
                        tma_load_1d(wring + (size_t)s * STAGE_BYTES, half_base + (size_t)src_tile * WTILE_BYTES + rank * bytes,
                                    bytes, &w_full[s]);
                        if (++s == W_STAGES) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && rank != 0) {
        // ===== peer CTA of a pair: tell the leader's issuer when this CTA's half of a weight stage has landed =====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long g = group0; g < n_groups; g += group_stride)
                for (int t = 0; t < NM * NHALF * KT * TILES_PER_KC; ++t) {
                    mbar_wait(&w_full[s], ph);
                    mbar_arrive_cluster(&w_peer[s], 0);
                    if (++s == W_STAGES) { s = 0; ph ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (the leader CTA of a pair issues for both) =====
        if (lane == 0) {
            const uint64_t b_desc_base = make_desc_w(smem_u32(wring));
            const uint64_t xlo_desc_base = make_desc(smem_u32(xlo));
            constexpr uint32_t GRAN_DESC = GRAN_BYTES >> 4, WTILE_DESC = STAGE_BYTES >> 4;
            const uint32_t d_tmem = tmem_base + TMEM_D_COL, a_tmem = tmem_base + TMEM_A_COL;
            int s = 0;
            uint32_t ph = 0, use = 0, duse = 0;  // use: MMA layers issued so far, duse: accumulator uses so far
            auto wait_stage = [&]() {            // this stage's weights are in shared memory (of both CTAs of a pair)
                DBG_T0();
                mbar_wait(&w_full[s], ph);
                if (NCTA == 2)
                    mbar_wait_cluster(&w_peer[s], ph);
                DBG_ADD(1);
                tc_fence_after();
            };
            auto next_stage = [&]() {
                mma_commit<NCTA>(&w_empty[s]);
                if (++s == W_STAGES) { s = 0; ph ^= 1; }
            };
#ifdef IKB_TC_DEBUG
            const long long _tstart = clock64();
#endif
            for (long long g = group0; g < n_groups; g += group_stride) {
                for (int m = 0; m < NM; ++m, ++use) {
                    for (int nh = 0; nh < NHALF; ++nh, ++duse) {
                        const uint32_t nfeat = (uint32_t)min(256, HP - 256 * nh);
                        // D fp32, A/B fp16, K-major, N = features of this half, M = 128 rows per CTA
                        const uint32_t idesc = (1u << 4) | ((nfeat >> 3) << 17) | (((128u * NCTA) >> 4) << 24);
                        // the epilogue (of both CTAs) has drained the accumulator
                        { DBG_T0(); if (NCTA == 2) mbar_wait_cluster(d_empty, (duse & 1) ^ 1); else mbar_wait(d_empty, (duse & 1) ^ 1); DBG_ADD(3); }
                        tc_fence_after();
                        for (int kt = 0; kt < KT; ++kt) {
                            const int k0 = kt * WK;
                            if (nh == 0 && (k0 & 127) == 0) {  // features k0 .. k0 + 127 of this layer's input are stored
                                { DBG_T0(); if (NCTA == 2) mbar_wait_cluster(&act_ready[k0 >> 7], use & 1); else mbar_wait(&act_ready[k0 >> 7], use & 1); DBG_ADD(2); }
                                tc_fence_after();
                            }
                            const uint32_t a_cols = a_tmem + (k0 >> 1);  // two k per 32-bit column
                            const uint64_t xlo_desc = xlo_desc_base + (uint64_t)((kt / WSUB) * GRAN_DESC + (kt % WSUB) * (WK >> 3));
                            // w_hi tile
                            wait_stage();
                            uint64_t b_desc = b_desc_base + (uint64_t)(s * WTILE_DESC);
#if IKB_TS_CORR_FIRST
                            // corrections first: x_lo (smem) w_hi opens the accumulator, x_hi w_hi waits for the second pass
#pragma unroll
                            for (int ks = 0; ks < WKSTEPS; ++ks)
                                mma_ss<NCTA>(d_tmem, xlo_desc + 2 * ks, b_desc + 2 * ks, idesc, (kt | ks) != 0);
#else
                            // x_hi (TMEM) and x_lo (smem) both multiply it
#pragma unroll
                            for (int ks = 0; ks < WKSTEPS; ++ks)
                                mma_ts<NCTA>(d_tmem, a_cols + ks * 8, b_desc + 2 * ks, idesc, (kt | ks) != 0);
#pragma unroll
                            for (int ks = 0; ks < WKSTEPS; ++ks)
                                mma_ss<NCTA>(d_tmem, xlo_desc + 2 * ks, b_desc + 2 * ks, idesc, 1);
#endif
                            next_stage();
                            // w_lo tile: only x_hi multiplies it (x_lo w_lo is below fp32 resolution)
                            wait_stage();
                            b_desc = b_desc_base + (uint64_t)(s * WTILE_DESC);
#pragma unroll
                            for (int ks = 0; ks < WKSTEPS; ++ks)
                                mma_ts<NCTA>(d_tmem, a_cols + ks * 8, b_desc + 2 * ks, idesc, 1);
                            next_stage();
                        }
#if IKB_TS_CORR_FIRST
                        // main product last: its 32 steps are the only ones that truncate at the sum's full size
                        for (int kt = 0; kt < KT; ++kt) {
                            const uint32_t a_cols = a_tmem + ((kt * WK) >> 1);
                            wait_stage();
                            const uint64_t b_desc = b_desc_base + (uint64_t)(s * WTILE_DESC);
#pragma unroll
                            for (int ks = 0; ks < WKSTEPS; ++ks)
                                mma_ts<NCTA>(d_tmem, a_cols + ks * 8, b_desc + 2 * ks, idesc, 1);
                            next_stage();
                        }
#endif
                        mma_commit<NCTA>(d_full);
                    }
                    mma_commit<NCTA>(a_free);  // every read of this layer's x_hi / x_lo has completed
                }
            }
#ifdef IKB_TC_DEBUG
            if (blockIdx.x == 0) g_tc2_dbg[0] += (unsigned long long)(clock64() - _tstart);
#endif
        }
    } else {
        // ===== epilogue warps: thread = (batch row = TMEM lane, half of the columns of the N half) =====
        const int ew = warp - 2;
        const int sub = warp & 3;             // TMEM sub-partition this warp may access
        const int ch = ew >> 2;               // column group (0 .. CW-1) within the N half
        const int row = sub * 32 + lane;      // batch row within the tile
        const int et = ew * 32 + lane;
        const uint32_t lane_base = (uint32_t)(sub * 32) << 16;
        uint32_t use = 0, duse = 0;
        for (long long g = group0; g < n_groups; g += group_stride) {
            const long long row0 = (g * NCTA + rank) * ROWS;   // past the end for the odd CTA of the last pair: an empty tile
            // ---- inputs: x_scaler.transform in fp64 -> fp32 (ann.py:72), workspace limits (inverse.py:154) ----
            if (et < ROWS) {
                const long long i = row0 + et;
                float v0 = 0.f, v1 = 0.f, v2 = 0.f;
                if (i < a.n) {
                    double x, y, z;
                    ikb_load_xyz(a.xyz, a.xyz_f64, i, x, y, z);
                    if (ikb_out_of_limits(a.rc, x, y, z))
                        atomicMin(&a.stats->first_out_of_limits, a.index_base + i);
                    v0 = (float)((x - net.mean_x[0]) / net.scale_x[0]);
                    v1 = (float)((y - net.mean_x[1]) / net.scale_x[1]);
                    v2 = (float)((z - net.mean_x[2]) / net.scale_x[2]);
                }
                s_io[et * 4 + 0] = v0; s_io[et * 4 + 1] = v1; s_io[et * 4 + 2] = v2;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32));
            const float in0 = s_io[row * 4], in1 = s_io[row * 4 + 1], in2 = s_io[row * 4 + 2];
            float out_acc[4] = {0.f, 0.f, 0.f, 0.f};  // output layer partial sums of this thread's features

            EpiCtx cx;
            cx.net = &net; cx.xlo = xlo; cx.a_free = a_free; cx.act_ready = act_ready; cx.tmem_base = tmem_base;
            cx.lane_base = lane_base; cx.row = row; cx.lane = lane; cx.HP = HP; cx.in0 = in0; cx.in1 = in1; cx.in2 = in2;
            cx.pair_peer = rank != 0;
            // ---- layer 1 (3 -> HP) on the CUDA cores ----
            for (int nh = 0; nh < NHALF; ++nh) {
                const int ngroups = min(256, HP - 256 * nh) / (32 * CW);
                for (int q = 0; q < ngroups; ++q)
                    first_layer_half(cx, 256 * nh + q * 32 * CW + 32 * ch, 32, NM == 0, 2 * nh + q, out_acc);
            }
            // ---- hidden layers 2..NH ----
            for (int m = 0; m < NM; ++m, ++use) {
                const float oscale = __ldg(net.out_scale + m);
                const float *bias = net.b_hidden + (size_t)(m + 1) * HP;
                for (int nh = 0; nh < NHALF; ++nh, ++duse) {
                    const int ngroups = min(256, HP - 256 * nh) / (32 * CW);
                    { DBG_T0(); mbar_wait(d_full, duse & 1); if (warp == 2 && lane == 0) DBG_ADD(4); }
                    tc_fence_after();
                    uint32_t d[QMAX][32];
#ifdef IKB_TC_DEBUG
                    const long long _te = clock64();
#endif
                    const uint32_t taddr = tmem_base + lane_base + TMEM_D_COL + 32 * ch;
#pragma unroll
                    for (int q = 0; q < QMAX; ++q)
                        if (q < ngroups)
                            tmem_ld32(taddr + q * 32 * CW, d[q]);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        epi_arrive(cx, d_empty);
                    finish_half(cx, d, 256 * nh + 32 * ch, ngroups, m == NM - 1, nh + 1 < NHALF, use & 1, 2 * nh, oscale,
                                bias, out_acc);
#ifdef IKB_TC_DEBUG
                    if (blockIdx.x == 0 && warp == 2 && lane == 0) g_tc2_dbg[6] += (unsigned long long)(clock64() - _te);
#endif
                }
            }
            // ---- output layer: combine the two column halves, bias, y_scaler.inverse_transform (ann.py:71-75) ----
            asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32));  // s_io inputs are no longer needed
            // column groups 1 .. CW-1 add their partial sums into s_io one after the other, group 0 finishes
#pragma unroll
            for (int g = 1; g < CW; ++g) {
                if (ch == g) {
                    float4 *slot = reinterpret_cast<float4 *>(s_io + row * 4);
                    float4 v = make_float4(out_acc[0], out_acc[1], out_acc[2], out_acc[3]);
                    if (g > 1) {
                        const float4 o = *slot;
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    *slot = v;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32));
            }
            if (ch == 0) {
                const bool live = row0 + row < a.n;
                const float4 other = *reinterpret_cast<const float4 *>(s_io + row * 4);
                float yv[4] = {(out_acc[0] + other.x) * (1.0f / X_SCALE), (out_acc[1] + other.y) * (1.0f / X_SCALE),
                               (out_acc[2] + other.z) * (1.0f / X_SCALE), (out_acc[3] + other.w) * (1.0f / X_SCALE)};
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    yv[o] += net.b_last[o];
                    yv[o] = __fmul_rn(yv[o], net.scale_y[o]);
                    yv[o] = __fadd_rn(yv[o], net.mean_y[o]);
                }
                if (live)
                    reinterpret_cast<float4 *>(a.out)[row0 + row] = make_float4(yv[0], yv[1], yv[2], yv[3]);
                if (a.fk_err || a.fk_stats)  // the ch == 0 threads are four whole warps: warp-uniform branch
                    fused_fk_error(a, row0 + row, live, yv);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (NCTA == 2)
        cluster_sync_all();   // the peer may still read this CTA's shared memory / arrive on its barriers until here
    if (warp == 1) {
        if (NCTA == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

size_t tc2_smem_bytes(int hp)
{
    return (size_t)(hp / 64) * GRAN_BYTES + (size_t)W_RING_BYTES + ROWS * 4 * sizeof(float) + 48 * 8 + 64;
}

}  // namespace

struct IkbMlpTc2 {
    bool usable = false;
    std::string why;
    Tc2Net net;
    void *arena = nullptr;
};

IkbMlpTc2 *ikb_mlp_tc2_new() { return new IkbMlpTc2(); }
void ikb_mlp_tc2_delete(IkbMlpTc2 *t)
{
    if (t) {
        if (t->arena)
            cudaFree(t->arena);
        delete t;
    }
}

int ikb_mlp_tc2_pack(IkbMlpTc2 &t, int n_layers, const int *dims, const float *const *weights,
                     const float *const *biases, const double mean_x[3], const double scale_x[3],
                     const double mean_y[4], const double scale_y[4], std::string &err)
{
    if (t.arena)
        cudaFree(t.arena);
    t.arena = nullptr;
    t.usable = false;
    const int nh = n_layers - 1;
    if (nh < 1) {
        t.why = "needs at least one hidden layer";
        return IKB_OK;
    }
    int hmax = 0;
    for (int l = 1; l <= nh; ++l)
        hmax = dims[l] > hmax ? dims[l] : hmax;
    const int hp = ((hmax + 127) / 128) * 128;
    const int KT = hp / WK, NHALF = (hp + 255) / 256, NM = nh - 1;
    const size_t tile_halfs = WTILE_BYTES / sizeof(__half);
    const size_t n_tiles = (size_t)NM * NHALF * KT * 2;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_tiles = take((n_tiles ? n_tiles : 1) * WTILE_BYTES);
    const size_t o_first = take((size_t)3 * hp * sizeof(float));
    const size_t o_bias = take((size_t)(1 + NM) * hp * sizeof(float));
    const size_t o_scale = take((size_t)(NM > 0 ? NM : 1) * sizeof(float));
    const size_t o_last = take((size_t)hp * 4 * sizeof(float));
    std::vector<unsigned char> host(off, 0);
    __half *tiles = reinterpret_cast<__half *>(host.data() + o_tiles);
    float *wfirst = reinterpret_cast<float *>(host.data() + o_first);
    float *bias = reinterpret_cast<float *>(host.data() + o_bias);
    float *oscale = reinterpret_cast<float *>(host.data() + o_scale);
    float *wlast = reinterpret_cast<float *>(host.data() + o_last);
    for (int k = 0; k < 3; ++k)
        for (int f = 0; f < dims[1]; ++f)
            wfirst[(size_t)k * hp + f] = weights[0][(size_t)k * dims[1] + f];
    // When the padded width has a spare column, feature `hmax` is kept at a constant 1 (its own weight makes the
    // pre-activation 20, tanh(20) == 1.0f) and every layer's biases ride along as weight row `hmax`: the MMA adds
    // them for free and the epilogue needs no per-element bias loads.
    const int bias_row = (hp > hmax && NM > 0) ? hmax : -1;
    constexpr float kConstPre = 20.0f;
    for (int l = 0; l < nh; ++l)
        for (int f = 0; f < dims[l + 1]; ++f)
            bias[(size_t)l * hp + f] = (bias_row >= 0 && l > 0) ? 0.f : biases[l][f];
    if (bias_row >= 0)
        bias[bias_row] = kConstPre;  // layer 1 (CUDA cores) produces the constant feature through its bias
    auto weight_at = [&](int l, int k, int f) -> float {  // true (unscaled) weight of MMA layer l at (k, f)
        const int fin = dims[l], fout = dims[l + 1];
        if (k < fin && f < fout)
            return weights[l][(size_t)k * fout + f];
        if (k == bias_row)
            return f < fout ? biases[l][f] : (f == bias_row ? kConstPre : 0.f);
        return 0.f;
    };
    for (int m = 0; m < NM; ++m) {
        const int l = m + 1, fin = dims[l], fout = dims[l + 1];
        float wmax = bias_row >= 0 ? kConstPre : 0.f;
        for (size_t i = 0; i < (size_t)fin * fout; ++i)
            wmax = std::fmax(wmax, std::fabs(weights[l][i]));
        if (bias_row >= 0)
            for (int f = 0; f < fout; ++f)
                wmax = std::fmax(wmax, std::fabs(biases[l][f]));
        int e = 0;
        if (wmax > 0.f)
            e = 12 - (int)std::ceil(std::log2(wmax));  // largest weight near 2^12: w_lo stays normal, sums stay small
        e = e > 24 ? 24 : (e < -8 ? -8 : e);
        const float sw = std::ldexp(1.0f, e);
        // truncating steps at the sum's full size (see ikb_tc_truncation_compensation): the main product's alone when
        // the two corrections are accumulated first, all three products' when they are interleaved
        const int acc_steps = (IKB_TS_CORR_FIRST ? 1 : 3) * ((fin + (bias_row >= 0 ? 1 : 0) + 15) / 16);
        oscale[m] = (float)(ikb_tc_truncation_compensation(acc_steps) / ((double)sw * X_SCALE));
        for (int nhalf = 0; nhalf < NHALF; ++nhalf)
            for (int kt = 0; kt < KT; ++kt) {
                __half *hi = tiles + (((size_t)(m * NHALF + nhalf) * KT + kt) * 2 + 0) * tile_halfs;
                __half *lo = hi + tile_halfs;
                const int nfeat = std::min(256, hp - 256 * nhalf);
                for (int r = 0; r < nfeat; ++r)
                    for (int c = 0; c < WK; ++c) {
                        const int f = nhalf * 256 + r, k = kt * WK + c;
                        const float w = weight_at(l, k, f) * sw;
                        const __half h = __float2half_rn(w);
                        const int o = wswz_off(r, c) / 2;
                        hi[o] = h;
                        lo[o] = __float2half_rn(w - __half2float(h));
                    }
            }
    }
    for (int k = 0; k < dims[nh]; ++k)
        for (int o = 0; o < 4; ++o)
            wlast[(size_t)k * 4 + o] = weights[nh][(size_t)k * 4 + o];
    cudaError_t ce = cudaMalloc(&t.arena, off);
    if (ce == cudaSuccess)
        ce = cudaMemcpy(t.arena, host.data(), off, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        err = std::string("ikb_mlp_load (TS pack): ") + cudaGetErrorString(ce);
        return IKB_ERR_CUDA;
    }
    Tc2Net &n = t.net;
    memset(&n, 0, sizeof n);
    n.n_mma_layers = NM;
    n.hp = hp;
    n.bias_in_mma = bias_row >= 0 ? 1 : 0;
    n.w_tiles = reinterpret_cast<const __half *>((char *)t.arena + o_tiles);
    n.w_first = reinterpret_cast<const float *>((char *)t.arena + o_first);
    n.b_hidden = reinterpret_cast<const float *>((char *)t.arena + o_bias);
    n.out_scale = reinterpret_cast<const float *>((char *)t.arena + o_scale);
    n.w_last = reinterpret_cast<const float *>((char *)t.arena + o_last);
    for (int o = 0; o < 4; ++o) {
        n.b_last[o] = biases[nh][o];
        n.mean_y[o] = (float)mean_y[o];
        n.scale_y[o] = (float)scale_y[o];
    }
    for (int j = 0; j < 3; ++j) {
        n.mean_x[j] = mean_x[j];
        n.scale_x[j] = scale_x[j];
    }
    ce = cudaFuncSetAttribute(mlp_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2_smem_bytes(hp));
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(mlp_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2_smem_bytes(hp));
    if (ce != cudaSuccess) {
        err = std::string("ikb_mlp_load (TS smem attribute): ") + cudaGetErrorString(ce);
        return IKB_ERR_CUDA;
    }
    t.usable = true;
    return IKB_OK;
}

int ikb_mlp_tc2_launch(const IkbMlpTc2 &t, const void *xyz, int xyz_f64, long long n, long long index_base,
                       float *angles_out, float *fk_err_out, int fk_stats, IkbDeviceStats *stats,
                       const IkbRobot &rc, int num_sms, cudaStream_t stream, std::string &err)
{
    if (!t.usable) {
        err = "IKB_MLP_FP16X3_TS: this network cannot use the tensor-core path (" + t.why + ")";
        return IKB_ERR_UNSUPPORTED;
    }
    Tc2Args a;
    a.xyz = xyz; a.xyz_f64 = xyz_f64; a.n = n; a.index_base = index_base; a.out = angles_out;
    a.fk_err = fk_err_out; a.fk_stats = fk_stats;
    a.stats = stats; a.rc = rc; a.net = t.net;
    const long long tiles = (n + ROWS - 1) / ROWS;
    // CTA pairs (IKB_TS_CLUSTER=2; off by default: measured slower): clusters of two CTAs on the
    // SMs of one TPC, an even grid
    static const bool pairs = [] { const char *v = std::getenv("IKB_TS_CLUSTER"); return v && v[0] == '2'; }();
    cudaError_t ce;
    if (pairs && num_sms >= 2) {
        const long long groups = (tiles + 1) / 2;
        const unsigned grid = 2u * (unsigned)(groups < num_sms / 2 ? groups : num_sms / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = tc2_smem_bytes(t.net.hp);
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ce = cudaLaunchKernelEx(&cfg, mlp_tc2_kernel<2>, a);
    } else {
        const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);
        mlp_tc2_kernel<1><<<grid, THREADS, tc2_smem_bytes(t.net.hp), stream>>>(a);
        ce = cudaGetLastError();
    }
    if (ce != cudaSuccess) {
        err = std::string("mlp_tc2_kernel launch: ") + cudaGetErrorString(ce);
        return IKB_ERR_CUDA;
    }
    return IKB_OK;
}

#ifdef IKB_TC_DEBUG
extern "C" void ikbdbg_tc2_counters(unsigned long long *out, int reset)
{
    cudaMemcpyFromSymbol(out, g_tc2_dbg, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_tc2_dbg, z, sizeof z);
    }
}
#endif
