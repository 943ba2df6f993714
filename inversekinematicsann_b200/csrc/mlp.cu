// mlp.cu -- K2: fused StandardScaler -> Dense MLP (tanh) -> StandardScaler inference (sm_100a).
//
// Replaces reference ANN.predict (ann.py:70-76) behind AnnInverseKinematics.ikine (inverse.py:152-155):
//   y_scaler.inverse_transform(model.predict(x_scaler.transform(points)))
// for the Sequential model of ann.py:46-56 (Input(3), 12 x Dense(500, tanh), Dense(4)).
//
// IKB_MLP_FP32_SIMT (this file): one persistent CTA owns a tile of 64 targets; their activations
// stay in shared memory, transposed ([k][row]), through every layer; each thread keeps an
// 8 x 16 block of the layer's outputs in registers.  Weight slabs (16 input rows x padded width,
// contiguous in HBM/L2) are streamed into a two-stage shared-memory ring by the TMA engine
// (cp.async.bulk + mbarrier complete_tx), so the FFMA loop only touches shared memory.
// Both scalers are applied inside the kernel (input load / output store): no separate passes.
#include <vector>

#include "mlp.cuh"

namespace {

constexpr int TM = 64;            // targets per CTA tile
constexpr int KC = 16;            // weight rows per slab
constexpr int THREADS = 256;
constexpr int MAXW = IKB_MLP_MAX_WIDTH;
constexpr size_t SMEM_A = (size_t)MAXW * TM * sizeof(float);          // 128 KB activations
constexpr size_t SMEM_W = (size_t)2 * KC * MAXW * sizeof(float);      // 64 KB weight ring
constexpr size_t SMEM_MISC = 2048;                                    // barriers + output staging
constexpr size_t SMEM_TOTAL = SMEM_A + SMEM_W + SMEM_MISC;

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP), completion on an mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

struct MlpArgs {
    const void *xyz;
    int xyz_f64;
    long long n;
    long long index_base;
    float *out;
    IkbDeviceStats *stats;
    IkbRobot rc;
    IkbMlpDevice net;
};

__global__ void __launch_bounds__(THREADS, 1) mlp_simt_kernel(const MlpArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sA = reinterpret_cast<float *>(smem_raw);                   // [MAXW][TM]
    float *sW = reinterpret_cast<float *>(smem_raw + SMEM_A);          // [2][KC][np]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + SMEM_A + SMEM_W);  // [2]
    float *sOut = reinterpret_cast<float *>(smem_raw + SMEM_A + SMEM_W + 64);   // [TM][4]

    const int tid = threadIdx.x;
    const int rg = tid & 7;    // row group: rows rg*8 .. rg*8+7
    const int cg = tid >> 3;   // column group: columns j*128 + cg*4 .. +3, j = 0..3
    const IkbMlpDevice &net = a.net;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};

    const long long n_tiles = (a.n + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        // ---- input stage: x_scaler.transform in fp64, cast to fp32 (ann.py:72, Keras float32) ----
        if (tid < TM) {
            const long long i = row0 + tid;
            float v[3] = {0.f, 0.f, 0.f};
            if (i < a.n) {
                double x, y, z;
                ikb_load_xyz(a.xyz, a.xyz_f64, i, x, y, z);
                if (ikb_out_of_limits(a.rc, x, y, z))
                    atomicMin(&a.stats->first_out_of_limits, a.index_base + i);
                v[0] = (float)((x - net.mean_x[0]) / net.scale_x[0]);
                v[1] = (float)((y - net.mean_x[1]) / net.scale_x[1]);
                v[2] = (float)((z - net.mean_x[2]) / net.scale_x[2]);
            }
            for (int k = 0; k < net.kp[0]; ++k)
                sA[k * TM + tid] = k < 3 ? v[k] : 0.f;
        }
        __syncthreads();

        // ---- hidden layers: tanh(h @ W + b), activations updated in place --------------------
        for (int l = 0; l + 1 < net.n_layers; ++l) {
            const int kp = net.kp[l], np = net.np[l];
            const int J = np >> 7;  // column blocks of 128
            const int n_slabs = kp / KC;
            const uint32_t slab_bytes = (uint32_t)(KC * np * sizeof(float));
            const float *Wl = net.W[l];
            float acc[8][16];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    acc[r][q] = 0.f;
            if (tid == 0) {
                mbar_expect_tx(&bars[0], slab_bytes);
                tma_load_1d(sW, Wl, slab_bytes, &bars[0]);
            }
            for (int s = 0; s < n_slabs; ++s) {
                const int st = s & 1;
                mbar_wait(&bars[st], phase[st]);
                phase[st] ^= 1;
                __syncthreads();  // everyone is done with slab s-1, its stage may be refilled
                if (tid == 0 && s + 1 < n_slabs) {
                    mbar_expect_tx(&bars[st ^ 1], slab_bytes);
                    tma_load_1d(sW + (size_t)(st ^ 1) * KC * MAXW, Wl + (size_t)(s + 1) * KC * np, slab_bytes,
                                &bars[st ^ 1]);
                }
                const float *w_st = sW + (size_t)st * KC * MAXW;
                const float *a_st = sA + (size_t)s * KC * TM + rg * 8;
#pragma unroll 4
                for (int kk = 0; kk < KC; ++kk) {
                    const float4 a0 = *reinterpret_cast<const float4 *>(a_st + kk * TM);
                    const float4 a1 = *reinterpret_cast<const float4 *>(a_st + kk * TM + 4);
                    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < J) {
                            const float4 w4 =
                                *reinterpret_cast<const float4 *>(w_st + kk * np + j * 128 + cg * 4);
                            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int r = 0; r < 8; ++r)
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    acc[r][j * 4 + q] = fmaf(av[r], wv[q], acc[r][j * 4 + q]);
                        }
                    }
                }
            }
            __syncthreads();  // all reads of this layer's input activations are done
            const float *bl = net.b[l];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < J) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int col = j * 128 + cg * 4 + q;
                        const float bias = __ldg(bl + col);
                        float4 o0, o1;
                        o0.x = tanhf(acc[0][j * 4 + q] + bias); o0.y = tanhf(acc[1][j * 4 + q] + bias);
                        o0.z = tanhf(acc[2][j * 4 + q] + bias); o0.w = tanhf(acc[3][j * 4 + q] + bias);
                        o1.x = tanhf(acc[4][j * 4 + q] + bias); o1.y = tanhf(acc[5][j * 4 + q] + bias);
                        o1.z = tanhf(acc[6][j * 4 + q] + bias); o1.w = tanhf(acc[7][j * 4 + q] + bias);
                        *reinterpret_cast<float4 *>(sA + col * TM + rg * 8) = o0;
                        *reinterpret_cast<float4 *>(sA + col * TM + rg * 8 + 4) = o1;
                    }
                }
            }
            __syncthreads();
        }

        // ---- output layer (linear) + y_scaler.inverse_transform on fp32 (ann.py:71-75) ----------
        {
            const int l = net.n_layers - 1;
            const int row = tid & (TM - 1), nout = tid >> 6;  // 64 rows x 4 outputs = 256 threads
            const float *Wl = net.W[l];
            float y = 0.f;
            for (int k = 0; k < net.kp[l]; ++k)
                y = fmaf(sA[k * TM + row], __ldg(Wl + k * 4 + nout), y);
            y += __ldg(net.b[l] + nout);
            y = __fmul_rn(y, net.scale_y[nout]);
            y = __fadd_rn(y, net.mean_y[nout]);
            sOut[row * 4 + nout] = y;
        }
        __syncthreads();
        if (tid < TM && row0 + tid < a.n)
            reinterpret_cast<float4 *>(a.out)[row0 + tid] = *reinterpret_cast<const float4 *>(sOut + tid * 4);
        __syncthreads();
    }
}

}  // namespace

void ikb_mlp_free(IkbMlp &m)
{
    if (m.arena)
        cudaFree(m.arena);
    m.arena = nullptr;
    m.loaded = false;
    ikb_mlp_tc_delete(m.tc);
    m.tc = nullptr;
    ikb_mlp_tc2_delete(m.tc2);
    m.tc2 = nullptr;
}

int ikb_mlp_upload(IkbMlp &m, int n_layers, const int *dims, const float *const *weights,
                   const float *const *biases, const double mean_x[3], const double scale_x[3],
                   const double mean_y[4], const double scale_y[4], std::string &err)
{
    if (n_layers < 2 || n_layers > IKB_MLP_MAX_LAYERS) {
        err = "ikb_mlp_load: need 2..16 Dense layers";
        return IKB_ERR_UNSUPPORTED;
    }
    if (dims[0] != 3 || dims[n_layers] != 4) {
        err = "ikb_mlp_load: the network must map 3 inputs to 4 outputs (ann.py:44,56)";
        return IKB_ERR_UNSUPPORTED;
    }
    for (int l = 1; l < n_layers; ++l)
        if (dims[l] < 1 || dims[l] > IKB_MLP_MAX_WIDTH) {
            err = "ikb_mlp_load: hidden width must be 1..512";
            return IKB_ERR_UNSUPPORTED;
        }
    ikb_mlp_free(m);
    IkbMlpDevice d;
    memset(&d, 0, sizeof d);
    d.n_layers = n_layers;
    size_t total = 0;
    std::vector<size_t> w_off(n_layers), b_off(n_layers);
    for (int l = 0; l < n_layers; ++l) {
        d.in_dim[l] = dims[l];
        d.out_dim[l] = dims[l + 1];
        // fan-in = previous layer's padded width so padded activations (exact zeros) meet zero rows
        d.kp[l] = l == 0 ? KC : d.np[l - 1];
        d.np[l] = l + 1 < n_layers ? ((dims[l + 1] + 127) / 128) * 128 : 4;
        w_off[l] = total;
        total += (size_t)d.kp[l] * d.np[l] * sizeof(float);
        b_off[l] = total;
        total += (((size_t)d.np[l] * sizeof(float)) + 127) / 128 * 128;
    }
    std::vector<unsigned char> host(total, 0);
    long long macs = 0;
    for (int l = 0; l < n_layers; ++l) {
        float *W = reinterpret_cast<float *>(host.data() + w_off[l]);
        float *b = reinterpret_cast<float *>(host.data() + b_off[l]);
        for (int k = 0; k < dims[l]; ++k)
            for (int n = 0; n < dims[l + 1]; ++n)
                W[(size_t)k * d.np[l] + n] = weights[l][(size_t)k * dims[l + 1] + n];
        for (int n = 0; n < dims[l + 1]; ++n)
            b[n] = biases[l][n];
        macs += (long long)dims[l] * dims[l + 1];
    }
    cudaError_t ce = cudaMalloc(&m.arena, total);
    if (ce != cudaSuccess) {
        err = std::string("ikb_mlp_load: cudaMalloc: ") + cudaGetErrorString(ce);
        return IKB_ERR_CUDA;
    }
    ce = cudaMemcpy(m.arena, host.data(), total, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        err = std::string("ikb_mlp_load: cudaMemcpy: ") + cudaGetErrorString(ce);
        return IKB_ERR_CUDA;
    }
    for (int l = 0; l < n_layers; ++l) {
        d.W[l] = reinterpret_cast<const float *>((char *)m.arena + w_off[l]);
        d.b[l] = reinterpret_cast<const float *>((char *)m.arena + b_off[l]);
    }
    for (int j = 0; j < 3; ++j) {
        d.mean_x[j] = mean_x[j];
        d.scale_x[j] = scale_x[j];
    }
    for (int j = 0; j < 4; ++j) {
        d.mean_y[j] = (float)mean_y[j];  // sklearn casts scale_/mean_ to the fp32 array's dtype
        d.scale_y[j] = (float)scale_y[j];
    }
    m.tc = ikb_mlp_tc_new();
    {
        const int rc = ikb_mlp_tc_pack(*m.tc, n_layers, dims, weights, biases, mean_x, scale_x, mean_y, scale_y, err);
        if (rc != IKB_OK)
            return rc;
    }
    m.tc2 = ikb_mlp_tc2_new();
    {
        const int rc = ikb_mlp_tc2_pack(*m.tc2, n_layers, dims, weights, biases, mean_x, scale_x, mean_y, scale_y, err);
        if (rc != IKB_OK)
            return rc;
    }
    m.dev = d;
    m.arena_bytes = total;
    m.macs_per_row = macs;
    m.loaded = true;
    static bool attr_set = false;
    if (!attr_set) {
        ce = cudaFuncSetAttribute(mlp_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL);
        if (ce != cudaSuccess) {
            err = std::string("ikb_mlp_load: cudaFuncSetAttribute: ") + cudaGetErrorString(ce);
            return IKB_ERR_CUDA;
        }
        attr_set = true;
    }
    return IKB_OK;
}

cudaError_t ikb_launch_fk(const void *angles, int angles_f64, long long n, long long index_base,
                          void *pos_out, const void *targets, int xyz_f64, void *err_out,
                          IkbDeviceStats *stats, const IkbRobot &rc, int num_sms, cudaStream_t stream);

int ikb_mlp_launch(const IkbMlp &m, const void *xyz, int xyz_f64, long long n, long long index_base,
                   float *angles_out, float *fk_err_out, int fk_stats, int mode, IkbDeviceStats *stats,
                   const IkbRobot &rc, int num_sms, cudaStream_t stream, std::string &err, int &launches)
{
    launches = 0;
    if (n <= 0)
        return IKB_OK;
    if (mode == IKB_MLP_FP16X3_TS) {  // default mode: the FK error is part of the kernel's output stage
        const int rc_tc = ikb_mlp_tc2_launch(*m.tc2, xyz, xyz_f64, n, index_base, angles_out, fk_err_out, fk_stats,
                                             stats, rc, num_sms, stream, err);
        launches = rc_tc == IKB_OK ? 1 : 0;
        return rc_tc;
    }
    if (mode == IKB_MLP_FP16X3_TC) {
        const int rc_tc = ikb_mlp_tc_launch(*m.tc, xyz, xyz_f64, n, index_base, angles_out, stats, rc, num_sms,
                                            stream, err);
        if (rc_tc != IKB_OK)
            return rc_tc;
        launches = 1;
    } else if (mode == IKB_MLP_FP32_SIMT) {
        MlpArgs a;
        a.xyz = xyz; a.xyz_f64 = xyz_f64; a.n = n; a.index_base = index_base; a.out = angles_out;
        a.stats = stats; a.rc = rc; a.net = m.dev;
        const long long tiles = (n + TM - 1) / TM;
        const unsigned grid = (unsigned)(tiles < num_sms ? tiles : num_sms);
        mlp_simt_kernel<<<grid, THREADS, SMEM_TOTAL, stream>>>(a);
        cudaError_t ce = cudaGetLastError();
        if (ce != cudaSuccess) {
            err = std::string("mlp_simt_kernel launch: ") + cudaGetErrorString(ce);
            return IKB_ERR_CUDA;
        }
        launches = 1;
    } else {
        err = "ikb_ann_solve: unknown MLP mode";
        return IKB_ERR_INVALID;
    }
    if (fk_err_out || fk_stats) {  // the cross-check modes run K3 as a second launch on the same stream
        cudaError_t ce = ikb_launch_fk(angles_out, 0, n, index_base, nullptr, xyz, xyz_f64, fk_err_out, stats, rc,
                                       num_sms, stream);
        if (ce != cudaSuccess) {
            err = std::string("fk_kernel launch: ") + cudaGetErrorString(ce);
            return IKB_ERR_CUDA;
        }
        ++launches;
    }
    return IKB_OK;
}
