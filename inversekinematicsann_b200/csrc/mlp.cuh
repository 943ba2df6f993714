// mlp.cuh -- host-side state of the fused scaler -> MLP -> scaler inference path (K2).
#pragma once
#include <string>

#include "ikb_common.cuh"

#define IKB_MLP_MAX_LAYERS 16
#define IKB_MLP_MAX_WIDTH 512

// Device view of the network handed to the kernels by value.
struct IkbMlpDevice {
    int n_layers;                       // Dense layers (hidden tanh layers + 1 linear output layer)
    int in_dim[IKB_MLP_MAX_LAYERS];     // real fan-in
    int out_dim[IKB_MLP_MAX_LAYERS];    // real fan-out
    int kp[IKB_MLP_MAX_LAYERS];         // fan-in padded to a multiple of 16 (zero rows)
    int np[IKB_MLP_MAX_LAYERS];         // fan-out padded to a multiple of 128 (hidden) / 4 (output)
    const float *W[IKB_MLP_MAX_LAYERS]; // [kp][np] row-major fp32, zero padded (Keras kernel layout)
    const float *b[IKB_MLP_MAX_LAYERS]; // [np]
    double mean_x[3], scale_x[3];       // StandardScaler.transform, applied in fp64 (ann.py:72)
    float mean_y[4], scale_y[4];        // StandardScaler.inverse_transform on the fp32 output
};

struct IkbMlpTc;  // tensor-core packing of the same network (mlp_tc.cu)
IkbMlpTc *ikb_mlp_tc_new();
void ikb_mlp_tc_delete(IkbMlpTc *t);
int ikb_mlp_tc_pack(IkbMlpTc &t, int n_layers, const int *dims, const float *const *weights,
                    const float *const *biases, const double mean_x[3], const double scale_x[3],
                    const double mean_y[4], const double scale_y[4], std::string &err);
int ikb_mlp_tc_launch(const IkbMlpTc &t, const void *xyz, int xyz_f64, long long n, long long index_base,
                      float *angles_out, IkbDeviceStats *stats, const IkbRobot &rc, int num_sms,
                      cudaStream_t stream, std::string &err);

struct IkbMlpTc2;  // second tensor-core layout: activations as a TMEM A operand (mlp_tc2.cu)
IkbMlpTc2 *ikb_mlp_tc2_new();
void ikb_mlp_tc2_delete(IkbMlpTc2 *t);
int ikb_mlp_tc2_pack(IkbMlpTc2 &t, int n_layers, const int *dims, const float *const *weights,
                     const float *const *biases, const double mean_x[3], const double scale_x[3],
                     const double mean_y[4], const double scale_y[4], std::string &err);
int ikb_mlp_tc2_launch(const IkbMlpTc2 &t, const void *xyz, int xyz_f64, long long n, long long index_base,
                       float *angles_out, IkbDeviceStats *stats, const IkbRobot &rc, int num_sms,
                       cudaStream_t stream, std::string &err);

struct IkbMlp {
    bool loaded = false;
    IkbMlpTc *tc = nullptr;
    IkbMlpTc2 *tc2 = nullptr;
    IkbMlpDevice dev;
    void *arena = nullptr;  // one device allocation holding all padded weights and biases
    size_t arena_bytes = 0;
    long long macs_per_row = 0;
};

int ikb_mlp_upload(IkbMlp &m, int n_layers, const int *dims, const float *const *weights,
                   const float *const *biases, const double mean_x[3], const double scale_x[3],
                   const double mean_y[4], const double scale_y[4], std::string &err);
void ikb_mlp_free(IkbMlp &m);
int ikb_mlp_launch(const IkbMlp &m, const void *xyz, int xyz_f64, long long n, long long index_base,
                   float *angles_out, int mode, IkbDeviceStats *stats, const IkbRobot &rc, int num_sms,
                   cudaStream_t stream, std::string &err, int &launches);
