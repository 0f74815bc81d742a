// Device-side parameter block of the SubNet kernel (see subnet.cu, pack in api.cu).
#pragma once
namespace gd {
struct SubnetParams {
    const float* conv[8];      // per conv: [Cin][9][Cout] BN-folded weights followed by [Cout] folded biases
    const float *l1w, *l1b;    // Linear(1025, 64)
    const float *l2w, *l2b;    // Linear(64, 64)
    const float *l3w, *l3b;    // Linear(64, n_out)
    int n_out;
};
}  // namespace gd
