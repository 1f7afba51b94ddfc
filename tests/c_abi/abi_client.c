/* A plain-C client of include/sqgpu.h: what a reference-side shim does, without Python or torch.
 * Builds the 3-qubit structure [U3(0) U3(1) CRY(0,1) U3(2) CNOT(2,0) RZ(1)], evaluates cost + gradient for two parameter
 * vectors on U = identity and prints them; tests/test_gpu_parity.py compares the numbers with the Python binding and the
 * oracle. Usage: abi_client [n_devices [mode]] -- n_devices > 1 (accelerator_num = G) makes ONE handle over G GPUs with
 * sqgpu_create_multi (mode: 0 auto, 1 batch, 2 columns); every other call is the same. Prints "cost[b] ..." / "grad[b] ..." lines
 * (single device: also "shift<s>[b] ..." = cost(params_b + shifts[s] e_p) for every p, sqgpu_cost_shifted_batched). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sqgpu.h"

#define CHECK(call)                                                                    \
    do {                                                                               \
        int rc_ = (call);                                                              \
        if (rc_ != SQGPU_OK) {                                                         \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, sqgpu_last_error());  \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

static sqgpu_gate_desc gate(int type, int target, int control, int param_start, int n_params) {
    sqgpu_gate_desc g;
    memset(&g, 0, sizeof(g));
    g.type = type;
    g.target = target;
    g.control = control;
    g.target2 = g.control2 = -1;
    g.param_start = param_start;
    g.n_params = n_params;
    return g;
}

int main(int argc, char** argv) {
    int ndev = 0;
    const int want = argc > 1 ? atoi(argv[1]) : 1, mode = argc > 2 ? atoi(argv[2]) : SQGPU_SHARD_AUTO;
    CHECK(sqgpu_device_count(&ndev));
    if (ndev < want || want < 1) {
        fprintf(stderr, "%d device(s) visible, %d wanted\n", ndev, want);
        return 2;
    }
    sqgpu_handle_t h;
    if (want > 1) {
        int n_in_handle = 0, mode_in_force = -1;
        CHECK(sqgpu_create_multi(want, NULL, mode, &h));
        CHECK(sqgpu_multi_info(h, &n_in_handle, &mode_in_force));
        if (n_in_handle != want) return 3;
    } else {
        CHECK(sqgpu_create(0, &h));
    }
    const int n = 3, dim = 8, P = 11, B = 2;
    sqgpu_gate_desc gates[6];
    gates[0] = gate(SQGPU_U3, 0, -1, 0, 3);
    gates[1] = gate(SQGPU_U3, 1, -1, 3, 3);
    gates[2] = gate(SQGPU_CRY, 0, 1, 6, 1);
    gates[3] = gate(SQGPU_U3, 2, -1, 7, 3);
    gates[4] = gate(SQGPU_CNOT, 2, 0, 10, 0);
    gates[5] = gate(SQGPU_RZ, 1, -1, 10, 1);
    double U[2 * 64];
    memset(U, 0, sizeof(U));
    for (int i = 0; i < dim; ++i) U[2 * (i * dim + i)] = 1.0;
    CHECK(sqgpu_upload_matrix(h, U, dim, dim, dim));
    CHECK(sqgpu_set_circuit(h, gates, 6, P, n, NULL, 0));
    CHECK(sqgpu_set_cost(h, SQGPU_FROBENIUS_NORM, 0, 1.0, 1.0 / 1.7, 0.5));
    double params[2 * 11], cost[2], grad[2 * 11];
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p) params[b * P + p] = 0.1 * (p + 1) + 0.37 * b;
    CHECK(sqgpu_cost_grad_batched(h, params, B, cost, grad));
    for (int b = 0; b < B; ++b) {
        printf("cost[%d] %.17g\n", b, cost[b]);
        printf("grad[%d]", b);
        for (int p = 0; p < P; ++p) printf(" %.17g", grad[b * P + p]);
        printf("\n");
    }
    if (want == 1) {  /* the shift batch of the parameter-shift engines from one sweep (single-device handles) */
        const double shifts[2] = {1.5707963267948966, 3.141592653589793};
        double shifted[2 * 2 * 11];
        CHECK(sqgpu_cost_shifted_batched(h, params, B, shifts, 2, cost, shifted));
        for (int s = 0; s < 2; ++s)
            for (int b = 0; b < B; ++b) {
                printf("shift%d[%d]", s, b);
                for (int p = 0; p < P; ++p) printf(" %.17g", shifted[(s * B + b) * P + p]);
                printf("\n");
            }
    }
    CHECK(sqgpu_destroy(h));
    return 0;
}
