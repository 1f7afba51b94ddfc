/*
 * sq_oracle.c -- CPU restatement (plain C11, scalar, no FMA contraction) of the SQUANDER decomposition hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see sq_oracle.h). Parity status: PINNED against oracle/_ref/libsqref.so (the
 * reference's own translation units) by tests/test_oracle_vs_reference.py and against tests/golden/ fixtures.
 *
 * Every function cites the reference lines it follows; paths are relative to
 * /root/reference/squander/src-cpp/ .
 */
#define _GNU_SOURCE
#include "sq_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cplx;

/* scalar product with the reference's rounding order, common/common.cpp:298-306 */
static inline cplx cmul(cplx a, cplx b) {
    cplx r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}

static void kset(double* k, int idx, double re, double im) { k[2 * idx] = re; k[2 * idx + 1] = im; }

/* ------------------------------------------------------------------------------------------------------------ */
/* kernel builders: gates/include/gate_kernel_templates.h                                                        */
/* ------------------------------------------------------------------------------------------------------------ */

/* calc_one_qubit_u3_from_trig_to, gate_kernel_templates.h:508-522 */
static void u3_from_trig(double* k, double st, double ct, double sp, double cp, double sl, double cl) {
    const double spl = sp * cl + cp * sl;
    const double cpl = cp * cl - sp * sl;
    kset(k, 0, ct, 0.0);
    kset(k, 1, -st * cl, -st * sl);
    kset(k, 2, st * cp, st * sp);
    kset(k, 3, ct * cpl, ct * spl);
}

/* u3_derivative_kernel_{theta,phi,lambda}_from_trig_to, gate_kernel_templates.h:554-609 */
static void u3_deriv_from_trig(double* k, int pidx, double st, double ct, double sp, double cp, double sl, double cl) {
    const double spl = sp * cl + cp * sl;
    const double cpl = cp * cl - sp * sl;
    if (pidx == 0) {
        kset(k, 0, -st, 0.0);
        kset(k, 1, -ct * cl, -ct * sl);
        kset(k, 2, ct * cp, ct * sp);
        kset(k, 3, -st * cpl, -st * spl);
    } else if (pidx == 1) {
        kset(k, 0, 0.0, 0.0);
        kset(k, 1, 0.0, 0.0);
        kset(k, 2, -st * sp, st * cp);
        kset(k, 3, -ct * spl, ct * cpl);
    } else {
        kset(k, 0, 0.0, 0.0);
        kset(k, 1, st * sl, -st * cl);
        kset(k, 2, 0.0, 0.0);
        kset(k, 3, -ct * spl, ct * cpl);
    }
}

/* multiply_2x2_by_phase, gate_kernel_templates.h:611-619 */
static void phase2x2(double* k, double sg, double cg) {
    for (int i = 0; i < 4; ++i) {
        const double re = k[2 * i], im = k[2 * i + 1];
        k[2 * i] = re * cg - im * sg;
        k[2 * i + 1] = re * sg + im * cg;
    }
}

static void zero4x4(double* k) { memset(k, 0, sizeof(double) * 32); }

int sqo_gate_param_count(int type) {
    switch (type) {
        case SQGPU_U3: return 3;
        case SQGPU_CU: return 4;
        case SQGPU_U2: case SQGPU_R: case SQGPU_CR: case SQGPU_CROT: return 2;
        case SQGPU_RX: case SQGPU_RY: case SQGPU_RZ: case SQGPU_U1: case SQGPU_CRY: case SQGPU_CRX: case SQGPU_CRZ:
        case SQGPU_CP: case SQGPU_ADAPTIVE: case SQGPU_RXX: case SQGPU_RYY: case SQGPU_RZZ: return 1;
        case SQGPU_GENERAL: case SQGPU_CZ: case SQGPU_CNOT: case SQGPU_CH: case SQGPU_X: case SQGPU_Y: case SQGPU_Z:
        case SQGPU_H: case SQGPU_S: case SQGPU_SDG: case SQGPU_T: case SQGPU_TDG: case SQGPU_SX: case SQGPU_SXDG:
        case SQGPU_SYC: case SQGPU_CCX: case SQGPU_SWAP: case SQGPU_CSWAP: return 0;
        default: return -1;
    }
}

/* Parameters reach the kernels as they are stored: U3-family "theta" slots hold theta/2
 * (get_parameter_multipliers, U3.cpp:49-51, RY.cpp:23-25) and sincos is taken of the stored value
 * (Gate::precompute_sincos, Gate.cpp:1430-1446; glibc sincos, common/include/qgd_math.h:117-127).
 * Adaptive passes its parameter through activation_function == identity (common/common.cpp:35-38). */
int sqo_gate_kernel(int type, const double* p, double* k) {
    double s0 = 0, c0 = 1, s1 = 0, c1 = 1, s2 = 0, c2 = 1, s3 = 0, c3 = 1;
    const int np = sqo_gate_param_count(type);
    if (np < 0) return -1;
    if (np > 0) sincos(p[0], &s0, &c0);
    if (np > 1) sincos(p[1], &s1, &c1);
    if (np > 2) sincos(p[2], &s2, &c2);
    if (np > 3) sincos(p[3], &s3, &c3);
    const double sq = M_SQRT1_2;
    switch (type) {
        case SQGPU_U3: u3_from_trig(k, s0, c0, s1, c1, s2, c2); return 2;                   /* U3.cpp:89-93 */
        case SQGPU_CU: u3_from_trig(k, s0, c0, s1, c1, s2, c2); phase2x2(k, s3, c3); return 2; /* :621-625 */
        case SQGPU_RX: case SQGPU_CRX:                                                       /* :192-199 */
            kset(k, 0, c0, 0); kset(k, 1, 0, -s0); kset(k, 2, 0, -s0); kset(k, 3, c0, 0); return 2;
        case SQGPU_RY: case SQGPU_CRY: case SQGPU_ADAPTIVE:                                   /* :236-243 */
            kset(k, 0, c0, 0); kset(k, 1, -s0, 0); kset(k, 2, s0, 0); kset(k, 3, c0, 0); return 2;
        case SQGPU_RZ: case SQGPU_CRZ:                                                       /* :280-287 */
            kset(k, 0, c0, -s0); kset(k, 1, 0, 0); kset(k, 2, 0, 0); kset(k, 3, c0, s0); return 2;
        case SQGPU_U1: case SQGPU_CP:                                                        /* :388-395 */
            kset(k, 0, 1, 0); kset(k, 1, 0, 0); kset(k, 2, 0, 0); kset(k, 3, c0, s0); return 2;
        case SQGPU_U2: {                                                                     /* :432-442 */
            const double spl = s0 * c1 + c0 * s1, cpl = c0 * c1 - s0 * s1;
            kset(k, 0, sq, 0); kset(k, 1, -sq * c1, -sq * s1); kset(k, 2, sq * c0, sq * s0);
            kset(k, 3, sq * cpl, sq * spl); return 2;
        }
        case SQGPU_R: case SQGPU_CR:                                                         /* :324-335 */
            kset(k, 0, c0, 0); kset(k, 1, -s0 * s1, -s0 * c1); kset(k, 2, s0 * s1, -s0 * c1); kset(k, 3, c0, 0);
            return 2;
        case SQGPU_X: case SQGPU_CNOT: case SQGPU_CCX:                                        /* :42-49 */
            kset(k, 0, 0, 0); kset(k, 1, 1, 0); kset(k, 2, 1, 0); kset(k, 3, 0, 0); return 2;
        case SQGPU_Y: kset(k, 0, 0, 0); kset(k, 1, 0, -1); kset(k, 2, 0, 1); kset(k, 3, 0, 0); return 2; /* :58-65 */
        case SQGPU_Z: case SQGPU_CZ:                                                         /* :74-81 */
            kset(k, 0, 1, 0); kset(k, 1, 0, 0); kset(k, 2, 0, 0); kset(k, 3, -1, 0); return 2;
        case SQGPU_H: case SQGPU_CH:                                                         /* :25-33 */
            kset(k, 0, sq, 0); kset(k, 1, sq, 0); kset(k, 2, sq, 0); kset(k, 3, -sq, 0); return 2;
        case SQGPU_S: kset(k, 0, 1, 0); kset(k, 1, 0, 0); kset(k, 2, 0, 0); kset(k, 3, 0, 1); return 2;   /* :90-97 */
        case SQGPU_SDG: kset(k, 0, 1, 0); kset(k, 1, 0, 0); kset(k, 2, 0, 0); kset(k, 3, 0, -1); return 2;
        case SQGPU_T: kset(k, 0, 1, 0); kset(k, 1, 0, 0); kset(k, 2, 0, 0); kset(k, 3, sq, sq); return 2; /* :122-130 */
        case SQGPU_TDG: kset(k, 0, 1, 0); kset(k, 1, 0, 0); kset(k, 2, 0, 0); kset(k, 3, sq, -sq); return 2;
        case SQGPU_SX:                                                                       /* :156-163 */
            kset(k, 0, .5, .5); kset(k, 1, .5, -.5); kset(k, 2, .5, -.5); kset(k, 3, .5, .5); return 2;
        case SQGPU_SXDG:
            kset(k, 0, .5, -.5); kset(k, 1, .5, .5); kset(k, 2, .5, .5); kset(k, 3, .5, -.5); return 2;
        case SQGPU_RXX:                                                                      /* :669-678 */
            zero4x4(k); k[0] = c0; k[2 * 3 + 1] = -s0; k[2 * 5] = c0; k[2 * 6 + 1] = -s0; k[2 * 10] = c0;
            k[2 * 9 + 1] = -s0; k[2 * 15] = c0; k[2 * 12 + 1] = -s0; return 4;
        case SQGPU_RYY:                                                                      /* :704-713 */
            zero4x4(k); k[0] = c0; k[2 * 3 + 1] = s0; k[2 * 5] = c0; k[2 * 6 + 1] = -s0; k[2 * 10] = c0;
            k[2 * 9 + 1] = -s0; k[2 * 15] = c0; k[2 * 12 + 1] = s0; return 4;
        case SQGPU_RZZ:                                                                      /* :739-748 */
            zero4x4(k); k[0] = c0; k[1] = -s0; k[2 * 5] = c0; k[2 * 5 + 1] = s0; k[2 * 10] = c0; k[2 * 10 + 1] = s0;
            k[2 * 15] = c0; k[2 * 15 + 1] = -s0; return 4;
        case SQGPU_SWAP: case SQGPU_CSWAP: /* permutation |q1 q0> -> |q0 q1>, kernels/apply_dedicated_gate_kernel_to_input.cpp (SWAP) */
            zero4x4(k); k[0] = 1; k[2 * 6] = 1; k[2 * 9] = 1; k[2 * 15] = 1; return 4;
        default: return -1;
    }
}

int sqo_gate_derivative_kernel(int type, const double* p, int pidx, double* k) {
    double s0 = 0, c0 = 1, s1 = 0, c1 = 1, s2 = 0, c2 = 1, s3 = 0, c3 = 1;
    const int np = sqo_gate_param_count(type);
    if (np <= 0 || pidx < 0 || pidx >= np) return -1;
    if (np > 0) sincos(p[0], &s0, &c0);
    if (np > 1) sincos(p[1], &s1, &c1);
    if (np > 2) sincos(p[2], &s2, &c2);
    if (np > 3) sincos(p[3], &s3, &c3);
    const double sq = M_SQRT1_2;
    switch (type) {
        case SQGPU_U3: u3_deriv_from_trig(k, pidx, s0, c0, s1, c1, s2, c2); return 2;         /* U3.cpp:113-129 */
        case SQGPU_CU:                                                                       /* CU.cpp:144-180 */
            if (pidx < 3) {
                u3_deriv_from_trig(k, pidx, s0, c0, s1, c1, s2, c2);
                phase2x2(k, s3, c3);
            } else { /* cu_derivative_kernel_gamma_from_trig_to, :647-656: multiply by i */
                u3_from_trig(k, s0, c0, s1, c1, s2, c2);
                phase2x2(k, s3, c3);
                for (int i = 0; i < 4; ++i) { const double re = k[2 * i], im = k[2 * i + 1]; k[2 * i] = -im; k[2 * i + 1] = re; }
            }
            return 2;
        case SQGPU_RX: case SQGPU_CRX:                                                       /* :220-227 */
            kset(k, 0, -s0, 0); kset(k, 1, 0, -c0); kset(k, 2, 0, -c0); kset(k, 3, -s0, 0); return 2;
        case SQGPU_RY: case SQGPU_CRY: case SQGPU_ADAPTIVE:                                   /* :264-271 */
            kset(k, 0, -s0, 0); kset(k, 1, -c0, 0); kset(k, 2, c0, 0); kset(k, 3, -s0, 0); return 2;
        case SQGPU_RZ: case SQGPU_CRZ:                                                       /* :308-315 */
            kset(k, 0, -s0, -c0); kset(k, 1, 0, 0); kset(k, 2, 0, 0); kset(k, 3, -s0, c0); return 2;
        case SQGPU_U1: case SQGPU_CP:                                                        /* :416-423 */
            kset(k, 0, 0, 0); kset(k, 1, 0, 0); kset(k, 2, 0, 0); kset(k, 3, -s0, c0); return 2;
        case SQGPU_U2: {                                                                     /* :470-499 */
            const double spl = s0 * c1 + c0 * s1, cpl = c0 * c1 - s0 * s1;
            if (pidx == 0) {
                kset(k, 0, 0, 0); kset(k, 1, 0, 0); kset(k, 2, -sq * s0, sq * c0); kset(k, 3, -sq * spl, sq * cpl);
            } else {
                kset(k, 0, 0, 0); kset(k, 1, sq * s1, -sq * c1); kset(k, 2, 0, 0); kset(k, 3, -sq * spl, sq * cpl);
            }
            return 2;
        }
        case SQGPU_R: case SQGPU_CR:                                                         /* :356-379 */
            if (pidx == 0) {
                kset(k, 0, -s0, 0); kset(k, 1, -c0 * s1, -c0 * c1); kset(k, 2, c0 * s1, -c0 * c1); kset(k, 3, -s0, 0);
            } else {
                kset(k, 0, 0, 0); kset(k, 1, -s0 * c1, s0 * s1); kset(k, 2, s0 * c1, s0 * s1); kset(k, 3, 0, 0);
            }
            return 2;
        case SQGPU_RXX:                                                                      /* :687-695 */
            zero4x4(k); k[0] = -s0; k[2 * 3 + 1] = -c0; k[2 * 5] = -s0; k[2 * 6 + 1] = -c0; k[2 * 10] = -s0;
            k[2 * 9 + 1] = -c0; k[2 * 15] = -s0; k[2 * 12 + 1] = -c0; return 4;
        case SQGPU_RYY:                                                                      /* :722-730 */
            zero4x4(k); k[0] = -s0; k[2 * 3 + 1] = c0; k[2 * 5] = -s0; k[2 * 6 + 1] = -c0; k[2 * 10] = -s0;
            k[2 * 9 + 1] = -c0; k[2 * 15] = -s0; k[2 * 12 + 1] = c0; return 4;
        case SQGPU_RZZ:                                                                      /* :757-765 */
            zero4x4(k); k[0] = -s0; k[1] = -c0; k[2 * 5] = -s0; k[2 * 5 + 1] = c0; k[2 * 10] = -s0;
            k[2 * 10 + 1] = c0; k[2 * 15] = -s0; k[2 * 15 + 1] = -c0; return 4;
        default: return -1;
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* gate application                                                                                              */
/* ------------------------------------------------------------------------------------------------------------ */

/* kernels/apply_kernel_to_input.cpp:33-115 (the scalar reference kernel); CCX adds a second control bit
 * (apply_X_kernel_to_input, kernels/apply_dedicated_gate_kernel_to_input.cpp:45-92: all control bits must be set) */
void sqo_apply_kernel_to_input(const double* kk, double* input, int rows, int cols, int stride, int deriv, int target,
                               int control, int control2) {
    const cplx* k = (const cplx*)kk;
    cplx* in = (cplx*)input;
    const int step = 1 << target;
    for (int base = 0; base < rows; base += (step << 1)) {
        for (int idx = 0; idx < step; ++idx) {
            const int r0 = base + idx, r1 = r0 + step;
            int active = 1;
            if (control >= 0 && !((r0 >> control) & 1)) active = 0;
            if (control2 >= 0 && !((r0 >> control2) & 1)) active = 0;
            cplx* row0 = in + (size_t)r0 * stride;
            cplx* row1 = in + (size_t)r1 * stride;
            if (active) {
                for (int c = 0; c < cols; ++c) {
                    const cplx e0 = row0[c], e1 = row1[c];
                    cplx t1 = cmul(k[0], e0), t2 = cmul(k[1], e1);
                    row0[c].re = t1.re + t2.re;
                    row0[c].im = t1.im + t2.im;
                    t1 = cmul(k[2], e0);
                    t2 = cmul(k[3], e1);
                    row1[c].re = t1.re + t2.re;
                    row1[c].im = t1.im + t2.im;
                }
            } else if (deriv) {
                memset(row0, 0, sizeof(cplx) * cols);
                memset(row1, 0, sizeof(cplx) * cols);
            }
        }
    }
}

/* kernels/apply_large_kernel_to_input.cpp:123-213: local index bit j <-> qubits[j] (ascending). `control` >= 0 adds
 * a control bit (CSWAP); deriv zero-fills inactive groups like the 1-qubit kernel. */
void sqo_apply_large_kernel_to_input(const double* kernel, double* input, int rows, int cols, int stride,
                                     const int* qubits, int k, int control, int deriv) {
    const cplx* K = (const cplx*)kernel;
    cplx* in = (cplx*)input;
    const int dim = 1 << k;
    int pattern[32];
    int mask = 0;
    for (int j = 0; j < k; ++j) mask |= 1 << qubits[j];
    for (int l = 0; l < dim; ++l) {
        int idx = 0;
        for (int b = 0; b < k; ++b) if (l & (1 << b)) idx |= 1 << qubits[b];
        pattern[l] = idx;
    }
    cplx src[32], out[32];
    for (int base = 0; base < rows; ++base) {
        if (base & mask) continue;
        const int active = (control < 0) || ((base >> control) & 1);
        for (int c = 0; c < cols; ++c) {
            if (!active) {
                if (deriv) for (int l = 0; l < dim; ++l) { in[(size_t)(base | pattern[l]) * stride + c].re = 0; in[(size_t)(base | pattern[l]) * stride + c].im = 0; }
                continue;
            }
            for (int l = 0; l < dim; ++l) src[l] = in[(size_t)(base | pattern[l]) * stride + c];
            for (int o = 0; o < dim; ++o) {
                cplx acc = {0.0, 0.0};
                for (int i = 0; i < dim; ++i) {
                    const cplx ke = K[o * dim + i], se = src[i];
                    acc.re += ke.re * se.re - ke.im * se.im;
                    acc.im += ke.re * se.im + ke.im * se.re;
                }
                out[o] = acc;
            }
            for (int l = 0; l < dim; ++l) in[(size_t)(base | pattern[l]) * stride + c] = out[l];
        }
    }
}

/* SYC (Sycamore fSim(pi/2, pi/6)), kernels/apply_dedicated_gate_kernel_to_input.cpp:582-640: rows with (target, control)
 * bits (1,0) and (0,1) are exchanged and multiplied by -i, rows with both bits set by exp(-i pi/6); symmetric in the two
 * qubits. */
static void sqo_apply_syc(double* input, int rows, int cols, int stride, int target, int control) {
    cplx* in = (cplx*)input;
    const int tb = 1 << target, cb = 1 << control;
    const double pr = sqrt(3.0) / 2.0, pi_ = -0.5;
    for (int r = 0; r < rows; ++r) {
        if ((r & tb) || (r & cb)) continue; /* r = the (0,0) row of its group */
        cplx* r01 = in + (size_t)(r | tb) * stride;
        cplx* r10 = in + (size_t)(r | cb) * stride;
        cplx* r11 = in + (size_t)(r | tb | cb) * stride;
        for (int c = 0; c < cols; ++c) {
            const cplx e01 = r01[c], e10 = r10[c], e11 = r11[c];
            r01[c].re = e10.im;  r01[c].im = -e10.re;
            r10[c].re = e01.im;  r10[c].im = -e01.re;
            r11[c].re = pr * e11.re - pi_ * e11.im;
            r11[c].im = pr * e11.im + pi_ * e11.re;
        }
    }
}

/* CROT kernels, gates/include/gate_kernel_templates.h:779-877: the control = 0 branch is U3(theta, phi - pi/2, -phi + pi/2),
 * the control = 1 branch the same with -theta (Gate.cpp:1568-1580 hands the "inverse" kernel to the control = 1 rows of
 * apply_crot_kernel_to_matrix_input, kernels/apply_large_kernel_to_input.cpp:436-505). which: -1 forward, 0 d/dtheta
 * (theta -> theta + pi/2 on both branches), 1 d/dphi (U3(+-theta, phi, -phi) with the diagonal zeroed). */
static void crot_kernels(const double* p, int which, double* k0, double* k1) {
    double st, ct, sp, cp;
    sincos(p[0], &st, &ct);
    sincos(p[1], &sp, &cp);
    if (which == 0) { const double t = st; st = ct; ct = -t; }
    if (which <= 0) {
        u3_from_trig(k0, st, ct, -cp, sp, cp, sp);
        u3_from_trig(k1, -st, ct, -cp, sp, cp, sp);
    } else {
        u3_from_trig(k0, st, ct, sp, cp, -sp, cp);
        u3_from_trig(k1, -st, ct, sp, cp, -sp, cp);
        kset(k0, 0, 0, 0); kset(k0, 3, 0, 0); kset(k1, 0, 0, 0); kset(k1, 3, 0, 0);
    }
}

static void sqo_apply_crot(const double* k0, const double* k1, double* input, int rows, int cols, int stride, int target, int control) {
    cplx* in = (cplx*)input;
    const int step = 1 << target;
    for (int base = 0; base < rows; base += (step << 1))
        for (int idx = 0; idx < step; ++idx) {
            const int r0 = base + idx, r1 = r0 + step;
            const cplx* k = (const cplx*)(((r0 >> control) & 1) ? k1 : k0);
            cplx* row0 = in + (size_t)r0 * stride;
            cplx* row1 = in + (size_t)r1 * stride;
            for (int c = 0; c < cols; ++c) {
                const cplx e0 = row0[c], e1 = row1[c];
                cplx t1 = cmul(k[0], e0), t2 = cmul(k[1], e1);
                row0[c].re = t1.re + t2.re;
                row0[c].im = t1.im + t2.im;
                t1 = cmul(k[2], e0);
                t2 = cmul(k[3], e1);
                row1[c].re = t1.re + t2.re;
                row1[c].im = t1.im + t2.im;
            }
        }
}

/* Gate::apply_to / apply_to_inner -> gate_kernel_to -> apply_kernel_to (Gate.cpp:432-570, 1477-1768) and
 * Gate::apply_derivative_to_precomputed (Gate.cpp:644-706; deriv == true). */
int sqo_apply_gate(const sqgpu_gate_desc* g, const double* params, const double* pool, int deriv_param, double* input,
                   int rows, int cols, int stride) {
    double k[32];
    const double* gp = params ? params + g->param_start : NULL;
    const int deriv = deriv_param >= 0;
    if (g->type == SQGPU_GENERAL) {
        if (deriv) return -1;
        if (g->n_qubits == 1) {
            sqo_apply_kernel_to_input(pool + 2 * g->matrix_off, input, rows, cols, stride, 0, g->qubits[0], -1, -1);
        } else {
            sqo_apply_large_kernel_to_input(pool + 2 * g->matrix_off, input, rows, cols, stride, g->qubits,
                                            g->n_qubits, -1, 0);
        }
        return 0;
    }
    if (g->type == SQGPU_SYC) {
        if (deriv) return -1;
        sqo_apply_syc(input, rows, cols, stride, g->target, g->control);
        return 0;
    }
    if (g->type == SQGPU_CROT) {
        double k1[8];
        crot_kernels(gp, deriv ? deriv_param : -1, k, k1);
        sqo_apply_crot(k, k1, input, rows, cols, stride, g->target, g->control);
        return 0;
    }
    const int dim = deriv ? sqo_gate_derivative_kernel(g->type, gp, deriv_param, k) : sqo_gate_kernel(g->type, gp, k);
    if (dim == 2) {
        sqo_apply_kernel_to_input(k, input, rows, cols, stride, deriv, g->target, g->control, g->control2);
        return 0;
    }
    if (dim == 4) {
        int q[2];
        q[0] = g->target < g->target2 ? g->target : g->target2;
        q[1] = g->target < g->target2 ? g->target2 : g->target;
        sqo_apply_large_kernel_to_input(k, input, rows, cols, stride, q, 2, g->control, deriv);
        return 0;
    }
    return -1;
}

/* Gates_block::apply_to_inner forward loop, Gates_block.cpp:683-708 */
int sqo_apply_circuit(const sqgpu_gate_desc* gates, int n_gates, const double* params, const double* pool,
                      double* input, int rows, int cols, int stride) {
    for (int i = 0; i < n_gates; ++i) {
        const int rc = sqo_apply_gate(&gates[i], params, pool, -1, input, rows, cols, stride);
        if (rc) return rc;
    }
    return 0;
}

/* Gates_block::apply_derivate_to, prefix-only route (Gates_block.cpp:430-470, 1060-1133): for every gate, the prefix
 * state, the gate's derivative kernels, then the remaining gates one by one. */
int sqo_apply_derivate(const sqgpu_gate_desc* gates, int n_gates, int n_params, const double* params,
                       const double* pool, const double* input, int rows, int cols, int stride, double* out) {
    const size_t msz = (size_t)rows * cols * 2;
    double* prefix = (double*)malloc(sizeof(double) * msz);
    if (!prefix) return -1;
    for (int r = 0; r < rows; ++r) memcpy(prefix + 2 * (size_t)r * cols, input + 2 * (size_t)r * stride, sizeof(double) * 2 * cols);
    memset(out, 0, sizeof(double) * msz * n_params);
    for (int gi = 0; gi < n_gates; ++gi) {
        const sqgpu_gate_desc* g = &gates[gi];
        for (int p = 0; p < g->n_params; ++p) {
            double* d = out + msz * (size_t)(g->param_start + p);
            memcpy(d, prefix, sizeof(double) * msz);
            int rc = sqo_apply_gate(g, params, pool, p, d, rows, cols, cols);
            for (int gj = gi + 1; gj < n_gates && !rc; ++gj) rc = sqo_apply_gate(&gates[gj], params, pool, -1, d, rows, cols, cols);
            if (rc) { free(prefix); return rc; }
        }
        const int rc = sqo_apply_gate(g, params, pool, -1, prefix, rows, cols, cols);
        if (rc) { free(prefix); return rc; }
    }
    free(prefix);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* cost functions: decomposition/N_Qubit_Decomposition_Cost_Function.cpp                                         */
/* ------------------------------------------------------------------------------------------------------------ */

/* get_cost_function :73-162 (t=0 real part, offset), get_cost_function_with_correction{,2} :191-404 (t=1,2),
 * get_trace :482-499, get_trace_with_correction{,2} :564-664 (complex; the reference ignores the offset there --
 * callers pass trace_offset = 0 for those variants). */
void sqo_traces(const double* mtx, int rows, int cols, int stride, int qbit_num, int trace_offset, double* out) {
    const cplx* m = (const cplx*)mtx;
    (void)rows;
    double re = 0, im = 0;
    for (int j = 0; j < cols; ++j) { re += m[(size_t)(j + trace_offset) * stride + j].re; im += m[(size_t)(j + trace_offset) * stride + j].im; }
    out[0] = re; out[1] = im;
    re = 0; im = 0;
    for (int q = 0; q < qbit_num; ++q) {
        const int mask = 1 << q;
        for (int j = 0; j < cols; ++j) { const int r = (j + trace_offset) ^ mask; re += m[(size_t)r * stride + j].re; im += m[(size_t)r * stride + j].im; }
    }
    out[2] = re; out[3] = im;
    re = 0; im = 0;
    for (int q = 0; q < qbit_num - 1; ++q)
        for (int q2 = q + 1; q2 < qbit_num; ++q2) {
            const int mask = (1 << q) + (1 << q2);
            for (int j = 0; j < cols; ++j) { const int r = (j + trace_offset) ^ mask; re += m[(size_t)r * stride + j].re; im += m[(size_t)r * stride + j].im; }
        }
    out[4] = re; out[5] = im;
}

/* Optimization_Interface::calculate_cost_function, decomposition/Optimization_Interface.cpp:677-735 */
double sqo_cost_from_traces(int variant, const double* t, int cols, double prev, double c1, double c2) {
    const double n = (double)cols;
    switch (variant) {
        case SQGPU_FROBENIUS_NORM: return 1.0 - t[0] / n;
        case SQGPU_FROBENIUS_NORM_CORRECTION1: return (1.0 - t[0] / n) - sqrt(prev) * (t[2] / n) * c1;
        case SQGPU_FROBENIUS_NORM_CORRECTION2: return (1.0 - t[0] / n) - sqrt(prev) * ((t[2] / n) * c1 + (t[4] / n) * c2);
        case SQGPU_HILBERT_SCHMIDT_TEST: { const double d = 1.0 / n; return 1.0 - d * d * (t[0] * t[0] + t[1] * t[1]); }
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1: {
            const double d = 1.0 / n;
            return 1 - d * d * (t[0] * t[0] + t[1] * t[1] + sqrt(prev) * c1 * (t[2] * t[2] + t[3] * t[3]));
        }
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2: {
            const double d = 1.0 / n;
            return 1 - d * d * (t[0] * t[0] + t[1] * t[1] + sqrt(prev) * (c1 * (t[2] * t[2] + t[3] * t[3]) + c2 * (t[4] * t[4] + t[5] * t[5])));
        }
        case SQGPU_INFIDELITY: return 1.0 - ((t[0] * t[0] + t[1] * t[1]) / n + 1) / (n + 1);
        default: return NAN;
    }
}

/* gradient component formulas, decomposition/Optimization_Interface.cpp:1397-1458 */
double sqo_grad_from_traces(int variant, const double* t, const double* dt, int cols, double prev, double c1, double c2) {
    const double n = (double)cols;
    switch (variant) {
        case SQGPU_FROBENIUS_NORM: return (1.0 - dt[0] / n) - 1.0;
        case SQGPU_FROBENIUS_NORM_CORRECTION1: return (1.0 - dt[0] / n) - sqrt(prev) * (dt[2] / n) * c1 - 1.0;
        case SQGPU_FROBENIUS_NORM_CORRECTION2:
            return (1.0 - dt[0] / n) - sqrt(prev) * ((dt[2] / n) * c1 + (dt[4] / n) * c2) - 1.0;
        case SQGPU_HILBERT_SCHMIDT_TEST: {
            const double d = 1.0 / n;
            return -2.0 * d * d * t[0] * dt[0] - 2.0 * d * d * t[1] * dt[1];
        }
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION1: {
            const double d = 1.0 / n;
            return -2.0 * d * d * (t[0] * dt[0] + t[1] * dt[1] + sqrt(prev) * c1 * (t[2] * dt[2] + t[3] * dt[3]));
        }
        case SQGPU_HILBERT_SCHMIDT_TEST_CORRECTION2: {
            const double d = 1.0 / n;
            return -2.0 * d * d * (t[0] * dt[0] + t[1] * dt[1] + sqrt(prev) * (c1 * (t[2] * dt[2] + t[3] * dt[3]) + c2 * (t[4] * dt[4] + t[5] * dt[5])));
        }
        case SQGPU_INFIDELITY: return -2.0 / n / (n + 1) * t[0] * dt[0] - 2.0 / n / (n + 1) * t[1] * dt[1];
        default: return NAN;
    }
}

/* get_cost_function_sum_of_squares, N_Qubit_Decomposition_Cost_Function.cpp:443-458 */
double sqo_cost_sum_of_squares(const double* mtx, int rows, int cols, int stride) {
    const cplx* m = (const cplx*)mtx;
    double ret = 0.0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const cplx e = m[(size_t)r * stride + c];
            if (r == c) ret += (e.re - 1.0) * (e.re - 1.0) + e.im * e.im;
            else ret += e.re * e.re + e.im * e.im;
        }
    return ret;
}

static double* copy_compact(const double* src, int rows, int cols, int stride) {
    double* m = (double*)malloc(sizeof(double) * 2 * (size_t)rows * cols);
    if (!m) return NULL;
    for (int r = 0; r < rows; ++r) memcpy(m + 2 * (size_t)r * cols, src + 2 * (size_t)r * stride, sizeof(double) * 2 * cols);
    return m;
}

/* Optimization_Interface::optimization_problem, decomposition/Optimization_Interface.cpp:634-668 */
int sqo_cost(const sqgpu_gate_desc* gates, int n_gates, const double* params, const double* pool, const double* umtx,
             int rows, int cols, int stride, int qbit_num, int variant, int trace_offset, double prev, double c1,
             double c2, double* cost) {
    double* m = copy_compact(umtx, rows, cols, stride);
    if (!m) return -1;
    int rc = sqo_apply_circuit(gates, n_gates, params, pool, m, rows, cols, cols);
    if (!rc) {
        if (variant == SQGPU_SUM_OF_SQUARES) {
            *cost = sqo_cost_sum_of_squares(m, rows, cols, cols);
        } else {
            double t[6];
            const int frob = variant <= SQGPU_FROBENIUS_NORM_CORRECTION2;
            sqo_traces(m, rows, cols, cols, qbit_num, frob ? trace_offset : 0, t);
            *cost = sqo_cost_from_traces(variant, t, cols, prev, c1, c2);
        }
    }
    free(m);
    return rc;
}

/* Optimization_Interface::optimization_problem_combined_non_static, decomposition/Optimization_Interface.cpp:1145-1490
 * (apply_to_combined -> f0 from element 0, grad[i] from the traces of derivative matrix i). */
int sqo_cost_grad(const sqgpu_gate_desc* gates, int n_gates, int n_params, const double* params, const double* pool,
                  const double* umtx, int rows, int cols, int stride, int qbit_num, int variant, int trace_offset,
                  double prev, double c1, double c2, double* cost, double* grad) {
    const size_t msz = (size_t)rows * cols * 2;
    double* m = copy_compact(umtx, rows, cols, stride);
    double* d = (double*)malloc(sizeof(double) * msz * (size_t)(n_params > 0 ? n_params : 1));
    if (!m || !d) { free(m); free(d); return -1; }
    int rc = sqo_apply_derivate(gates, n_gates, n_params, params, pool, m, rows, cols, cols, d);
    if (!rc) rc = sqo_apply_circuit(gates, n_gates, params, pool, m, rows, cols, cols);
    if (!rc) {
        const int frob = variant <= SQGPU_FROBENIUS_NORM_CORRECTION2;
        const int off = frob ? trace_offset : 0;
        if (variant == SQGPU_SUM_OF_SQUARES) {
            /* get_deriv_sum_of_squares :459-475 and real_trace_conj_dot :1451 */
            *cost = sqo_cost_sum_of_squares(m, rows, cols, cols);
            for (int p = 0; p < n_params; ++p) {
                const double* dm = d + msz * (size_t)p;
                double acc = 0.0;
                for (int r = 0; r < rows; ++r)
                    for (int c = 0; c < cols; ++c) {
                        const size_t o = 2 * ((size_t)r * cols + c);
                        const double ur = 2 * (m[o] - (r == c ? 1.0 : 0.0)), ui = 2 * m[o + 1];
                        acc += ur * dm[o] + ui * dm[o + 1];
                    }
                grad[p] = acc;
            }
        } else {
            double t[6], dt[6];
            sqo_traces(m, rows, cols, cols, qbit_num, off, t);
            *cost = sqo_cost_from_traces(variant, t, cols, prev, c1, c2);
            for (int p = 0; p < n_params; ++p) {
                sqo_traces(d + msz * (size_t)p, rows, cols, cols, qbit_num, off, dt);
                grad[p] = sqo_grad_from_traces(variant, t, dt, cols, prev, c1, c2);
            }
        }
    }
    free(m);
    free(d);
    return rc;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* VQE state-vector path: variational_quantum_eigensolver/Variational_Quantum_Eigensolver_Base.cpp               */
/* ------------------------------------------------------------------------------------------------------------ */

/* mult(Matrix_sparse, Matrix&), common/common.cpp:403-436 */
void sqo_csr_matvec(int n_rows, const int32_t* indptr, const int32_t* indices, const double* values, const double* x,
                    double* y) {
    const cplx* v = (const cplx*)values;
    const cplx* xx = (const cplx*)x;
    cplx* yy = (cplx*)y;
    for (int r = 0; r < n_rows; ++r) {
        cplx acc = {0.0, 0.0};
        for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
            const cplx t = cmul(v[e], xx[indices[e]]);
            acc.re += t.re;
            acc.im += t.im;
        }
        yy[r] = acc;
    }
}

/* Expectation_value_of_energy_real :584-624 : sum_i Re(conj(left_i) * (H right)_i) */
static double expectation(int n, const double* left, const double* hright) {
    double e = 0.0;
    for (int i = 0; i < n; ++i) e += left[2 * i] * hright[2 * i] + left[2 * i + 1] * hright[2 * i + 1];
    return e;
}

int sqo_vqe_energy(const sqgpu_gate_desc* gates, int n_gates, const double* params, const double* pool,
                   const double* state0, int n_rows, const int32_t* indptr, const int32_t* indices,
                   const double* values, double* energy) {
    double* psi = copy_compact(state0, n_rows, 1, 1);
    double* hpsi = (double*)malloc(sizeof(double) * 2 * (size_t)n_rows);
    if (!psi || !hpsi) { free(psi); free(hpsi); return -1; }
    int rc = sqo_apply_circuit(gates, n_gates, params, pool, psi, n_rows, 1, 1);
    if (!rc) {
        sqo_csr_matvec(n_rows, indptr, indices, values, psi, hpsi);
        *energy = expectation(n_rows, psi, hpsi);
    }
    free(psi);
    free(hpsi);
    return rc;
}

/* optimization_problem_combined_non_static :1131-1199 */
int sqo_vqe_energy_grad(const sqgpu_gate_desc* gates, int n_gates, int n_params, const double* params,
                        const double* pool, const double* state0, int n_rows, const int32_t* indptr,
                        const int32_t* indices, const double* values, double* energy, double* grad) {
    double* psi = copy_compact(state0, n_rows, 1, 1);
    double* hpsi = (double*)malloc(sizeof(double) * 2 * (size_t)n_rows);
    double* d = (double*)malloc(sizeof(double) * 2 * (size_t)n_rows * (size_t)(n_params > 0 ? n_params : 1));
    if (!psi || !hpsi || !d) { free(psi); free(hpsi); free(d); return -1; }
    int rc = sqo_apply_derivate(gates, n_gates, n_params, params, pool, psi, n_rows, 1, 1, d);
    if (!rc) rc = sqo_apply_circuit(gates, n_gates, params, pool, psi, n_rows, 1, 1);
    if (!rc) {
        sqo_csr_matvec(n_rows, indptr, indices, values, psi, hpsi);
        *energy = expectation(n_rows, psi, hpsi);
        for (int p = 0; p < n_params; ++p) grad[p] = 2 * expectation(n_rows, d + 2 * (size_t)n_rows * p, hpsi);
    }
    free(psi);
    free(hpsi);
    free(d);
    return rc;
}

/* The same gradient restricted to a sample of parameters: the derivative state of parameter p is the prefix state, the
 * derivative kernel of p's gate, then the remaining gates (the prefix-only route of sqo_apply_derivate), evaluated only for
 * p in sample[0..n_sample). Used at sizes where all P derivative states are out of reach (n = 20: P = 1140 passes over
 * 2^20 amplitudes). grad[i] belongs to sample[i]. */
int sqo_vqe_energy_grad_sampled(const sqgpu_gate_desc* gates, int n_gates, const double* params, const double* pool,
                                const double* state0, int n_rows, const int32_t* indptr, const int32_t* indices,
                                const double* values, const int32_t* sample, int n_sample, double* energy, double* grad) {
    const size_t vsz = 2 * (size_t)n_rows;
    double* prefix = copy_compact(state0, n_rows, 1, 1);
    double* hpsi = (double*)malloc(sizeof(double) * vsz);
    double* d = (double*)malloc(sizeof(double) * vsz * (size_t)(n_sample > 0 ? n_sample : 1));
    if (!prefix || !hpsi || !d) { free(prefix); free(hpsi); free(d); return -1; }
    int rc = 0;
    for (int gi = 0; gi < n_gates && !rc; ++gi) {
        const sqgpu_gate_desc* g = &gates[gi];
        for (int s = 0; s < n_sample && !rc; ++s) {
            const int p = sample[s] - g->param_start;
            if (p < 0 || p >= g->n_params) continue;
            double* ds = d + vsz * (size_t)s;
            memcpy(ds, prefix, sizeof(double) * vsz);
            rc = sqo_apply_gate(g, params, pool, p, ds, n_rows, 1, 1);
            for (int gj = gi + 1; gj < n_gates && !rc; ++gj) rc = sqo_apply_gate(&gates[gj], params, pool, -1, ds, n_rows, 1, 1);
        }
        if (!rc) rc = sqo_apply_gate(g, params, pool, -1, prefix, n_rows, 1, 1);
    }
    if (!rc) {
        sqo_csr_matvec(n_rows, indptr, indices, values, prefix, hpsi);
        *energy = expectation(n_rows, prefix, hpsi);
        for (int s = 0; s < n_sample; ++s) grad[s] = 2 * expectation(n_rows, d + vsz * (size_t)s, hpsi);
    }
    free(prefix);
    free(hpsi);
    free(d);
    return rc;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* Adam::update, common/Adam.cpp:120-262, in its SEQUENTIAL semantics: the bias-correction products beta1_t / beta2_t */
/* are class members that the reference advances inside the per-parameter loop (Adam.cpp:226-229), i.e. once per     */
/* parameter and update (under TBB that loop is racy; one thread, ascending idx, is the deterministic reading).      */
/* Host mirror of csrc/optim.cuh: adam_update_kernel for the trajectory tests. state = sqo_adam_state.               */
/* ------------------------------------------------------------------------------------------------------------ */
void sqo_adam_reset(sqo_adam_state* s) {
    memset(s, 0, sizeof(*s));
    s->beta1_t = 1.0;
    s->beta2_t = 1.0;
    s->decreasing_test = -1.0;
    s->f0_prev = 1.7976931348623157e308;
    for (int i = 0; i < 20; ++i) s->decreasing_vec[i] = -1;
}

int sqo_adam_update(sqo_adam_state* s, double* params, const double* grad, double* mom, double* var, int n, double f0,
                    double eta, double beta1, double beta2, double epsilon) {
    s->f0_mean = s->f0_mean + (f0 - s->f0_vec[s->f0_idx]) / 100.0;
    s->f0_vec[s->f0_idx] = f0;
    s->f0_idx = (s->f0_idx + 1) % 100;
    double var_f0 = 0.0;
    for (int i = 0; i < 100; ++i) var_f0 = var_f0 + (s->f0_vec[i] - s->f0_mean) * (s->f0_vec[i] - s->f0_mean);
    var_f0 = sqrt(var_f0) / 100.0;
    if (f0 < s->f0_prev) {
        if (s->decreasing_vec[s->decreasing_idx] != 1) s->decreasing_test = s->decreasing_test + 2.0 / 20.0;
        s->decreasing_vec[s->decreasing_idx] = 1;
    } else {
        if (s->decreasing_vec[s->decreasing_idx] == 1) s->decreasing_test = s->decreasing_test - 2.0 / 20.0;
        s->decreasing_vec[s->decreasing_idx] = -1;
    }
    s->decreasing_idx = (s->decreasing_idx + 1) % 20;
    s->f0_prev = f0;
    double grad_var = 0.0;
    for (int i = 0; i < n; ++i) grad_var += var[i];
    const int barren_plateau = (grad_var < epsilon && s->decreasing_test > 0.7) ? 1 : 0;
    for (int i = 0; i < n; ++i) {
        mom[i] = beta1 * mom[i] + (1 - beta1) * grad[i];
        var[i] = beta2 * var[i] + (1 - beta2) * grad[i] * grad[i];
        s->beta1_t = s->beta1_t * beta1;
        const double mom_bias_corr = mom[i] / (1 - s->beta1_t);
        s->beta2_t = s->beta2_t * beta2;
        const double var_bias_corr = var[i] / (1 - s->beta2_t);
        if (barren_plateau) params[i] = params[i] - eta * mom_bias_corr / (sqrt(var_bias_corr) + epsilon / 100);
        else params[i] = params[i] - eta * mom_bias_corr / (sqrt(var_bias_corr) + epsilon);
    }
    s->iter_t++;
    return (fabs(s->f0_mean - f0) < 1e-6 && s->decreasing_test <= 0.7 && var_f0 / s->f0_mean < 1e-6) ? 1 : 0;
}
