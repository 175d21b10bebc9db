// Kernel-tree programs: wire format (include/agp_b200.h) -> device instruction stream.
//
// The wire format is the reference's postfix `unroll` order (src/GP.jl:111-113).  The device
// interpreter (agp_eval.cuh) keeps its operand stack in registers, so the host re-orders the
// evaluation of each binary node (deeper subtree first, Sethi-Ullman) to bound the stack at
// AGP_MAX_STACK.  Plus and Times commute bit-for-bit in IEEE-754; ChangePoint gets a
// "swapped" opcode so left/right keep their meaning.  Per-node scalars that the reference
// computes once outside its broadcasts (lengthscale^2, pi/period, -2/lengthscale^2:
// src/GP.jl:243, 332, 334) are computed here once in host FP64, exactly as Julia does.
#pragma once
#include <stdint.h>

#define AGP_MAX_STACK 8
#define AGP_GRAD_TAPE_BIG 512  // tape levels of the big variant of the gradient interpreter (agp_eval.cuh)

enum AgpDevOp : int32_t {
    AGP_I_CONST = 0,   // a = value
    AGP_I_LINEAR = 1,  // a = intercept, b = bias, c = amplitude
    AGP_I_SE = 2,      // a = lengthscale^2, b = amplitude, c = 1 / lengthscale^2, d = lengthscale
    AGP_I_GE = 3,      // a = lengthscale, b = gamma, c = amplitude, d = 1 / lengthscale
    AGP_I_PER = 4,     // a = pi/period, b = -2/lengthscale^2, c = amplitude, d = lengthscale, reserved = period
    AGP_I_WN = 5,      // a = value
    AGP_I_PLUS = 6,
    AGP_I_TIMES = 7,
    AGP_I_CP = 8,      // a = location, b = scale, c = 1 / scale; stack: s1 = left, s0 = right
    AGP_I_CP_SWAP = 9, // same, stack: s1 = right, s0 = left
    // flag in AgpInstr::op above the opcode byte: the node's divisor is a normal number of moderate
    // magnitude, so x / divisor may run as reciprocal + exact-remainder corrections (agp_math.cuh)
    AGP_I_FASTDIV = 1 << 8
};

struct __attribute__((aligned(16))) AgpInstr {
    int32_t op;   // AgpDevOp | flags
    int32_t pad;  // index of the node's first parameter in the particle's params[] slice (wire order)
    double a, b, c, d;
    double reserved;  // sizeof == 48 (three 16-byte shared-memory loads); Periodic keeps its period here
};
#define AGP_INSTR_DOUBLES 6  // sizeof(AgpInstr) / 8

#ifdef __cplusplus
#include <string>
#include <vector>

// Compile one program.  Appends to `out`; returns 0 or an AGP_ERR_* code (message in `err`).
// `need` receives the register-stack depth the program needs (1..AGP_MAX_STACK).
int agp_compile_program(const int32_t* ops, const int32_t* param_off, int32_t m, const double* params,
                        int32_t n_params, std::vector<AgpInstr>& out, int* need, std::string& err);
#endif
