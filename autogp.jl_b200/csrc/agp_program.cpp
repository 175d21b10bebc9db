// Host-side compiler: postfix kernel-tree program -> register-stack device program.
// See agp_program.h.  Field orders follow src/GP.jl:131-133,157-159,185-192,228-234,
// 269-277,315-322,466-473; opcodes follow GP.GPConfig (src/GP.jl:1101-1108).
#include "agp_program.h"

#include <algorithm>
#include <cmath>

#include "../../include/agp_b200.h"

namespace {

struct TNode {
    int32_t op;
    int32_t off;
    int left = -1, right = -1;
    int need = 1;
};

int n_node_params(int32_t op) {
    switch (op) {
        case AGP_OP_CONSTANT: return 1;
        case AGP_OP_LINEAR: return 3;
        case AGP_OP_SQUARED_EXPONENTIAL: return 2;
        case AGP_OP_GAMMA_EXPONENTIAL: return 3;
        case AGP_OP_PERIODIC: return 3;
        case AGP_OP_WHITE_NOISE: return 1;
        case AGP_OP_CHANGEPOINT: return 2;
        default: return 0;
    }
}

// x / a may run as x * (1/a) + two exact-remainder corrections (agp_math.cuh: div_const_v) when neither
// a nor 1/a is near the ends of the exponent range; otherwise the device divides generically.
bool fast_divisor(double a) {
    const double m = std::fabs(a);
    return std::isfinite(a) && m > 0x1p-500 && m < 0x1p500;
}

void emit(const std::vector<TNode>& t, int id, const double* params, std::vector<AgpInstr>& out) {
    // explicit stack instead of recursion: trees can be deep (max_depth = -1, src/GP.jl:1127)
    struct Frame { int id; int state; bool swapped; };
    std::vector<Frame> st;
    st.push_back({id, 0, false});
    while (!st.empty()) {
        Frame& f = st.back();
        const TNode& nd = t[f.id];
        const double* p = params + nd.off;
        AgpInstr in{};
        in.pad = nd.off;  // where this node's parameters sit in the caller's array (gradient output order)
        if (nd.left < 0) {
            switch (nd.op) {
                case AGP_OP_CONSTANT: in.op = AGP_I_CONST; in.a = p[0]; break;
                case AGP_OP_LINEAR: in.op = AGP_I_LINEAR; in.a = p[0]; in.b = p[1]; in.c = p[2]; break;
                case AGP_OP_SQUARED_EXPONENTIAL:
                    in.op = AGP_I_SE; in.a = p[0] * p[0]; in.b = p[1];
                    in.c = 1.0 / in.a;
                    in.d = p[0];  // the lengthscale itself (gradient interpreter)
                    if (fast_divisor(in.a)) in.op |= AGP_I_FASTDIV;
                    break;
                case AGP_OP_GAMMA_EXPONENTIAL:
                    in.op = AGP_I_GE; in.a = p[0]; in.b = p[1]; in.c = p[2];
                    in.d = 1.0 / in.a;
                    if (fast_divisor(in.a)) in.op |= AGP_I_FASTDIV;
                    break;
                case AGP_OP_PERIODIC:
                    in.op = AGP_I_PER;
                    in.a = M_PI / p[1];
                    in.b = -2.0 / (p[0] * p[0]);
                    in.c = p[2];
                    in.d = p[0];         // lengthscale and period themselves (gradient interpreter)
                    in.reserved = p[1];
                    break;
                default: in.op = AGP_I_WN; in.a = p[0]; break;
            }
            out.push_back(in);
            st.pop_back();
            continue;
        }
        if (f.state == 0) {
            f.swapped = t[nd.right].need > t[nd.left].need;
            f.state = 1;
            int first = f.swapped ? nd.right : nd.left;
            st.push_back({first, 0, false});
        } else if (f.state == 1) {
            f.state = 2;
            int second = f.swapped ? nd.left : nd.right;
            st.push_back({second, 0, false});
        } else {
            if (nd.op == AGP_OP_PLUS) in.op = AGP_I_PLUS;
            else if (nd.op == AGP_OP_TIMES) in.op = AGP_I_TIMES;
            else {
                in.op = f.swapped ? AGP_I_CP_SWAP : AGP_I_CP; in.a = p[0]; in.b = p[1];
                in.c = 1.0 / in.b;
                if (fast_divisor(in.b)) in.op |= AGP_I_FASTDIV;
            }
            out.push_back(in);
            st.pop_back();
        }
    }
}

}  // namespace

int agp_compile_program(const int32_t* ops, const int32_t* param_off, int32_t m, const double* params,
                        int32_t n_params, std::vector<AgpInstr>& out, int* need, std::string& err) {
    if (m <= 0) { err = "empty kernel program"; return AGP_ERR_PROGRAM; }
    std::vector<TNode> t;
    t.reserve(m);
    std::vector<int> stack;
    for (int32_t q = 0; q < m; ++q) {
        int32_t op = ops[q];
        if (op < AGP_OP_CONSTANT || op > AGP_OP_WHITE_NOISE) {
            err = "unknown node-type code " + std::to_string(op) + " at program position " + std::to_string(q);
            return AGP_ERR_PROGRAM;
        }
        TNode nd;
        nd.op = op;
        nd.off = 0;
        int np = n_node_params(op);
        if (np > 0) {
            nd.off = param_off[q];
            if (nd.off < 0 || nd.off + np > n_params) {
                err = "param_off out of range at program position " + std::to_string(q);
                return AGP_ERR_PROGRAM;
            }
        }
        if (op == AGP_OP_GAMMA_EXPONENTIAL) {
            double g = params[nd.off + 1];
            if (!(g > 0.0 && g <= 2.0)) {  // GammaExponential constructor assert, src/GP.jl:274
                err = "GammaExponential requires 0 < gamma <= 2";
                return AGP_ERR_PROGRAM;
            }
        }
        bool binary = (op == AGP_OP_PLUS || op == AGP_OP_TIMES || op == AGP_OP_CHANGEPOINT);
        if (binary) {
            if (stack.size() < 2) { err = "operand stack underflow at program position " + std::to_string(q); return AGP_ERR_PROGRAM; }
            nd.right = stack.back(); stack.pop_back();
            nd.left = stack.back(); stack.pop_back();
            int a = t[nd.left].need, b = t[nd.right].need;
            nd.need = (a == b) ? a + 1 : std::max(a, b);
        }
        t.push_back(nd);
        stack.push_back((int)t.size() - 1);
    }
    if (stack.size() != 1) { err = "kernel program does not reduce to a single tree"; return AGP_ERR_PROGRAM; }
    int root = stack.back();
    if (t[root].need > AGP_MAX_STACK) {
        err = "kernel tree needs an operand stack deeper than " + std::to_string(AGP_MAX_STACK);
        return AGP_ERR_PROGRAM;
    }
    *need = t[root].need;
    emit(t, root, params, out);
    return AGP_OK;
}
