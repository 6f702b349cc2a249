// CGMLP (lgn/models/lgn_levels.py:191-227): the scalar MLP applied to every particle's (0,0) features,
// rows = B*N, re/im interleaved on the feature axis: Linear(2C'->w), [Linear(w->w)] x (hidden-1), Linear(w->2C'),
// LeakyReLU(0.01) after all but the last.  This is the one genuinely dense GEMM chain of the LGAE
// ((B*N) x w x w, fp64), so it runs on the fp64 tensor-core MMA (DMMA m8n8k4).
//
// Forward: a warp owns 8*MT rows and keeps them in registers through the whole chain.  The accumulator
// fragment of layer l (row g, columns 2q,2q+1 of every 8-column tile) is reused directly as the A operand of
// layer l+1 by enumerating the reduction index in the order the fragments already have (k-step (tile, e) <->
// column 8*tile + 2q + e); the weights are staged in shared memory pre-permuted into that fragment order, so
// no shuffles or shared-memory round trips are needed between layers.
// Backward: per layer, the data gradient uses the same trick with the transposed weights; the weight gradient
// dW = dZ^T H is a second DMMA GEMM over the chunk's rows with dZ and H staged in shared memory.
#include <cstring>

#include "lgae_common.cuh"

namespace lgae {

struct MlpArgs {
    const double* theta;
    int64_t off_w[LGAE_MAX_LINEAR], off_b[LGAE_MAX_LINEAR];
    int n_lin;   // number of Linear layers (hidden + 1)
    int nin;     // 2C'
    int width;   // hidden width w
    const double* x;  // (rows, nin)
    double* acts;     // (n_lin-1, rows, 8*NTW)
    double* y;        // (rows, nin)
    int64_t rows;
    const double* g_y;
    double* g_x;
    double* part;          // (gridDim.x, part_stride) per-CTA rows of parameter-gradient partials (zeroed by the host)
    int64_t part_stride;
    int64_t po_w[LGAE_MAX_LINEAR], po_b[LGAE_MAX_LINEAR];   // column offsets inside a row
    double slope;
};

// Stage W (out x in, row-major, from theta) into fragment order for  D[row][n] += A[row][k] W[n][k]:
//   Wp[((kt*2+e)*NO + nt)*32 + q*8 + g] = W[8nt+g][8kt+2q+e]
LGAE_DEV void stage_w_fwd(const double* w, int nout, int nink, int KT, int NO, double* Wp) {
    const int total = KT * 2 * NO * 32;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int g = t & 7, q = (t >> 3) & 3, r = t >> 5, nt = r % NO, ke = r / NO, kt = ke >> 1, e = ke & 1;
        const int n = 8 * nt + g, k = 8 * kt + 2 * q + e;
        Wp[t] = (n < nout && k < nink) ? w[(int64_t)n * nink + k] : 0.0;
    }
}
// Transposed use  Gin[row][k] += Gout[row][n] W[n][k]:
//   Wq[((nt*2+e)*KI + kt)*32 + q*8 + g] = W[8nt+2q+e][8kt+g]
LGAE_DEV void stage_w_bwd(const double* w, int nout, int nink, int NO, int KI, double* Wq) {
    const int total = NO * 2 * KI * 32;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int g = t & 7, q = (t >> 3) & 3, r = t >> 5, kt = r % KI, ne = r / KI, nt = ne >> 1, e = ne & 1;
        const int n = 8 * nt + 2 * q + e, k = 8 * kt + g;
        Wq[t] = (n < nout && k < nink) ? w[(int64_t)n * nink + k] : 0.0;
    }
}

template <int MT, int KT, int NO>
LGAE_DEV void layer_mma(const double (&act)[MT][KT][2], double (&acc)[MT][NO][2], const double* Wp, int q, int g) {
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int nt = 0; nt < NO; ++nt) {
                const double bv = Wp[((kt * 2 + e) * NO + nt) * 32 + q * 8 + g];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma(acc[mt][nt][0], acc[mt][nt][1], act[mt][kt][e], bv);
            }
}

template <int NTW, int NTI, int MT>
__global__ void __launch_bounds__(256) mlp_fwd_kernel(const MlpArgs a) {
    extern __shared__ __align__(128) double smem[];
    double* Wp = smem;                                   // NTW*NTW*64
    double* bias_s = smem + NTW * NTW * 64;              // 8*NTW
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int RW = nwarps * 8 * MT, WP = 8 * NTW;
    const int last = a.n_lin - 1;
    for (int64_t r0 = (int64_t)blockIdx.x * RW; r0 < a.rows; r0 += (int64_t)gridDim.x * RW) {
        const int64_t rbase = r0 + warp * 8 * MT;
        double in0[MT][NTI][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTI; ++nt) {
                const int64_t row = rbase + mt * 8 + g;
                const int col = 8 * nt + 2 * q;
                double2 v = make_double2(0.0, 0.0);
                if (row < a.rows && col < a.nin) v = *reinterpret_cast<const double2*>(a.x + row * a.nin + col);
                in0[mt][nt][0] = v.x;
                in0[mt][nt][1] = v.y;
            }
        double act[MT][NTW][2], acc[MT][NTW][2];
        auto init_bias = [&](int nout, int off) {
            __syncthreads();
            for (int t = tid; t < WP; t += blockDim.x) bias_s[t] = t < nout ? a.theta[off + t] : 0.0;
        };
        auto finish_hidden = [&](int l) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt) {
                    act[mt][nt][0] = leaky(acc[mt][nt][0], a.slope);
                    act[mt][nt][1] = leaky(acc[mt][nt][1], a.slope);
                    const int64_t row = rbase + mt * 8 + g;
                    if (row < a.rows)
                        *reinterpret_cast<double2*>(a.acts + ((int64_t)l * a.rows + row) * WP + 8 * nt + 2 * q) =
                            make_double2(act[mt][nt][0], act[mt][nt][1]);
                }
        };
        // ---- first layer: nin -> w ----
        init_bias(a.width, a.off_b[0]);
        stage_w_fwd(a.theta + a.off_w[0], a.width, a.nin, NTI, NTW, Wp);
        __syncthreads();
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) { acc[mt][nt][0] = bias_s[8 * nt + 2 * q]; acc[mt][nt][1] = bias_s[8 * nt + 2 * q + 1]; }
        layer_mma<MT, NTI, NTW>(in0, acc, Wp, q, g);
        finish_hidden(0);
        // ---- hidden layers: w -> w ----
        for (int l = 1; l < last; ++l) {
            init_bias(a.width, a.off_b[l]);
            stage_w_fwd(a.theta + a.off_w[l], a.width, a.width, NTW, NTW, Wp);
            __syncthreads();
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt) { acc[mt][nt][0] = bias_s[8 * nt + 2 * q]; acc[mt][nt][1] = bias_s[8 * nt + 2 * q + 1]; }
            layer_mma<MT, NTW, NTW>(act, acc, Wp, q, g);
            finish_hidden(l);
        }
        // ---- last layer: w -> nin, no activation ----
        init_bias(a.nin, a.off_b[last]);
        stage_w_fwd(a.theta + a.off_w[last], a.nin, a.width, NTW, NTI, Wp);
        __syncthreads();
        double out[MT][NTI][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTI; ++nt) { out[mt][nt][0] = bias_s[8 * nt + 2 * q]; out[mt][nt][1] = bias_s[8 * nt + 2 * q + 1]; }
        layer_mma<MT, NTW, NTI>(act, out, Wp, q, g);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTI; ++nt) {
                const int64_t row = rbase + mt * 8 + g;
                const int col = 8 * nt + 2 * q;
                if (row < a.rows && col < a.nin)
                    *reinterpret_cast<double2*>(a.y + row * a.nin + col) = make_double2(out[mt][nt][0], out[mt][nt][1]);
            }
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
template <int NTW, int NTI, int MT>
__global__ void __launch_bounds__(256) mlp_bwd_kernel(const MlpArgs a) {
    extern __shared__ __align__(128) double smem[];
    constexpr int WS = 8 * NTW + 4;  // padded row stride of the staged tiles (bank-conflict-free fragment loads)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int RW = nwarps * 8 * MT, WP = 8 * NTW;
    double* Wq = smem;                        // NTW*NTW*64
    double* dz_s = Wq + NTW * NTW * 64;       // RW * WS
    double* h_s = dz_s + RW * WS;             // RW * WS
    const int last = a.n_lin - 1;
    double* part = a.part + (int64_t)blockIdx.x * a.part_stride;

    for (int64_t r0 = (int64_t)blockIdx.x * RW; r0 < a.rows; r0 += (int64_t)gridDim.x * RW) {
        const int rl0 = warp * 8 * MT;          // first local row of this warp
        const int64_t rbase = r0 + rl0;
        // gradient wrt the output of the last layer, accumulator-fragment layout
        double dz[MT][NTW][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) {
                dz[mt][nt][0] = dz[mt][nt][1] = 0.0;
                if (nt < NTI) {
                    const int64_t row = rbase + mt * 8 + g;
                    const int col = 8 * nt + 2 * q;
                    if (row < a.rows && col < a.nin) {
                        const double2 v = *reinterpret_cast<const double2*>(a.g_y + row * a.nin + col);
                        dz[mt][nt][0] = v.x;
                        dz[mt][nt][1] = v.y;
                    }
                }
            }
        for (int l = last; l >= 0; --l) {
            const int nout = l == last ? a.nin : a.width, nink = l == 0 ? a.nin : a.width;
            const int NOt = l == last ? NTI : NTW, KIt = l == 0 ? NTI : NTW;
            __syncthreads();
            // ---- stage dZ_l, the layer input H_l and the transposed weights ----
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt)
                    if (nt < NOt)
                        *reinterpret_cast<double2*>(dz_s + (rl0 + mt * 8 + g) * WS + 8 * nt + 2 * q) = make_double2(dz[mt][nt][0], dz[mt][nt][1]);
            {
                const double* src = l == 0 ? a.x : a.acts + (int64_t)(l - 1) * a.rows * WP;
                const int ld = l == 0 ? a.nin : WP, ncols = 8 * KIt;
                for (int t = tid; t < RW * ncols; t += blockDim.x) {
                    const int r = t / ncols, cc = t % ncols;
                    const int64_t row = r0 + r;
                    h_s[r * WS + cc] = (row < a.rows && cc < (l == 0 ? a.nin : WP)) ? src[row * ld + cc] : 0.0;
                }
            }
            stage_w_bwd(a.theta + a.off_w[l], nout, nink, NOt, KIt, Wq);
            __syncthreads();
            // ---- bias gradient: column sums of dZ ----
            for (int n = tid; n < nout; n += blockDim.x) {
                double s = 0.0;
                for (int r = 0; r < RW; ++r) s += dz_s[r * WS + n];
                part[a.po_b[l] + n] += s;
            }
            // ---- weight gradient: dW[n][k] += sum_rows dZ[row][n] H[row][k] ----
            for (int t = warp; t < NOt * KIt; t += nwarps) {
                const int mt = t / KIt, nt = t % KIt;
                double c0 = 0.0, c1 = 0.0;
                for (int ks = 0; ks < RW / 4; ++ks) {
                    const double av = dz_s[(4 * ks + q) * WS + 8 * mt + g];
                    const double bv = h_s[(4 * ks + q) * WS + 8 * nt + g];
                    dmma(c0, c1, av, bv);
                }
                const int n = 8 * mt + g, k = 8 * nt + 2 * q;
                if (n < nout) {
                    if (k < nink) part[a.po_w[l] + (int64_t)n * nink + k] += c0;
                    if (k + 1 < nink) part[a.po_w[l] + (int64_t)n * nink + k + 1] += c1;
                }
            }
            // ---- data gradient: Gin[row][k] = sum_n dZ[row][n] W[n][k], then through the LeakyReLU of layer l-1 ----
            double gin[MT][NTW][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int kt = 0; kt < NTW; ++kt) gin[mt][kt][0] = gin[mt][kt][1] = 0.0;
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) {
                if (nt < NOt) {
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int kt = 0; kt < NTW; ++kt) {
                            if (kt < KIt) {
                                const double bv = Wq[((nt * 2 + e) * KIt + kt) * 32 + q * 8 + g];
#pragma unroll
                                for (int mt = 0; mt < MT; ++mt) dmma(gin[mt][kt][0], gin[mt][kt][1], dz[mt][nt][e], bv);
                            }
                        }
                }
            }
            if (l > 0) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int kt = 0; kt < NTW; ++kt) {
                        const double2 h = *reinterpret_cast<const double2*>(h_s + (rl0 + mt * 8 + g) * WS + 8 * kt + 2 * q);
                        dz[mt][kt][0] = gin[mt][kt][0] * (h.x > 0.0 ? 1.0 : a.slope);
                        dz[mt][kt][1] = gin[mt][kt][1] * (h.y > 0.0 ? 1.0 : a.slope);
                    }
            } else if (a.g_x) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int kt = 0; kt < NTI; ++kt) {
                        const int64_t row = rbase + mt * 8 + g;
                        const int col = 8 * kt + 2 * q;
                        if (row < a.rows && col < a.nin)
                            *reinterpret_cast<double2*>(a.g_x + row * a.nin + col) = make_double2(gin[mt][kt][0], gin[mt][kt][1]);
                    }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
template <int NTW, int NTI, int MT>
static int launch_mlp(const MlpArgs& a, bool bwd, cudaStream_t st) {
    const int threads = 256, RW = (threads / 32) * 8 * MT;
    if (!bwd) {
        const size_t bytes = (size_t)(NTW * NTW * 64 + 8 * NTW) * sizeof(double);
        auto kern = mlp_fwd_kernel<NTW, NTI, MT>;
        if (int rc = ensure_smem((const void*)kern, bytes)) return rc;
        int64_t chunks = (a.rows + RW - 1) / RW;
        int grid = (int)(chunks < (int64_t)2 * sm_count() ? chunks : (int64_t)2 * sm_count());
        kern<<<grid, threads, bytes, st>>>(a);
        count_launch();
        return check_launch("mlp_fwd");
    }
    const size_t bytes = (size_t)(NTW * NTW * 64 + 2 * RW * (8 * NTW + 4)) * sizeof(double);
    if (bytes > 227 * 1024) return LGAE_E_UNSUPPORTED;
    auto kern = mlp_bwd_kernel<NTW, NTI, MT>;
    if (int rc = ensure_smem((const void*)kern, bytes)) return rc;
    kern<<<sm_count(), threads, bytes, st>>>(a);
    count_launch();
    return check_launch("mlp_bwd");
}

int run_mlp(const LgaeModelDesc* d, int level, const double* theta, const double* x, int64_t rows, double* acts,
            double* y, const double* g_y, double* g_x, PartPlan* plan, bool bwd, cudaStream_t st) {
    if (!d || level < 0 || level >= d->n_levels || !d->has_mlp) return LGAE_E_BADARG;
    MlpArgs a;
    memset(&a, 0, sizeof(a));
    a.theta = theta;
    a.n_lin = d->mlp_hidden + 1;
    if (a.n_lin < 2 || a.n_lin > LGAE_MAX_LINEAR) return LGAE_E_UNSUPPORTED;
    for (int i = 0; i < a.n_lin; ++i) { a.off_w[i] = d->off_mlp_w[level][i]; a.off_b[i] = d->off_mlp_b[level][i]; }
    a.nin = 2 * d->channels[level + 1];
    a.width = d->mlp_width[level];
    a.x = x; a.acts = acts; a.y = y; a.rows = rows; a.g_y = g_y; a.g_x = g_x;
    a.slope = 0.01;
    if (rows <= 0) return LGAE_OK;
    if (bwd) {
        if (!plan) return LGAE_E_BADARG;
        const int grid = sm_count();
        int64_t w = 0;
        for (int i = 0; i < a.n_lin; ++i) {
            const int nout = i == a.n_lin - 1 ? a.nin : a.width, nink = i == 0 ? a.nin : a.width;
            a.po_w[i] = w; w += (int64_t)nout * nink;
            a.po_b[i] = w; w += nout;
        }
        const int64_t off = plan->block(grid, w);
        a.part = plan->base + off;
        a.part_stride = w;
        for (int i = 0; i < a.n_lin; ++i) {
            const int nout = i == a.n_lin - 1 ? a.nin : a.width, nink = i == 0 ? a.nin : a.width;
            if (int rc = plan->seg(a.off_w[i], off, w, a.po_w[i], (int64_t)nout * nink, grid)) return rc;
            if (int rc = plan->seg(a.off_b[i], off, w, a.po_b[i], nout, grid)) return rc;
        }
        if (cudaMemsetAsync(a.part, 0, (size_t)grid * w * sizeof(double), st) != cudaSuccess) return check_launch("memset mlp partials");
    }
    const int ntw = (a.width + 7) / 8, nti = (a.nin + 7) / 8;
#define LGAE_MLP_CASE(W, I, M) \
    if (ntw == W && nti == I) return launch_mlp<W, I, M>(a, bwd, st);
    LGAE_MLP_CASE(1, 1, 2) LGAE_MLP_CASE(2, 1, 2) LGAE_MLP_CASE(3, 1, 2) LGAE_MLP_CASE(4, 1, 2)
    LGAE_MLP_CASE(5, 1, 2) LGAE_MLP_CASE(6, 1, 2)
    LGAE_MLP_CASE(8, 2, 1) LGAE_MLP_CASE(9, 2, 1) LGAE_MLP_CASE(11, 2, 1) LGAE_MLP_CASE(12, 2, 1)
#undef LGAE_MLP_CASE
    return LGAE_E_UNSUPPORTED;
}

// Width (doubles) of one row of partials of the MLP adjoint.
int64_t mlp_part_width(const LgaeModelDesc* d, int level) {
    const int n_lin = d->mlp_hidden + 1, nin = 2 * d->channels[level + 1], width = d->mlp_width[level];
    int64_t w = 0;
    for (int i = 0; i < n_lin; ++i) {
        const int nout = i == n_lin - 1 ? nin : width, nink = i == 0 ? nin : width;
        w += (int64_t)nout * nink + nout;
    }
    return w;
}
int mlp_bwd_grid() { return sm_count(); }

int mlp_padded_width(const LgaeModelDesc* d, int level) { return 8 * ((d->mlp_width[level] + 7) / 8); }

}  // namespace lgae
