// Adam on the flat parameter buffers of the training step (the caller of the hot path: optimizer_encoder.step() /
// optimizer_decoder.step(), utils/train.py:342-343, torch.optim.Adam built in utils/initialize.py:152-158).
// FusedTrainStep leaves the gradients of both models in one flat bucket and the parameters of each model in one flat
// buffer, so the update of up to two models is ONE element-wise launch instead of ~130 per-tensor updates; the step count
// lives on the device (bias corrections computed in the kernel), so the update can sit in the same CUDA graph as the step.
#include "lgae_common.cuh"

namespace lgae {

struct AdamArgs {
    double* theta[2];
    const double* grad[2];
    double* m[2];
    double* v[2];
    int64_t n[2];
    double lr, beta1, beta2, eps, weight_decay;
    int64_t* step;            // device: number of updates done so far
    unsigned int* counter;    // device: completion counter of the launch (zero between launches)
};

// Same arithmetic as torch.optim.Adam (amsgrad = False, maximize = False): L2 weight decay added to the gradient,
// exp_avg <- exp_avg + (g - exp_avg)(1 - beta1), exp_avg_sq <- beta2 exp_avg_sq + (1 - beta2) g^2,
// theta <- theta - (lr / (1 - beta1^t)) exp_avg / (sqrt(exp_avg_sq) / sqrt(1 - beta2^t) + eps).
__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
    // no early launch_dependents: the next step's kernels read theta in their prologues, which must not overlap this update
    pdl_wait();
    const double t = (double)(a.step[0] + 1);
    const double bc1 = 1.0 - pow(a.beta1, t), bc2 = 1.0 - pow(a.beta2, t);
    const double step_size = a.lr / bc1, rs2 = sqrt(bc2);
    const int64_t total = a.n[0] + a.n[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = i < a.n[0] ? 0 : 1;
        const int64_t k = s ? i - a.n[0] : i;
        double g = a.grad[s][k];
        const double p = a.theta[s][k];
        if (a.weight_decay != 0.0) g = fma(a.weight_decay, p, g);
        double m = a.m[s][k], v = a.v[s][k];
        m = m + (g - m) * (1.0 - a.beta1);
        v = a.beta2 * v + (1.0 - a.beta2) * g * g;
        a.m[s][k] = m;
        a.v[s][k] = v;
        a.theta[s][k] = p - step_size * (m / (sqrt(v) / rs2 + a.eps));
    }
    // the last block to finish advances the step count (every block has read it by then) and re-arms the counter
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.counter, 1u) == gridDim.x - 1) {
            a.step[0] += 1;
            *a.counter = 0u;
        }
    }
}

// torch.optim.RMSprop (centered = False): square_avg <- alpha square_avg + (1 - alpha) g^2; avg = sqrt(square_avg) + eps;
// momentum > 0: buf <- momentum buf + g / avg, theta <- theta - lr buf; else theta <- theta - lr g / avg.
struct RmsArgs {
    double* theta[2];
    const double* grad[2];
    double* sq[2];
    double* buf[2];
    int64_t n[2];
    double lr, alpha, eps, momentum, weight_decay;
};
__global__ void __launch_bounds__(256) rmsprop_kernel(const RmsArgs a) {
    // no early launch_dependents: the next step's kernels read theta in their prologues, which must not overlap this update
    pdl_wait();
    const int64_t total = a.n[0] + a.n[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = i < a.n[0] ? 0 : 1;
        const int64_t k = s ? i - a.n[0] : i;
        double g = a.grad[s][k];
        const double p = a.theta[s][k];
        if (a.weight_decay != 0.0) g = fma(a.weight_decay, p, g);
        const double sq = a.alpha * a.sq[s][k] + (1.0 - a.alpha) * g * g;
        a.sq[s][k] = sq;
        const double avg = sqrt(sq) + a.eps;
        if (a.momentum > 0.0) {
            const double b = a.momentum * a.buf[s][k] + g / avg;
            a.buf[s][k] = b;
            a.theta[s][k] = p - a.lr * b;
        } else {
            a.theta[s][k] = p - a.lr * (g / avg);
        }
    }
}

}  // namespace lgae

using namespace lgae;

extern "C" int lgae_rmsprop_step(double* theta_a, const double* grad_a, double* square_avg_a, double* momentum_buf_a, int64_t n_a,
                                 double* theta_b, const double* grad_b, double* square_avg_b, double* momentum_buf_b, int64_t n_b, double lr,
                                 double alpha, double eps, double momentum, double weight_decay, void* stream) {
    if (n_a < 0 || n_b < 0) return LGAE_E_BADARG;
    if (!(lr >= 0.0) || !(alpha >= 0.0) || !(eps >= 0.0) || !(momentum >= 0.0)) return LGAE_E_BADARG;
    if (n_a > 0 && (!theta_a || !grad_a || !square_avg_a || (momentum > 0.0 && !momentum_buf_a))) return LGAE_E_BADARG;
    if (n_b > 0 && (!theta_b || !grad_b || !square_avg_b || (momentum > 0.0 && !momentum_buf_b))) return LGAE_E_BADARG;
    if (n_a + n_b == 0) return LGAE_OK;
    RmsArgs a;
    a.theta[0] = theta_a; a.grad[0] = grad_a; a.sq[0] = square_avg_a; a.buf[0] = momentum_buf_a; a.n[0] = n_a;
    a.theta[1] = theta_b; a.grad[1] = grad_b; a.sq[1] = square_avg_b; a.buf[1] = momentum_buf_b; a.n[1] = n_b;
    a.lr = lr; a.alpha = alpha; a.eps = eps; a.momentum = momentum; a.weight_decay = weight_decay;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = (int)((n_a + n_b + 255) / 256);
    const int cap = 2 * sm_count();
    grid = grid > cap ? cap : grid;
    LaunchScope ls_("rmsprop", st);
    launch_k(rmsprop_kernel, dim3(grid), dim3(256), 0, st, a);
    return check_launch("rmsprop");
}

extern "C" int lgae_adam_step(double* theta_a, const double* grad_a, double* exp_avg_a, double* exp_avg_sq_a, int64_t n_a, double* theta_b,
                              const double* grad_b, double* exp_avg_b, double* exp_avg_sq_b, int64_t n_b, double lr, double beta1, double beta2,
                              double eps, double weight_decay, int64_t* step_state, void* stream) {
    if (n_a < 0 || n_b < 0 || !step_state) return LGAE_E_BADARG;
    if (!(lr >= 0.0) || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0)) return LGAE_E_BADARG;
    if (n_a > 0 && (!theta_a || !grad_a || !exp_avg_a || !exp_avg_sq_a)) return LGAE_E_BADARG;
    if (n_b > 0 && (!theta_b || !grad_b || !exp_avg_b || !exp_avg_sq_b)) return LGAE_E_BADARG;
    if (n_a + n_b == 0) return LGAE_OK;
    AdamArgs a;
    a.theta[0] = theta_a; a.grad[0] = grad_a; a.m[0] = exp_avg_a; a.v[0] = exp_avg_sq_a; a.n[0] = n_a;
    a.theta[1] = theta_b; a.grad[1] = grad_b; a.m[1] = exp_avg_b; a.v[1] = exp_avg_sq_b; a.n[1] = n_b;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.step = step_state;
    a.counter = reinterpret_cast<unsigned int*>(step_state + 1);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = n_a + n_b;
    int grid = (int)((total + 255) / 256);
    const int cap = 2 * sm_count();
    grid = grid > cap ? cap : grid;
    LaunchScope ls_("adam", st);
    launch_k(adam_kernel, dim3(grid), dim3(256), 0, st, a);
    return check_launch("adam");
}
