// One LGN message-passing level at maxdim 2, fused per jet:
//   radial functions -> edge features -> CG aggregation over the fully connected graph -> CG self product ->
//   concat [ag | node | sq] -> per-irrep complex channel mix,
// forward and hand-written adjoint.  Replaces LGNNodeLevel.forward (lgn/models/lgn_levels.py:96-121) and what
// it calls: cg_product / complex_kron_product (lgn/cg_lib/cg_ops.py:135-298), CatMixReps (lgn/nn/g_nn.py:260-278),
// RadPolyTrig.forward (lgn/nn/position_levels.py:118-209), the edge-feature product (lgn/models/lgn_cg.py:167)
// and the pairwise zonal functions / norms (lgn/cg_lib/zonal_functions.py:123-248).
//
// Thread mapping: lane = particle i (one warp covers a block of 32 particles), warp = channel c.  The radial
// weights R^l_ij[c] = Linear_l(phi(n_ij)) of a tile of neighbours j are produced with the fp64 tensor-core
// MMA (phi (pairs x K) times W^T (K x 4C)) into shared memory and consumed by the (i, c) threads; the neighbour
// sums stay in registers.  Nothing of size O(N^2) ever goes to HBM.
#include "lgae_common.cuh"

namespace lgae {

struct LevelArgs {
    const double* theta;
    int64_t off_a, off_b, off_c, off_w0, off_b0, off_w1, off_b1, off_m00, off_m11;
    const double* p;           // encoder: (B,N,4) real Cartesian ; decoder: (B,N,4,2) complex canonical
    const uint8_t* node_mask;  // encoder: (B,N) or nullptr
    const double* s_in;        // (B,N,C,2)
    const double* v_in;        // (B,N,C,4,2)
    double* sums;              // (B,N,C,10,2)   [A0V(4), A0S, A1Y(4), A1E]
    double* s_pre;             // (B,N,C',2)
    double* v_out;             // (B,N,C',4,2)
    // backward only
    const double* g_s_pre;
    const double* g_v_out;
    double* g_s_in;
    double* g_v_in;
    double* g_y;               // decoder: (B,N,4,2), accumulated
    double* partials;          // (gridDim.x, n_params)
    int64_t n_params;
    int B, N, C, Cout, K;
};

constexpr int TJ = 8;  // neighbours per shared-memory tile of radial weights

// Pair norm n_ij = s / sqrt|s|, s = (p_i - p_j)^2 + 1e-16, with the reference's rounding sequence
// (zonal_functions.py:142-144, 201-248).
LGAE_DEV double pair_norm(const double* pi, const double* pj) {
    const double d0 = pi[0] - pj[0], d1 = pi[1] - pj[1], d2 = pi[2] - pj[2], d3 = pi[3] - pj[3];
    const double s = __dadd_rn(minkowski_sq(d0, d1, d2, d3), 1e-16);
    return s != 0.0 ? __ddiv_rn(s, __dsqrt_rn(fabs(s))) : s;
}

LGAE_DEV double bell(double a, double b, double c, double n) {
    const double cn = c * n;
    const double d = 1.0 + cn * cn + 1e-16;
    return fma(b, 1.0 / d, a);
}

// Column `col` of the stacked radial linear maps: col = l*2C + 2c + (re|im)  (position_levels.py:171-176)
LGAE_DEV double radial_w(const LevelArgs& a, int col, int k) {
    if (k >= a.K || col >= 4 * a.C) return 0.0;
    return col < 2 * a.C ? a.theta[a.off_w0 + (int64_t)col * a.K + k] : a.theta[a.off_w1 + (int64_t)(col - 2 * a.C) * a.K + k];
}
LGAE_DEV double radial_bias(const LevelArgs& a, int col) {
    if (col >= 4 * a.C) return 0.0;
    return col < 2 * a.C ? a.theta[a.off_b0 + col] : a.theta[a.off_b1 + col - 2 * a.C];
}

// Radial weights of the pairs (i in block i0..i0+31, j in j0..j0+tj-1) -> Rs[((jj*C + c)*32 + il)*4 + 2l + ri].
template <int NT, int KS>
LGAE_DEV void radial_tile(const double* p_s, const uint8_t* msk_s, int N, int C, int i0, int j0, int tj,
                          const double* abc_s, const double (&wf)[KS][NT], const double (&bf)[NT][2], double* Rs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    constexpr int KP = 4 * KS;
    for (int grp = warp; grp < 4 * tj; grp += nwarps) {
        const int jj = grp >> 2, il = (grp & 3) * 8 + g, i = i0 + il, j = j0 + jj;
        double n = 0.0;
        bool m = false;
        if (i < N) {
            n = pair_norm(p_s + 4 * i, p_s + 4 * j);
            m = msk_s[i] && msk_s[j] && n != 0.0;
        }
        double acc[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = bf[nt][0]; acc[nt][1] = bf[nt][1]; }
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const int k = 4 * s + q;
            const double phi = m ? bell(abc_s[k], abc_s[KP + k], abc_s[2 * KP + k], n) : 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) dmma(acc[nt][0], acc[nt][1], phi, wf[s][nt]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = 8 * nt + 2 * q;
            if (col < 4 * C) {
                const int l = col >= 2 * C ? 1 : 0;
                const int cc = (col - l * 2 * C) >> 1;
                *reinterpret_cast<double2*>(Rs + ((size_t)((jj * C + cc) * 32 + il)) * 4 + 2 * l) = make_double2(acc[nt][0], acc[nt][1]);
            }
        }
    }
}

template <int NT, int KS>
LGAE_DEV void load_radial_frags(const LevelArgs& a, double (&wf)[KS][NT], double (&bf)[NT][2], double* abc_s) {
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    constexpr int KP = 4 * KS;
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) wf[s][nt] = radial_w(a, 8 * nt + g, 4 * s + q);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        bf[nt][0] = radial_bias(a, 8 * nt + 2 * q);
        bf[nt][1] = radial_bias(a, 8 * nt + 2 * q + 1);
    }
    for (int k = threadIdx.x; k < KP; k += blockDim.x) {
        abc_s[k] = k < a.K ? a.theta[a.off_a + k] : 0.0;
        abc_s[KP + k] = k < a.K ? a.theta[a.off_b + k] : 0.0;
        abc_s[2 * KP + k] = k < a.K ? a.theta[a.off_c + k] : 0.0;
    }
}

// Shared-memory carve-up (in doubles) common to forward and backward.
struct LevelSmem {
    int p, msk, S, V, abc, m00, m11, big;  // offsets
    int total;
};
__host__ __device__ inline LevelSmem level_smem(bool enc, int N, int C, int Cout, int KS, bool bwd) {
    LevelSmem s;
    int o = 0;
    s.p = o; o += enc ? 4 * N : 8 * N;
    s.msk = o; o += ((N + 7) / 8 + 2) & ~1;
    s.S = o; o += 2 * N * C;
    s.V = o; o += 8 * N * C;
    s.abc = o; o += (3 * 4 * KS + 1) & ~1;
    s.m00 = o; o += 2 * Cout * 5 * C;
    s.m11 = o; o += 2 * Cout * 5 * C;
    o = (o + 3) & ~3;
    s.big = o;
    const int cat = 2 * 25 * C * 32;                       // cat / gcat: [(k*5+comp)][32] complex
    const int tile = (enc ? (bwd ? 2 : 1) : 0) * TJ * C * 32 * 4;  // Rs (+ gRs)
    o += cat > tile ? cat : tile;
    if (bwd) {
        o += 2 * Cout * 5 * 32;      // gout_s
        o += 2 * C * 10 * 32;        // gA_s
        o += 2 * 2 * Cout * 5 * C;   // gm_s (m00, m11 gradient accumulators)
        o += 2 * 4 * 32;             // gy_s
    }
    s.total = o;
    return s;
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <bool ENC, int NT, int KS>
__global__ void __launch_bounds__(256) level_fwd_kernel(const LevelArgs a) {
    extern __shared__ __align__(128) double smem[];
    const int N = a.N, C = a.C, Cout = a.Cout;
    const int nib = (N + 31) / 32;
    const int b = blockIdx.x / nib, i0 = (blockIdx.x % nib) * 32;
    const int tid = threadIdx.x, lane = tid & 31, c = tid >> 5;
    const LevelSmem L = level_smem(ENC, N, C, Cout, KS, false);
    double* p_s = smem + L.p;
    uint8_t* msk_s = reinterpret_cast<uint8_t*>(smem + L.msk);
    cplx* S_s = reinterpret_cast<cplx*>(smem + L.S);
    cplx* V_s = reinterpret_cast<cplx*>(smem + L.V);
    double* abc_s = smem + L.abc;
    cplx* m00_s = reinterpret_cast<cplx*>(smem + L.m00);
    cplx* m11_s = reinterpret_cast<cplx*>(smem + L.m11);
    double* Rs = smem + L.big;
    cplx* cat_s = reinterpret_cast<cplx*>(smem + L.big);

    // ---- stage the jet ----
    {
        const int np = ENC ? 4 * N : 8 * N;
        const double* src = a.p + (int64_t)b * np;
        for (int t = tid; t < np; t += blockDim.x) p_s[t] = src[t];
        if (ENC)
            for (int t = tid; t < N; t += blockDim.x)
                msk_s[t] = a.node_mask ? a.node_mask[(int64_t)b * N + t] : (src[4 * t] != 0.0);
        const double* ss = a.s_in + (int64_t)b * N * C * 2;
        for (int t = tid; t < 2 * N * C; t += blockDim.x) (smem + L.S)[t] = ss[t];
        const double* vs = a.v_in + (int64_t)b * N * C * 8;
        for (int t = tid; t < 8 * N * C; t += blockDim.x) (smem + L.V)[t] = vs[t];
        const int nm = Cout * 5 * C;
        for (int t = tid; t < nm; t += blockDim.x) {
            m00_s[t] = cmake(a.theta[a.off_m00 + t], a.theta[a.off_m00 + nm + t]);
            m11_s[t] = cmake(a.theta[a.off_m11 + t], a.theta[a.off_m11 + nm + t]);
        }
    }
    double wf[KS][NT], bf[NT][2];
    cplx R0c = czero(), R1c = czero();
    if (ENC) {
        load_radial_frags<NT, KS>(a, wf, bf, abc_s);
    } else {
        // decoder: all-zero edge mask => R^l[c] = bias_l[c] * (1+i)   
        const double b0 = a.theta[a.off_b0 + c], b1 = a.theta[a.off_b1 + c];
        R0c = cmake(b0, b0);
        R1c = cmake(b1, b1);
    }
    __syncthreads();

    const int i = i0 + lane;
    const bool live = i < N;
    double pi[4] = {0, 0, 0, 0};
    cplx yi[4] = {czero(), czero(), czero(), czero()};
    if (live) {
        if (ENC) {
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) pi[mu] = p_s[4 * i + mu];
        } else {
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) yi[mu] = reinterpret_cast<const cplx*>(p_s)[4 * i + mu];
        }
    }
    cplx A0V[4], A0S = czero(), A1Y[4], A1E = czero();
#pragma unroll
    for (int mu = 0; mu < 4; ++mu) { A0V[mu] = czero(); A1Y[mu] = czero(); }

    for (int j0 = 0; j0 < N; j0 += TJ) {
        const int tj = min(TJ, N - j0);
        if (ENC) {
            radial_tile<NT, KS>(p_s, msk_s, N, C, i0, j0, tj, abc_s, wf, bf, Rs);
            __syncthreads();
        }
        for (int jj = 0; jj < tj; ++jj) {
            const int j = j0 + jj;
            cplx R0 = R0c, R1 = R1c;
            if (ENC) {
                const double4 r = *reinterpret_cast<const double4*>(Rs + ((size_t)((jj * C + c) * 32 + lane)) * 4);
                R0 = cmake(r.x, r.y);
                R1 = cmake(r.z, r.w);
            }
            const cplx Sj = S_s[j * C + c];
            cplx Vj[4];
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) Vj[mu] = V_s[(j * C + c) * 4 + mu];
            cplx Y[4];
            if (ENC) {
                double d[4];
#pragma unroll
                for (int mu = 0; mu < 4; ++mu) d[mu] = pi[mu] - p_s[4 * j + mu];
                canon_from_real(d, Y);
            } else {
#pragma unroll
                for (int mu = 0; mu < 4; ++mu) Y[mu] = csub(yi[mu], reinterpret_cast<const cplx*>(p_s)[4 * j + mu]);
            }
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) cfma(A0V[mu], R0, Vj[mu]);
            cfma(A0S, R0, Sj);
            const cplx t = cmul(R1, Sj);
            cplx e;
            if (ENC) {
                // Y0, Y2 are real, Y3 = -conj... structured: saves a third of the multiplies
                cfmar(A1Y[0], t, Y[0].x);
                cfma(A1Y[1], t, Y[1]);
                cfmar(A1Y[2], t, Y[2].x);
                cfma(A1Y[3], t, Y[3]);
                e = cscale(Vj[0], Y[0].x);
                cfma(e, Vj[1], Y[3]);
                cfmar(e, Vj[2], -Y[2].x);
                cfma(e, Vj[3], Y[1]);
            } else {
#pragma unroll
                for (int mu = 0; mu < 4; ++mu) cfma(A1Y[mu], t, Y[mu]);
                e = ceta(Vj, Y);
            }
            cfma(A1E, R1, e);
        }
        if (ENC) __syncthreads();
    }

    // ---- keep the neighbour sums for the backward pass; build cat = [ag | node | sq] in shared memory ----
    __syncthreads();  // Rs is dead, cat_s aliases it
    if (live) {
        cplx* dst = reinterpret_cast<cplx*>(a.sums) + ((int64_t)(b * N + i) * C + c) * 10;
#pragma unroll
        for (int mu = 0; mu < 4; ++mu) { dst[mu] = A0V[mu]; dst[5 + mu] = A1Y[mu]; }
        dst[4] = A0S;
        dst[9] = A1E;
    }
    {
        const cplx Si = live ? S_s[i * C + c] : czero();
        cplx Vi[4];
#pragma unroll
        for (int mu = 0; mu < 4; ++mu) Vi[mu] = live ? V_s[(i * C + c) * 4 + mu] : czero();
        auto put = [&](int k, int comp, cplx v) { cat_s[(k * 5 + comp) * 32 + lane] = v; };
        put(c, 0, cscale(A1E, 0.5));
        put(C + c, 0, cmul_1pi(A0S));
        put(2 * C + c, 0, Si);
        put(3 * C + c, 0, cscale(ceta(Vi, Vi), 0.5));
        put(4 * C + c, 0, cmul(Si, Si));
#pragma unroll
        for (int mu = 0; mu < 4; ++mu) {
            put(c, 1 + mu, cmul_1pi(A0V[mu]));
            put(C + c, 1 + mu, A1Y[mu]);
            put(2 * C + c, 1 + mu, Vi[mu]);
            const cplx sv = cmul(Vi[mu], Si);
            put(3 * C + c, 1 + mu, sv);
            put(4 * C + c, 1 + mu, sv);
        }
    }
    __syncthreads();
    // ---- complex channel mix (lgn/nn/g_nn.py:95-117): out[c'] = sum_k W[c'][k] cat[k] ----
    for (int it = tid; it < 32 * Cout * 5; it += blockDim.x) {
        const int il = it & 31, r = it >> 5, co = r / 5, comp = r % 5;
        const int ii = i0 + il;
        if (ii >= N) continue;
        const cplx* w = (comp == 0 ? m00_s : m11_s) + co * 5 * C;
        cplx acc = czero();
        for (int k = 0; k < 5 * C; ++k) cfma(acc, w[k], cat_s[(k * 5 + comp) * 32 + il]);
        if (comp == 0)
            reinterpret_cast<cplx*>(a.s_pre)[(int64_t)(b * N + ii) * Cout + co] = acc;
        else
            reinterpret_cast<cplx*>(a.v_out)[((int64_t)(b * N + ii) * Cout + co) * 4 + comp - 1] = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
// Adjoint of the radial functions for one tile: consumes gRs (same layout as Rs), accumulates
//   gw[mt][nt2] : d/dW[col][k] (+ bias in column k == K), as MMA accumulators (col = 8mt+g, k = 8nt2+2q+e)
//   gabc[3][nt2][2] : d/da_k, d/db_k, d/dc_k for k = 8nt2+2q+e, partial over this lane's pairs
template <int NT, int KS, int NT2>
LGAE_DEV void radial_tile_bwd(const double* p_s, const uint8_t* msk_s, int N, int C, int K, int i0, int j0, int tj,
                              const double* abc_s, const double (&w2)[2 * NT][NT2], const double* gRs,
                              double (&gw)[NT][NT2][2], double (&gabc)[3][NT2][2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    constexpr int KP = 4 * KS;
    auto gr_at = [&](int jj, int il, int col) -> double {
        if (col >= 4 * C) return 0.0;
        const int l = col >= 2 * C ? 1 : 0, rem = col - l * 2 * C;
        return gRs[((size_t)((jj * C + (rem >> 1)) * 32 + il)) * 4 + 2 * l + (rem & 1)];
    };
    for (int grp = warp; grp < 4 * tj; grp += nwarps) {
        const int jj = grp >> 2, ib8 = (grp & 3) * 8, il = ib8 + g, i = i0 + il, j = j0 + jj;
        double n = 0.0;
        bool m = false;
        if (i < N) {
            n = pair_norm(p_s + 4 * i, p_s + 4 * j);
            m = msk_s[i] && msk_s[j] && n != 0.0;
        }
        const bool valid = i < N;
        // d/dphi[pair g][k] = sum_col gR[pair][col] W[col][k]
        double gphi[NT2][2];
#pragma unroll
        for (int nt = 0; nt < NT2; ++nt) gphi[nt][0] = gphi[nt][1] = 0.0;
#pragma unroll
        for (int s = 0; s < 2 * NT; ++s) {
            const double av = valid ? gr_at(jj, il, 4 * s + q) : 0.0;
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt) dmma(gphi[nt][0], gphi[nt][1], av, w2[s][nt]);
        }
        if (m) {
            const double nn = n * n;
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * nt + 2 * q + e;
                    if (k < K) {
                        const double ck = abc_s[2 * KP + k], bk = abc_s[KP + k];
                        const double cn = ck * n;
                        const double rd = 1.0 / (1.0 + cn * cn + 1e-16);
                        const double gp = gphi[nt][e];
                        gabc[0][nt][e] += gp;
                        gabc[1][nt][e] = fma(gp, rd, gabc[1][nt][e]);
                        gabc[2][nt][e] = fma(gp, -2.0 * bk * rd * rd * ck * nn, gabc[2][nt][e]);
                    }
                }
        }
        // d/dW[col][k] += sum_pairs gR[pair][col] phi[pair][k]   (k == K: the bias column, phi == 1)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int pp = q + 4 * ks;
            const double n2 = __shfl_sync(0xffffffffu, n, pp * 4);
            const int m2 = __shfl_sync(0xffffffffu, (int)m, pp * 4);
            const bool valid2 = (i0 + ib8 + pp) < N;
            double a3[NT];
#pragma unroll
            for (int mt = 0; mt < NT; ++mt) a3[mt] = valid2 ? gr_at(jj, ib8 + pp, 8 * mt + g) : 0.0;
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt) {
                const int k = 8 * nt + g;
                double phi = 0.0;
                if (k < K) {
                    if (m2) phi = bell(abc_s[k], abc_s[KP + k], abc_s[2 * KP + k], n2);
                } else if (k == K) {
                    phi = 1.0;
                }
#pragma unroll
                for (int mt = 0; mt < NT; ++mt) dmma(gw[mt][nt][0], gw[mt][nt][1], a3[mt], phi);
            }
        }
    }
}

template <bool ENC, int NT, int KS>
__global__ void __launch_bounds__(256) level_bwd_kernel(const LevelArgs a) {
    extern __shared__ __align__(128) double smem[];
    constexpr int NT2 = KS / 2 + 1;
    constexpr int KP = 4 * KS;
    const int N = a.N, C = a.C, Cout = a.Cout, K = a.K;
    const int tid = threadIdx.x, lane = tid & 31, c = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const LevelSmem L = level_smem(ENC, N, C, Cout, KS, true);
    double* p_s = smem + L.p;
    uint8_t* msk_s = reinterpret_cast<uint8_t*>(smem + L.msk);
    cplx* S_s = reinterpret_cast<cplx*>(smem + L.S);
    cplx* V_s = reinterpret_cast<cplx*>(smem + L.V);
    double* abc_s = smem + L.abc;
    cplx* m00_s = reinterpret_cast<cplx*>(smem + L.m00);
    cplx* m11_s = reinterpret_cast<cplx*>(smem + L.m11);
    double* Rs = smem + L.big;
    double* gRs = Rs + TJ * C * 32 * 4;
    cplx* cat_s = reinterpret_cast<cplx*>(smem + L.big);
    const int cat_d = 2 * 25 * C * 32, tile_d = (ENC ? 2 : 0) * TJ * C * 32 * 4;
    double* after = smem + L.big + (cat_d > tile_d ? cat_d : tile_d);
    cplx* gout_s = reinterpret_cast<cplx*>(after);                 // [(c'*5+comp)][32]
    cplx* gA_s = gout_s + Cout * 5 * 32;                            // [(c*10+e)][32]
    cplx* gm_s = gA_s + C * 10 * 32;                                // [2][Cout*5C]
    cplx* gy_s = gm_s + 2 * Cout * 5 * C;                           // [4][32]
    const int nm = Cout * 5 * C;

    for (int t = tid; t < nm; t += blockDim.x) {
        m00_s[t] = cmake(a.theta[a.off_m00 + t], a.theta[a.off_m00 + nm + t]);
        m11_s[t] = cmake(a.theta[a.off_m11 + t], a.theta[a.off_m11 + nm + t]);
    }
    for (int t = tid; t < 2 * nm; t += blockDim.x) gm_s[t] = czero();

    double wf[KS][NT], bf[NT][2], w2[2 * NT][NT2];
    double gw[NT][NT2][2], gabc[3][NT2][2];
    cplx R0c = czero(), R1c = czero(), gR0c = czero(), gR1c = czero();
    if (ENC) {
        load_radial_frags<NT, KS>(a, wf, bf, abc_s);
#pragma unroll
        for (int s = 0; s < 2 * NT; ++s)
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt) w2[s][nt] = radial_w(a, 4 * s + q, 8 * nt + g);
#pragma unroll
        for (int mt = 0; mt < NT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt) gw[mt][nt][0] = gw[mt][nt][1] = 0.0;
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt) gabc[x][nt][0] = gabc[x][nt][1] = 0.0;
    } else {
        const double b0 = a.theta[a.off_b0 + c], b1 = a.theta[a.off_b1 + c];
        R0c = cmake(b0, b0);
        R1c = cmake(b1, b1);
    }

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();
        // ---- stage the jet and the incoming gradients ----
        {
            const int np = ENC ? 4 * N : 8 * N;
            const double* src = a.p + (int64_t)b * np;
            for (int t = tid; t < np; t += blockDim.x) p_s[t] = src[t];
            if (ENC)
                for (int t = tid; t < N; t += blockDim.x)
                    msk_s[t] = a.node_mask ? a.node_mask[(int64_t)b * N + t] : (src[4 * t] != 0.0);
            const double* ss = a.s_in + (int64_t)b * N * C * 2;
            for (int t = tid; t < 2 * N * C; t += blockDim.x) (smem + L.S)[t] = ss[t];
            const double* vs = a.v_in + (int64_t)b * N * C * 8;
            for (int t = tid; t < 8 * N * C; t += blockDim.x) (smem + L.V)[t] = vs[t];
            for (int it = tid; it < 32 * Cout * 5; it += blockDim.x) {
                const int il = it & 31, r = it >> 5, co = r / 5, comp = r % 5;
                cplx v = czero();
                if (il < N) {
                    if (comp == 0) {
                        if (a.g_s_pre) v = reinterpret_cast<const cplx*>(a.g_s_pre)[(int64_t)(b * N + il) * Cout + co];
                    } else {
                        v = reinterpret_cast<const cplx*>(a.g_v_out)[((int64_t)(b * N + il) * Cout + co) * 4 + comp - 1];
                    }
                }
                gout_s[r * 32 + il] = v;
            }
            if (!ENC)
                for (int t = tid; t < 4 * 32; t += blockDim.x) gy_s[t] = czero();
        }
        __syncthreads();
        const int i = lane;
        const bool live = i < N;
        // ---- rebuild cat from the saved neighbour sums ----
        {
            cplx A0V[4], A0S = czero(), A1Y[4], A1E = czero();
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) { A0V[mu] = czero(); A1Y[mu] = czero(); }
            if (live) {
                const cplx* src = reinterpret_cast<const cplx*>(a.sums) + ((int64_t)(b * N + i) * C + c) * 10;
#pragma unroll
                for (int mu = 0; mu < 4; ++mu) { A0V[mu] = src[mu]; A1Y[mu] = src[5 + mu]; }
                A0S = src[4];
                A1E = src[9];
            }
            const cplx Si = live ? S_s[i * C + c] : czero();
            cplx Vi[4];
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) Vi[mu] = live ? V_s[(i * C + c) * 4 + mu] : czero();
            auto put = [&](int k, int comp, cplx v) { cat_s[(k * 5 + comp) * 32 + lane] = v; };
            put(c, 0, cscale(A1E, 0.5));
            put(C + c, 0, cmul_1pi(A0S));
            put(2 * C + c, 0, Si);
            put(3 * C + c, 0, cscale(ceta(Vi, Vi), 0.5));
            put(4 * C + c, 0, cmul(Si, Si));
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) {
                put(c, 1 + mu, cmul_1pi(A0V[mu]));
                put(C + c, 1 + mu, A1Y[mu]);
                put(2 * C + c, 1 + mu, Vi[mu]);
                const cplx sv = cmul(Vi[mu], Si);
                put(3 * C + c, 1 + mu, sv);
                put(4 * C + c, 1 + mu, sv);
            }
        }
        __syncthreads();
        // ---- mix-weight gradients: g_m[c'][k] += sum_{i,comp} gout[c'][comp][i] conj(cat[k][comp][i]) ----
        {
            const int nw = blockDim.x >> 5;
            for (int item = c; item < 2 * nm; item += nw) {
                const int irr = item / nm, r = item % nm, co = r / (5 * C), k = r % (5 * C);
                cplx acc = czero();
                if (irr == 0) {
                    cfmac(acc, cat_s[(k * 5) * 32 + lane], gout_s[(co * 5) * 32 + lane]);
                } else {
#pragma unroll
                    for (int comp = 1; comp < 5; ++comp)
                        cfmac(acc, cat_s[(k * 5 + comp) * 32 + lane], gout_s[(co * 5 + comp) * 32 + lane]);
                }
                acc.x = warp_sum(acc.x);
                acc.y = warp_sum(acc.y);
                if (lane == 0) gm_s[item] = cadd(gm_s[item], acc);
            }
        }
        __syncthreads();
        // ---- gcat = W^H gout, overwriting cat ----
        for (int it = tid; it < 32 * 25 * C; it += blockDim.x) {
            const int il = it & 31, r = it >> 5, k = r / 5, comp = r % 5;
            const cplx* w = (comp == 0 ? m00_s : m11_s) + k;
            cplx acc = czero();
            for (int co = 0; co < Cout; ++co) cfmac(acc, w[co * 5 * C], gout_s[(co * 5 + comp) * 32 + il]);
            cat_s[r * 32 + il] = acc;
        }
        __syncthreads();
        // ---- adjoint of the cat assembly: direct gS/gV, and the adjoints of the four neighbour sums ----
        cplx gS = czero(), gV[4];
        {
            const cplx Si = live ? S_s[i * C + c] : czero();
            cplx Vi[4], gh[4];
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) Vi[mu] = live ? V_s[(i * C + c) * 4 + mu] : czero();
            cghat(Vi, gh);
            auto get = [&](int k, int comp) { return cat_s[(k * 5 + comp) * 32 + lane]; };
            const cplx g_sq00a = get(3 * C + c, 0), g_sq00b = get(4 * C + c, 0);
            gS = get(2 * C + c, 0);
            cfmac(gS, Si, cscale(g_sq00b, 2.0));
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) {
                const cplx g_sq11 = cadd(get(3 * C + c, 1 + mu), get(4 * C + c, 1 + mu));
                gV[mu] = get(2 * C + c, 1 + mu);
                cfmac(gV[mu], Si, g_sq11);
                cfmac(gV[mu], gh[mu], g_sq00a);
                cfmac(gS, Vi[mu], g_sq11);
            }
            // gA: [0..3] A0V, [4] A0S, [5..8] A1Y, [9] A1E
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) {
                gA_s[(c * 10 + mu) * 32 + lane] = cmul_1mi(get(c, 1 + mu));
                gA_s[(c * 10 + 5 + mu) * 32 + lane] = get(C + c, 1 + mu);
            }
            gA_s[(c * 10 + 4) * 32 + lane] = cmul_1mi(get(C + c, 0));
            gA_s[(c * 10 + 9) * 32 + lane] = cscale(get(c, 0), 0.5);
        }
        __syncthreads();  // gcat consumed; Rs/gRs may now overwrite it; gA_s visible
        cplx gy[4] = {czero(), czero(), czero(), czero()};
        // ---- pair loop: own index = lane, other index o runs over tiles ----
        for (int o0 = 0; o0 < N; o0 += TJ) {
            const int to = min(TJ, N - o0);
            if (ENC) {
                radial_tile<NT, KS>(p_s, msk_s, N, C, 0, o0, to, abc_s, wf, bf, Rs);
                __syncthreads();
            }
            {
                // own quantities, reloaded per tile so that they are dead during the radial adjoint
                double pa[4] = {0, 0, 0, 0};
                cplx ya[4] = {czero(), czero(), czero(), czero()};
                if (live) {
                    if (ENC) {
#pragma unroll
                        for (int mu = 0; mu < 4; ++mu) pa[mu] = p_s[4 * i + mu];
                    } else {
#pragma unroll
                        for (int mu = 0; mu < 4; ++mu) ya[mu] = reinterpret_cast<const cplx*>(p_s)[4 * i + mu];
                    }
                }
                const cplx Sa = live ? S_s[i * C + c] : czero();
                cplx Va[4], gAa[10];
#pragma unroll
                for (int mu = 0; mu < 4; ++mu) Va[mu] = live ? V_s[(i * C + c) * 4 + mu] : czero();
#pragma unroll
                for (int e = 0; e < 10; ++e) gAa[e] = gA_s[(c * 10 + e) * 32 + lane];
                for (int oo = 0; oo < to; ++oo) {
                    const int o = o0 + oo;
                    cplx R0 = R0c, R1 = R1c;
                    if (ENC) {
                        const double4 r = *reinterpret_cast<const double4*>(Rs + ((size_t)((oo * C + c) * 32 + lane)) * 4);
                        R0 = cmake(r.x, r.y);
                        R1 = cmake(r.z, r.w);
                    }
                    const cplx So = S_s[o * C + c];
                    cplx Vo[4], Y[4];
#pragma unroll
                    for (int mu = 0; mu < 4; ++mu) Vo[mu] = V_s[(o * C + c) * 4 + mu];
                    if (ENC) {
                        double d[4];
#pragma unroll
                        for (int mu = 0; mu < 4; ++mu) d[mu] = pa[mu] - p_s[4 * o + mu];
                        canon_from_real(d, Y);
                    } else {
#pragma unroll
                        for (int mu = 0; mu < 4; ++mu) Y[mu] = csub(ya[mu], reinterpret_cast<const cplx*>(p_s)[4 * o + mu]);
                    }
                    // role 1: own = receiving node i, other = neighbour j.  Y = Y_ij.
                    {
                        cplx w = czero(), gR0 = czero();
#pragma unroll
                        for (int mu = 0; mu < 4; ++mu) {
                            cfmac(w, Y[mu], gAa[5 + mu]);
                            cfmac(gR0, Vo[mu], gAa[mu]);
                        }
                        cfmac(gR0, So, gAa[4]);
                        const cplx e = ceta(Vo, Y);
                        cplx gR1 = cmulc(So, w);
                        cfmac(gR1, e, gAa[9]);
                        if (ENC) {
                            *reinterpret_cast<double4*>(gRs + ((size_t)((oo * C + c) * 32 + lane)) * 4) =
                                make_double4(gR0.x, gR0.y, gR1.x, gR1.y);
                        } else {
                            gR0c = cadd(gR0c, live ? gR0 : czero());
                            gR1c = cadd(gR1c, live ? gR1 : czero());
                            // gy_i += conj(R1 S_j) gA1Y_i + conj(R1 ghat(V_j)) gA1E_i
                            const cplx rs = cmul(R1, So);
                            cplx gh[4];
                            cghat(Vo, gh);
#pragma unroll
                            for (int mu = 0; mu < 4; ++mu) {
                                cfmac(gy[mu], rs, gAa[5 + mu]);
                                cfmac(gy[mu], cmul(R1, gh[mu]), gAa[9]);
                            }
                        }
                    }
                    // role 2: own = neighbour j, other = receiving node i.  Y_ij = -Y ; R_ij = R_ji (encoder).
                    {
                        cplx gAo[10];
#pragma unroll
                        for (int e = 0; e < 10; ++e) gAo[e] = gA_s[(c * 10 + e) * 32 + o];
                        cplx Yn[4], gh[4];
#pragma unroll
                        for (int mu = 0; mu < 4; ++mu) Yn[mu] = cneg(Y[mu]);
                        cghat(Yn, gh);
                        cplx w = czero();
#pragma unroll
                        for (int mu = 0; mu < 4; ++mu) cfmac(w, Yn[mu], gAo[5 + mu]);
                        const cplx r1g = cmulc(R1, gAo[9]);   // conj(R1) gA1E_i
#pragma unroll
                        for (int mu = 0; mu < 4; ++mu) {
                            cfmac(gV[mu], R0, gAo[mu]);
                            cfmac(gV[mu], gh[mu], r1g);
                        }
                        cfmac(gS, R0, gAo[4]);
                        cfmac(gS, R1, w);
                        if (!ENC) {
                            // gy_j -= conj(R1 S_j) gA1Y_i + conj(R1 ghat(V_j)) gA1E_i
                            const cplx rs = cmul(R1, Sa);
                            cplx gha[4];
                            cghat(Va, gha);
#pragma unroll
                            for (int mu = 0; mu < 4; ++mu) {
                                cplx t = cmulc(rs, gAo[5 + mu]);
                                cfmac(t, cmul(R1, gha[mu]), gAo[9]);
                                gy[mu] = csub(gy[mu], t);
                            }
                        }
                    }
                }
            }
            if (ENC) {
                __syncthreads();
                radial_tile_bwd<NT, KS, NT2>(p_s, msk_s, N, C, K, 0, o0, to, abc_s, w2, gRs, gw, gabc);
                __syncthreads();
            }
        }
        // ---- results for this jet ----
        if (live) {
            reinterpret_cast<cplx*>(a.g_s_in)[(int64_t)(b * N + i) * C + c] = gS;
#pragma unroll
            for (int mu = 0; mu < 4; ++mu) reinterpret_cast<cplx*>(a.g_v_in)[((int64_t)(b * N + i) * C + c) * 4 + mu] = gV[mu];
        }
        if (!ENC) {
            if (live) {
#pragma unroll
                for (int mu = 0; mu < 4; ++mu) {
                    atomicAdd(&gy_s[mu * 32 + lane].x, gy[mu].x);
                    atomicAdd(&gy_s[mu * 32 + lane].y, gy[mu].y);
                }
            }
            __syncthreads();
            for (int t = tid; t < 4 * N; t += blockDim.x) {
                const int ii = t >> 2, mu = t & 3;
                cplx* dst = reinterpret_cast<cplx*>(a.g_y) + (int64_t)(b * N + ii) * 4 + mu;
                *dst = cadd(*dst, gy_s[mu * 32 + ii]);
            }
        }
    }

    // ---- per-CTA partial parameter gradients ----
    __syncthreads();
    double* part = a.partials + (int64_t)blockIdx.x * a.n_params;
    for (int t = tid; t < nm; t += blockDim.x) {
        part[a.off_m00 + t] = gm_s[t].x;
        part[a.off_m00 + nm + t] = gm_s[t].y;
        part[a.off_m11 + t] = gm_s[nm + t].x;
        part[a.off_m11 + nm + t] = gm_s[nm + t].y;
    }
    double* red = smem + L.big;  // reuse: [8*NT cols][8*NT2 + 1 ...] accumulators
    if (ENC) {
        constexpr int NK = 8 * NT2;
        const int ncol = 8 * NT;
        for (int t = tid; t < ncol * NK + 3 * NK; t += blockDim.x) red[t] = 0.0;
        __syncthreads();
#pragma unroll
        for (int mt = 0; mt < NT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) atomicAdd(&red[(8 * mt + g) * NK + 8 * nt + 2 * q + e], gw[mt][nt][e]);
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double v = gabc[x][nt][e];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (g == 0) atomicAdd(&red[ncol * NK + x * NK + 8 * nt + 2 * q + e], v);
                }
        __syncthreads();
        for (int t = tid; t < 4 * C * (K + 1); t += blockDim.x) {
            const int col = t / (K + 1), k = t % (K + 1);
            const double v = red[col * NK + k];
            const int l = col >= 2 * C ? 1 : 0, o = col - l * 2 * C;
            if (k < K)
                part[(l ? a.off_w1 : a.off_w0) + (int64_t)o * K + k] = v;
            else
                part[(l ? a.off_b1 : a.off_b0) + o] = v;
        }
        for (int t = tid; t < 3 * K; t += blockDim.x) {
            const int x = t / K, k = t % K;
            part[(x == 0 ? a.off_a : x == 1 ? a.off_b : a.off_c) + k] = red[ncol * NK + x * NK + k];
        }
    } else {
        // decoder: only the biases learn; R^l[c] = bias (1+i)  =>  g_bias = Re(gR) + Im(gR)
        double v0 = warp_sum(gR0c.x + gR0c.y), v1 = warp_sum(gR1c.x + gR1c.y);
        if (lane == 0) {
            part[a.off_b0 + c] = v0;
            part[a.off_b1 + c] = v1;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------------------
static int pick_ks(int K) { return K <= 12 ? 3 : (K <= 20 ? 5 : (K <= 32 ? 8 : -1)); }

template <bool ENC, int NT, int KS>
static int launch_level(const LevelArgs& a, bool bwd, cudaStream_t st) {
    const LevelSmem L = level_smem(ENC, a.N, a.C, a.Cout, KS, bwd);
    const size_t bytes = (size_t)L.total * sizeof(double);
    if (bytes > 227 * 1024) return LGAE_E_UNSUPPORTED;
    const int threads = 32 * a.C;
    if (!bwd) {
        auto kern = level_fwd_kernel<ENC, NT, KS>;
        if (int rc = ensure_smem((const void*)kern, bytes)) return rc;
        const int nib = (a.N + 31) / 32;
        kern<<<a.B * nib, threads, bytes, st>>>(a);
    } else {
        if (a.N > 32) return LGAE_E_UNSUPPORTED;
        auto kern = level_bwd_kernel<ENC, NT, KS>;
        if (int rc = ensure_smem((const void*)kern, bytes)) return rc;
        kern<<<sm_count(), threads, bytes, st>>>(a);
    }
    count_launch();
    return check_launch(bwd ? "level_bwd" : "level_fwd");
}

template <bool ENC>
static int dispatch_level(const LevelArgs& a, bool bwd, cudaStream_t st) {
    if (a.C < 1 || a.C > LGAE_MAX_CHANNELS || a.Cout < 1 || a.Cout > LGAE_MAX_CHANNELS) return LGAE_E_UNSUPPORTED;
    if (!ENC) return launch_level<false, 1, 3>(a, bwd, st);
    const int nt = (4 * a.C + 7) / 8, ks = pick_ks(a.K);
    if (ks < 0) return LGAE_E_UNSUPPORTED;
#define LGAE_CASE(NTV, KSV) \
    if (nt == NTV && ks == KSV) return launch_level<true, NTV, KSV>(a, bwd, st);
    LGAE_CASE(1, 3) LGAE_CASE(2, 3) LGAE_CASE(3, 3) LGAE_CASE(4, 3)
    LGAE_CASE(1, 5) LGAE_CASE(2, 5) LGAE_CASE(3, 5) LGAE_CASE(4, 5)
    LGAE_CASE(1, 8) LGAE_CASE(2, 8) LGAE_CASE(3, 8) LGAE_CASE(4, 8)
#undef LGAE_CASE
    return LGAE_E_UNSUPPORTED;
}

int run_level(const LgaeModelDesc* d, int level, const double* theta, const double* p_or_y, const uint8_t* node_mask,
              int batch, const double* s_in, const double* v_in, double* sums, double* s_pre, double* v_out,
              const double* g_s_pre, const double* g_v_out, double* g_s_in, double* g_v_in, double* g_y,
              double* partials, bool bwd, cudaStream_t st) {
    if (!d || level < 0 || level >= d->n_levels) return LGAE_E_BADARG;
    LevelArgs a;
    a.theta = theta;
    a.off_a = d->off_rad_a[level]; a.off_b = d->off_rad_b[level]; a.off_c = d->off_rad_c[level];
    a.off_w0 = d->off_rad_w0[level]; a.off_b0 = d->off_rad_b0[level];
    a.off_w1 = d->off_rad_w1[level]; a.off_b1 = d->off_rad_b1[level];
    a.off_m00 = d->off_mix00[level]; a.off_m11 = d->off_mix11[level];
    a.p = p_or_y; a.node_mask = node_mask; a.s_in = s_in; a.v_in = v_in; a.sums = sums; a.s_pre = s_pre; a.v_out = v_out;
    a.g_s_pre = g_s_pre; a.g_v_out = g_v_out; a.g_s_in = g_s_in; a.g_v_in = g_v_in; a.g_y = g_y;
    a.partials = partials; a.n_params = d->n_params;
    a.B = batch; a.N = d->n_particles; a.C = d->channels[level]; a.Cout = d->channels[level + 1]; a.K = d->n_basis;
    if (batch <= 0) return LGAE_OK;
    return d->is_decoder ? dispatch_level<false>(a, bwd, st) : dispatch_level<true>(a, bwd, st);
}

}  // namespace lgae
