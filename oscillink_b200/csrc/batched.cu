// K3: batched serving kernel -- many small lattices of equal (N, D, k) settled concurrently.
//
// Every CG reduction of solver.py:22-36 is per column, so a CTA that owns an 8-column slab of
// ALL N rows of one lattice can run the whole recurrence on chip:
//   p (the only vector that is gathered)  -> shared memory  [N][8] fp32
//   ELL graph (u16 neighbour, fp32 W)     -> shared memory  (re-used by every iteration)
//   x, r, Ap                              -> registers (each thread owns fixed (row, 4-col) tasks)
// Only the stop test (max over ALL D columns, solver.py:29-31) crosses CTAs: the G = D/8 CTAs of
// a lattice form a group that exchanges one float per iteration through global atomics (the
// grid is launched cooperatively so a group is always co-resident).
//
// The same kernel optionally chains: settle (lattice.py:170-207) -> stationary solve
// (lattice.py:245-265) -> deltaH (receipts.py:21-25), touching HBM only for Y in and U / U* out.
#include <cooperative_groups.h>

#include "common.cuh"

namespace osc {

constexpr int BC = 8;        // columns per slab
constexpr int BT_MAX = 1024;  // threads per CTA are chosen per shape (multiple of 32)
constexpr int BW_MAX = BT_MAX / 32;
#define BT ((int)blockDim.x)
#define BW ((int)(blockDim.x >> 5))

struct BatchedK {
  const int32_t* nbr;
  const float* W;
  const int32_t* deg;
  const float* Y;
  const float* U_in;
  const float* psi;
  const float* gates;
  float* U_out;
  float* Ustar_out;
  float* stats;
  double* dh_part;
  unsigned* sync;
  int64_t batch, N;
  int k, kp, D, G, groups, maxit;
  int do_settle, do_ustar, do_dh;
  float lamG, lamC, lamQ, dt;
  double tol_settle, tol_ustar;
  int max_iters_settle, max_iters_ustar;
};

struct SolveCoef {
  float diag0, diag1;  // operator diagonal = diag0 + diag1 * b_i
  float offc;
  float lamG, lamQ, dt;
  int settle, kq;
};

__device__ __forceinline__ float md_of(const SolveCoef& c, float b) {
  const float base = __fadd_rn(c.lamG, __fmul_rn(c.lamQ, b));
  return c.settle ? __fadd_rn(1.0f, __fmul_rn(c.dt, base)) : base;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_fma(float a, float4 x, float4 y) {
  return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// sum over all tasks of the CTA that share this thread's column half; one __syncthreads.
__device__ __forceinline__ float4 block_colsum(float4 v, float4* red, int lane, int warp, int half) {
#pragma unroll
  for (int o = 2; o < 32; o <<= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
    v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  if (lane < 2) red[warp * 2 + lane] = v;
  __syncthreads();
  float4 t = f4_zero();
  const int nw = BW;
#pragma unroll 4
  for (int w = 0; w < nw; ++w) t = f4_add(t, red[w * 2 + half]);
  return t;
}

template <int TPT>
struct Slab {
  float4 X[TPT], R[TPT], AP[TPT];
};

// A(p) for one task: diag*p_own - offc * sum_t W_t p[nbr_t]
template <int KQ>
__device__ __forceinline__ float4 apply_task(const float4* p_s, const ushort4* nbr_s,
                                             const float4* w_s, int row, int half, int kq_rt,
                                             float diag, float offc) {
  float4 acc = f4_zero();
  const int kq = KQ > 0 ? KQ : kq_rt;
#pragma unroll
  for (int c = 0; c < kq; ++c) {
    const ushort4 jj = nbr_s[row * kq + c];
    const float4 ww = w_s[row * kq + c];
    acc = f4_fma(ww.x, p_s[jj.x * 2 + half], acc);
    acc = f4_fma(ww.y, p_s[jj.y * 2 + half], acc);
    acc = f4_fma(ww.z, p_s[jj.z * 2 + half], acc);
    acc = f4_fma(ww.w, p_s[jj.w * 2 + half], acc);
  }
  const float4 own = p_s[row * 2 + half];
  return make_float4(diag * own.x - offc * acc.x, diag * own.y - offc * acc.y,
                     diag * own.z - offc * acc.z, diag * own.w - offc * acc.w);
}

// One PCG solve for this CTA's slab.  On exit st.X holds the solution.  Returns iterations;
// *res_out the last max-column residual (group-wide).
template <int TPT, int KQ>
__device__ int slab_solve(Slab<TPT>& st, const SolveCoef& c, double tol, int max_iters, float4* p_s,
                          const float* gates_s, const ushort4* nbr_s, const float4* w_s,
                          float4* red, unsigned* sync_base, int G, const bool (&act)[TPT],
                          float* res_out, unsigned* flag_s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = tid & 1;
  // ---- r0 = b - A x0 ; p = z0 ; rz
  __syncthreads();  // previous users of p_s are done
#pragma unroll
  for (int m = 0; m < TPT; ++m)
    if (act[m]) p_s[tid + BT * m] = st.X[m];
  __syncthreads();
  float4 z0[TPT];
  float4 part = f4_zero();
#pragma unroll
  for (int m = 0; m < TPT; ++m) {
    z0[m] = f4_zero();
    if (act[m]) {
      const int q = tid + BT * m, row = q >> 1;
      const float b = gates_s[row];
      const float4 a = apply_task<KQ>(p_s, nbr_s, w_s, row, half, c.kq, c.diag0 + c.diag1 * b,
                                  c.offc);
      const float md = md_of(c, b) + 1e-12f;
      float4 r = st.R[m];
      r = make_float4(r.x - a.x, r.y - a.y, r.z - a.z, r.w - a.w);
      st.R[m] = r;
      z0[m] = make_float4(__fdiv_rn(r.x, md), __fdiv_rn(r.y, md), __fdiv_rn(r.z, md),
                          __fdiv_rn(r.w, md));
      part = f4_add(part, f4_mul(r, z0[m]));
    }
  }
  float4 rz = block_colsum(part, red, lane, warp, half);  // sync: all gathers of x0 finished
#pragma unroll
  for (int m = 0; m < TPT; ++m)
    if (act[m]) p_s[tid + BT * m] = z0[m];
  __syncthreads();

  int it = 0;
  float res = __int_as_float(0x7fc00000);
  for (it = 1; it <= max_iters; ++it) {
    float4* redA = red + 1 * (BW_MAX * 2);
    float4* redB = red + 2 * (BW_MAX * 2);  // two consecutive buffers (rr, rz')
    // ---- Ap, p.Ap
    part = f4_zero();
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      if (act[m]) {
        const int q = tid + BT * m, row = q >> 1;
        const float b = gates_s[row];
        st.AP[m] = apply_task<KQ>(p_s, nbr_s, w_s, row, half, c.kq, c.diag0 + c.diag1 * b,
                              c.offc);
        part = f4_add(part, f4_mul(p_s[q], st.AP[m]));
      }
    }
    const float4 pap = block_colsum(part, redA, lane, warp, half);
    const float4 alpha = make_float4(__fdiv_rn(rz.x, pap.x + 1e-18f), __fdiv_rn(rz.y, pap.y + 1e-18f),
                                     __fdiv_rn(rz.z, pap.z + 1e-18f), __fdiv_rn(rz.w, pap.w + 1e-18f));
    // ---- x, r update; rr and rz'
    float4 prr = f4_zero(), prz = f4_zero();
    float4 zz[TPT];
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      zz[m] = f4_zero();
      if (act[m]) {
        const int q = tid + BT * m, row = q >> 1;
        const float4 p = p_s[q];
        float4 x = st.X[m], r = st.R[m];
        const float4 ap = st.AP[m];
        x = make_float4(__fadd_rn(x.x, __fmul_rn(p.x, alpha.x)), __fadd_rn(x.y, __fmul_rn(p.y, alpha.y)),
                        __fadd_rn(x.z, __fmul_rn(p.z, alpha.z)), __fadd_rn(x.w, __fmul_rn(p.w, alpha.w)));
        r = make_float4(__fsub_rn(r.x, __fmul_rn(ap.x, alpha.x)), __fsub_rn(r.y, __fmul_rn(ap.y, alpha.y)),
                        __fsub_rn(r.z, __fmul_rn(ap.z, alpha.z)), __fsub_rn(r.w, __fmul_rn(ap.w, alpha.w)));
        st.X[m] = x;
        st.R[m] = r;
        const float md = md_of(c, gates_s[row]) + 1e-12f;
        const float4 z = make_float4(__fdiv_rn(r.x, md), __fdiv_rn(r.y, md), __fdiv_rn(r.z, md),
                                     __fdiv_rn(r.w, md));
        zz[m] = z;
        prr = f4_add(prr, f4_mul(r, r));
        prz = f4_add(prz, f4_mul(r, z));
      }
    }
    // two reductions, one barrier: lanes 0/1 of each warp publish both
#pragma unroll
    for (int o = 2; o < 32; o <<= 1) {
      prr.x += __shfl_xor_sync(0xffffffffu, prr.x, o);
      prr.y += __shfl_xor_sync(0xffffffffu, prr.y, o);
      prr.z += __shfl_xor_sync(0xffffffffu, prr.z, o);
      prr.w += __shfl_xor_sync(0xffffffffu, prr.w, o);
      prz.x += __shfl_xor_sync(0xffffffffu, prz.x, o);
      prz.y += __shfl_xor_sync(0xffffffffu, prz.y, o);
      prz.z += __shfl_xor_sync(0xffffffffu, prz.z, o);
      prz.w += __shfl_xor_sync(0xffffffffu, prz.w, o);
    }
    if (lane < 2) {
      redB[warp * 2 + lane] = prr;
      redB[BW_MAX * 2 + warp * 2 + lane] = prz;
    }
    __syncthreads();
    float4 rr = f4_zero(), rzn = f4_zero();
    const int nw = BW;
#pragma unroll 4
    for (int w = 0; w < nw; ++w) {
      rr = f4_add(rr, redB[w * 2 + half]);
      rzn = f4_add(rzn, redB[BW_MAX * 2 + w * 2 + half]);
    }
    // slab max of the column residuals -> group-wide max (solver.py:29)
    float mx = fmaxf(fmaxf(rr.x, rr.y), fmaxf(rr.z, rr.w));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    unsigned* sres = sync_base + 2 * it;
    if (tid == 0) {
      atomicMax(sres, __float_as_uint(fmaxf(mx, 0.f)));
      __threadfence();
      atomicAdd(sres + 1, 1u);
    }
    // ---- p = z + beta p (harmless if this turns out to be the last iteration)
    const float4 beta = make_float4(__fdiv_rn(rzn.x, rz.x + 1e-18f), __fdiv_rn(rzn.y, rz.y + 1e-18f),
                                    __fdiv_rn(rzn.z, rz.z + 1e-18f), __fdiv_rn(rzn.w, rz.w + 1e-18f));
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      if (act[m]) {
        const int q = tid + BT * m;
        const float4 p = p_s[q];
        const float4 z = zz[m];
        p_s[q] = make_float4(__fadd_rn(z.x, __fmul_rn(p.x, beta.x)), __fadd_rn(z.y, __fmul_rn(p.y, beta.y)),
                             __fadd_rn(z.z, __fmul_rn(p.z, beta.z)), __fadd_rn(z.w, __fmul_rn(p.w, beta.w)));
      }
    }
    rz = rzn;
    if (tid == 0) {
      volatile unsigned* cnt = sres + 1;
      while (*cnt < (unsigned)G) {
      }
      __threadfence();
      *flag_s = *((volatile unsigned*)sres);
    }
    __syncthreads();
    res = __fsqrt_rn(__uint_as_float(*flag_s));
    if ((double)res <= tol) break;
    if (it == max_iters) break;
  }
  *res_out = res;
  return it;
}

template <int TPT, int KQ>
__global__ void __launch_bounds__(TPT == 3 ? 800 : (TPT == 4 ? 640 : 1024), 1)
batched_settle_kernel(BatchedK P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = (int)P.N, kq = P.kp / 4;
  float4* p_s = reinterpret_cast<float4*>(smem_raw);                 // [N][2]
  float4* w_s = p_s + (size_t)N * 2;                                  // [N][kq]
  float4* red = w_s + (size_t)N * kq;                                 // [4][BW_MAX*2]
  ushort4* nbr_s = reinterpret_cast<ushort4*>(red + 4 * BW_MAX * 2);  // [N][kq]
  float* gates_s = reinterpret_cast<float*>(nbr_s + (size_t)N * kq);  // [N]
  unsigned* flag_s = reinterpret_cast<unsigned*>(gates_s + N);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = tid & 1;
  const int gid = blockIdx.x / P.G, slab = blockIdx.x % P.G;
  const int col = slab * BC + half * 4;
  const bool col_ok = col + 3 < P.D;
  bool act[TPT];
#pragma unroll
  for (int m = 0; m < TPT; ++m) act[m] = col_ok && ((tid + BT * m) >> 1) < N;

  for (int64_t b = gid; b < P.batch; b += P.groups) {
    __syncthreads();
    // ---- stage graph + gates (one 4-neighbour chunk per thread-iteration, 16 B global loads)
    {
      const int32_t* nb = P.nbr + b * P.N * P.k;
      const float* wt = P.W + b * P.N * P.k;
      const int32_t* dg = P.deg + b * P.N;
      const bool vec = (P.k & 3) == 0;
      for (int e = tid; e < N * kq; e += BT) {
        const int row = e / kq, c = e - row * kq;
        const int d = dg[row];
        int j[4];
        float w[4];
        if (vec) {
          const int4 jj = *reinterpret_cast<const int4*>(nb + (int64_t)row * P.k + 4 * c);
          const float4 ww = *reinterpret_cast<const float4*>(wt + (int64_t)row * P.k + 4 * c);
          j[0] = jj.x; j[1] = jj.y; j[2] = jj.z; j[3] = jj.w;
          w[0] = ww.x; w[1] = ww.y; w[2] = ww.z; w[3] = ww.w;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int t = 4 * c + u;
            j[u] = t < P.k ? nb[(int64_t)row * P.k + t] : -1;
            w[u] = t < P.k ? wt[(int64_t)row * P.k + t] : 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = (4 * c + u) < d;
          if (!ok) {
            j[u] = row;  // padding gathers the row itself with weight 0
            w[u] = 0.f;
          }
        }
        nbr_s[e] = make_ushort4((unsigned short)j[0], (unsigned short)j[1], (unsigned short)j[2],
                                (unsigned short)j[3]);
        w_s[e] = make_float4(w[0], w[1], w[2], w[3]);
      }
      for (int e = tid; e < N; e += BT) gates_s[e] = P.gates ? P.gates[b * P.N + e] : 1.0f;
    }
    const float* Yb = P.Y + b * P.N * P.D;
    const float* Ub = (P.U_in ? P.U_in : P.Y) + b * P.N * P.D;
    float* Uo = P.U_out ? P.U_out + b * P.N * P.D : nullptr;
    const float4 psi4 = col_ok ? *reinterpret_cast<const float4*>(P.psi + b * P.D + col) : f4_zero();
    __syncthreads();

    Slab<TPT> st;
    unsigned* sync_b = P.sync + (size_t)b * 2 * (P.maxit + 1) * 2;
    float res = 0.f;
    int iters = 0;
    // ---------------- settle: (I + dt M) U+ = U + dt (lamG Y + lamQ b psi^T), x0 = U
    if (P.do_settle) {
      SolveCoef c;
      c.lamG = P.lamG; c.lamQ = P.lamQ; c.dt = P.dt;
      c.settle = 1;
      c.kq = kq;
      c.diag0 = 1.0f + P.dt * (P.lamG + P.lamC);
      c.diag1 = P.dt * P.lamQ;
      c.offc = P.dt * P.lamC;
#pragma unroll
      for (int m = 0; m < TPT; ++m) {
        st.X[m] = st.R[m] = st.AP[m] = f4_zero();
        if (act[m]) {
          const int row = (tid + BT * m) >> 1;
          const float4 y = *reinterpret_cast<const float4*>(Yb + (int64_t)row * P.D + col);
          const float4 u = *reinterpret_cast<const float4*>(Ub + (int64_t)row * P.D + col);
          const float bq = gates_s[row];
          const float4 rhs = make_float4(
              __fadd_rn(__fmul_rn(P.lamG, y.x), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.x))),
              __fadd_rn(__fmul_rn(P.lamG, y.y), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.y))),
              __fadd_rn(__fmul_rn(P.lamG, y.z), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.z))),
              __fadd_rn(__fmul_rn(P.lamG, y.w), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.w))));
          st.X[m] = u;
          st.R[m] = make_float4(__fadd_rn(u.x, __fmul_rn(P.dt, rhs.x)), __fadd_rn(u.y, __fmul_rn(P.dt, rhs.y)),
                                __fadd_rn(u.z, __fmul_rn(P.dt, rhs.z)), __fadd_rn(u.w, __fmul_rn(P.dt, rhs.w)));
        }
      }
      iters = slab_solve<TPT, KQ>(st, c, P.tol_settle, P.max_iters_settle, p_s, gates_s, nbr_s, w_s, red,
                              sync_b, P.G, act, &res, flag_s);
      if (Uo != nullptr) {
#pragma unroll
        for (int m = 0; m < TPT; ++m)
          if (act[m])
            *reinterpret_cast<float4*>(Uo + (int64_t)((tid + BT * m) >> 1) * P.D + col) = st.X[m];
      }
      if (slab == 0 && tid == 0 && P.stats) {
        P.stats[b * 4 + 0] = (float)iters;
        P.stats[b * 4 + 1] = res;
      }
    }
    // ---------------- stationary: M U* = lamG Y + lamQ b psi^T, x0 = Y
    if (P.do_ustar) {
      SolveCoef c;
      c.lamG = P.lamG; c.lamQ = P.lamQ; c.dt = 0.f;
      c.settle = 0;
      c.kq = kq;
      c.diag0 = P.lamG + P.lamC;
      c.diag1 = P.lamQ;
      c.offc = P.lamC;
#pragma unroll
      for (int m = 0; m < TPT; ++m) {
        st.X[m] = st.R[m] = st.AP[m] = f4_zero();
        if (act[m]) {
          const int row = (tid + BT * m) >> 1;
          const float4 y = *reinterpret_cast<const float4*>(Yb + (int64_t)row * P.D + col);
          const float bq = gates_s[row];
          st.X[m] = y;
          st.R[m] = make_float4(
              __fadd_rn(__fmul_rn(P.lamG, y.x), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.x))),
              __fadd_rn(__fmul_rn(P.lamG, y.y), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.y))),
              __fadd_rn(__fmul_rn(P.lamG, y.z), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.z))),
              __fadd_rn(__fmul_rn(P.lamG, y.w), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.w))));
        }
      }
      iters = slab_solve<TPT, KQ>(st, c, P.tol_ustar, P.max_iters_ustar, p_s, gates_s, nbr_s, w_s, red,
                              sync_b + (P.maxit + 1) * 2, P.G, act, &res, flag_s);
      if (P.Ustar_out != nullptr) {
        float* So = P.Ustar_out + b * P.N * P.D;
#pragma unroll
        for (int m = 0; m < TPT; ++m)
          if (act[m])
            *reinterpret_cast<float4*>(So + (int64_t)((tid + BT * m) >> 1) * P.D + col) = st.X[m];
      }
      if (slab == 0 && tid == 0 && P.stats) {
        P.stats[b * 4 + 2] = (float)iters;
        P.stats[b * 4 + 3] = res;
      }
      // ---------------- deltaH = <U - U*, M (U - U*)>
      if (P.do_dh) {
        __syncthreads();
#pragma unroll
        for (int m = 0; m < TPT; ++m) {
          if (act[m]) {
            // the settled state: just written by this very thread (U_out) or the caller's U
            const float* usrc = P.do_settle ? Uo : Ub;
            const float4 u = *reinterpret_cast<const float4*>(usrc + (int64_t)((tid + BT * m) >> 1) * P.D + col);
            const float4 s = st.X[m];
            p_s[tid + BT * m] = make_float4(__fsub_rn(u.x, s.x), __fsub_rn(u.y, s.y),
                                            __fsub_rn(u.z, s.z), __fsub_rn(u.w, s.w));
          }
        }
        __syncthreads();
        float4 part = f4_zero();
#pragma unroll
        for (int m = 0; m < TPT; ++m) {
          if (act[m]) {
            const int q = tid + BT * m, row = q >> 1;
            const float4 a = apply_task<KQ>(p_s, nbr_s, w_s, row, half, kq,
                                        c.diag0 + c.diag1 * gates_s[row], c.offc);
            part = f4_add(part, f4_mul(p_s[q], a));
          }
        }
        const float4 tot = block_colsum(part, red, lane, warp, half);
        float s4 = (tot.x + tot.y) + (tot.z + tot.w);
        s4 += __shfl_xor_sync(0xffffffffu, s4, 1);
        if (tid == 0) P.dh_part[b * P.G + slab] = (double)s4;
      }
    }
  }
}

__global__ void batched_dh_reduce_kernel(const double* __restrict__ part, int G, int64_t batch,
                                         double* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double s = 0.0;
  for (int g = 0; g < G; ++g) s += part[b * G + g];
  out[b] = s;
}

// ================================================================= host side
// tasks per thread / threads per CTA: 2N float4 tasks spread over at most 1024 threads, keeping
// x, r, Ap (12 registers per task) inside the per-thread register budget
static int tpt_for(int64_t N) {
  const int64_t tasks = 2 * N;
  if (tasks <= 1024) return 1;
  if (tasks <= 2 * 1024) return 2;
  if (tasks <= 3 * 800) return 3;
  if (tasks <= 4 * 640) return 4;
  return 0;
}
static int threads_for(int64_t N, int tpt) {
  int t = (int)((2 * N + tpt - 1) / tpt);
  t = (t + 31) / 32 * 32;
  return t < 64 ? 64 : t;
}

static size_t batched_smem(int64_t N, int k) {
  const int kp = (k + 3) / 4 * 4;
  return (size_t)N * 32 + (size_t)N * kp * 4 + (size_t)N * kp * 2 + 4 * BW_MAX * 2 * 16 + (size_t)N * 4 +
         16;
}

int batched_supported(int64_t N, int D, int k) {
  if (N < 1 || N > 65535 || tpt_for(N) == 0) return 0;
  if (D % 4 != 0 || D < 4) return 0;
  const int G = (D + BC - 1) / BC;
  if (G > sm_count()) return 0;
  if (k < 1 || batched_smem(N, k) > 227 * 1024) return 0;
  return 1;
}

int batched_workspace(int64_t batch, int64_t N, int D, size_t* bytes) {
  (void)N;
  const int G = (D + BC - 1) / BC;
  const int maxit = 256;
  *bytes = align_up((size_t)batch * 2 * (maxit + 1) * 2 * sizeof(unsigned)) +
           align_up((size_t)batch * G * sizeof(double)) + 1024;
  return OSC_OK;
}

int batched_settle(const osc_graph_t* g, const osc_params_t* prm, const osc_batched_args_t* a,
                   void* workspace, size_t ws_bytes, cudaStream_t st) {
  OSC_REQUIRE(g != nullptr && prm != nullptr && a != nullptr, "batched_settle: NULL argument");
  if (!batched_supported(g->N, a->D, g->k))
    return fail(OSC_ERR_UNSUPPORTED, "batched_settle: shape not covered (need N<=1280, D%4==0, D<=8*SMs)");
  if (prm->chain_present) return fail(OSC_ERR_UNSUPPORTED, "batched_settle: chain prior not supported");
  OSC_REQUIRE(a->Y != nullptr && a->psi != nullptr, "batched_settle: Y/psi NULL");
  OSC_REQUIRE(!a->do_deltaH || (a->do_ustar && a->deltaH != nullptr), "deltaH needs do_ustar");
  OSC_REQUIRE(!(a->do_settle && a->do_deltaH) || a->U_out != nullptr, "deltaH after settle needs U_out");
  const int maxit = a->max_iters_settle > a->max_iters_ustar ? a->max_iters_settle : a->max_iters_ustar;
  OSC_REQUIRE(maxit >= 1 && maxit <= 256, "batched_settle: max_iters must be in [1,256]");
  if (g->batch == 0) return OSC_OK;
  size_t need = 0;
  batched_workspace(g->batch, g->N, a->D, &need);
  if (ws_bytes < need) return fail(OSC_ERR_WORKSPACE, "batched_settle: workspace too small");

  BatchedK P;
  P.nbr = g->nbr; P.W = g->W; P.deg = g->deg;
  P.Y = a->Y; P.U_in = a->U_in; P.psi = a->psi; P.gates = a->gates;
  P.U_out = a->U_out; P.Ustar_out = a->Ustar_out; P.stats = a->stats;
  P.batch = g->batch; P.N = g->N; P.k = g->k; P.kp = (g->k + 3) / 4 * 4; P.D = a->D;
  P.G = (a->D + BC - 1) / BC;
  int groups = sm_count() / P.G;
  if ((int64_t)groups > g->batch) groups = (int)g->batch;
  P.groups = groups;
  P.maxit = 256;
  P.do_settle = a->do_settle; P.do_ustar = a->do_ustar; P.do_dh = a->do_deltaH;
  P.lamG = prm->lamG; P.lamC = prm->lamC; P.lamQ = prm->lamQ; P.dt = a->dt;
  P.tol_settle = a->tol_settle; P.tol_ustar = a->tol_ustar;
  P.max_iters_settle = a->max_iters_settle; P.max_iters_ustar = a->max_iters_ustar;
  Arena ar(workspace, ws_bytes);
  const size_t sync_n = (size_t)g->batch * 2 * (P.maxit + 1) * 2;
  P.sync = ar.take<unsigned>(sync_n);
  P.dh_part = ar.take<double>((size_t)g->batch * P.G);
  if (!ar.ok) return fail(OSC_ERR_WORKSPACE, "batched_settle: workspace too small");
  OSC_CUDA(cudaMemsetAsync(P.sync, 0, sync_n * sizeof(unsigned), st));

  const size_t smem = batched_smem(g->N, g->k);
  const int tpt = tpt_for(g->N);
  const int kq = P.kp / 4;
  void* args[] = {&P};
  const dim3 grid(groups * P.G), block(threads_for(g->N, tpt));
  const void* fn = nullptr;
#define OSC_PICK(T)                                                         \
  (kq == 1 ? (const void*)batched_settle_kernel<T, 1>                        \
           : kq == 2 ? (const void*)batched_settle_kernel<T, 2>              \
                     : kq == 4 ? (const void*)batched_settle_kernel<T, 4>    \
                               : (const void*)batched_settle_kernel<T, 0>)
  if (tpt == 1) fn = OSC_PICK(1);
  else if (tpt == 2) fn = OSC_PICK(2);
  else if (tpt == 3) fn = OSC_PICK(3);
  else fn = OSC_PICK(4);
#undef OSC_PICK
  OSC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  OSC_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, st));
  if (a->do_deltaH) {
    batched_dh_reduce_kernel<<<(unsigned)((g->batch + 127) / 128), 128, 0, st>>>(P.dh_part, P.G,
                                                                                  g->batch, a->deltaH);
    OSC_LAUNCH_CHECK("batched_dh_reduce_kernel");
  }
  return OSC_OK;
}

}  // namespace osc
