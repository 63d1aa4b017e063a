// K2: multi-RHS Jacobi-PCG on the ELL graph (single lattice, vectors in HBM).
//
//   lattice.py:171-192 / :245-259  operator, right-hand side, Jacobi diagonal
//   solver.py:15-37                per-column alpha/beta recurrences, max-column stop test
//
// Sparse operator, chi' = chain_present && lamP > 0:
//   M x_i      = (lamG + lamC + chi' lamP + lamQ b_i) x_i - lamC sum_j W_ij x_j - chi' lamP sum_j Wp_ij x_j
//   settle     : Aop = I + dt M ;  Mdiag_i = 1 + dt (lamG + lamQ b_i + chi lamP)
//   stationary : Aop = M        ;  Mdiag_i =          lamG + lamQ b_i + chi lamP
//
// HBM-bound.  Thread layout: blockDim = (CX column groups of VEC floats, RY row lanes); a
// block owns a contiguous row range, every thread keeps fp64 partial sums for its columns,
// partials go to part[block][D] and a second tiny kernel reduces them in a fixed order
// (run-to-run deterministic -- the stop test is a knife edge at large N).
#include <cstdlib>
#include <map>

#include "pcg.cuh"

namespace osc {

struct Coef {
  float lamG, lamC, lamQ, lamP_op, lamP_md, dt;
  int settle, jacobi;
};

static Coef make_coef(const osc_params_t* p, int mode, float dt, int jacobi) {
  Coef c;
  c.lamG = p->lamG;
  c.lamC = p->lamC;
  c.lamQ = p->lamQ;
  c.lamP_op = (p->chain_present && p->lamP > 0.f) ? p->lamP : 0.f;
  c.lamP_md = p->chain_present ? p->lamP : 0.f;
  c.dt = dt;
  c.settle = (mode == OSC_MODE_SETTLE);
  c.jacobi = jacobi;
  return c;
}

__device__ __forceinline__ float op_diag(const Coef& c, float b) {
  const float m = (c.lamG + c.lamC + c.lamP_op) + c.lamQ * b;
  return c.settle ? 1.0f + c.dt * m : m;
}
__device__ __forceinline__ float op_offc(const Coef& c) { return c.settle ? c.dt * c.lamC : c.lamC; }
__device__ __forceinline__ float op_offp(const Coef& c) {
  return c.settle ? c.dt * c.lamP_op : c.lamP_op;
}
// lattice.py:187-192 / :257-259 (evaluated in the reference's order)
__device__ __forceinline__ float md_diag(const Coef& c, float b) {
  const float base = __fadd_rn(__fadd_rn(c.lamG, __fmul_rn(c.lamQ, b)), c.lamP_md);
  return c.settle ? __fadd_rn(1.0f, __fmul_rn(c.dt, base)) : base;
}
__device__ __forceinline__ float precond(const Coef& c, float r, float md) {
  return c.jacobi ? __fdiv_rn(r, md + 1e-12f) : r;
}

template <int VEC>
__device__ __forceinline__ void ldv(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    *p = v[0];
  }
}

struct Dims {
  int64_t N, row0, n_local;
  int D, n_blocks;
};

// reduce per-thread fp64 partials over the RY row lanes and store part[block][cols]
template <int VEC>
__device__ __forceinline__ void flush_partial(double (&acc)[VEC], double* sh, double* part, int D,
                                              int cg, bool valid) {
  const int CX = blockDim.x, RY = blockDim.y;
#pragma unroll
  for (int v = 0; v < VEC; ++v) sh[((size_t)threadIdx.y * CX + threadIdx.x) * VEC + v] = acc[v];
  __syncthreads();
  if (threadIdx.y == 0 && valid) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      double s = 0.0;
      for (int y = 0; y < RY; ++y) s += sh[((size_t)y * CX + threadIdx.x) * VEC + v];
      part[(size_t)blockIdx.x * D + cg * VEC + v] = s;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------- setup: x0 and right-hand side
template <int VEC>
__global__ void pcg_setup_kernel(Dims dm, Coef c, int warm, float w, const float* __restrict__ Y,
                                 const float* __restrict__ U, const float* __restrict__ psi,
                                 const float* __restrict__ gates, float* __restrict__ X,
                                 float* __restrict__ Bv) {
  const int CG = dm.D / VEC;
  const int64_t total = dm.n_local * CG;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / CG;
    const int cg = (int)(e - i * CG);
    const int64_t o = i * dm.D + cg * VEC;
    float y[VEC], u[VEC], q[VEC], x0[VEC], bv[VEC];
    ldv<VEC>(Y + o, y);
    ldv<VEC>(U + o, u);
    ldv<VEC>(psi + cg * VEC, q);
    const float b = gates ? gates[i] : 1.0f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float rhs = __fadd_rn(__fmul_rn(c.lamG, y[v]), __fmul_rn(c.lamQ, __fmul_rn(b, q[v])));
      bv[v] = c.settle ? __fadd_rn(u[v], __fmul_rn(c.dt, rhs)) : rhs;
      if (!c.settle || !warm) x0[v] = y[v];
      else if (w <= 0.f) x0[v] = u[v];
      else x0[v] = __fadd_rn(__fmul_rn(1.0f - w, y[v]), __fmul_rn(w, u[v]));
    }
    stv<VEC>(X + o, x0);
    stv<VEC>(Bv + o, bv);
  }
}

// ---------------------------------------------------------------- SpMM (+dot / +initial residual)
struct GraphView {
  const int32_t* nbr;
  const float* W;
  const int32_t* deg;
  int k;
};
struct ChainView {
  const int32_t* rowptr;
  const int32_t* col;
  const float* Wp;
  const int32_t* slot;  // nullptr => no chain
};

// VecView (pcg.cuh): where the gathered vector lives.
__device__ __forceinline__ const float* row_ptr(const VecView& v, int64_t j, int D) {
  if (v.peers == nullptr) return v.all + j * D;
  const int64_t g = j / v.shard;
  return v.peers[g] + (j - g * v.shard) * D;
}

// RES0 = false: AP = Aop(V); part = sum_i V_i * AP_i
// RES0 = true : R  = Bv - Aop(V) (in place over RBv); P = precond(R); part = sum_i R_i * Z_i
//
// A block owns a contiguous row range and walks it in chunks of SPMM_RCH rows: the chunk's
// neighbour ids / weights / degrees are staged in shared memory with one coalesced pass, so the
// gather loop has no dependent index load in front of every row fetch and can keep 8 row fetches
// per thread in flight (the kernel is latency-bound on random 4*D-byte rows otherwise).
// blockIdx.y selects the 256-column-group panel when D/VEC > 256.
// The chunk is SPMM_RCH rows; very wide ELL rows (dense adjacencies adopted by from_state) shrink it so
// that the staged chunk still fits in shared memory (`rch` argument).
constexpr int SPMM_RCH = 32;

template <int VEC, bool RES0>
__global__ void __launch_bounds__(256)
pcg_spmm_kernel(Dims dm, Coef c, GraphView g, ChainView ch, const float* __restrict__ gates,
                VecView vv, float* __restrict__ out, float* __restrict__ Pout,
                double* __restrict__ part, int rch, const int* __restrict__ done) {
  extern __shared__ double sh[];
  if (done != nullptr && *done != 0) return;  // the solve has stopped (pcg_decide): nothing to do
  const int CG = dm.D / VEC;
  const int nthr = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  // staged per chunk: the ADDRESS of every neighbour row (resolved once per chunk, not per column
  // thread), its weight, the row degrees
  const float** s_nb = reinterpret_cast<const float**>(sh + (size_t)nthr * VEC);
  float* s_w = reinterpret_cast<float*>(s_nb + (size_t)rch * g.k);
  int32_t* s_deg = reinterpret_cast<int32_t*>(s_w + (size_t)rch * g.k);
  const int64_t rpb = (dm.n_local + dm.n_blocks - 1) / dm.n_blocks;
  const int64_t r_beg = (int64_t)blockIdx.x * rpb;
  const int64_t r_end = min(dm.n_local, r_beg + rpb);
  const float offc = op_offc(c), offp = op_offp(c);
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  const bool col_ok = cg < CG;
  const int co = cg * VEC;
  double acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
  for (int64_t c0 = r_beg; c0 < r_end; c0 += rch) {
    const int rows = (int)min((int64_t)rch, r_end - c0);
    __syncthreads();  // the previous chunk's readers are done
    for (int e = tid; e < rows * g.k; e += nthr) {
      const int32_t j = g.nbr[c0 * g.k + e];
      s_nb[e] = j >= 0 ? (vv.local_ids ? vv.all + (int64_t)j * dm.D : row_ptr(vv, j, dm.D)) : nullptr;
      s_w[e] = g.W[c0 * g.k + e];
    }
    for (int e = tid; e < rows; e += nthr) s_deg[e] = g.deg[c0 + e];
    __syncthreads();
    if (!col_ok) continue;
    for (int lr = threadIdx.y; lr < rows; lr += blockDim.y) {
      const int64_t i = c0 + lr;
      const int64_t gi = dm.row0 + i;
      float own[VEC], s[VEC];
      ldv<VEC>((vv.local_ids ? vv.all + i * dm.D : row_ptr(vv, gi, dm.D)) + co, own);
#pragma unroll
      for (int v = 0; v < VEC; ++v) s[v] = 0.f;
      const int n = s_deg[lr];
      const float* const* nb = s_nb + lr * g.k;
      const float* wt = s_w + lr * g.k;
      int t = 0;
      for (; t + 8 <= n; t += 8) {
        float x[8][VEC];
#pragma unroll
        for (int u = 0; u < 8; ++u) ldv<VEC>(nb[t + u] + co, x[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float w = wt[t + u];
#pragma unroll
          for (int v = 0; v < VEC; ++v) s[v] = fmaf(w, x[u][v], s[v]);
        }
      }
      // remainder (1..7 neighbours): ONE masked batch instead of a 4-batch plus single dependent loads --
      // at k = 10 a row with 9 or 10 neighbours used to take two or three DRAM round trips, now two.
      // Masked lanes re-read the row's own segment (just fetched, an L1 hit) with weight 0.
      if (t < n) {
        const float* own_p = (vv.local_ids ? vv.all + i * dm.D : row_ptr(vv, gi, dm.D)) + co;
        if (n - t > 4) {
          float x[8][VEC];
#pragma unroll
          for (int u = 0; u < 8; ++u) ldv<VEC>(t + u < n ? nb[t + u] + co : own_p, x[u]);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float w = t + u < n ? wt[t + u] : 0.f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) s[v] = fmaf(w, x[u][v], s[v]);
          }
        } else {
          float x[4][VEC];
#pragma unroll
          for (int u = 0; u < 4; ++u) ldv<VEC>(t + u < n ? nb[t + u] + co : own_p, x[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float w = t + u < n ? wt[t + u] : 0.f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) s[v] = fmaf(w, x[u][v], s[v]);
          }
        }
      }
      const float b = gates ? gates[i] : 1.0f;
      const float dg = op_diag(c, b);
      float o[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) o[v] = dg * own[v] - offc * s[v];
      if (ch.slot != nullptr && offp != 0.f) {
        const int sl = ch.slot[gi];
        if (sl >= 0) {
          float sp[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) sp[v] = 0.f;
          for (int e = ch.rowptr[sl]; e < ch.rowptr[sl + 1]; ++e) {
            float x[VEC];
            const int32_t cj = ch.col[e];
            ldv<VEC>((vv.local_ids ? vv.all + (int64_t)cj * dm.D : row_ptr(vv, cj, dm.D)) + co, x);
            const float w = ch.Wp[e];
#pragma unroll
            for (int v = 0; v < VEC; ++v) sp[v] = fmaf(w, x[v], sp[v]);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) o[v] -= offp * sp[v];
        }
      }
      if constexpr (RES0) {
        float bv[VEC], r[VEC], z[VEC];
        ldv<VEC>(out + i * dm.D + co, bv);
        const float md = md_diag(c, b);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          r[v] = bv[v] - o[v];
          z[v] = precond(c, r[v], md);
          acc[v] += (double)r[v] * (double)z[v];
        }
        stv<VEC>(out + i * dm.D + co, r);
        stv<VEC>(Pout + i * dm.D + co, z);
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] += (double)own[v] * (double)o[v];
        stv<VEC>(out + i * dm.D + co, o);
      }
    }
  }
  flush_partial<VEC>(acc, sh, part, dm.D, cg, col_ok);
}

// ---------------------------------------------------------------- SpMM, double-buffered staging
// Same arithmetic as pcg_spmm_kernel for a gathered vector in ONE buffer (every case but the fused-P2P
// view).  The graph chunk of the NEXT row block (neighbour ids, weights, degrees: contiguous in global
// memory) is copied into the second shared-memory buffer with cp.async while the current chunk's rows are
// being gathered: one barrier per chunk instead of two, and the staging latency (a dependent global load in
// front of every chunk) disappears behind the gathers.  Row addresses are formed at the point of use
// (one 64-bit multiply-add per fetch) instead of being staged as pointers: 8 instead of 12 bytes per entry.
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(src) : "memory");
}

// RES0: 0 = A p and p.Ap; 1 = first residual, `out` holds the right-hand side on entry; 2 = first residual with
// the right-hand side formed in place from Y, U and psi (the arithmetic of pcg_setup_kernel), gathered vector =
// the start vector itself (Y or U): no setup pass, no x0 / b round trip through HBM (5 vector streams less).
template <int VEC, int RES0>
__global__ void __launch_bounds__(256)
pcg_spmm2_kernel(Dims dm, Coef c, GraphView g, ChainView ch, const float* __restrict__ gates,
                 const float* __restrict__ vec, int local_ids, float* __restrict__ out,
                 float* __restrict__ Pout, double* __restrict__ part, int rch,
                 const int* __restrict__ done, InitSrc src) {
  extern __shared__ double sh[];
  if (done != nullptr && *done != 0) return;
  const int CG = dm.D / VEC;
  const int nthr = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  int32_t* s_id0 = reinterpret_cast<int32_t*>(sh + (size_t)nthr * VEC);
  const size_t per_buf = (size_t)rch * (2 * g.k + 1);  // ids, weights, degrees (4-byte words)
  const int64_t rpb = (dm.n_local + dm.n_blocks - 1) / dm.n_blocks;
  const int64_t r_beg = (int64_t)blockIdx.x * rpb;
  const int64_t r_end = min(dm.n_local, r_beg + rpb);
  const float offc = op_offc(c), offp = op_offp(c);
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  const bool col_ok = cg < CG;
  const int co = cg * VEC;
  auto stage = [&](int buf, int64_t c0, int rows) {
    int32_t* ids = s_id0 + buf * per_buf;
    float* ws = reinterpret_cast<float*>(ids + (size_t)rch * g.k);
    int32_t* dg = reinterpret_cast<int32_t*>(ws + (size_t)rch * g.k);
    const int n = rows * g.k;
    for (int e = tid; e < n; e += nthr) {
      cp_async4(ids + e, g.nbr + c0 * g.k + e);
      cp_async4(ws + e, g.W + c0 * g.k + e);
    }
    for (int e = tid; e < rows; e += nthr) cp_async4(dg + e, g.deg + c0 + e);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  double acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.0;
  float psiv[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) psiv[v] = 0.f;
  if constexpr (RES0 == 2) {
    if (col_ok) ldv<VEC>(src.psi + co, psiv);
  }
  if (r_beg < r_end) stage(0, r_beg, (int)min((int64_t)rch, r_end - r_beg));
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  int buf = 0;
  for (int64_t c0 = r_beg; c0 < r_end; c0 += rch, buf ^= 1) {
    const int rows = (int)min((int64_t)rch, r_end - c0);
    if (c0 + rch < r_end) stage(buf ^ 1, c0 + rch, (int)min((int64_t)rch, r_end - (c0 + rch)));
    const int32_t* ids = s_id0 + buf * per_buf;
    const float* ws = reinterpret_cast<const float*>(ids + (size_t)rch * g.k);
    const int32_t* dgs = reinterpret_cast<const int32_t*>(ws + (size_t)rch * g.k);
    if (col_ok) {
      for (int lr = threadIdx.y; lr < rows; lr += blockDim.y) {
        const int64_t i = c0 + lr;
        const int64_t gi = dm.row0 + i;
        const float* own_p = vec + (local_ids ? i : gi) * dm.D + co;
        float own[VEC], s[VEC], bv[VEC];
        ldv<VEC>(own_p, own);
        if constexpr (RES0 == 1) ldv<VEC>(out + i * dm.D + co, bv);  // the right-hand side travels with the gathers
        if constexpr (RES0 == 2) {
          // a row of Y / U that is not the gathered vector itself is only PREFETCHED here (into L2) and loaded
          // after the gathers: held in registers across them it costs 40 registers and half the resident warps
          if (!src.y_is_x0) prefetch_l2(src.Y + i * dm.D + co);
          if (!src.u_is_x0 && c.settle) prefetch_l2(src.U + i * dm.D + co);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] = 0.f;
        const int n = dgs[lr];
        const int32_t* nb = ids + lr * g.k;
        const float* wt = ws + lr * g.k;
        int t = 0;
        for (; t + 8 <= n; t += 8) {
          float x[8][VEC];
#pragma unroll
          for (int u = 0; u < 8; ++u) ldv<VEC>(vec + (int64_t)nb[t + u] * dm.D + co, x[u]);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float w = wt[t + u];
#pragma unroll
            for (int v = 0; v < VEC; ++v) s[v] = fmaf(w, x[u][v], s[v]);
          }
        }
        if (t < n) {  // masked remainder batch (see pcg_spmm_kernel)
          if (n - t > 4) {
            float x[8][VEC];
#pragma unroll
            for (int u = 0; u < 8; ++u) ldv<VEC>(t + u < n ? vec + (int64_t)nb[t + u] * dm.D + co : own_p, x[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float w = t + u < n ? wt[t + u] : 0.f;
#pragma unroll
              for (int v = 0; v < VEC; ++v) s[v] = fmaf(w, x[u][v], s[v]);
            }
          } else {
            float x[4][VEC];
#pragma unroll
            for (int u = 0; u < 4; ++u) ldv<VEC>(t + u < n ? vec + (int64_t)nb[t + u] * dm.D + co : own_p, x[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float w = t + u < n ? wt[t + u] : 0.f;
#pragma unroll
              for (int v = 0; v < VEC; ++v) s[v] = fmaf(w, x[u][v], s[v]);
            }
          }
        }
        const float b = gates ? gates[i] : 1.0f;
        const float dg = op_diag(c, b);
        float o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) o[v] = dg * own[v] - offc * s[v];
        if (ch.slot != nullptr && offp != 0.f) {
          const int sl = ch.slot[gi];
          if (sl >= 0) {
            float sp[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) sp[v] = 0.f;
            for (int e = ch.rowptr[sl]; e < ch.rowptr[sl + 1]; ++e) {
              float x[VEC];
              ldv<VEC>(vec + (int64_t)ch.col[e] * dm.D + co, x);
              const float w = ch.Wp[e];
#pragma unroll
              for (int v = 0; v < VEC; ++v) sp[v] = fmaf(w, x[v], sp[v]);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) o[v] -= offp * sp[v];
          }
        }
        if constexpr (RES0 == 2) {  // pcg_setup_kernel's right-hand side, same operations in the same order
          float yv[VEC], uv[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) yv[v] = uv[v] = own[v];
          if (!src.y_is_x0) ldv<VEC>(src.Y + i * dm.D + co, yv);
          if (!src.u_is_x0 && c.settle) ldv<VEC>(src.U + i * dm.D + co, uv);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float y = yv[v];
            const float u = uv[v];
            const float rhs = __fadd_rn(__fmul_rn(c.lamG, y), __fmul_rn(c.lamQ, __fmul_rn(b, psiv[v])));
            bv[v] = c.settle ? __fadd_rn(u, __fmul_rn(c.dt, rhs)) : rhs;
          }
        }
        if constexpr (RES0 != 0) {
          float r[VEC], z[VEC];
          const float md = md_diag(c, b);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            r[v] = bv[v] - o[v];
            z[v] = precond(c, r[v], md);
            acc[v] += (double)r[v] * (double)z[v];
          }
          stv<VEC>(out + i * dm.D + co, r);
          stv<VEC>(Pout + i * dm.D + co, z);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += (double)own[v] * (double)o[v];
          stv<VEC>(out + i * dm.D + co, o);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();  // the next chunk has landed; every reader of this buffer is done
  }
  flush_partial<VEC>(acc, sh, part, dm.D, cg, col_ok);
}

// ---------------------------------------------------------------- x, r update + partial rr, rz'
// (256, 3): <= 80 registers, i.e. four 192-thread blocks per SM -- the kernel is bound by the bytes it keeps
// in flight (ncu: 124 registers / 12 warps per SM left it at 5.2 of 6.5 TB/s)
template <int VEC, bool WITH_X>
__global__ void __launch_bounds__(256, 3)
pcg_update_kernel(Dims dm, Coef c, const float* __restrict__ gates, const float* __restrict__ rz,
                  const float* __restrict__ pap, const float* __restrict__ P,
                  const float* __restrict__ AP, float* __restrict__ X, float* __restrict__ R,
                  double* __restrict__ part_rr, double* __restrict__ part_rz,
                  const int* __restrict__ done) {
  extern __shared__ double sh[];
  if (done != nullptr && *done != 0) return;
  const int CG = dm.D / VEC;
  const int64_t rpb = (dm.n_local + dm.n_blocks - 1) / dm.n_blocks;
  const int64_t r_beg = (int64_t)blockIdx.x * rpb;
  const int64_t r_end = min(dm.n_local, r_beg + rpb);
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  const bool col_ok = cg < CG;
  double arr[VEC], arz[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) arr[v] = arz[v] = 0.0;
  if (col_ok) {
    const int co = cg * VEC;
    float alpha[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) alpha[v] = __fdiv_rn(rz[co + v], pap[co + v] + 1e-18f);
    int64_t i = r_beg + threadIdx.y;
    if constexpr (!WITH_X) {
      // r only: two loads per row do not cover the HBM latency -- two rows per trip (four spill at 80 registers)
      constexpr int U = 2;
      for (; i + (U - 1) * (int64_t)blockDim.y < r_end; i += U * (int64_t)blockDim.y) {
        float ap[U][VEC], r[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t o = (i + u * (int64_t)blockDim.y) * dm.D + co;
          ldv<VEC>(AP + o, ap[u]);
          ldv<VEC>(R + o, r[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t iu = i + u * (int64_t)blockDim.y;
          const float md = md_diag(c, gates ? gates[iu] : 1.0f);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            r[u][v] = __fsub_rn(r[u][v], __fmul_rn(ap[u][v], alpha[v]));
            const float z = precond(c, r[u][v], md);
            arr[v] += (double)r[u][v] * (double)r[u][v];
            arz[v] += (double)r[u][v] * (double)z;
          }
          stv<VEC>(R + iu * dm.D + co, r[u]);
        }
      }
    }
    for (; i < r_end; i += blockDim.y) {
      const int64_t o = i * dm.D + co;
      float ap[VEC], r[VEC];
      ldv<VEC>(AP + o, ap);
      ldv<VEC>(R + o, r);
      if constexpr (WITH_X) {  // otherwise the x update rides with the p update (pcg_pupdate_x_kernel)
        float p[VEC], x[VEC];
        ldv<VEC>(P + o, p);
        ldv<VEC>(X + o, x);
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[v] = __fadd_rn(x[v], __fmul_rn(p[v], alpha[v]));
        stv<VEC>(X + o, x);
      }
      const float md = md_diag(c, gates ? gates[i] : 1.0f);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        r[v] = __fsub_rn(r[v], __fmul_rn(ap[v], alpha[v]));
        const float z = precond(c, r[v], md);
        arr[v] += (double)r[v] * (double)r[v];
        arz[v] += (double)r[v] * (double)z;
      }
      stv<VEC>(R + o, r);
    }
  }
  flush_partial<VEC>(arr, sh, part_rr, dm.D, cg, col_ok);
  flush_partial<VEC>(arz, sh, part_rz, dm.D, cg, col_ok);
}

// ---------------------------------------------------------------- p = z + beta p
template <int VEC>
__global__ void pcg_pupdate_kernel(Dims dm, Coef c, const float* __restrict__ gates,
                                   const float* __restrict__ rz_new,
                                   const float* __restrict__ rz_old, const float* __restrict__ R,
                                   float* __restrict__ P, const int* __restrict__ done) {
  if (done != nullptr && *done != 0) return;
  const int CG = dm.D / VEC;
  const int64_t total = dm.n_local * CG;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / CG;
    const int cg = (int)(e - i * CG);
    const int co = cg * VEC;
    const int64_t o = i * dm.D + co;
    float r[VEC], p[VEC];
    ldv<VEC>(R + o, r);
    ldv<VEC>(P + o, p);
    const float md = md_diag(c, gates ? gates[i] : 1.0f);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float beta = __fdiv_rn(rz_new[co + v], rz_old[co + v] + 1e-18f);
      p[v] = __fadd_rn(precond(c, r[v], md), __fmul_rn(p[v], beta));
    }
    stv<VEC>(P + o, p);
  }
}

// x += alpha p and p = z + beta p in ONE pass over p (solver.py:25,33-35): the x update of an iteration does
// not feed its stop test, so it can wait for the pass that rewrites p anyway -- 8 instead of 9 vector streams
// per iteration next to the SpMM.  Same operations on the same operands, so x is bit-identical.  The kernel
// runs after the iteration's verdict: finished in THIS iteration -> x only; finished earlier -> nothing.
template <int VEC>
__global__ void pcg_pupdate_x_kernel(Dims dm, Coef c, const float* __restrict__ gates,
                                     const float* __restrict__ rz_new, const float* __restrict__ rz_old,
                                     const float* __restrict__ pap, const float* __restrict__ R,
                                     float* __restrict__ P, float* __restrict__ X, const PcgCtl* __restrict__ ctl,
                                     int it, int x_only, const float* __restrict__ Xsrc) {
  bool fin = x_only != 0;  // ctl == NULL: the host knows whether this was the last iteration
  if (ctl != nullptr) {
    fin = ctl->done != 0;
    if (fin && ctl->iters != it) return;
  }
  const int CG = dm.D / VEC;
  const int64_t total = dm.n_local * CG;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 2;  // two chunks per trip: six loads in flight per thread (the kernel is latency bound)
  for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < total; e0 += U * stride) {
    float p[U][VEC], x[U][VEC], r[U][VEC];
    int64_t row[U], off[U];
    int co[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t e = e0 + u * stride;
      ok[u] = e < total;
      row[u] = ok[u] ? e / CG : 0;
      co[u] = ok[u] ? (int)(e - row[u] * CG) * VEC : 0;
      off[u] = row[u] * dm.D + co[u];
      if (ok[u]) {
        ldv<VEC>(P + off[u], p[u]);
        ldv<VEC>(Xsrc + off[u], x[u]);  // X itself, or the start vector in the first iteration of a fused start
        if (!fin) ldv<VEC>(R + off[u], r[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float alpha = __fdiv_rn(rz_old[co[u] + v], pap[co[u] + v] + 1e-18f);
        x[u][v] = __fadd_rn(x[u][v], __fmul_rn(p[u][v], alpha));
      }
      stv<VEC>(X + off[u], x[u]);
      if (fin) continue;
      const float md = md_diag(c, gates ? gates[row[u]] : 1.0f);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float beta = __fdiv_rn(rz_new[co[u] + v], rz_old[co[u] + v] + 1e-18f);
        p[u][v] = __fadd_rn(precond(c, r[u][v], md), __fmul_rn(p[u][v], beta));
      }
      stv<VEC>(P + off[u], p[u]);
    }
  }
}

// diff = U - Ustar (receipts.py:21)
__global__ void diff_kernel(const float* __restrict__ a, const float* __restrict__ b,
                            float* __restrict__ out, int64_t n) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x)
    out[e] = __fsub_rn(a[e], b[e]);
}

// ---------------------------------------------------------------- column reduction of partials
// out[c] = sum_b part[b][c] in a FIXED order (32 row lanes stride the row blocks, then a fixed
// 32-term sum): run-to-run deterministic, which the knife-edge stop test needs.
// One block per 32 columns; the lattice-wide max is merged with an order-independent atomicMax
// on the bit pattern (values are >= 0; *d_max is zeroed by the host wrapper beforehand).
__global__ void __launch_bounds__(1024)
pcg_reduce_kernel(const double* __restrict__ part, int n_blocks, int D, float* __restrict__ out,
                  float* __restrict__ d_max, double* __restrict__ out64, const int* __restrict__ done) {
  __shared__ double ssum[32][33];
  if (done != nullptr && *done != 0) return;
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int cidx = blockIdx.x * 32 + cx;
  double s = 0.0;
  if (cidx < D)
    for (int b = ry; b < n_blocks; b += 32) s += part[(size_t)b * D + cidx];
  ssum[ry][cx] = s;
  __syncthreads();
  if (ry != 0) return;
  double t = 0.0;
#pragma unroll
  for (int y = 0; y < 32; ++y) t += ssum[y][cx];
  float mx = 0.f;
  if (cidx < D) {
    const float f = (float)t;
    out[cidx] = f;
    if (out64 != nullptr) out64[cidx] = t;
    mx = __fsqrt_rn(fmaxf(f, 0.f));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (cx == 0 && d_max != nullptr) atomicMax(reinterpret_cast<int*>(d_max), __float_as_int(mx));
}

// fixed-order total of D doubles (deltaH)
__global__ void __launch_bounds__(1024) sum_doubles_kernel(const double* __restrict__ v, int D,
                                                          double* __restrict__ total) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (int i = threadIdx.x; i < D; i += blockDim.x) s += v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = sh[0];
}

// ================================================================= host side
static void block_shape(int D, int& vec, dim3& blk) {
  vec = (D % 4 == 0) ? 4 : 1;
  const int CG = D / vec;
  const int cx = CG < 256 ? CG : 256;
  int ry = 256 / cx;
  if (ry < 1) ry = 1;
  blk = dim3(cx, ry, 1);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static Dims to_dims(const osc_pcg_dims_t* d) {
  Dims r;
  r.N = d->N;
  r.row0 = d->row0;
  r.n_local = d->n_local;
  r.D = d->D;
  r.n_blocks = d->n_blocks;
  return r;
}

static GraphView gview(const osc_graph_t* g) { return GraphView{g->nbr, g->W, g->deg, g->k}; }
static ChainView cview(const osc_chain_t* c) {
  if (c == nullptr || c->n_rows == 0) return ChainView{nullptr, nullptr, nullptr, nullptr};
  return ChainView{c->rowptr, c->col, c->Wp, c->slot};
}

static int ew_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int pcg_plan(osc_pcg_dims_t* d, size_t* ws) {
  OSC_REQUIRE(d != nullptr, "dims is NULL");
  OSC_REQUIRE(d->D >= 1 && d->N >= 0 && d->n_local >= 0 && d->row0 >= 0, "bad dims");
  int vec;
  dim3 blk;
  block_shape(d->D, vec, blk);
  int64_t nb = (d->n_local + blk.y - 1) / blk.y;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  d->n_blocks = (int)nb;
  if (ws != nullptr) {
    const size_t vecb = align_up((size_t)d->n_local * d->D * sizeof(float));
    const size_t partb = align_up((size_t)nb * d->D * sizeof(double));
    const size_t colb = align_up((size_t)d->D * sizeof(float));
    *ws = 3 * vecb + 3 * partb + 5 * colb + 1024;
  }
  return OSC_OK;
}

#define OSC_VEC_DISPATCH(vec, ...)   \
  if ((vec) == 4) {                  \
    constexpr int VEC = 4;           \
    __VA_ARGS__                      \
  } else {                           \
    constexpr int VEC = 1;           \
    __VA_ARGS__                      \
  }

int pcg_setup(const osc_pcg_dims_t* d, const osc_params_t* prm, int mode, float dt, int warm,
              float inertia, const float* Y, const float* U, const float* psi, const float* gates,
              float* X, float* Bv, cudaStream_t st) {
  if (d->n_local == 0) return OSC_OK;
  int vec;
  dim3 blk;
  block_shape(d->D, vec, blk);
  if (!(aligned16(Y) && aligned16(U) && aligned16(psi) && aligned16(X) && aligned16(Bv))) vec = 1;
  Coef c = make_coef(prm, mode, dt, 1);
  float w = inertia < 0.f ? 0.f : (inertia > 1.f ? 1.f : inertia);
  const int64_t total = d->n_local * (d->D / vec);
  OSC_VEC_DISPATCH(vec, pcg_setup_kernel<VEC><<<ew_grid(total), 256, 0, st>>>(
                            to_dims(d), c, warm, w, Y, U, psi, gates, X, Bv);)
  OSC_LAUNCH_CHECK("pcg_setup_kernel");
  return OSC_OK;
}

// dynamic shared memory of pcg_spmm_kernel: fp64 partials + one staged graph chunk of `rch` rows
static constexpr size_t kSpmmSmemMax = 227 * 1024;
static size_t spmm_smem_bytes(const dim3& blk, int vec, int k, int rch) {
  return (size_t)blk.x * blk.y * vec * sizeof(double) + (size_t)rch * ((size_t)k * 12 + 4);
}
// rows per staged chunk: SPMM_RCH when that fits the 48 KB default, else the largest power of two that
// does; ELL rows too wide even for one row per chunk use the opt-in limit.  0 = does not fit at all.
static int spmm_chunk_rows(const dim3& blk, int vec, int k) {
  // Narrow rows (column slabs of a sharded lattice: D/G floats) put many row lanes in a block: a chunk is
  // then a whole number of rows per lane and at least 4 of them, so that the two barriers per chunk are
  // amortised and no lane idles in the last pass (D = 48: 21 lanes, 84 rows per chunk instead of 32, where
  // the second pass ran 11 of 21 lanes; measured 5.6 ms -> see DESIGN.md at N = 10M).
  int want = SPMM_RCH;
  const int ry = (int)blk.y;
  if (ry > 2) want = ry * (ry > 8 ? 4 : (SPMM_RCH + ry - 1) / ry);
  if (spmm_smem_bytes(blk, vec, k, want) <= 48 * 1024) return want;
  for (int rch = SPMM_RCH; rch >= 1; rch >>= 1)
    if (spmm_smem_bytes(blk, vec, k, rch) <= 48 * 1024) return rch;
  return spmm_smem_bytes(blk, vec, k, 1) <= kSpmmSmemMax ? 1 : 0;
}

// largest ELL width the staged SpMM can launch for D columns (graph loaders check it up front)
int pcg_max_ell_width(int D) {
  int vec;
  dim3 blk;
  block_shape(D, vec, blk);
  const size_t fixed = (size_t)blk.x * blk.y * vec * sizeof(double) + 4;
  return (int)((kSpmmSmemMax - fixed) / 12);
}

// does the double-buffered kernel (one buffer holds the gathered vector) serve this launch?
static bool spmm_two_buffer(const dim3& blk, int vec, int k, int rch, const VecView& vv, size_t* smem2) {
  bool two = vv.peers == nullptr && vv.all != nullptr && rch >= 2;
  static const int off = [] { const char* e = getenv("OSC_SPMM2"); return (e && atoi(e) == 0) ? 1 : 0; }();
  if (off) two = false;  // dev-only A/B switch
  // same bytes as the pointer staging: 2 buffers x rch x (8 k + 4) <= rch x (12 k + 4) + ... for k >= 1
  *smem2 = (size_t)blk.x * blk.y * vec * sizeof(double) + 2 * (size_t)rch * ((size_t)k * 8 + 4);
  return two && *smem2 <= kSpmmSmemMax;
}

// the fused first residual (InitSrc) exists in the double-buffered kernel only
bool pcg_fused_init_ok(const osc_pcg_dims_t* d, const osc_graph_t* g) {
  const char* e = getenv("OSC_PCG_FUSE_INIT");  // dev-only A/B switch (read per call: tests flip it)
  if ((e && atoi(e) == 0) || d->n_local == 0) return false;
  int vec;
  dim3 blk;
  block_shape(d->D, vec, blk);
  const int rch = spmm_chunk_rows(blk, vec, g->k);
  size_t smem2 = 0;
  const float dummy = 0.f;
  return rch != 0 && spmm_two_buffer(blk, vec, g->k, rch, VecView{&dummy, nullptr, 0, 0}, &smem2);
}

int spmm_launch(bool res0, const osc_pcg_dims_t* d, const osc_graph_t* g, const osc_chain_t* chain,
                const osc_params_t* prm, int mode, float dt, int jacobi, const float* gates, VecView vv,
                float* out, float* Pout, double* part, cudaStream_t st, const int* done, const InitSrc* init) {
  if (d->n_local == 0) return OSC_OK;
  int vec;
  dim3 blk;
  block_shape(d->D, vec, blk);
  if (!((vv.all == nullptr || aligned16(vv.all)) && aligned16(out) && (Pout == nullptr || aligned16(Pout)) &&
        (init == nullptr || (aligned16(init->Y) && aligned16(init->U) && aligned16(init->psi))))) {
    if (vec == 4) return fail(OSC_ERR_INVALID, "pcg: vectors must be 16-byte aligned when D % 4 == 0");
  }
  Coef c = make_coef(prm, mode, dt, jacobi);
  const int rch = spmm_chunk_rows(blk, vec, g->k);
  if (rch == 0) return fail(OSC_ERR_UNSUPPORTED, "pcg: ELL width too large for the staged SpMM");
  const size_t smem = spmm_smem_bytes(blk, vec, g->k, rch);
  const dim3 grid((unsigned)d->n_blocks, (unsigned)((d->D / vec + (int)blk.x - 1) / (int)blk.x), 1);
  size_t smem2 = 0;
  const bool two = spmm_two_buffer(blk, vec, g->k, rch, vv, &smem2);
  if (init != nullptr && !(two && res0))
    return fail(OSC_ERR_UNSUPPORTED, "pcg: fused first residual needs the double-buffered SpMM (pcg_fused_init_ok)");
  if (two) {
    const InitSrc none{nullptr, nullptr, nullptr, 0, 0};
#define OSC_SPMM2_LAUNCH(R0, SRC)                                                                         \
  OSC_VEC_DISPATCH(vec, {                                                                                 \
    if (smem2 > 48 * 1024)                                                                                \
      OSC_CUDA(cudaFuncSetAttribute((const void*)pcg_spmm2_kernel<VEC, R0>,                               \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmmSmemMax));    \
    pcg_spmm2_kernel<VEC, R0><<<grid, blk, smem2, st>>>(to_dims(d), c, gview(g), cview(chain), gates,     \
                                                       vv.all, vv.local_ids, out, Pout, part, rch, done, \
                                                       SRC);                                              \
  })
    if (init != nullptr) {
      OSC_SPMM2_LAUNCH(2, *init)
    } else if (res0) {
      OSC_SPMM2_LAUNCH(1, none)
    } else {
      OSC_SPMM2_LAUNCH(0, none)
    }
#undef OSC_SPMM2_LAUNCH
    OSC_LAUNCH_CHECK("pcg_spmm2_kernel");
    return OSC_OK;
  }
#define OSC_SPMM_LAUNCH(R0)                                                                               \
  OSC_VEC_DISPATCH(vec, {                                                                                 \
    if (smem > 48 * 1024)                                                                                 \
      OSC_CUDA(cudaFuncSetAttribute((const void*)pcg_spmm_kernel<VEC, R0>,                                \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpmmSmemMax));    \
    pcg_spmm_kernel<VEC, R0><<<grid, blk, smem, st>>>(to_dims(d), c, gview(g), cview(chain), gates, vv,  \
                                                      out, Pout, part, rch, done);                        \
  })
  if (res0) {
    OSC_SPMM_LAUNCH(true)
  } else {
    OSC_SPMM_LAUNCH(false)
  }
#undef OSC_SPMM_LAUNCH
  OSC_LAUNCH_CHECK("pcg_spmm_kernel");
  return OSC_OK;
}

int pcg_residual0(const osc_pcg_dims_t* d, const osc_graph_t* g, const osc_chain_t* chain,
                  const osc_params_t* prm, int mode, float dt, int jacobi, const float* gates,
                  const float* Xall, float* RBv, float* P, double* part_rz, cudaStream_t st) {
  return spmm_launch(true, d, g, chain, prm, mode, dt, jacobi, gates, VecView{Xall, nullptr, 0, 0}, RBv, P,
                     part_rz, st);
}

int pcg_residual0_p2p(const osc_pcg_dims_t* d, const osc_graph_t* g, const osc_chain_t* chain,
                      const osc_params_t* prm, int mode, float dt, int jacobi, const float* gates,
                      const float* const* peers, int64_t shard, float* RBv, float* P, double* part_rz,
                      cudaStream_t st) {
  return spmm_launch(true, d, g, chain, prm, mode, dt, jacobi, gates, VecView{nullptr, peers, shard, 0}, RBv, P,
                     part_rz, st);
}

int pcg_spmm_dot(const osc_pcg_dims_t* d, const osc_graph_t* g, const osc_chain_t* chain,
                 const osc_params_t* prm, int mode, float dt, const float* gates, const float* Pall,
                 float* AP, double* part_pap, cudaStream_t st, const int* done) {
  return spmm_launch(false, d, g, chain, prm, mode, dt, 1, gates, VecView{Pall, nullptr, 0, 0}, AP, nullptr,
                     part_pap, st, done);
}

int pcg_spmm_dot_p2p(const osc_pcg_dims_t* d, const osc_graph_t* g, const osc_chain_t* chain,
                     const osc_params_t* prm, int mode, float dt, const float* gates,
                     const float* const* peers, int64_t shard, float* AP, double* part_pap, cudaStream_t st) {
  return spmm_launch(false, d, g, chain, prm, mode, dt, 1, gates, VecView{nullptr, peers, shard, 0}, AP, nullptr,
                     part_pap, st);
}

int pcg_reduce(const double* part, int n_blocks, int D, float* out, float* d_max, double* out64,
               cudaStream_t st, const int* done) {
  if (d_max != nullptr) OSC_CUDA(cudaMemsetAsync(d_max, 0, sizeof(float), st));
  pcg_reduce_kernel<<<(D + 31) / 32, dim3(32, 32, 1), 0, st>>>(part, n_blocks, D, out, d_max, out64, done);
  OSC_LAUNCH_CHECK("pcg_reduce_kernel");
  return OSC_OK;
}

int pcg_update(const osc_pcg_dims_t* d, const osc_params_t* prm, int mode, float dt, int jacobi,
               const float* gates, const float* rz, const float* pap, const float* P,
               const float* AP, float* X, float* R, double* part_rr, double* part_rz,
               cudaStream_t st, const int* done) {
  if (d->n_local == 0) return OSC_OK;
  int vec;
  dim3 blk;
  block_shape(d->D, vec, blk);
  Coef c = make_coef(prm, mode, dt, jacobi);
  const size_t smem = (size_t)blk.x * blk.y * vec * sizeof(double);
  const dim3 grid((unsigned)d->n_blocks, (unsigned)((d->D / vec + (int)blk.x - 1) / (int)blk.x), 1);
  if (X != nullptr) {
    OSC_VEC_DISPATCH(vec, pcg_update_kernel<VEC, true><<<grid, blk, smem, st>>>(
                              to_dims(d), c, gates, rz, pap, P, AP, X, R, part_rr, part_rz, done);)
  } else {
    OSC_VEC_DISPATCH(vec, pcg_update_kernel<VEC, false><<<grid, blk, smem, st>>>(
                              to_dims(d), c, gates, rz, pap, P, AP, X, R, part_rr, part_rz, done);)
  }
  OSC_LAUNCH_CHECK("pcg_update_kernel");
  return OSC_OK;
}

int pcg_pupdate(const osc_pcg_dims_t* d, const osc_params_t* prm, int mode, float dt, int jacobi,
                const float* gates, const float* rz_new, const float* rz_old, const float* R,
                float* P, cudaStream_t st, const int* done) {
  if (d->n_local == 0) return OSC_OK;
  int vec;
  dim3 blk;
  block_shape(d->D, vec, blk);
  Coef c = make_coef(prm, mode, dt, jacobi);
  const int64_t total = d->n_local * (d->D / vec);
  OSC_VEC_DISPATCH(vec, pcg_pupdate_kernel<VEC><<<ew_grid(total), 256, 0, st>>>(
                            to_dims(d), c, gates, rz_new, rz_old, R, P, done);)
  OSC_LAUNCH_CHECK("pcg_pupdate_kernel");
  return OSC_OK;
}

int pcg_pupdate_x(const osc_pcg_dims_t* d, const osc_params_t* prm, int mode, float dt, int jacobi,
                  const float* gates, const float* rz_new, const float* rz_old, const float* pap, const float* R,
                  float* P, float* X, const PcgCtl* ctl, int it, int x_only, cudaStream_t st, const float* Xsrc) {
  if (d->n_local == 0) return OSC_OK;
  int vec;
  dim3 blk;
  block_shape(d->D, vec, blk);
  Coef c = make_coef(prm, mode, dt, jacobi);
  const int64_t total = d->n_local * (d->D / vec);
  OSC_VEC_DISPATCH(vec, pcg_pupdate_x_kernel<VEC><<<ew_grid(total), 256, 0, st>>>(
                            to_dims(d), c, gates, rz_new, rz_old, pap, R, P, X, ctl, it, x_only,
                            Xsrc != nullptr ? Xsrc : X);)
  OSC_LAUNCH_CHECK("pcg_pupdate_x_kernel");
  return OSC_OK;
}

// ---------------------------------------------------------------- stop test on the device
__global__ void __launch_bounds__(256) pcg_decide_kernel(PcgCtl* ctl, const float* __restrict__ rr,
                                                         const float* __restrict__ d_res, int D, double tol,
                                                         int it, int max_iters) {
  if (ctl->done != 0) return;
  __shared__ float smax[256];
  float mx = 0.f;
  if (rr != nullptr) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) mx = fmaxf(mx, __fsqrt_rn(fmaxf(rr[c], 0.f)));
  } else if (threadIdx.x == 0) {
    mx = *d_res;
  }
  smax[threadIdx.x] = mx;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float res = smax[0];
    ctl->res = res;
    ctl->iters = it;
    if ((double)res <= tol || it >= max_iters) ctl->done = 1;  // solver.py:29-31
  }
}

int pcg_decide(PcgCtl* ctl, const float* rr, const float* d_res, int D, double tol, int it, int max_iters,
               cudaStream_t st) {
  pcg_decide_kernel<<<1, 256, 0, st>>>(ctl, rr, d_res, D, tol, it, max_iters);
  OSC_LAUNCH_CHECK("pcg_decide_kernel");
  return OSC_OK;
}

int launch_diff(const float* a, const float* b, float* out, int64_t n, cudaStream_t st) {
  if (n == 0) return OSC_OK;
  diff_kernel<<<ew_grid(n), 256, 0, st>>>(a, b, out, n);
  OSC_LAUNCH_CHECK("diff_kernel");
  return OSC_OK;
}

int launch_sum_doubles(const double* v, int D, double* total, cudaStream_t st) {
  sum_doubles_kernel<<<1, 1024, 0, st>>>(v, D, total);
  OSC_LAUNCH_CHECK("sum_doubles_kernel");
  return OSC_OK;
}

// dev switch: OSC_PCG_FUSE_X=0 runs x += alpha p inside the r update again (A/B of the deferred x update)
bool pcg_fuse_x() {
  const char* e = getenv("OSC_PCG_FUSE_X");
  return !(e != nullptr && atoi(e) == 0);
}

// ---------------------------------------------------------------- lagged host poll of the control block
int CtlPoll::init() {
  if (h != nullptr) return OSC_OK;
  OSC_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h), 2 * sizeof(PcgCtl), cudaHostAllocDefault));
  for (int i = 0; i < 2; ++i) OSC_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  return OSC_OK;
}
int CtlPoll::record(int it, const PcgCtl* d_ctl, cudaStream_t st) {
  OSC_CUDA(cudaMemcpyAsync(&h[it & 1], d_ctl, sizeof(PcgCtl), cudaMemcpyDeviceToHost, st));
  OSC_CUDA(cudaEventRecord(ev[it & 1], st));
  return OSC_OK;
}
int CtlPoll::wait(int it, PcgCtl* out) {
  OSC_CUDA(cudaEventSynchronize(ev[it & 1]));
  *out = h[it & 1];
  return OSC_OK;
}
CtlPoll* ctl_poll() {
  // one set of pinned slots + events per host thread and device; never freed (a few bytes per thread)
  static thread_local std::map<int, CtlPoll> polls;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("cudaGetDevice failed");
    return nullptr;
  }
  CtlPoll& p = polls[dev];
  if (p.init() != OSC_OK) return nullptr;
  return &p;
}

// The recurrences of solver.py:19-37 from a caller-supplied start: on entry X = x0 and R = right-hand
// side; on exit X = last iterate, R = recurrence residual.  Workspace: P, AP, partials, column sums.
// The stop test runs on the device (pcg_decide) and the host enqueues ONE iteration ahead of it: the
// kernels of an iteration that follows the stop return at once, so the result is the reference's and
// the stream never drains between iterations.
// `init` (optional): fused start -- X and R are NOT initialised by the caller; x0 = init->x0, the right-hand
// side is formed inside the first residual, and X is first written by the x update of iteration 1.
static int pcg_core(const osc_pcg_dims_t& d, const osc_graph_t* g, const osc_chain_t* chain,
                    const osc_params_t* prm, int mode, float dt, int jacobi, double tol, int max_iters,
                    const float* gates, float* X, float* R, Arena& ar, int* h_iters, float* h_res,
                    cudaStream_t st, const InitSrc* init = nullptr, const float* x0 = nullptr) {
  const int D = d.D;
  const size_t nd = (size_t)g->N * D;
  float* P = ar.take<float>(nd);
  float* AP = ar.take<float>(nd);
  double* part_a = ar.take<double>((size_t)d.n_blocks * D);
  double* part_b = ar.take<double>((size_t)d.n_blocks * D);
  float* rz = ar.take<float>(D);
  float* rz_new = ar.take<float>(D);
  float* pap = ar.take<float>(D);
  float* rr = ar.take<float>(D);
  float* d_res = ar.take<float>(32);
  PcgCtl* ctl = ar.take<PcgCtl>(1);
  if (!ar.ok) return fail(OSC_ERR_WORKSPACE, "pcg: workspace too small");
  CtlPoll* poll = ctl_poll();
  if (poll == nullptr) return OSC_ERR_CUDA;
  const int* done = &ctl->done;
  const bool fuse_x = pcg_fuse_x();
  int rc;
  OSC_CUDA(cudaMemsetAsync(ctl, 0, sizeof(PcgCtl), st));
  if (init != nullptr) {
    if ((rc = spmm_launch(true, &d, g, chain, prm, mode, dt, jacobi, gates, VecView{x0, nullptr, 0, 0}, R, P, part_a,
                          st, nullptr, init)))
      return rc;
  } else if ((rc = pcg_residual0(&d, g, chain, prm, mode, dt, jacobi, gates, X, R, P, part_a, st))) {
    return rc;
  }
  if ((rc = pcg_reduce(part_a, d.n_blocks, D, rz, nullptr, nullptr, st))) return rc;
  PcgCtl h{0, 0, __builtin_nanf(""), 0};
  int it = 0;
  for (it = 1; it <= max_iters; ++it) {
    if ((rc = pcg_spmm_dot(&d, g, chain, prm, mode, dt, gates, P, AP, part_a, st, done))) return rc;
    if ((rc = pcg_reduce(part_a, d.n_blocks, D, pap, nullptr, nullptr, st, done))) return rc;
    if ((rc = pcg_update(&d, prm, mode, dt, jacobi, gates, rz, pap, P, AP, fuse_x ? nullptr : X, R, part_a, part_b,
                         st, done)))
      return rc;
    if ((rc = pcg_reduce(part_a, d.n_blocks, D, rr, d_res, nullptr, st, done))) return rc;
    if ((rc = pcg_reduce(part_b, d.n_blocks, D, rz_new, nullptr, nullptr, st, done))) return rc;
    if ((rc = pcg_decide(ctl, nullptr, d_res, D, tol, it, max_iters, st))) return rc;
    if ((rc = poll->record(it, ctl, st))) return rc;
    // x += alpha p always belongs to this iteration; p = z + beta p only if it was not the last
    if (fuse_x && (rc = pcg_pupdate_x(&d, prm, mode, dt, jacobi, gates, rz_new, rz, pap, R, P, X, ctl, it, 0, st,
                                      (init != nullptr && it == 1) ? x0 : nullptr)))
      return rc;
    if (it > 1) {  // the previous iteration's verdict has landed (or lands while this one runs)
      if ((rc = poll->wait(it - 1, &h))) return rc;
      if (h.done) break;
    }
    if (it == max_iters) break;
    if (!fuse_x && (rc = pcg_pupdate(&d, prm, mode, dt, jacobi, gates, rz_new, rz, R, P, st, done))) return rc;
    float* t = rz;
    rz = rz_new;
    rz_new = t;
  }
  if (it > max_iters) it = max_iters;
  if (!h.done && (rc = poll->wait(it, &h))) return rc;  // the last enqueued iteration decides
  if (h_iters) *h_iters = h.iters;
  if (h_res) *h_res = h.res;
  return OSC_OK;
}

int pcg_solve(const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm, int mode,
              float dt, int warm, float inertia, int jacobi, double tol, int max_iters,
              const float* Y, const float* U, const float* psi, const float* gates, int D, float* X,
              int* h_iters, float* h_res, void* workspace, size_t ws_bytes, cudaStream_t st) {
  OSC_REQUIRE(g != nullptr && prm != nullptr && Y != nullptr && X != nullptr && psi != nullptr,
              "pcg_solve: NULL argument");
  OSC_REQUIRE(g->batch == 1, "pcg_solve handles one lattice (use osc_batched_settle)");
  if (U == nullptr) U = Y;
  osc_pcg_dims_t d{g->N, 0, g->N, D, 0};
  size_t need = 0;
  int rc = pcg_plan(&d, &need);
  if (rc) return rc;
  if (ws_bytes < need) return fail(OSC_ERR_WORKSPACE, "pcg_solve: workspace too small");
  if (h_iters) *h_iters = 0;
  if (h_res) *h_res = __builtin_nanf("");
  if (g->N == 0) return OSC_OK;
  Arena ar(workspace, ws_bytes);
  float* R = ar.take<float>((size_t)g->N * D);
  if (!ar.ok) return fail(OSC_ERR_WORKSPACE, "pcg_solve: workspace too small");
  // start vector of lattice.py:751-758: Y (stationary solve, cold start), U (warm, no inertia), else a mix
  const bool settle = mode == OSC_MODE_SETTLE;
  const bool mixed = settle && warm && inertia > 0.f;
  const float* x0 = (!settle || !warm) ? Y : U;
  if (max_iters >= 1 && !mixed && pcg_fuse_x() && pcg_fused_init_ok(&d, g) && x0 != X) {
    // no setup pass: the first residual forms the right-hand side itself and gathers the start vector in place
    const InitSrc init{Y, U, psi, x0 == Y ? 1 : 0, x0 == U ? 1 : 0};
    return pcg_core(d, g, chain, prm, mode, dt, jacobi, tol, max_iters, gates, X, R, ar, h_iters, h_res, st, &init,
                    x0);
  }
  if ((rc = pcg_setup(&d, prm, mode, dt, warm, inertia, Y, U, psi, gates, X, R, st))) return rc;
  if (max_iters < 1) return OSC_OK;  // no iteration: X = x0 (never uninitialised memory)
  return pcg_core(d, g, chain, prm, mode, dt, jacobi, tol, max_iters, gates, X, R, ar, h_iters, h_res, st);
}

// Same recurrences for an arbitrary right-hand side and start vector (used by the screened-diffusion
// gate solve, preprocess/diffusion.py:132-150): X holds x0 on entry, B the right-hand side (clobbered).
int pcg_solve_system(const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm, int mode,
                     float dt, int jacobi, double tol, int max_iters, const float* gates, int D, float* X,
                     float* B, int* h_iters, float* h_res, void* workspace, size_t ws_bytes,
                     cudaStream_t st) {
  OSC_REQUIRE(g != nullptr && prm != nullptr && X != nullptr && B != nullptr, "pcg_solve_system: NULL argument");
  OSC_REQUIRE(g->batch == 1, "pcg_solve_system handles one lattice");
  osc_pcg_dims_t d{g->N, 0, g->N, D, 0};
  size_t need = 0;
  int rc = pcg_plan(&d, &need);
  if (rc) return rc;
  if (ws_bytes < need) return fail(OSC_ERR_WORKSPACE, "pcg_solve_system: workspace too small");
  if (h_iters) *h_iters = 0;
  if (h_res) *h_res = __builtin_nanf("");
  if (g->N == 0 || max_iters < 1) return OSC_OK;
  Arena ar(workspace, ws_bytes);
  return pcg_core(d, g, chain, prm, mode, dt, jacobi, tol, max_iters, gates, X, B, ar, h_iters, h_res, st);
}

// deltaH = sum_c diff_c . (M diff)_c   (receipts.py:21-25)
int delta_h(const osc_graph_t* g, const osc_chain_t* chain, const osc_params_t* prm, const float* U,
            const float* Ustar, const float* gates, int D, double* h_out, void* workspace,
            size_t ws_bytes, cudaStream_t st) {
  OSC_REQUIRE(g != nullptr && prm != nullptr && U != nullptr && Ustar != nullptr && h_out != nullptr,
              "delta_h: NULL argument");
  OSC_REQUIRE(g->batch == 1, "delta_h handles one lattice");
  osc_pcg_dims_t d{g->N, 0, g->N, D, 0};
  size_t need = 0;
  int rc = pcg_plan(&d, &need);
  if (rc) return rc;
  if (ws_bytes < need) return fail(OSC_ERR_WORKSPACE, "delta_h: workspace too small");
  *h_out = 0.0;
  if (g->N == 0) return OSC_OK;
  Arena ar(workspace, ws_bytes);
  const size_t nd = (size_t)g->N * D;
  float* diff = ar.take<float>(nd);
  float* Md = ar.take<float>(nd);
  double* part = ar.take<double>((size_t)d.n_blocks * D);
  float* colsum = ar.take<float>(D);
  double* col64 = ar.take<double>(D);
  double* total = ar.take<double>(8);
  if (!ar.ok) return fail(OSC_ERR_WORKSPACE, "delta_h: workspace too small");
  diff_kernel<<<ew_grid((int64_t)nd), 256, 0, st>>>(U, Ustar, diff, (int64_t)nd);
  OSC_LAUNCH_CHECK("diff_kernel");
  if ((rc = pcg_spmm_dot(&d, g, chain, prm, OSC_MODE_STATIONARY, 0.f, gates, diff, Md, part, st)))
    return rc;
  if ((rc = pcg_reduce(part, d.n_blocks, D, colsum, nullptr, col64, st))) return rc;
  sum_doubles_kernel<<<1, 1024, 0, st>>>(col64, D, total);
  OSC_LAUNCH_CHECK("sum_doubles_kernel");
  OSC_CUDA(cudaMemcpyAsync(h_out, total, sizeof(double), cudaMemcpyDeviceToHost, st));
  OSC_CUDA(cudaStreamSynchronize(st));
  return OSC_OK;
}

}  // namespace osc
