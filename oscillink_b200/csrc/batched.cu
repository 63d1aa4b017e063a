// K3: batched serving kernel -- many small lattices of equal (N, D, k) settled concurrently.
//
// Every CG reduction of solver.py:22-36 is per column, so a CTA that owns a 4-column slab of ALL
// N rows of one lattice runs the whole recurrence on chip with no other CTA involved:
//   p (the only vector that is gathered)  -> shared memory  [N] float4
//   ELL graph (u16 byte offset, fp32 W)   -> shared memory  (staged once per chunk of slabs)
//   x, r, Ap                              -> registers (thread owns fixed rows, one float4 each)
//
// The one thing that couples the columns of a lattice is the stop test (max over ALL D columns,
// solver.py:29-31).  It is resolved WITHOUT any cross-CTA synchronisation inside the iteration:
//   pass 1   every slab iterates until its OWN 4 columns satisfy the test and records
//            {iterations, max_c ||r_c||^2}.  The lattice's iteration count is T = max over slabs
//            (the reference stops at the first iteration where every column is below tol).
//   resolve  slabs that stopped before T are put on a fix list ...
//   pass 2   ... and re-run for exactly T iterations (same kernel, list mode).  For the usual
//            case -- all slabs of a lattice need the same count -- the list is empty.
//   finalize per-lattice {T, res = sqrt(max rr)}, deltaH = sum of slab partials (fixed order), and
//            an `unresolved` flag for the pathological non-monotone case (a re-run slab is above
//            tol again at T); the host mirror then settles that lattice with the HBM-resident
//            PCG (csrc/pcg.cu), which has the global test built in.
//
// Because slabs are independent, two CTAs share an SM: while one sits in a reduction/barrier
// latency chain the other one keeps the shared-memory pipe busy with its gathers.
//
// The kernel chains settle (lattice.py:170-207) -> stationary solve (lattice.py:245-265) -> deltaH
// (receipts.py:21-25), touching HBM only for Y in and U / U* out.
#include <cstdio>
#include <cstdlib>

#include "batched_common.cuh"

namespace osc {

// One PCG solve for this CTA's slab (solver.py:15-37).  On exit st.X holds the iterate of the
// returned iteration.  forced == 0: stop at the first iteration whose slab residual is <= tol
// (or at max_iters); forced > 0: run exactly `forced` iterations.  *rr_out = max_c ||r_c||^2 there.
//
// GATES = false (no gates given, b == 1): the operator diagonal and the Jacobi preconditioner are
// the same number for every row, so z = im_u * r and r.z = im_u * r.r -- the r.z reduction and the
// z vector disappear (one 4-value reduction per phase instead of 4 + 8).
template <int TPT, int KQ, int T, bool GATES>
__device__ int slab_solve(Slab<TPT>& st, const SolveCoef& c, double tol, int max_iters, int forced,
                          int N, int kq, float4* p_s, const float* diag_s, const float* im_s, const ushort4* nbr_s,
                          const float4* w_s, float4* red, const bool (&act)[TPT], float* rr_out,
                          float4* acc_out, const float4* acc_in) {
  // acc_out != nullptr: the gather sums of x0 are written there (thread-private rows);
  // acc_in  != nullptr: they are read back instead of gathering x0 again (same x0, same graph:
  //                     bit-identical to recomputing them).
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nw = T >> 5;
  float4* redA = red;               // init r.z, then p.Ap
  float4* redB = red + RED_F4;      // r.r
  float4* redC = red + 2 * RED_F4;  // r.z'
  const float noffc = -c.offc;
  // Rows tid + T*m >= N are PAD rows: zero state, zero weights, neighbour offset 0.  They run through
  // every phase unguarded (all their values stay 0 and add 0 to the reductions), which keeps the
  // per-row work free of per-thread branches so the compiler interleaves the TPT rows' load chains.
  if (acc_in == nullptr) {
#pragma unroll
    for (int m = 0; m < TPT; ++m) sts_v4(p_s + tid + T * m, st.X[m]);
  }
  __syncthreads();  // x0 visible; diag_s, im_s of this solve visible
  // ---- r0 = b - A x0 ; z0 ; rz
  V4 part = v4_zero();
#pragma unroll
  for (int m = 0; m < TPT; ++m) {
    const int row = tid + T * m;
    V4 g;
    if (acc_in != nullptr) {
      g = act[m] ? lds_v4(acc_in + row) : v4_zero();
    } else {
      g = gather_row<KQ>(p_s, nbr_s, w_s, row, N, kq);
      if (acc_out != nullptr && act[m]) sts_v4(acc_out + row, g);
    }
    const V4 a = combine_row(st.X[m], g, GATES ? diag_s[row] : c.diag_u, noffc);
    const V4 r = v4_sub(st.R[m], a);
    st.R[m] = r;
    if (GATES) {
      part = v4_fma(r, v4_mul_s(im_s[row], r), part);
    } else {
      part = v4_fma(r, r, part);
    }
  }
  warp_reduce4(to_f4(part), redA + warp, lane);
  __syncthreads();  // also: every gather of x0 has completed
  float rz = block_total_c(redA, nw, lane);  // component lane&3
  if (!GATES) rz *= c.im_u;
  // The thread's own rows of p stay in registers from the p update to the next gather phase: they sit
  // in the AP slots, which are dead between the r update and the next A(p).  With the x update moved
  // next to the p update (where p is read anyway) the own-row traffic on the shared-memory pipe is one
  // load + one store per row and iteration instead of three loads + one store.
#pragma unroll
  for (int m = 0; m < TPT; ++m) {
    const float imr = GATES ? im_s[tid + T * m] : c.im_u;
    st.AP[m] = v4_mul_s(imr, st.R[m]);  // p0 = z0
    sts_v4(p_s + tid + T * m, st.AP[m]);
  }
  __syncthreads();

  int it = 1;
  float mx;
  while (true) {
    // ---- A: Ap, p.Ap
    part = v4_zero();
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      const int row = tid + T * m;
      const V4 own = st.AP[m];
      st.AP[m] = combine_row(own, gather_row<KQ>(p_s, nbr_s, w_s, row, N, kq),
                             GATES ? diag_s[row] : c.diag_u, noffc);
      part = v4_fma(own, st.AP[m], part);
    }
    warp_reduce4(to_f4(part), redA + warp, lane);
    __syncthreads();
    const float pap = block_total_c(redA, nw, lane);
    const float4 alpha = bcast4(__fdiv_rn(rz, pap + 1e-18f));
    const V4 al = to_v4(alpha);
    const V4 nal = to_v4(make_float4(-alpha.x, -alpha.y, -alpha.z, -alpha.w));
    // ---- C: r update; rr and rz'   (x += alpha p is applied in E / at the stop, same arithmetic)
    V4 prr = v4_zero(), prz = v4_zero();
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      const V4 r = v4_fma(st.AP[m], nal, st.R[m]);
      st.R[m] = r;
      prr = v4_fma(r, r, prr);
      if (GATES) prz = v4_fma(r, v4_mul_s(im_s[tid + T * m], r), prz);
    }
    float rr, rzn;
    if (GATES) {
      warp_reduce8(to_f4(prr), to_f4(prz), redB + warp, redC + warp, lane);
      __syncthreads();
      rr = block_total_c(redB, nw, lane);
      rzn = block_total_c(redC, nw, lane);
    } else {
      warp_reduce4(to_f4(prr), redB + warp, lane);
      __syncthreads();
      rr = block_total_c(redB, nw, lane);
      rzn = rr * c.im_u;
    }
    mx = fmaxf(rr, __shfl_xor_sync(0xffffffffu, rr, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    // identical in every thread (same summation order) -> uniform branch
    const bool stop = forced > 0 ? (it >= forced)
                                 : ((double)__fsqrt_rn(mx) <= tol || it >= max_iters);
    if (stop) {
#pragma unroll
      for (int m = 0; m < TPT; ++m) st.X[m] = v4_fma(lds_v4(p_s + tid + T * m), al, st.X[m]);
      break;
    }
    // ---- E: x += alpha p ; p = z + beta p   (z = r / (Mdiag + 1e-12), recomputed from r)
    const V4 beta = to_v4(bcast4(__fdiv_rn(rzn, rz + 1e-18f)));
    rz = rzn;
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      const int row = tid + T * m;
      const float im = GATES ? im_s[row] : c.im_u;
      const V4 p = lds_v4(p_s + row);
      st.X[m] = v4_fma(p, al, st.X[m]);
      st.AP[m] = v4_fma(p, beta, v4_mul_s(im, st.R[m]));
      sts_v4(p_s + row, st.AP[m]);
    }
    __syncthreads();
    ++it;
  }
  *rr_out = mx;
  return it;
}

// ---------------------------------------------------------------------------------------------
// Graph packing pre-pass: ELL (int32 nbr / fp32 W / deg) -> the shared-memory image of the slab
// kernel: slot-major [kq][N] ushort4 byte offsets (16*j) + [kq][N] float4 weights, kp = 4*kq.
//
// Bank conflicts: a 128-bit shared load is served per quarter-warp = 8 consecutive lattice rows.
// Row j of p occupies banks 4*(j mod 8)..+3, so the 8 rows collide whenever two of their t-th
// neighbours agree mod 8 (2.5 wavefronts per gather for random graphs).  The ORDER in which a
// row visits its neighbours is free, and padding slots (weight 0) may point at any row, so each
// group of 8 rows greedily schedules its neighbour lists such that the residues met at every
// step are distinct where possible (measured 1.34 wavefronts per gather at N=1200, k=8).
// Tried and dropped (B200, B=1440): (a) an exact bipartite edge colouring of the visit order -- 1.38,
// the residue imbalance inside a group (a residue met more than kp times) is what remains, not the
// greedy order; (b) permuting rows inside their block of 8 so that every group meets the residues
// evenly (1.12 in simulation): the slab kernel gained 3 % (20.0 -> 19.4 ms) but the balancing pass and
// the position-indirect packer cost 1.6 ms more than the plain packer.
__global__ void __launch_bounds__(128)
batched_pack_kernel(const int32_t* __restrict__ nbr, const float* __restrict__ W,
                    const int32_t* __restrict__ deg, int64_t batch, int N, int k, int kp,
                    unsigned short* __restrict__ out_nbr, float* __restrict__ out_w) {
  const int groups = (N + 7) / 8;
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= batch * groups) return;
  const int64_t b = gidx / groups;
  const int g = (int)(gidx - b * groups);
  const int kq = kp / 4;
  int idx[8][PK_MAXK];
  int cnt[8], left[8];
  unsigned used[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = 8 * g + r;
    cnt[r] = 0;
    used[r] = 0;
    if (row < N) {
      const int d = min(deg[b * N + row], k);
      cnt[r] = d;
      for (int t = 0; t < d; ++t) idx[r][t] = nbr[(b * N + row) * k + t];
    }
    left[r] = cnt[r];
  }
  for (int t = 0; t < kp; ++t) {
    unsigned taken = 0;
    int pick[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) pick[r] = -1;
    for (int ps = 0; ps < 2; ++ps) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (pick[r] >= 0 || left[r] == 0) continue;
        const bool must = left[r] >= (kp - t);
        if ((ps == 0) != must) continue;
        int rc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int e = 0; e < cnt[r]; ++e)
          if (!((used[r] >> e) & 1u)) rc[idx[r][e] & 7]++;
        int best = -1, sel = -1, first = -1;
        for (int e = 0; e < cnt[r]; ++e) {
          if ((used[r] >> e) & 1u) continue;
          if (first < 0) first = e;
          const int res = idx[r][e] & 7;
          if (!((taken >> res) & 1u) && rc[res] > best) {
            best = rc[res];
            sel = e;
          }
        }
        if (sel < 0 && must) sel = first;
        if (sel >= 0) {
          pick[r] = sel;
          used[r] |= 1u << sel;
          left[r]--;
          taken |= 1u << (idx[r][sel] & 7);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = 8 * g + r;
      if (row >= N) continue;
      int j;
      float w;
      if (pick[r] >= 0) {
        j = idx[r][pick[r]];
        w = W[(b * N + row) * k + pick[r]];
      } else {
        int res = 0;  // padding: any row whose residue is still free at this step
        while (res < 7 && ((taken >> res) & 1u)) ++res;
        taken |= 1u << res;
        j = res < N ? res : 0;
        w = 0.f;
      }
      const int64_t o = ((b * kq + (t >> 2)) * N + row) * 4 + (t & 3);
      out_nbr[o] = (unsigned short)(j * 16);
      out_w[o] = w;
    }
  }
}

// Register-resident variant of the packer for ELL widths <= 8 (the serving configuration): the same greedy
// schedule as batched_pack_kernel, but with every row / entry index a compile-time constant, so the 8 x 8
// neighbour ids, the per-row bookkeeping and the picks live in registers instead of local memory (the
// generic kernel above spends its time on local-memory traffic: 3.4 ms per 4096 lattices of N = 1200).
// Residue counts of a row's unused entries are kept as eight 4-bit counters in one word.
__global__ void __launch_bounds__(128)
batched_pack8_kernel(const int32_t* __restrict__ nbr, const float* __restrict__ W,
                     const int32_t* __restrict__ deg, int64_t batch, int N, int k,
                     unsigned short* __restrict__ out_nbr, float* __restrict__ out_w) {
  constexpr int KP = 8, KQ = 2;
  const int groups = (N + 7) / 8;
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= batch * groups) return;
  const int64_t b = gidx / groups;
  const int g = (int)(gidx - b * groups);
  int idx[8][8];
  int cnt[8], left[8];
  unsigned used[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int row = 8 * g + r;
    cnt[r] = 0;
    used[r] = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) idx[r][e] = 0;
    if (row < N) {
      const int d = min(deg[b * N + row], k);
      cnt[r] = d;
      const int32_t* src = nbr + (b * N + row) * (int64_t)k;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (e < d) idx[r][e] = src[e];
    }
    left[r] = cnt[r];
  }
  // 4-bit count per residue (column & 7) over a row's UNUSED entries, kept up to date as entries are placed
  unsigned rcw[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    rcw[r] = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (e < cnt[r]) rcw[r] += 1u << ((idx[r][e] & 7) * 4);
  }
  ushort4 jbuf[8];
  float4 wbuf[8];
#pragma unroll 1
  for (int t = 0; t < KP; ++t) {
    unsigned taken = 0;
    int pick_j[8], pick_e[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) pick_e[r] = -1;
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const bool must = left[r] >= (KP - t);
        if (pick_e[r] < 0 && left[r] != 0 && ((ps == 0) == must)) {
          int best = -1, sel = -1, sel_j = 0, first = -1, first_j = 0;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (e < cnt[r] && !((used[r] >> e) & 1u)) {
              if (first < 0) {
                first = e;
                first_j = idx[r][e];
              }
              const int res = idx[r][e] & 7;
              const int c = (int)((rcw[r] >> (res * 4)) & 15u);
              if (!((taken >> res) & 1u) && c > best) {
                best = c;
                sel = e;
                sel_j = idx[r][e];
              }
            }
          }
          if (sel < 0 && must) {
            sel = first;
            sel_j = first_j;
          }
          if (sel >= 0) {
            pick_e[r] = sel;
            pick_j[r] = sel_j;
            used[r] |= 1u << sel;
            rcw[r] -= 1u << ((sel_j & 7) * 4);
            left[r]--;
            taken |= 1u << (sel_j & 7);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = 8 * g + r;
      int j;
      float w = 0.f;
      if (pick_e[r] >= 0) {
        j = pick_j[r];
        if (row < N) w = W[(b * N + row) * (int64_t)k + pick_e[r]];
      } else {
        int res = 0;  // padding: any row whose residue is still free at this step
        while (res < 7 && ((taken >> res) & 1u)) ++res;
        taken |= 1u << res;
        j = res < N ? res : 0;
      }
      const unsigned short off = (unsigned short)(j * 16);
      // slot t of this row: component t & 3 of vector t >> 2
      const int comp = t & 3;
      if (comp == 0) { jbuf[r].x = off; wbuf[r].x = w; }
      else if (comp == 1) { jbuf[r].y = off; wbuf[r].y = w; }
      else if (comp == 2) { jbuf[r].z = off; wbuf[r].z = w; }
      else { jbuf[r].w = off; wbuf[r].w = w; }
      if (comp == 3 && row < N) {
        const int64_t o = (b * KQ + (t >> 2)) * (int64_t)N + row;
        reinterpret_cast<ushort4*>(out_nbr)[o] = jbuf[r];
        reinterpret_cast<float4*>(out_w)[o] = wbuf[r];
      }
    }
  }
}

template <int TPT, int KQ, int MAXT, int MINB, bool GATES>
__global__ void __launch_bounds__(MAXT, MINB) batched_settle_kernel(BatchedK P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = P.N, kq = P.kq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the block size is the template's MAXT: every row address is tid*16 + an immediate
  constexpr int T = MAXT;
  constexpr int Np = T * TPT;  // rows incl. pad rows (zero state / zero weights), see slab_solve
  // p lives in STATIC shared memory: its address is a link-time constant, so a gather is
  // LDS.128 [offset + imm] straight from the packed u16 byte offset (no address arithmetic)
  __shared__ __align__(16) float4 p_static[MAXT * TPT];
  float4* p_s = p_static;                                               // [Np]
  float4* w_s = reinterpret_cast<float4*>(smem_raw);                    // [kq][Np]
  float4* red = w_s + (size_t)Np * kq;                                  // 3 x RED_F4
  ushort4* nbr_s = reinterpret_cast<ushort4*>(red + 3 * RED_F4);        // [kq][Np]
  // with gates: [Np] operator diagonal and [Np] 1/(Mdiag + 1e-12) (not allocated otherwise)
  float* diag_s = reinterpret_cast<float*>(nbr_s + (size_t)Np * kq);
  float* im_s = diag_s + Np;
  // Y-slab prefetch buffer [Np] float4 (P.use_ybuf): the 4 columns of Y this CTA works on next are
  // copied with cp.async while the current slab is still iterating, so the set-up phases of a slab
  // (settle RHS, U* RHS / x0, the u0 of the deltaH identity) read shared memory instead of waiting on
  // HBM three times per slab.  Every thread copies and reads only ITS OWN rows: no barrier involved.
  float4* ybuf = reinterpret_cast<float4*>(GATES ? im_s + Np : diag_s);
  // pad rows are written here once and never again (staging and set-up touch rows < N only)
  for (int e = tid; e < Np * kq; e += T) {
    w_s[e] = f4_zero();
    nbr_s[e] = make_ushort4(0, 0, 0, 0);
  }
  for (int e = tid; e < Np; e += T) {
    p_s[e] = f4_zero();
    if (GATES) {
      diag_s[e] = 0.f;
      im_s[e] = 0.f;
    }
  }
  bool act[TPT];
#pragma unroll
  for (int m = 0; m < TPT; ++m) act[m] = (tid + T * m) < N;

  // x0 = Y for both solves when the caller gave no U_in: one gather pass over Y serves both initial
  // residuals.  With settle + U* + deltaH in one call, M(U - U*) follows from the two final
  // residuals (see below) and the deltaH SpMM disappears as well: 12 -> 10 SpMMs per lattice.
  const bool share_r0 = P.do_settle && P.do_ustar && P.U_in == nullptr;
  const bool dh_fast = P.do_settle && P.do_ustar && P.do_dh && P.dt != 0.f;  // dt == 0: SpMM deltaH below
  float4* scr_acc = P.scratch + (size_t)blockIdx.x * 2 * N;
  float4* scr_t1 = scr_acc + N;
  const bool list_mode = P.fix_list != nullptr;
  const int64_t n_work = list_mode ? (int64_t)(*P.fix_count) : P.n_work;
  const bool use_ybuf = P.use_ybuf != 0;
  bool ybuf_ahead = false;  // ybuf already holds / is receiving the slab about to be processed
  auto ybuf_fetch = [&](int64_t fb, int fs) {
    const float* src = P.Y + fb * (int64_t)N * P.D + fs * SC;
#pragma unroll
    for (int m = 0; m < TPT; ++m)
      if (act[m]) cp_async16(ybuf + tid + T * m, src + (int64_t)(tid + T * m) * P.D);
    cp_async_commit();
  };
  for (int64_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
    int64_t b;
    int s0, s1, Fs = 0, Fu = 0;
    if (list_mode) {
      const int4 e = P.fix_list[wk];
      b = e.x;
      s0 = e.y;
      s1 = s0 + 1;
      Fs = e.z;
      Fu = e.w;
    } else {
      b = wk / P.cpl;
      s0 = (int)(wk - b * P.cpl) * P.CH;
      s1 = min(s0 + P.CH, P.G);
    }
    __syncthreads();  // readers of the previous graph image are done
    {
      const uint4* src_w = reinterpret_cast<const uint4*>(P.pk_w) + b * N * kq;
      uint4* dst_w = reinterpret_cast<uint4*>(w_s);
      const uint2* src_n = reinterpret_cast<const uint2*>(P.pk_nbr) + b * N * kq;
      uint2* dst_n = reinterpret_cast<uint2*>(nbr_s);
      for (int c = 0; c < kq; ++c) {
        for (int e = tid; e < N; e += T) {
          dst_w[c * Np + e] = __ldg(src_w + c * N + e);
          dst_n[c * Np + e] = __ldg(src_n + c * N + e);
        }
      }
    }
    const float* Yb = P.Y + b * (int64_t)N * P.D;
    const float* Ub = (P.U_in ? P.U_in : P.Y) + b * (int64_t)N * P.D;
    float* Uo = P.U_out ? P.U_out + b * (int64_t)N * P.D : nullptr;
    const float* gb = GATES ? P.gates + b * N : nullptr;

    for (int s = s0; s < s1; ++s) {
      const int col = s * SC;
      const float4 psi4 = *reinterpret_cast<const float4*>(P.psi + b * P.D + col);
      Slab<TPT> st;
      float rr = 0.f;
      if (use_ybuf) {
        if (!ybuf_ahead) ybuf_fetch(b, s);
        cp_async_wait_all();
        ybuf_ahead = false;
      }
      // issued once this slab has read Y for the last time: the next slab of this work item, or the
      // first slab of this CTA's next work item
      auto ybuf_next = [&]() {
        if (!use_ybuf) return;
        if (s + 1 < s1) {
          ybuf_fetch(b, s + 1);
          ybuf_ahead = true;
        } else if (!list_mode && wk + gridDim.x < n_work) {
          const int64_t wn = wk + gridDim.x;
          const int64_t bn = wn / P.cpl;
          ybuf_fetch(bn, (int)(wn - bn * P.cpl) * P.CH);
          ybuf_ahead = true;
        }
      };
      // ---------------- settle: (I + dt M) U+ = U + dt (lamG Y + lamQ b psi^T), x0 = U
      if (P.do_settle) {
        SolveCoef c;
        c.lamG = P.lamG; c.lamQ = P.lamQ; c.dt = P.dt;
        c.settle = 1;
        c.diag0 = 1.0f + P.dt * (P.lamG + P.lamC);
        c.diag1 = P.dt * P.lamQ;
        c.offc = P.dt * P.lamC;
        c.diag_u = c.diag0 + c.diag1 * 1.0f;
        c.im_u = __fdiv_rn(1.0f, md_of(c, 1.0f) + 1e-12f);
        __syncthreads();  // previous solve's readers of diag_s, im_s, p_s are done
        if (GATES) {
          for (int e = tid; e < N; e += T) {
            const float bq = gb[e];
            diag_s[e] = c.diag0 + c.diag1 * bq;
            im_s[e] = __fdiv_rn(1.0f, md_of(c, bq) + 1e-12f);
          }
        }
#pragma unroll
        for (int m = 0; m < TPT; ++m) {
          st.X[m] = st.R[m] = st.AP[m] = v4_zero();
          if (act[m]) {
            const int row = tid + T * m;
            const float4 y = use_ybuf ? ybuf[row] : *reinterpret_cast<const float4*>(Yb + (int64_t)row * P.D + col);
            const float4 u = P.U_in ? *reinterpret_cast<const float4*>(Ub + (int64_t)row * P.D + col) : y;
            const float bq = GATES ? gb[row] : 1.0f;
            const float4 rhs = make_float4(
                __fadd_rn(__fmul_rn(P.lamG, y.x), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.x))),
                __fadd_rn(__fmul_rn(P.lamG, y.y), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.y))),
                __fadd_rn(__fmul_rn(P.lamG, y.z), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.z))),
                __fadd_rn(__fmul_rn(P.lamG, y.w), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.w))));
            st.X[m] = to_v4(u);
            st.R[m] = to_v4(make_float4(__fadd_rn(u.x, __fmul_rn(P.dt, rhs.x)), __fadd_rn(u.y, __fmul_rn(P.dt, rhs.y)),
                                        __fadd_rn(u.z, __fmul_rn(P.dt, rhs.z)), __fadd_rn(u.w, __fmul_rn(P.dt, rhs.w))));
          }
        }
        if (!P.do_ustar) ybuf_next();
        const int iters = slab_solve<TPT, KQ, T, GATES>(st, c, P.tol_settle, P.max_iters_settle, Fs, Np, kq,
                                                     p_s, diag_s, im_s, nbr_s, w_s, red, act, &rr,
                                                     share_r0 ? scr_acc : nullptr, nullptr);
        if (Uo != nullptr) {
#pragma unroll
          for (int m = 0; m < TPT; ++m)
            if (act[m]) *reinterpret_cast<float4*>(Uo + (int64_t)(tid + T * m) * P.D + col) = to_f4(st.X[m]);
        }
        if (dh_fast) {
          // (I + dt M) U = b - r_s with b = U_in + dt RHS  =>  M U = RHS + (U_in - U - r_s)/dt
          const float idt = __fdiv_rn(1.0f, P.dt);
#pragma unroll
          for (int m = 0; m < TPT; ++m) {
            if (act[m]) {
              const int row = tid + T * m;
              const float4 u0 = (use_ybuf && P.U_in == nullptr)
                                    ? ybuf[row]
                                    : *reinterpret_cast<const float4*>(Ub + (int64_t)row * P.D + col);
              const float4 x = to_f4(st.X[m]), r = to_f4(st.R[m]);
              scr_t1[row] = make_float4(((u0.x - x.x) - r.x) * idt, ((u0.y - x.y) - r.y) * idt,
                                        ((u0.z - x.z) - r.z) * idt, ((u0.w - x.w) - r.w) * idt);
            }
          }
        }
        if (tid == 0) P.rec[(b * 2 + 0) * P.G + s] = make_int2(iters, __float_as_int(rr));
      }
      // ---------------- stationary: M U* = lamG Y + lamQ b psi^T, x0 = Y
      if (P.do_ustar) {
        SolveCoef c;
        c.lamG = P.lamG; c.lamQ = P.lamQ; c.dt = 0.f;
        c.settle = 0;
        c.diag0 = P.lamG + P.lamC;
        c.diag1 = P.lamQ;
        c.offc = P.lamC;
        c.diag_u = c.diag0 + c.diag1 * 1.0f;
        c.im_u = __fdiv_rn(1.0f, md_of(c, 1.0f) + 1e-12f);
        __syncthreads();  // the settle solve's readers of diag_s, im_s are done
        if (GATES) {
          for (int e = tid; e < N; e += T) {
            const float bq = gb[e];
            diag_s[e] = c.diag0 + c.diag1 * bq;
            im_s[e] = __fdiv_rn(1.0f, md_of(c, bq) + 1e-12f);
          }
        }
#pragma unroll
        for (int m = 0; m < TPT; ++m) {
          st.X[m] = st.R[m] = st.AP[m] = v4_zero();
          if (act[m]) {
            const int row = tid + T * m;
            const float4 y = use_ybuf ? ybuf[row] : *reinterpret_cast<const float4*>(Yb + (int64_t)row * P.D + col);
            const float bq = GATES ? gb[row] : 1.0f;
            st.X[m] = to_v4(y);
            st.R[m] = to_v4(make_float4(
                __fadd_rn(__fmul_rn(P.lamG, y.x), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.x))),
                __fadd_rn(__fmul_rn(P.lamG, y.y), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.y))),
                __fadd_rn(__fmul_rn(P.lamG, y.z), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.z))),
                __fadd_rn(__fmul_rn(P.lamG, y.w), __fmul_rn(P.lamQ, __fmul_rn(bq, psi4.w)))));
          }
        }
        ybuf_next();
        const int iters = slab_solve<TPT, KQ, T, GATES>(st, c, P.tol_ustar, P.max_iters_ustar, Fu, Np, kq,
                                                     p_s, diag_s, im_s, nbr_s, w_s, red, act, &rr, nullptr,
                                                     share_r0 ? scr_acc : nullptr);
        if (P.Ustar_out != nullptr) {
          float* So = P.Ustar_out + b * (int64_t)N * P.D;
#pragma unroll
          for (int m = 0; m < TPT; ++m)
            if (act[m]) *reinterpret_cast<float4*>(So + (int64_t)(tid + T * m) * P.D + col) = to_f4(st.X[m]);
        }
        if (tid == 0) P.rec[(b * 2 + 1) * P.G + s] = make_int2(iters, __float_as_int(rr));
        // ---------------- deltaH = <U - U*, M (U - U*)>
        if (dh_fast) {
          // M U* = RHS - r_u  =>  M (U - U*) = (U_in - U - r_s)/dt + r_u : no SpMM, thread-local.
          // (The recurrence residuals differ from the true ones by ~eps*||A||*|U|, the same size as
          //  the rounding of a fresh fp32 matvec on U - U*.)
          float4 part = f4_zero();
#pragma unroll
          for (int m = 0; m < TPT; ++m) {
            if (act[m]) {
              const int row = tid + T * m;
              const float4 u = *reinterpret_cast<const float4*>(Uo + (int64_t)row * P.D + col);
              const float4 t1 = scr_t1[row];
              const float4 x = to_f4(st.X[m]), r = to_f4(st.R[m]);
              const float4 d = make_float4(__fsub_rn(u.x, x.x), __fsub_rn(u.y, x.y), __fsub_rn(u.z, x.z),
                                           __fsub_rn(u.w, x.w));
              part = f4_add(part, f4_mul(d, f4_add(t1, r)));
            }
          }
          __syncthreads();  // the solve's last readers of `red` are done
          warp_reduce4(part, red + warp, lane);
          __syncthreads();
          if (warp == 0) {
            const float4 tot = block_total(red, T >> 5, lane);
            if (lane == 0) P.dh_part[b * P.G + s] = (double)((tot.x + tot.y) + (tot.z + tot.w));
          }
        } else if (P.do_dh) {
#pragma unroll
          for (int m = 0; m < TPT; ++m) {
            if (act[m]) {
              const int row = tid + T * m;
              // the settled state: written by this very thread a moment ago (U_out), or the caller's U
              const float* usrc = P.do_settle ? Uo : Ub;
              const float4 u = *reinterpret_cast<const float4*>(usrc + (int64_t)row * P.D + col);
              const float4 x = to_f4(st.X[m]);
              p_s[row] = make_float4(__fsub_rn(u.x, x.x), __fsub_rn(u.y, x.y), __fsub_rn(u.z, x.z),
                                     __fsub_rn(u.w, x.w));
            }
          }
          __syncthreads();
          float4 part = f4_zero();
#pragma unroll
          for (int m = 0; m < TPT; ++m) {
            if (act[m]) {
              const int row = tid + T * m;
              const float4 a = to_f4(apply_row<KQ>(p_s, nbr_s, w_s, row, Np, kq, GATES ? diag_s[row] : c.diag_u, -c.offc));
              part = f4_add(part, f4_mul(p_s[row], a));
            }
          }
          warp_reduce4(part, red + warp, lane);
          __syncthreads();
          if (warp == 0) {
            const float4 tot = block_total(red, T >> 5, lane);
            if (lane == 0) P.dh_part[b * P.G + s] = (double)((tot.x + tot.y) + (tot.z + tot.w));
          }
        }
      }
    }
  }
}

// slabs whose own stop came before the lattice-wide count T = max over slabs go on the fix list
__global__ void batched_resolve_kernel(const int2* __restrict__ rec, int64_t batch, int G, int do_settle,
                                       int do_ustar, int4* __restrict__ list, int* __restrict__ count,
                                       int32_t* __restrict__ flags) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  int Ts = 0, Tu = 0;
  for (int s = 0; s < G; ++s) {
    if (do_settle) Ts = max(Ts, rec[(b * 2 + 0) * G + s].x);
    if (do_ustar) Tu = max(Tu, rec[(b * 2 + 1) * G + s].x);
  }
  int any = 0;
  for (int s = 0; s < G; ++s) {
    const bool bad = (do_settle && rec[(b * 2 + 0) * G + s].x != Ts) ||
                     (do_ustar && rec[(b * 2 + 1) * G + s].x != Tu);
    if (bad) {
      list[atomicAdd(count, 1)] = make_int4((int)b, s, Ts, Tu);
      any = 2;
    }
  }
  if (flags) flags[b] = any;  // bit 1: some slab of this lattice is re-run in pass 2
}

__global__ void batched_finalize_kernel(const int2* __restrict__ rec, const double* __restrict__ dh_part,
                                        int64_t batch, int G, int do_settle, int do_ustar, int do_dh,
                                        double tol_s, double tol_u, int max_s, int max_u,
                                        float* __restrict__ stats, double* __restrict__ deltaH,
                                        int32_t* __restrict__ unresolved, int force_bad) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  int bad = force_bad;
  for (int sv = 0; sv < 2; ++sv) {
    if (!(sv == 0 ? do_settle : do_ustar)) continue;
    int T = 0, Tmin = 0x7fffffff;
    float rr = 0.f;
    for (int s = 0; s < G; ++s) {
      const int2 e = rec[(b * 2 + sv) * G + s];
      T = max(T, e.x);
      Tmin = min(Tmin, e.x);
      rr = fmaxf(rr, __int_as_float(e.y));
    }
    const float res = __fsqrt_rn(rr);
    // after the fix pass every slab sits at T; a re-run slab above tol at T < max_iters means the
    // residual was not monotone there and the true stop is later -> host falls back for this lattice
    if (Tmin != T) bad = 1;
    if (T < (sv == 0 ? max_s : max_u) && !((double)res <= (sv == 0 ? tol_s : tol_u))) bad = 1;
    if (stats) {
      stats[b * 4 + 2 * sv] = (float)T;
      stats[b * 4 + 2 * sv + 1] = res;
    }
  }
  if (do_dh) {
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += dh_part[b * G + g];
    deltaH[b] = s;
  }
  if (unresolved) unresolved[b] = (unresolved[b] & 2) | bad;  // bit 0: host must fall back
}

// ================================================================= host side
// kernel variant serving N rows: MAXT threads (the launch block size) x TPT rows per thread
static int maxt_for(int64_t N) { return N <= 1280 ? 320 : 640; }
static int tpt_for(int64_t N) {
  const int t = maxt_for(N);
  return (int)((N + t - 1) / t);  // 1..4
}
static size_t batched_smem(int64_t N, int k, bool gates) {
  const int kp = (k + 3) / 4 * 4;
  const size_t Np = (size_t)maxt_for(N) * tpt_for(N);  // rows incl. pad rows
  return Np * kp * 4 + 3 * RED_F4 * 16 + Np * kp * 2 + (gates ? Np * 8 : 0);  // + p in static shared memory
}

// static shared memory of the kernel variant that serves N (p_static[MAXT * TPT])
static size_t batched_static_smem(int64_t N) { return (size_t)maxt_for(N) * tpt_for(N) * 16; }
// the optional Y-slab prefetch buffer (dynamic, behind the gates arrays)
static size_t batched_ybuf_smem(int64_t N) { return (size_t)maxt_for(N) * tpt_for(N) * 16; }

int batched_supported(int64_t N, int D, int k) {
  if (N < 1 || N > 2560) return 0;
  if (D % 4 != 0 || D < 4) return 0;
  if (k < 1 || k > PK_MAXK || batched_smem(N, k, true) + batched_static_smem(N) > 227 * 1024) return 0;
  return 1;
}

int batched_workspace(int64_t batch, int64_t N, int D, size_t* bytes) {
  const int G = D / SC;
  *bytes = align_up((size_t)batch * 2 * G * sizeof(int2)) + align_up((size_t)batch * G * sizeof(double)) +
           align_up((size_t)batch * G * sizeof(int4)) + align_up(sizeof(int)) +
           align_up((size_t)batch * N * PK_MAXK * sizeof(unsigned short)) +
           align_up((size_t)batch * N * PK_MAXK * sizeof(float)) +
           align_up((size_t)SCR_CTAS_PER_SM * sm_count() * 2 * N * sizeof(float4)) + 1024;
  return OSC_OK;
}

template <int TPT, int MAXT, int MINB, bool GATES>
static BatchedFn pick_kq2(int kq) {
  switch (kq) {
    case 1: return batched_settle_kernel<TPT, 1, MAXT, MINB, GATES>;
    case 2: return batched_settle_kernel<TPT, 2, MAXT, MINB, GATES>;
    case 3: return batched_settle_kernel<TPT, 3, MAXT, MINB, GATES>;
    default: return batched_settle_kernel<TPT, 4, MAXT, MINB, GATES>;
  }
}
template <int TPT, int MAXT, int MINB>
static BatchedFn pick_kq(int kq, bool gates) {
  return gates ? pick_kq2<TPT, MAXT, MINB, true>(kq) : pick_kq2<TPT, MAXT, MINB, false>(kq);
}

int batched_settle(const osc_graph_t* g, const osc_params_t* prm, const osc_batched_args_t* a,
                   void* workspace, size_t ws_bytes, cudaStream_t st) {
  OSC_REQUIRE(g != nullptr && prm != nullptr && a != nullptr, "batched_settle: NULL argument");
  if (!batched_supported(g->N, a->D, g->k))
    return fail(OSC_ERR_UNSUPPORTED, "batched_settle: shape not covered (need N<=2560, D%4==0, k<=16)");
  if (prm->chain_present) return fail(OSC_ERR_UNSUPPORTED, "batched_settle: chain prior not supported");
  OSC_REQUIRE(a->Y != nullptr && a->psi != nullptr, "batched_settle: Y/psi NULL");
  OSC_REQUIRE(a->do_settle || a->do_ustar, "batched_settle: nothing to do");
  OSC_REQUIRE(!a->do_deltaH || (a->do_ustar && a->deltaH != nullptr), "deltaH needs do_ustar");
  OSC_REQUIRE(!(a->do_settle && a->do_deltaH) || a->U_out != nullptr, "deltaH after settle needs U_out");
  OSC_REQUIRE(!a->do_settle || (a->max_iters_settle >= 1 && a->max_iters_settle <= 65535),
              "batched_settle: max_iters must be in [1,65535]");
  OSC_REQUIRE(!a->do_ustar || (a->max_iters_ustar >= 1 && a->max_iters_ustar <= 65535),
              "batched_settle: max_iters must be in [1,65535]");
  if (g->batch == 0) return OSC_OK;
  size_t need = 0;
  batched_workspace(g->batch, g->N, a->D, &need);
  if (ws_bytes < need) return fail(OSC_ERR_WORKSPACE, "batched_settle: workspace too small");

  const int N = (int)g->N, kp = (g->k + 3) / 4 * 4, kq = kp / 4, G = a->D / SC;
  Arena ar(workspace, ws_bytes);
  int2* rec = ar.take<int2>((size_t)g->batch * 2 * G);
  double* dh_part = ar.take<double>((size_t)g->batch * G);
  int4* fix_list = ar.take<int4>((size_t)g->batch * G);
  int* fix_count = ar.take<int>(1);
  unsigned short* pn = ar.take<unsigned short>((size_t)g->batch * N * kp);
  float* pw = ar.take<float>((size_t)g->batch * N * kp);
  float4* scratch = ar.take<float4>((size_t)SCR_CTAS_PER_SM * sm_count() * 2 * N);
  if (!ar.ok) return fail(OSC_ERR_WORKSPACE, "batched_settle: workspace too small");
  OSC_CUDA(cudaMemsetAsync(fix_count, 0, sizeof(int), st));
  {
    const int64_t groups8 = g->batch * ((N + 7) / 8);
    bool pack8 = kp == 8;
    {
      const char* e = getenv("OSC_BATCHED_PACK8");  // dev-only A/B switch
      if (e && atoi(e) == 0) pack8 = false;
    }
    if (pack8)
      batched_pack8_kernel<<<(unsigned)((groups8 + 127) / 128), 128, 0, st>>>(g->nbr, g->W, g->deg, g->batch, N,
                                                                              g->k, pn, pw);
    else
      batched_pack_kernel<<<(unsigned)((groups8 + 127) / 128), 128, 0, st>>>(g->nbr, g->W, g->deg, g->batch, N,
                                                                             g->k, kp, pn, pw);
    OSC_LAUNCH_CHECK("batched_pack_kernel");
  }

  BatchedK P;
  P.Y = a->Y; P.U_in = a->U_in; P.psi = a->psi; P.gates = a->gates;
  P.U_out = a->U_out; P.Ustar_out = a->Ustar_out;
  P.dh_part = dh_part; P.rec = rec;
  P.fix_list = nullptr; P.fix_count = nullptr;
  P.pk_nbr = pn; P.pk_w = pw;
  P.scratch = scratch;
  P.batch = g->batch; P.N = N; P.kq = kq; P.D = a->D; P.G = G;
  P.CH = 4;  // slabs per staged graph image (measured on B200 at B=1440: CH=2 23.07 ms, 4 22.65, 8 23.55)
  {
    const char* e = getenv("OSC_BATCHED_CHUNK");  // dev-only: slabs per staged graph image
    if (e && atoi(e) >= 1) P.CH = atoi(e);
  }
  if (P.CH > G) P.CH = G;
  P.cpl = (G + P.CH - 1) / P.CH;
  P.n_work = g->batch * P.cpl;
  P.do_settle = a->do_settle; P.do_ustar = a->do_ustar; P.do_dh = a->do_deltaH;
  P.lamG = prm->lamG; P.lamC = prm->lamC; P.lamQ = prm->lamQ; P.dt = a->dt;
  P.tol_settle = a->tol_settle; P.tol_ustar = a->tol_ustar;
  P.thr2_settle = batched_sq_threshold(a->tol_settle);
  P.thr2_ustar = batched_sq_threshold(a->tol_ustar);
  P.max_iters_settle = a->max_iters_settle; P.max_iters_ustar = a->max_iters_ustar;

  const bool gates = a->gates != nullptr;
  size_t smem = batched_smem(N, g->k, gates);
  const int tpt = tpt_for(N);
  int threads = maxt_for(N);
  BatchedFn fn = nullptr;
  // Fast path (batched_ms.cu): first settle (U = Y), uniform gates, settle + U* in one call -> the two
  // systems are shifts of one another and one multi-shift CG serves both.  Everything else (warm start,
  // gates, settle-only, N > 1280) takes the two-solve kernel below.
  bool ms_two_ctas = false;
  bool use_ms = a->do_settle && a->do_ustar && a->U_in == nullptr && !gates && a->U_out != nullptr &&
                a->dt > 0.f && prm->lamG + prm->lamQ > 0.f;
  {
    const char* e = getenv("OSC_BATCHED_MS");  // dev-only A/B switch
    if (e && atoi(e) == 0) use_ms = false;
  }
  if (use_ms) {
    int t_ms = 0, variant = 0;
    ms_two_ctas = false;
    size_t smem_ms = 0;
    {
      const char* e = getenv("OSC_BATCHED_MS_VARIANT");  // dev-only: 1 = shared-memory-graph kernel
      if (e) variant = atoi(e);
    }
    BatchedFn f = batched_ms_pick(N, kq, variant, &t_ms, &smem_ms, &ms_two_ctas);
    if (f != nullptr) {
      fn = f;
      threads = t_ms;
      smem = smem_ms;
      P.use_ybuf = 0;
      P.CH = 8;  // one CTA per SM: nothing overlaps the staging of a graph image, so stage it less often
      {
        const char* e = getenv("OSC_BATCHED_CHUNK");
        if (e && atoi(e) >= 1) P.CH = atoi(e);
      }
      if (P.CH > G) P.CH = G;
      P.cpl = (G + P.CH - 1) / P.CH;
      P.n_work = g->batch * P.cpl;
    } else {
      use_ms = false;
    }
  }
  int occ = 0;
  size_t smem_launch = smem;
  if (use_ms) {
    OSC_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OSC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)fn, threads, smem));
    if (ms_two_ctas) {
      // The tensor-memory variant is built for two CTAs per SM (256 TMEM columns each).  The occupancy
      // calculator reports 1 for a kernel that allocates tensor memory (measured: grid 148, 10 warps per SM),
      // although registers and shared memory admit two and two DO run side by side (13.9 vs 15.5 ms at
      // B = 1440 with the grid doubled): size the grid from the kernel's own resource figures.
      cudaFuncAttributes fa;
      OSC_CUDA(cudaFuncGetAttributes(&fa, (const void*)fn));
      const size_t per_cta = fa.sharedSizeBytes + smem + 1024;
      if (2 * per_cta <= 228 * 1024 && 2 * (size_t)fa.numRegs * threads <= 65536) occ = 2;
    }
  } else {
    if (threads == 640) {
      if (tpt <= 3) fn = pick_kq<3, 640, 1>(kq, gates);
      else fn = pick_kq<4, 640, 1>(kq, gates);
    } else if (tpt == 1) fn = pick_kq<1, 320, 2>(kq, gates);
    else if (tpt == 2) fn = pick_kq<2, 320, 2>(kq, gates);
    else if (tpt == 3) fn = pick_kq<3, 320, 2>(kq, gates);
    else fn = pick_kq<4, 320, 2>(kq, gates);
    // the prefetch buffer is taken only if it costs no resident CTA
    const size_t smem_y = smem + batched_ybuf_smem(N);
    const bool fits = smem_y + batched_static_smem(N) <= 227 * 1024;
    const char* e = getenv("OSC_BATCHED_YBUF");  // dev-only A/B switch
    const bool want = !(e && atoi(e) == 0);
    OSC_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(fits ? smem_y : smem)));
    OSC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)fn, threads, smem));
    int occ_y = 0;
    if (fits && want)
      OSC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_y, (const void*)fn, threads, smem_y));
    P.use_ybuf = (occ_y >= 1 && occ_y >= (occ < SCR_CTAS_PER_SM ? occ : SCR_CTAS_PER_SM)) ? 1 : 0;
    smem_launch = P.use_ybuf ? smem + batched_ybuf_smem(N) : smem;
  }
  {
    const char* e = getenv("OSC_BATCHED_OCC");  // dev-only: override the resident-CTA count per SM
    if (e && atoi(e) >= 1) occ = atoi(e);
  }
  if (occ < 1) return fail(OSC_ERR_UNSUPPORTED, "batched_settle: kernel does not fit on an SM");
  if (occ > SCR_CTAS_PER_SM) occ = SCR_CTAS_PER_SM;
  const int64_t resident = (int64_t)occ * sm_count();
  const unsigned grid = (unsigned)(P.n_work < resident ? P.n_work : resident);
  {
    const char* e = getenv("OSC_BATCHED_DEBUG");  // dev-only: launch geometry
    if (e && atoi(e) != 0)
      fprintf(stderr, "[osc] batched: ms=%d threads=%d smem_dyn=%zu occ=%d grid=%u n_work=%lld CH=%d\n", (int)use_ms,
              threads, smem_launch, occ, grid, (long long)P.n_work, P.CH);
  }
  fn<<<grid, threads, smem_launch, st>>>(P);
  OSC_LAUNCH_CHECK("batched_settle_kernel");

  batched_resolve_kernel<<<(unsigned)((g->batch + 127) / 128), 128, 0, st>>>(
      rec, g->batch, G, a->do_settle, a->do_ustar, fix_list, fix_count, a->unresolved);
  OSC_LAUNCH_CHECK("batched_resolve_kernel");
  {
    BatchedK Q = P;
    Q.fix_list = fix_list;
    Q.fix_count = fix_count;
    const int64_t worst = g->batch * G;
    const unsigned grid2 = (unsigned)(worst < resident ? worst : resident);
    fn<<<grid2, threads, smem_launch, st>>>(Q);
    OSC_LAUNCH_CHECK("batched_settle_kernel(fix)");
  }
  int force_bad = 0;
  {
    const char* e = getenv("OSC_BATCHED_FORCE_UNRESOLVED");  // test hook for the host fallback path
    if (e && atoi(e) != 0) force_bad = 1;
  }
  batched_finalize_kernel<<<(unsigned)((g->batch + 127) / 128), 128, 0, st>>>(
      rec, dh_part, g->batch, G, a->do_settle, a->do_ustar, a->do_deltaH, a->tol_settle, a->tol_ustar,
      a->max_iters_settle, a->max_iters_ustar, a->stats, a->deltaH, a->unresolved, force_bad);
  OSC_LAUNCH_CHECK("batched_finalize_kernel");
  return OSC_OK;
}

}  // namespace osc
