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

#include <cstdio>
#include <cstdlib>

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
  unsigned* sync;  // [batch][2 solves][maxit+1]{max bits, arrivals} (zeroed per launch)
  unsigned long long* prof;  // dev-only phase clocks (block 0, thread 0)
  const unsigned short* pk_nbr;  // packed graph image [batch][N][kp]
  const float* pk_w;
  int64_t batch, N;
  int k, kp, D, G, groups, maxit;
  int do_settle, do_ustar, do_dh;
  float lamG, lamC, lamQ, dt;
  double tol_settle, tol_ustar;
  int max_iters_settle, max_iters_ustar;
  int debug;  // dev-only experiment switches (OSC_BATCHED_DEBUG): 1 = no group wait, 2 = no gathers
};

struct SolveCoef {
  float diag0, diag1;  // operator diagonal = diag0 + diag1 * b_i
  float offc;
  float lamG, lamQ, dt;
  int settle, kq, debug;
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

#define OSC_TICK(slot)                                                        \
  do {                                                                        \
    if (prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0) {             \
      const long long _n = clock64();                                         \
      prof[slot] += (unsigned long long)(_n - tprev);                         \
      tprev = _n;                                                             \
    }                                                                         \
  } while (0)

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
  const int kq = (KQ > 0 ? KQ : kq_rt) & 0xff;
  if (!(kq_rt & 0x200))
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

// sum over lanes of equal parity (lane & 1); every lane ends with its parity's total
__device__ __forceinline__ float4 warp_parity_sum(float4 v) {
#pragma unroll
  for (int o = 2; o < 32; o <<= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
    v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  return v;
}
// second stage, executed redundantly by every warp: totals of the per-warp partials for this
// lane's parity (all lanes of a parity end with the same value)
__device__ __forceinline__ float4 warp0_collect(const float4* red, int lane, int nw) {
  float4 t = f4_zero();
  for (int w = lane >> 1; w < nw; w += 16) t = f4_add(t, red[w * 2 + (lane & 1)]);
  return warp_parity_sum(t);
}

// ---- shuffle-light block reductions ------------------------------------------------------------
// SHFL issues at ~1 warp-instruction/clk/SM, so reducing 8 separate floats with 4 butterfly steps
// each (32 SHFL) dominated the update phases.  Here the butterfly halves the number of live
// values at every step (4+2+1+1 = 8 SHFL for 8 values): after it each lane owns the 16-lane
// (equal-parity) total of ONE value, which it drops into shared memory.
__device__ __forceinline__ void reduce4_write(float4 v, float* redw, int lane) {
  const int b0 = (lane >> 1) & 1, b1 = (lane >> 2) & 1;
  float s0 = b0 ? v.x : v.z, s1 = b0 ? v.y : v.w;
  float k0 = b0 ? v.z : v.x, k1 = b0 ? v.w : v.y;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float s = b1 ? k0 : k1;
  float k = b1 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, s, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 16);
  if ((lane >> 3) == 0) redw[(lane & 1) * 4 + b0 * 2 + b1] = k;  // value b0*2+b1 of parity lane&1
}
// every warp folds the per-warp partials itself; returns the 4 column totals of this lane's parity
__device__ __forceinline__ float4 collect4(const float* red, float* tscr_w, int lane, int nw) {
  float t = 0.f;
  for (int w = lane >> 3; w < nw; w += 4) t += red[w * 8 + (lane & 7)];
  t += __shfl_xor_sync(0xffffffffu, t, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  if (lane < 8) tscr_w[lane] = t;
  __syncwarp();
  const float4 r = *reinterpret_cast<const float4*>(tscr_w + (lane & 1) * 4);
  __syncwarp();
  return r;
}
__device__ __forceinline__ void reduce8_write(float4 a, float4 b, float* redw, int lane) {
  const int b0 = (lane >> 1) & 1, b1 = (lane >> 2) & 1, b2 = (lane >> 3) & 1;
  const float4 s = b0 ? a : b;
  float4 k = b0 ? b : a;
  k.x += __shfl_xor_sync(0xffffffffu, s.x, 2);
  k.y += __shfl_xor_sync(0xffffffffu, s.y, 2);
  k.z += __shfl_xor_sync(0xffffffffu, s.z, 2);
  k.w += __shfl_xor_sync(0xffffffffu, s.w, 2);
  const float s0 = b1 ? k.x : k.z, s1 = b1 ? k.y : k.w;
  float k0 = b1 ? k.z : k.x, k1 = b1 ? k.w : k.y;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 4);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 4);
  const float ss = b2 ? k0 : k1;
  float kk = b2 ? k1 : k0;
  kk += __shfl_xor_sync(0xffffffffu, ss, 8);
  kk += __shfl_xor_sync(0xffffffffu, kk, 16);
  if ((lane >> 4) == 0) redw[(lane & 1) * 8 + b0 * 4 + b1 * 2 + b2] = kk;  // value b0*4+b1*2+b2
}
__device__ __forceinline__ void collect8(const float* red, float* tscr_w, int lane, int nw, float4& a,
                                         float4& b) {
  float t = 0.f;
  for (int w = lane >> 4; w < nw; w += 2) t += red[w * 16 + (lane & 15)];
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  if (lane < 16) tscr_w[lane] = t;
  __syncwarp();
  a = *reinterpret_cast<const float4*>(tscr_w + (lane & 1) * 8);
  b = *reinterpret_cast<const float4*>(tscr_w + (lane & 1) * 8 + 4);
  __syncwarp();
}

// ---- group exchange of the per-slab residual maxima -------------------------------------------
// Per (lattice, solve, iteration): {max bits, arrival count}.  The slab maximum is folded in with
// a relaxed RED as soon as it is known; the release-add on the counter (which fences) is issued
// only after the p update, when the first RED has long been acknowledged, so the fence is cheap.
__device__ __forceinline__ void group_publish_max(unsigned* slot, float mx) {
  asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(slot), "r"(__float_as_uint(fmaxf(mx, 0.f)))
               : "memory");
}
__device__ __forceinline__ void group_publish_arrive(unsigned* slot) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(slot + 1), "r"(1u) : "memory");
}
__device__ __forceinline__ unsigned group_wait(unsigned* slot, unsigned G) {
  unsigned cnt;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cnt) : "l"(slot + 1) : "memory");
  } while (cnt < G);
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
  return v;
}

// One PCG solve for this CTA's slab (solver.py:15-37).  On exit st.X holds the solution.
// Returns the iteration count; *res_out is the group-wide max-column residual at that iteration.
//
// Reductions: warp shuffle over equal-parity lanes -> per-warp partials in shared memory -> one
// barrier -> EVERY warp folds the partials itself (2 LDS + 16 SHFL), so no serial section and no
// second barrier: 3 barriers per iteration.
//
// The stop test of iteration `it` needs the maximum over all G slabs of the lattice.  Instead of
// stalling on it, the CTA publishes its slab maximum, goes straight on to the SpMM of iteration
// it+1 and only then consumes the group result (prefetched while the SpMM runs): if the solve had
// converged the speculative SpMM is dropped -- x and r are untouched at that point.
template <int TPT, int KQ>
__device__ int slab_solve(Slab<TPT>& st, const SolveCoef& c, double tol, int max_iters, float4* p_s,
                          const float2* rowc_s, const ushort4* nbr_s, const float4* w_s, float4* red,
                          unsigned* sync_base, int G, const bool (&act)[TPT], float* res_out,
                          unsigned* flag_s, unsigned long long* prof) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = tid & 1;
  const int nw = BW;
  long long tprev = clock64();
  float* redf = reinterpret_cast<float*>(red);
  float* red0 = redf;                 // [BW_MAX][8]   init r.z
  float* redA = redf + 256;           // [BW_MAX][8]   p.Ap
  float* redBC = redf + 512;          // [BW_MAX][16]  r.r | r.z'
  float* tscr = redf + 1024 + warp * 16;
  const int kqd = c.kq | ((c.debug & 2) << 8);
  // ---- r0 = b - A x0 ; p = z0 ; rz
  __syncthreads();  // previous users of p_s / rowc_s are done, rowc_s of this solve is written
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
      const float2 rc = rowc_s[row];
      const float4 a = apply_task<KQ>(p_s, nbr_s, w_s, row, half, kqd, rc.x, c.offc);
      float4 r = st.R[m];
      r = make_float4(r.x - a.x, r.y - a.y, r.z - a.z, r.w - a.w);
      st.R[m] = r;
      z0[m] = make_float4(r.x * rc.y, r.y * rc.y, r.z * rc.y, r.w * rc.y);
      part = f4_add(part, f4_mul(r, z0[m]));
    }
  }
  reduce4_write(part, red0 + warp * 8, lane);
  __syncthreads();  // also: every gather of x0 has completed
  float4 rz = collect4(red0, tscr, lane, nw);
#pragma unroll
  for (int m = 0; m < TPT; ++m)
    if (act[m]) p_s[tid + BT * m] = z0[m];
  __syncthreads();

  OSC_TICK(8);  // init (x0 -> r0, p0)
  int it = 1, done_it = 0;
  float res = __int_as_float(0x7fc00000);
  while (true) {
    // prefetch the group result of the previous iteration; it lands while the SpMM runs
    unsigned pre_cnt = 0, pre_val = 0x7f800000u;
    if (tid == 0 && it > 1 && !(c.debug & 1)) {
      unsigned* slot = sync_base + 2 * (it - 1);
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(pre_cnt) : "l"(slot + 1) : "memory");
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(pre_val) : "l"(slot) : "memory");
    }
    // ---- A: Ap, p.Ap
    part = f4_zero();
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      if (act[m]) {
        const int q = tid + BT * m, row = q >> 1;
        st.AP[m] = apply_task<KQ>(p_s, nbr_s, w_s, row, half, kqd, rowc_s[row].x, c.offc);
        part = f4_add(part, f4_mul(p_s[q], st.AP[m]));
      }
    }
    OSC_TICK(9);  // A: SpMM
    reduce4_write(part, redA + warp * 8, lane);
    if (tid == 0) {
      unsigned fl = 0x7f800000u;  // +inf: "not converged"
      if (it > 1 && !(c.debug & 1))
        fl = (pre_cnt >= (unsigned)G) ? pre_val : group_wait(sync_base + 2 * (it - 1), (unsigned)G);
      *flag_s = fl;
    }
    __syncthreads();
    if (it > 1) {
      res = __fsqrt_rn(__uint_as_float(*flag_s));
      if ((double)res <= tol) {  // solver.py:29-31 -- converged at it-1; drop the speculative Ap
        done_it = it - 1;
        break;
      }
    }
    OSC_TICK(10);  // barrier 1 + flag
    const float4 pap = collect4(redA, tscr, lane, nw);
    const float4 alpha = make_float4(__fdiv_rn(rz.x, pap.x + 1e-18f), __fdiv_rn(rz.y, pap.y + 1e-18f),
                                     __fdiv_rn(rz.z, pap.z + 1e-18f), __fdiv_rn(rz.w, pap.w + 1e-18f));
    // ---- C: x, r update; rr and rz'
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
        x = make_float4(fmaf(p.x, alpha.x, x.x), fmaf(p.y, alpha.y, x.y), fmaf(p.z, alpha.z, x.z),
                        fmaf(p.w, alpha.w, x.w));
        r = make_float4(fmaf(-ap.x, alpha.x, r.x), fmaf(-ap.y, alpha.y, r.y), fmaf(-ap.z, alpha.z, r.z),
                        fmaf(-ap.w, alpha.w, r.w));
        st.X[m] = x;
        st.R[m] = r;
        const float im = rowc_s[row].y;
        const float4 z = make_float4(r.x * im, r.y * im, r.z * im, r.w * im);
        zz[m] = z;
        prr = f4_add(prr, f4_mul(r, r));
        prz = f4_add(prz, f4_mul(r, z));
      }
    }
    OSC_TICK(11);  // C: update
    reduce8_write(prr, prz, redBC + warp * 16, lane);
    __syncthreads();
    OSC_TICK(12);  // barrier 2
    float4 rr, rzn;
    collect8(redBC, tscr, lane, nw, rr, rzn);
    if (warp == 0) {  // publish this slab's residual maximum to the group
      float mx = fmaxf(fmaxf(rr.x, rr.y), fmaxf(rr.z, rr.w));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      if (lane == 0) group_publish_max(sync_base + 2 * it, mx);
    }
    if (it == max_iters) {
      if (tid == 0) {
        group_publish_arrive(sync_base + 2 * it);
        *flag_s = (c.debug & 1) ? 0x7f800000u : group_wait(sync_base + 2 * it, (unsigned)G);
      }
      __syncthreads();
      res = __fsqrt_rn(__uint_as_float(*flag_s));
      done_it = it;
      break;
    }
    // ---- E: p = z + beta p
    const float4 beta = make_float4(__fdiv_rn(rzn.x, rz.x + 1e-18f), __fdiv_rn(rzn.y, rz.y + 1e-18f),
                                    __fdiv_rn(rzn.z, rz.z + 1e-18f), __fdiv_rn(rzn.w, rz.w + 1e-18f));
    rz = rzn;
#pragma unroll
    for (int m = 0; m < TPT; ++m) {
      if (act[m]) {
        const int q = tid + BT * m;
        const float4 p = p_s[q];
        const float4 z = zz[m];
        p_s[q] = make_float4(fmaf(p.x, beta.x, z.x), fmaf(p.y, beta.y, z.y), fmaf(p.z, beta.z, z.z),
                             fmaf(p.w, beta.w, z.w));
      }
    }
    if (tid == 0) group_publish_arrive(sync_base + 2 * it);
    __syncthreads();
    OSC_TICK(13);  // E: p update + barrier 3
    ++it;
  }
  *res_out = res;
  return done_it;
}


// ---------------------------------------------------------------------------------------------
// Graph packing pre-pass: ELL (int32 nbr / fp32 W / deg) -> the shared-memory image of the slab
// kernel ([N][kp] u16 neighbour + [N][kp] fp32 weight, kp = k rounded up to 4).
//
// Bank conflicts: a 128-bit shared load is served per quarter-warp = 4 lattice rows (2 lanes per
// row).  Row j of p occupies banks 8*(j mod 4)..+7, so the 4 rows of a quarter-warp collide
// whenever two of their t-th neighbours agree mod 4 (2.04 wavefronts per phase for random
// graphs).  The ORDER in which a row visits its neighbours is free, and padding slots (weight 0)
// may point at any row, so each group of 4 rows greedily schedules its neighbour lists such that
// the residues at every step are distinct (measured 1.18 wavefronts per phase).
constexpr int PK_MAXK = 16;

__global__ void __launch_bounds__(128)
batched_pack_kernel(const int32_t* __restrict__ nbr, const float* __restrict__ W,
                    const int32_t* __restrict__ deg, int64_t batch, int N, int k, int kp,
                    unsigned short* __restrict__ out_nbr, float* __restrict__ out_w) {
  const int groups = (N + 3) / 4;
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= batch * groups) return;
  const int64_t b = gidx / groups;
  const int g = (int)(gidx - b * groups);
  int idx[4][PK_MAXK];
  float wv[4][PK_MAXK];
  int cnt[4];
  unsigned used[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = 4 * g + r;
    cnt[r] = 0;
    used[r] = 0;
    if (row < N) {
      const int d = min(deg[b * N + row], k);
      cnt[r] = d;
      for (int t = 0; t < d; ++t) {
        idx[r][t] = nbr[(b * N + row) * k + t];
        wv[r][t] = W[(b * N + row) * k + t];
      }
    }
  }
  int left[4] = {cnt[0], cnt[1], cnt[2], cnt[3]};
  for (int t = 0; t < kp; ++t) {
    unsigned taken = 0;
    int pick[4] = {-1, -1, -1, -1};
    for (int ps = 0; ps < 2; ++ps) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (pick[r] >= 0 || left[r] == 0) continue;
        const bool must = left[r] >= (kp - t);
        if ((ps == 0) != must) continue;
        int rc[4] = {0, 0, 0, 0};
        for (int e = 0; e < cnt[r]; ++e)
          if (!((used[r] >> e) & 1u)) rc[idx[r][e] & 3]++;
        int best = -1, sel = -1, first = -1;
        for (int e = 0; e < cnt[r]; ++e) {
          if ((used[r] >> e) & 1u) continue;
          if (first < 0) first = e;
          const int res = idx[r][e] & 3;
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
          taken |= 1u << (idx[r][sel] & 3);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = 4 * g + r;
      if (row >= N) continue;
      int j;
      float w;
      if (pick[r] >= 0) {
        j = idx[r][pick[r]];
        w = wv[r][pick[r]];
      } else {
        int res = 0;  // padding: any row whose residue is still free at this step
        while (res < 3 && ((taken >> res) & 1u)) ++res;
        taken |= 1u << res;
        j = res < N ? res : 0;
        w = 0.f;
      }
      out_nbr[(b * N + row) * kp + t] = (unsigned short)j;
      out_w[(b * N + row) * kp + t] = w;
    }
  }
}

template <int TPT, int KQ>
__global__ void __launch_bounds__(TPT == 3 ? 800 : (TPT == 4 ? 640 : 1024), 1)
batched_settle_kernel(BatchedK P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = (int)P.N, kq = P.kp / 4;
  float4* p_s = reinterpret_cast<float4*>(smem_raw);                 // [N][2]
  float4* w_s = p_s + (size_t)N * 2;                                  // [N][kq]
  float4* red = w_s + (size_t)N * kq;                                 // 1536 floats of reduction scratch
  ushort4* nbr_s = reinterpret_cast<ushort4*>(red + 384);             // [N][kq]
  float2* rowc_s = reinterpret_cast<float2*>(nbr_s + (size_t)N * kq);  // [N] (diag, 1/Mdiag)
  float* gates_s = reinterpret_cast<float*>(rowc_s + N);              // [N]
  unsigned* flag_s = reinterpret_cast<unsigned*>(gates_s + N);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = tid & 1;
  const int gid = blockIdx.x / P.G, slab = blockIdx.x % P.G;
  const int col = slab * BC + half * 4;
  const bool col_ok = col + 3 < P.D;
  bool act[TPT];
#pragma unroll
  for (int m = 0; m < TPT; ++m) act[m] = col_ok && ((tid + BT * m) >> 1) < N;

  unsigned long long* prof = P.prof;
  long long tprev = clock64();
  for (int64_t b = gid; b < P.batch; b += P.groups) {
    __syncthreads();
    OSC_TICK(0);  // loop top barrier
    // ---- stage the packed graph image + gates (straight 16 B copies)
    {
      const uint4* src_w = reinterpret_cast<const uint4*>(P.pk_w + b * P.N * P.kp);
      uint4* dst_w = reinterpret_cast<uint4*>(w_s);
      for (int e = tid; e < N * kq; e += BT) dst_w[e] = src_w[e];
      const uint2* src_n = reinterpret_cast<const uint2*>(P.pk_nbr + b * P.N * P.kp);
      uint2* dst_n = reinterpret_cast<uint2*>(nbr_s);
      for (int e = tid; e < N * kq; e += BT) dst_n[e] = src_n[e];
      for (int e = tid; e < N; e += BT) gates_s[e] = P.gates ? P.gates[b * P.N + e] : 1.0f;
    }
    const float* Yb = P.Y + b * P.N * P.D;
    const float* Ub = (P.U_in ? P.U_in : P.Y) + b * P.N * P.D;
    float* Uo = P.U_out ? P.U_out + b * P.N * P.D : nullptr;
    const float4 psi4 = col_ok ? *reinterpret_cast<const float4*>(P.psi + b * P.D + col) : f4_zero();
    __syncthreads();

    OSC_TICK(1);  // graph staging
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
      c.debug = P.debug;
      c.diag0 = 1.0f + P.dt * (P.lamG + P.lamC);
      c.diag1 = P.dt * P.lamQ;
      c.offc = P.dt * P.lamC;
#pragma unroll
      for (int m = 0; m < TPT; ++m) {
        st.X[m] = st.R[m] = st.AP[m] = f4_zero();
        if (act[m]) {
          const int row = (tid + BT * m) >> 1;
          const float4 y = *reinterpret_cast<const float4*>(Yb + (int64_t)row * P.D + col);
          const float4 u = P.U_in ? *reinterpret_cast<const float4*>(Ub + (int64_t)row * P.D + col) : y;
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
      for (int e = tid; e < N; e += BT) {
        const float bq = gates_s[e];
        rowc_s[e] = make_float2(c.diag0 + c.diag1 * bq, __fdiv_rn(1.0f, md_of(c, bq) + 1e-12f));
      }
      OSC_TICK(2);  // settle: load Y/U, rhs
      iters = slab_solve<TPT, KQ>(st, c, P.tol_settle, P.max_iters_settle, p_s, rowc_s, nbr_s, w_s, red,
                                  sync_b, P.G, act, &res, flag_s, prof);
      tprev = clock64();
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
      c.debug = P.debug;
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
      OSC_TICK(3);  // store U, load Y, rhs (stationary)
      __syncthreads();  // the settle solve's readers of rowc_s are done
      for (int e = tid; e < N; e += BT) {
        const float bq = gates_s[e];
        rowc_s[e] = make_float2(c.diag0 + c.diag1 * bq, __fdiv_rn(1.0f, md_of(c, bq) + 1e-12f));
      }
      iters = slab_solve<TPT, KQ>(st, c, P.tol_ustar, P.max_iters_ustar, p_s, rowc_s, nbr_s, w_s, red,
                                  sync_b + (P.maxit + 1) * 2, P.G, act, &res, flag_s, prof);
      tprev = clock64();
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
            const float4 a = apply_task<KQ>(p_s, nbr_s, w_s, row, half, kq, rowc_s[row].x, c.offc);
            part = f4_add(part, f4_mul(p_s[q], a));
          }
        }
        const float4 tot = block_colsum(part, red, lane, warp, half);
        float s4 = (tot.x + tot.y) + (tot.z + tot.w);
        s4 += __shfl_xor_sync(0xffffffffu, s4, 1);
        if (tid == 0) P.dh_part[b * P.G + slab] = (double)s4;
      }
      OSC_TICK(4);  // store U*, deltaH
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
  return (size_t)N * 32 + (size_t)N * kp * 4 + (size_t)N * kp * 2 + 384 * 16 + (size_t)N * 12 +
         16;
}

int batched_supported(int64_t N, int D, int k) {
  if (N < 1 || N > 65535 || tpt_for(N) == 0) return 0;
  if (D % 4 != 0 || D < 4) return 0;
  const int G = (D + BC - 1) / BC;
  if (G > sm_count()) return 0;
  if (k < 1 || k > 16 || batched_smem(N, k) > 227 * 1024) return 0;
  return 1;
}

int batched_workspace(int64_t batch, int64_t N, int D, size_t* bytes) {
  const int G = (D + BC - 1) / BC;
  const int maxit = 256;
  *bytes = align_up((size_t)batch * 2 * (maxit + 1) * 2 * sizeof(unsigned)) +
           align_up((size_t)batch * G * sizeof(double)) + 2048 +
           align_up((size_t)batch * N * 16 * sizeof(unsigned short)) +
           align_up((size_t)batch * N * 16 * sizeof(float));
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
  {
    const char* dbg = getenv("OSC_BATCHED_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
  }
  Arena ar(workspace, ws_bytes);
  const size_t sync_n = (size_t)g->batch * 2 * (P.maxit + 1) * 2;
  P.sync = ar.take<unsigned>(sync_n);
  P.dh_part = ar.take<double>((size_t)g->batch * P.G);
  if (!ar.ok) return fail(OSC_ERR_WORKSPACE, "batched_settle: workspace too small");
  OSC_CUDA(cudaMemsetAsync(P.sync, 0, sync_n * sizeof(unsigned), st));
  {
    unsigned short* pn = ar.take<unsigned short>((size_t)g->batch * g->N * P.kp);
    float* pw = ar.take<float>((size_t)g->batch * g->N * P.kp);
    if (!ar.ok) return fail(OSC_ERR_WORKSPACE, "batched_settle: workspace too small");
    const int64_t groups4 = g->batch * ((g->N + 3) / 4);
    batched_pack_kernel<<<(unsigned)((groups4 + 127) / 128), 128, 0, st>>>(g->nbr, g->W, g->deg, g->batch,
                                                                           (int)g->N, g->k, P.kp, pn, pw);
    OSC_LAUNCH_CHECK("batched_pack_kernel");
    P.pk_nbr = pn;
    P.pk_w = pw;
  }
  P.prof = nullptr;
  if (P.debug & 16) {
    P.prof = ar.take<unsigned long long>(16);
    if (P.prof) OSC_CUDA(cudaMemsetAsync(P.prof, 0, 16 * sizeof(unsigned long long), st));
  }

  const size_t smem = batched_smem(g->N, g->k);
  int tpt = tpt_for(g->N);
  {
    const char* t = getenv("OSC_BATCHED_TPT");  // dev-only override
    if (t && atoi(t) >= tpt && atoi(t) <= 4) tpt = atoi(t);
  }
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
  if (P.prof != nullptr) {
    unsigned long long h[16];
    OSC_CUDA(cudaMemcpyAsync(h, P.prof, sizeof(h), cudaMemcpyDeviceToHost, st));
    OSC_CUDA(cudaStreamSynchronize(st));
    const char* names[16] = {"top", "stage", "load_settle", "store_load", "store_dh", "", "", "", "init", "A_spmm",
                             "bar1", "C_update", "bar2", "E_pupd", "", ""};
    unsigned long long tot = 0;
    for (int i = 0; i < 16; ++i) tot += h[i];
    for (int i = 0; i < 16; ++i)
      if (h[i]) fprintf(stderr, "[osc prof] %-12s %12llu clk %5.1f%%\n", names[i], h[i], 100.0 * h[i] / tot);
  }
  if (a->do_deltaH) {
    batched_dh_reduce_kernel<<<(unsigned)((g->batch + 127) / 128), 128, 0, st>>>(P.dh_part, P.G,
                                                                                  g->batch, a->deltaH);
    OSC_LAUNCH_CHECK("batched_dh_reduce_kernel");
  }
  return OSC_OK;
}

}  // namespace osc
