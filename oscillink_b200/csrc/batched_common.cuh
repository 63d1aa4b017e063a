// Shared device helpers of the batched serving kernels (batched.cu, batched_ms.cu).
#pragma once
#include "common.cuh"

namespace osc {

constexpr int SC = 4;  // columns per slab (one float4 per row)
constexpr int PK_MAXK = 16;
constexpr int RED_F4 = 32;  // float4 slots per reduction array (>= warps per CTA)
constexpr int SCR_CTAS_PER_SM = 8;  // scratch rows are provisioned for this many resident CTAs per SM

struct BatchedK {
  const float* Y;
  const float* U_in;
  const float* psi;
  const float* gates;
  float* U_out;
  float* Ustar_out;
  double* dh_part;               // [batch][G]
  int2* rec;                     // [batch][2][G] {iterations, float bits of max_c ||r_c||^2}
  const int4* fix_list;          // list mode: {lattice, slab, forced settle its, forced U* its}
  const int* fix_count;
  float4* scratch;               // [gridDim.x][2][N]: gather sums of Y (shared by the two initial
                                 // residuals) and (U_in - U - r_settle)/dt for the deltaH identity
  const unsigned short* pk_nbr;  // packed graph image [batch][kq][N] ushort4 (byte offsets j*16)
  const float* pk_w;             // [batch][kq][N] float4
  int64_t batch, n_work;
  int N, kq, D, G, CH, cpl;
  int do_settle, do_ustar, do_dh;
  int use_ybuf;  // the kernel was given Np float4 of shared memory for the Y-slab prefetch buffer
  float lamG, lamC, lamQ, dt;
  double tol_settle, tol_ustar;
  float thr2_settle, thr2_ustar;  // batched_sq_threshold(tol): the stop tests on squared norms (batched_ms.cu)
  int max_iters_settle, max_iters_ustar;
};

struct SolveCoef {
  float diag0, diag1;  // operator diagonal = diag0 + diag1 * b_i
  float diag_u, im_u;  // the same diagonal and 1/(Mdiag + 1e-12) for b_i == 1 (no gates given)
  float offc;
  float lamG, lamQ, dt;
  int settle;
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

// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only: the other half of the 32-byte sector
// belongs to the neighbouring slab and is picked up from L2 by whoever runs that one)
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- block reductions of per-column partials ---------------------------------------------------
// All 32 lanes of a warp hold partials of the SAME 4 columns.  The butterfly halves the number of
// live values at every step (2 + 1 + 3 = 6 SHFL for 4 values, 4 + 2 + 1 + 2 = 9 for 8), the
// per-warp totals go to shared memory, and after ONE barrier every thread sums the nw partials
// itself with broadcast LDS.128 (same order in every thread -> identical, deterministic totals).
__device__ __forceinline__ void warp_reduce4(float4 v, float4* red_w, int lane) {
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1;
  const float s0 = b4 ? v.x : v.z, s1 = b4 ? v.y : v.w;
  float k0 = b4 ? v.z : v.x, k1 = b4 ? v.w : v.y;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  const float s = b3 ? k0 : k1;
  float k = b3 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, s, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  if ((lane & 7) == 0) reinterpret_cast<float*>(red_w)[b4 * 2 + b3] = k;  // component b4*2+b3
}
__device__ __forceinline__ void warp_reduce8(float4 a, float4 b, float4* red_a, float4* red_b, int lane) {
  const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1;
  const float4 s = b4 ? a : b;
  float4 k = b4 ? b : a;
  k.x += __shfl_xor_sync(0xffffffffu, s.x, 16);
  k.y += __shfl_xor_sync(0xffffffffu, s.y, 16);
  k.z += __shfl_xor_sync(0xffffffffu, s.z, 16);
  k.w += __shfl_xor_sync(0xffffffffu, s.w, 16);
  const float s0 = b3 ? k.x : k.z, s1 = b3 ? k.y : k.w;
  float k0 = b3 ? k.z : k.x, k1 = b3 ? k.w : k.y;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 8);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 8);
  const float ss = b2 ? k0 : k1;
  float kk = b2 ? k1 : k0;
  kk += __shfl_xor_sync(0xffffffffu, ss, 4);
  kk += __shfl_xor_sync(0xffffffffu, kk, 2);
  kk += __shfl_xor_sync(0xffffffffu, kk, 1);
  // b4 picks the vector (0: a, 1: b), component = b3*2 + b2
  if ((lane & 3) == 0) reinterpret_cast<float*>(b4 ? red_b : red_a)[b3 * 2 + b2] = kk;
}
// Total of the nw per-warp partials, identical in every thread: lane l fetches component l&3 of
// warps l>>2, (l>>2)+8, ... (conflict-free LDS.32), a 3-step butterfly over lane bits 2-4 finishes
// the sum, and 4 SHFL hand every lane all four components.  (A broadcast LDS.128 costs 2 crossbar
// wavefronts; reading all nw partials per thread was 15 % of the kernel's shared-memory traffic.)
__device__ __forceinline__ float4 block_total(const float4* red, int nw, int lane) {
  const float* rf = reinterpret_cast<const float*>(red);
  float t = 0.f;
  for (int i = lane; i < nw * 4; i += 32) t += rf[i];
  t += __shfl_xor_sync(0xffffffffu, t, 4);
  t += __shfl_xor_sync(0xffffffffu, t, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  return make_float4(__shfl_sync(0xffffffffu, t, 0), __shfl_sync(0xffffffffu, t, 1),
                     __shfl_sync(0xffffffffu, t, 2), __shfl_sync(0xffffffffu, t, 3));
}

// Component-local variant: lane l gets the total of component l&3 only (no broadcast).  The CG scalars
// (rz, p.Ap, r.z') are per column, so each lane divides for ITS column and bcast4 hands the four
// quotients to everybody: one division per lane instead of four, same values as before.
__device__ __forceinline__ float block_total_c(const float4* red, int nw, int lane) {
  const float* rf = reinterpret_cast<const float*>(red);
  float t = 0.f;
  for (int i = lane; i < nw * 4; i += 32) t += rf[i];
  t += __shfl_xor_sync(0xffffffffu, t, 4);
  t += __shfl_xor_sync(0xffffffffu, t, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 16);
  return t;
}
__device__ __forceinline__ float4 bcast4(float q) {
  return make_float4(__shfl_sync(0xffffffffu, q, 0), __shfl_sync(0xffffffffu, q, 1),
                     __shfl_sync(0xffffffffu, q, 2), __shfl_sync(0xffffffffu, q, 3));
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2) ----------------------------------------
// The kernel issues about as many instructions as its shared-memory pipe can take wavefronts, so the
// float4 arithmetic runs on register PAIRS: fma.rn.f32x2 is two IEEE fp32 FMAs in one issue slot
// (bit-identical to fmaf per lane), and a scalar weight enters as the {w, w} broadcast operand.
typedef unsigned long long u64;
struct V4 {
  u64 lo, hi;  // {x, y}, {z, w}
};
__device__ __forceinline__ u64 pk2(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 upk2(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ V4 v4_zero() { return V4{0ull, 0ull}; }
__device__ __forceinline__ V4 to_v4(float4 f) { return V4{pk2(f.x, f.y), pk2(f.z, f.w)}; }
__device__ __forceinline__ float4 to_f4(V4 v) {
  const float2 a = upk2(v.lo), b = upk2(v.hi);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ V4 v4_bc(float s) {
  const u64 t = pk2(s, s);
  return V4{t, t};
}
__device__ __forceinline__ V4 v4_fma(V4 a, V4 b, V4 c) { return V4{fma2(a.lo, b.lo, c.lo), fma2(a.hi, b.hi, c.hi)}; }
__device__ __forceinline__ V4 v4_fma_s(float s, V4 b, V4 c) {
  const u64 t = pk2(s, s);
  return V4{fma2(t, b.lo, c.lo), fma2(t, b.hi, c.hi)};
}
__device__ __forceinline__ V4 v4_mul(V4 a, V4 b) { return V4{mul2(a.lo, b.lo), mul2(a.hi, b.hi)}; }
__device__ __forceinline__ V4 v4_mul_s(float s, V4 b) {
  const u64 t = pk2(s, s);
  return V4{mul2(t, b.lo), mul2(t, b.hi)};
}
__device__ __forceinline__ V4 v4_sub(V4 a, V4 b) { return V4{sub2(a.lo, b.lo), sub2(a.hi, b.hi)}; }
__device__ __forceinline__ V4 lds_v4(const void* p) {
  const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p);
  return V4{t.x, t.y};
}
__device__ __forceinline__ void sts_v4(void* p, V4 v) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(v.lo, v.hi); }

template <int TPT>
struct Slab {
  V4 X[TPT], R[TPT], AP[TPT];
};

// sum_t W_t p[nbr_t] for one row; graph image is slot-major [c][row]
template <int KQ>
__device__ __forceinline__ V4 gather_row(const float4* p_s, const ushort4* nbr_s, const float4* w_s,
                                         int row, int N, int kq_rt) {
  V4 acc = v4_zero();
  const int kq = KQ > 0 ? KQ : kq_rt;
  const char* pb = reinterpret_cast<const char*>(p_s);
#pragma unroll
  for (int c = 0; c < kq; ++c) {
    const ushort4 jj = nbr_s[c * N + row];
    const float4 ww = w_s[c * N + row];
    acc = v4_fma_s(ww.x, lds_v4(pb + jj.x), acc);
    acc = v4_fma_s(ww.y, lds_v4(pb + jj.y), acc);
    acc = v4_fma_s(ww.z, lds_v4(pb + jj.z), acc);
    acc = v4_fma_s(ww.w, lds_v4(pb + jj.w), acc);
  }
  return acc;
}
// diag * own - offc * acc   (noffc = -offc)
__device__ __forceinline__ V4 combine_row(V4 own, V4 acc, float diag, float noffc) {
  return v4_fma_s(noffc, acc, v4_mul_s(diag, own));
}
// A(p) for one row: diag*p_own - offc * sum_t W_t p[nbr_t]
template <int KQ>
__device__ __forceinline__ V4 apply_row(const float4* p_s, const ushort4* nbr_s, const float4* w_s,
                                        int row, int N, int kq_rt, float diag, float noffc) {
  return combine_row(lds_v4(p_s + row), gather_row<KQ>(p_s, nbr_s, w_s, row, N, kq_rt), diag, noffc);
}

// kernel entry type of the slab kernels and the selector of the multi-shift variant (batched_ms.cu)
typedef void (*BatchedFn)(BatchedK);
BatchedFn batched_ms_pick(int64_t N, int kq, int variant, int* threads, size_t* smem_dyn, bool* two_ctas);
float batched_sq_threshold(double tol);

}  // namespace osc
