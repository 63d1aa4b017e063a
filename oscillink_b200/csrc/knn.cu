// K1 (CUDA-core engine) + canonical re-scoring.
//
//  normalize_rows_kernel : graph.py:35   Yn = Y / (||Y|| + 1e-12), plus the TF32 hi/lo split
//  knn_simt_kernel       : graph.py:36-37,59  S = Yn Yn^T (diag excluded) with a fused per-row
//                          top-kc list kept in shared memory -- the N x N matrix never exists
//  knn_rescore_kernel    : graph.py:46-52  canonical order (fp32(exact dot) desc, column asc)
//
// The tensor-core engine (knn_tc.cu) produces the same candidate lists; both feed
// knn_rescore_kernel, which is what fixes the final neighbour sets.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace osc {

// ---------------------------------------------------------------- normalise + TF32 split
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ Y, int64_t rows,
                                                             int D, float* __restrict__ Yn,
                                                             float* __restrict__ hi,
                                                             float* __restrict__ lo,
                                                             __half* __restrict__ h16) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* y = Y + row * D;
  double ss = 0.0;
  for (int d = lane; d < D; d += 32) {
    double v = (double)y[d];
    ss = fma(v, v, ss);
  }
  ss = warp_sum(ss);
  const float den = (float)sqrt(ss) + 1e-12f;
  for (int d = lane; d < D; d += 32) {
    const float v = __fdiv_rn(y[d], den);
    Yn[row * D + d] = v;
    if (h16 != nullptr) h16[row * D + d] = __float2half_rn(v);
    if (hi != nullptr) {
      const float h = to_tf32(v);
      hi[row * D + d] = h;
      if (lo != nullptr) lo[row * D + d] = to_tf32(v - h);
    }
  }
}

// ---------------------------------------------------------------- SIMT candidate kernel
constexpr int TM = 64, TN = 64, BK = 16;

__device__ __forceinline__ void list_insert(float* lv, int* li, int& cnt, int kc, float s, int j) {
  int p = (cnt < kc) ? cnt : kc - 1;
  while (p > 0 && lv[p - 1] < s) {  // strict: an equal score keeps the earlier (smaller) column
    lv[p] = lv[p - 1];
    li[p] = li[p - 1];
    --p;
  }
  lv[p] = s;
  li[p] = j;
  if (cnt < kc) ++cnt;
}

__global__ void __launch_bounds__(256)
knn_simt_kernel(const float* __restrict__ Yq, const float* __restrict__ Yall, int64_t n_rows,
                int64_t row0, int64_t N, int D, int kc, int kcp, int32_t* __restrict__ cand_idx,
                float* __restrict__ cand_sim) {
  extern __shared__ float smem[];
  float* As = smem;                 // [BK][TM+1]
  float* Bs = As + BK * (TM + 1);   // [BK][TN+1]
  float* Ss = Bs + BK * (TN + 1);   // [TM][TN+1]
  float* Lv = Ss + TM * (TN + 1);   // [TM][kcp]
  int* Li = reinterpret_cast<int*>(Lv + TM * kcp);

  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int64_t b = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * TM;  // first query row (panel-local numbering)
  const float* q = Yq + b * n_rows * D;
  const float* all = Yall + b * N * D;

  int cnt = 0;  // meaningful for t < TM only
  const int lm = t >> 2;         // tile row loaded by this thread
  const int lk = (t & 3) * 4;    // first k offset loaded by this thread

  for (int64_t c0 = 0; c0 < N; c0 += TN) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < D; k0 += BK) {
      // stage A (query rows) and B (column rows), k-major
      {
        const int64_t ra = r0 + lm;
        const int64_t rb = c0 + lm;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int kk = k0 + lk + u;
          As[(lk + u) * (TM + 1) + lm] = (ra < n_rows && kk < D) ? q[ra * D + kk] : 0.f;
          Bs[(lk + u) * (TN + 1) + lm] = (rb < N && kk < D) ? all[rb * D + kk] : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk * (TM + 1) + ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bb[j] = Bs[kk * (TN + 1) + tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Ss[(ty * 4 + i) * (TN + 1) + tx * 4 + j] = acc[i][j];
    __syncthreads();
    if (t < TM) {
      const int64_t gi = row0 + r0 + t;  // global id of this query row
      if (r0 + t < n_rows) {
        float* lv = Lv + t * kcp;
        int* li = Li + t * kcp;
        const int lim = (int)min((int64_t)TN, N - c0);
        for (int c = 0; c < lim; ++c) {
          const int64_t j = c0 + c;
          if (j == gi) continue;
          const float s = Ss[t * (TN + 1) + c];
          if (cnt < kc || s > lv[kc - 1]) list_insert(lv, li, cnt, kc, s, (int)j);
        }
      }
    }
    __syncthreads();
  }
  if (t < TM && r0 + t < n_rows) {
    const int64_t o = (b * n_rows + r0 + t) * kc;
    for (int c = 0; c < kc; ++c) {
      cand_idx[o + c] = (c < cnt) ? Li[t * kcp + c] : -1;
      cand_sim[o + c] = (c < cnt) ? Lv[t * kcp + c] : -INFINITY;
    }
  }
}

// ---------------------------------------------------------------- canonical re-scoring
// Totals of four per-lane fp64 partials at once: a butterfly that halves the number of live values per
// step (8 + 4 + 3*2 = 18 SHFL and 6 DADD instead of 4 x (10 + 5)).  Lane 8*i ends up with the total
// of a[i].  Every pair sum is the one warp_sum() forms (x_l + x_{l^off}, IEEE addition commutes), so
// the totals are bit-identical to four warp_sum() calls -- knn_exact_rows_kernel relies on that.
__device__ __forceinline__ double warp_sum4(const double (&a)[4], int lane) {
  const bool b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1;
  double k0 = b4 ? a[2] : a[0], k1 = b4 ? a[3] : a[1];
  const double s0 = b4 ? a[0] : a[2], s1 = b4 ? a[1] : a[3];
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  double k = b3 ? k1 : k0;
  const double s = b3 ? k0 : k1;
  k += __shfl_xor_sync(0xffffffffu, s, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;  // lane: total of a[(b4 << 1) | b3]
}

// One warp per query row.  dynamic smem: warps * kc * (float + int).
// QC > 0: D <= 128 * QC and D % 4 == 0 -- the query row is converted to fp64 once and kept in
// registers (the pass is issue-bound: conversions and shuffles, not L2 bandwidth).  QC == 0: any D.
template <int QC>
__global__ void __launch_bounds__(256, QC == 0 ? 4 : 2)
knn_rescore_kernel(const float* __restrict__ Yq, const float* __restrict__ Yall, int64_t n_rows,
                   int64_t N, int D, const int32_t* __restrict__ cand_idx, int kc, int k,
                   int32_t* __restrict__ top_idx, float* __restrict__ top_sim,
                   float* __restrict__ gap, const float* __restrict__ cand_sim, float eps,
                   int64_t* __restrict__ flagged, int* __restrict__ n_flagged) {
  extern __shared__ float smem[];
  const int warps = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sv = smem + (size_t)w * kc;
  int* sj = reinterpret_cast<int*>(smem + (size_t)warps * kc) + (size_t)w * kc;
  const int64_t b = blockIdx.y;
  const int64_t r = (int64_t)blockIdx.x * warps + w;
  if (r >= n_rows) return;
  const float* yi = Yq + (b * n_rows + r) * D;
  const float* all = Yall + b * N * D;
  const int32_t* ci = cand_idx + (b * n_rows + r) * kc;
  // Pruning with the engine's error bound (cand_sim sorted descending): k candidates score at least
  // cs[k-1] approximately, so the exact k-th score is >= cs[k-1] - eps, and likewise the exact
  // (k+1)-th is >= cs[k] - eps; a candidate below cs[k] - 2 eps is exactly below both and is dropped
  // without fetching its row.
  const float* cs = cand_sim != nullptr ? cand_sim + (b * n_rows + r) * kc : nullptr;
  int n_keep = kc;
  if (cs != nullptr && kc > k) {
    const float thr = cs[k] - 2.0f * eps;
    const unsigned m = __ballot_sync(0xffffffffu, lane < kc && cs[lane < kc ? lane : 0] >= thr);
    int cnt = __popc(m);
    for (int c = 32 + lane; c < kc; c += 32) cnt += (cs[c] >= thr) ? 1 : 0;  // kc > 32 only
    if (kc > 32) {
      int extra = cnt - __popc(m);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) extra += __shfl_xor_sync(0xffffffffu, extra, off);
      cnt = __popc(m) + extra;
    }
    n_keep = cnt < k + 1 ? k + 1 : cnt;  // a prefix of the sorted list; -inf tails carry idx -1
    if (n_keep > kc) n_keep = kc;
  }
  // Four candidates at a time: their row fetches are independent, so a lane keeps up to 4 x D/128
  // 16-byte loads in flight.  Lane l owns elements {4l..4l+3} + 128 t of every row, for (i,j) and
  // (j,i) alike: S stays symmetric.  Slots past n_keep re-read the query row (always valid memory,
  // L1 hit) so no load is predicated; their totals are discarded.
  const bool v4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(yi) | reinterpret_cast<uintptr_t>(all)) % 16 == 0);
  if (QC > 0 && v4) {
    constexpr int QN = QC > 0 ? QC : 1;
    double q[QN][4];
#pragma unroll
    for (int t = 0; t < QC; ++t) {
      const int d = lane * 4 + 128 * t;
      const float4 f = d < D ? *reinterpret_cast<const float4*>(yi + d) : make_float4(0.f, 0.f, 0.f, 0.f);
      q[t][0] = (double)f.x; q[t][1] = (double)f.y; q[t][2] = (double)f.z; q[t][3] = (double)f.w;
    }
    for (int c0 = 0; c0 < n_keep; c0 += 4) {
      int jc[4];
      const float* yj[4];
      double acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        jc[u] = (c0 + u < n_keep) ? ci[c0 + u] : -1;
        yj[u] = jc[u] >= 0 ? all + (int64_t)jc[u] * D : yi;
        acc[u] = 0.0;
      }
      float4 x[QN][4];
#pragma unroll
      for (int t = 0; t < QC; ++t) {
        const int d = lane * 4 + 128 * t;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          x[t][u] = d < D ? *reinterpret_cast<const float4*>(yj[u] + d) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int t = 0; t < QC; ++t) {
        if (lane * 4 + 128 * t < D) {  // same FMA sequence per lane as the generic path
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[u] = fma(q[t][0], (double)x[t][u].x, acc[u]);
            acc[u] = fma(q[t][1], (double)x[t][u].y, acc[u]);
            acc[u] = fma(q[t][2], (double)x[t][u].z, acc[u]);
            acc[u] = fma(q[t][3], (double)x[t][u].w, acc[u]);
          }
        }
      }
      const double tot = warp_sum4(acc, lane);
      const int u = lane >> 3;
      const int ju = u == 0 ? jc[0] : (u == 1 ? jc[1] : (u == 2 ? jc[2] : jc[3]));
      if ((lane & 7) == 0 && c0 + u < n_keep) {
        sv[c0 + u] = ju >= 0 ? (float)tot : -INFINITY;
        sj[c0 + u] = ju;
      }
    }
  } else {
    for (int c0 = 0; c0 < n_keep; c0 += 4) {
      int jc[4];
      const float* yj[4];
      double acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        jc[u] = (c0 + u < n_keep) ? ci[c0 + u] : -1;
        yj[u] = jc[u] >= 0 ? all + (int64_t)jc[u] * D : yi;
        acc[u] = 0.0;
      }
      if (v4) {
        for (int d = lane * 4; d < D; d += 128) {
          const float4 q = *reinterpret_cast<const float4*>(yi + d);
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<const float4*>(yj[u] + d);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[u] = fma((double)q.x, (double)x[u].x, acc[u]);
            acc[u] = fma((double)q.y, (double)x[u].y, acc[u]);
            acc[u] = fma((double)q.z, (double)x[u].z, acc[u]);
            acc[u] = fma((double)q.w, (double)x[u].w, acc[u]);
          }
        }
      } else {
        for (int d = lane; d < D; d += 32) {
          const double q = (double)yi[d];
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = fma(q, (double)yj[u][d], acc[u]);
        }
      }
      const double tot = warp_sum4(acc, lane);
      const int u = lane >> 3;
      const int ju = u == 0 ? jc[0] : (u == 1 ? jc[1] : (u == 2 ? jc[2] : jc[3]));
      if ((lane & 7) == 0 && c0 + u < n_keep) {
        sv[c0 + u] = ju >= 0 ? (float)tot : -INFINITY;
        sj[c0 + u] = ju;
      }
    }
  }
  __syncwarp();
  const int64_t o = (b * n_rows + r) * k;
  float kth = INFINITY, nxt = -INFINITY;
  for (int c = lane; c < n_keep; c += 32) {
    const float s = sv[c];
    const int j = sj[c];
    if (j < 0) continue;
    int rank = 0;
    for (int c2 = 0; c2 < n_keep; ++c2) {
      const int j2 = sj[c2];
      if (j2 >= 0 && c2 != c && better(sv[c2], j2, s, j)) ++rank;
    }
    if (rank < k) {
      top_idx[o + rank] = j;
      top_sim[o + rank] = s;
    }
    if (rank == k - 1) kth = s;
    if (rank == k) nxt = s;
  }
  // exactly one lane holds each of kth / nxt
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    kth = fminf(kth, __shfl_xor_sync(0xffffffffu, kth, off));
    nxt = fmaxf(nxt, __shfl_xor_sync(0xffffffffu, nxt, off));
  }
  if (gap != nullptr && lane == 0) gap[b * n_rows + r] = (nxt == -INFINITY) ? INFINITY : kth - nxt;
  // ---- completeness check of the candidate list.  Every column outside the list scored at most
  // a_min (the smallest APPROXIMATE score in the list) in the approximate pass, i.e. at most
  // a_min + eps exactly; it can only belong to the true top-k if that reaches the exact k-th
  // score.  Such rows (and rows with fewer than k valid candidates) are re-done exhaustively.
  if (cand_sim != nullptr && flagged != nullptr && (int64_t)kc < N - 1) {
    float amin = INFINITY;
    for (int c = lane; c < kc; c += 32)
      if (ci[c] >= 0) amin = fminf(amin, cs[c]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, off));
    // kth == +inf: fewer than k valid candidates
    if (lane == 0 && (kth == INFINITY || amin + eps >= kth)) flagged[atomicAdd(n_flagged, 1)] = b * n_rows + r;
  }
}

// ---------------------------------------------------------------- re-scoring without duplicate dots
// In a square problem (every row of a lattice is a query) a mutual candidate pair is scored twice by
// knn_rescore_kernel: dot(i, j) by row i and dot(j, i) by row j -- the same fp64 sum, element for element
// (the lane ownership of the elements is identical), and ~80 % of the kept candidates are mutual.  Here the
// pair is computed ONCE, by the row with the smaller index:
//   row i OWNS candidate j  iff  j > i  or  i is not among row j's kept candidates
// (if j < i and i is kept by row j, then row j owns the pair because i > j).  Pass A computes the owned dots
// and stores them in S[row][slot]; a slot it does not own receives a NaN whose payload is the slot of i in
// row j's list.  Pass B resolves those slots from S[j][.] and ranks exactly as knn_rescore_kernel does.
// The set of row fetches drops by ~40 %; results are bit-identical (tests: OSC_RESCORE_DEDUP=0 vs 1).
// slot of `self` among the kept candidates of another row (-1: not kept).  The row's list is read with KC/4
// independent 16-byte loads per array (a runtime loop serialises them: one L2 round trip per entry).
template <int KC>
__device__ __forceinline__ int kept_slot_of(const int32_t* __restrict__ cj, const float* __restrict__ csj, int k,
                                            float eps, int32_t self) {
  int32_t id[KC];
  float sc[KC];
#pragma unroll
  for (int q = 0; q < KC / 4; ++q) {
    const int4 a = reinterpret_cast<const int4*>(cj)[q];
    const float4 f = reinterpret_cast<const float4*>(csj)[q];
    id[4 * q] = a.x; id[4 * q + 1] = a.y; id[4 * q + 2] = a.z; id[4 * q + 3] = a.w;
    sc[4 * q] = f.x; sc[4 * q + 1] = f.y; sc[4 * q + 2] = f.z; sc[4 * q + 3] = f.w;
  }
  const float thr = k < KC ? csj[k] - 2.0f * eps : -INFINITY;  // (a scalar load: indexing sc[] with the
                                                                 //  run-time k would move it to local memory)
  int cnt = 0, pos = -1;
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    cnt += (sc[c] >= thr) ? 1 : 0;
    if (id[c] == self) pos = c;
  }
  int nk = cnt < k + 1 ? k + 1 : cnt;
  if (nk > KC) nk = KC;
  return pos < nk ? pos : -1;
}

template <int KC>
__global__ void __launch_bounds__(256, 4)
knn_rescore_owned_kernel(const float* __restrict__ Yall, int64_t N, int D, const int32_t* __restrict__ cand_idx,
                         const float* __restrict__ cand_sim, int k, float eps, float* __restrict__ S) {
  constexpr int kc = KC;
  extern __shared__ float smem[];
  const int warps = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* own_j = reinterpret_cast<int*>(smem) + (size_t)w * 2 * kc;  // [kc] owned columns, [kc] their slots
  int* own_c = own_j + kc;
  const int64_t b = blockIdx.y;
  const int64_t r = (int64_t)blockIdx.x * warps + w;
  if (r >= N) return;
  const float* all = Yall + b * N * D;
  const float* yi = all + r * D;
  const int32_t* ci = cand_idx + (b * N + r) * kc;
  const float* cs = cand_sim + (b * N + r) * kc;
  float* Si = S + (b * N + r) * kc;
  // the row's own kept prefix (same rule as knn_rescore_kernel)
  int n_keep = kc;
  if (kc > k) {
    const float thr = cs[k] - 2.0f * eps;
    const unsigned m = __ballot_sync(0xffffffffu, lane < kc && cs[lane < kc ? lane : 0] >= thr);
    const int cnt = __popc(m);
    n_keep = cnt < k + 1 ? k + 1 : cnt;
    if (n_keep > kc) n_keep = kc;
  }
  // ownership of slot `lane`
  int j = -1, pos = -1;
  bool owned = false;
  if (lane < n_keep) {
    j = ci[lane];
    if (j >= 0) {
      owned = true;
      if (j < r) {
        pos = kept_slot_of<KC>(cand_idx + (b * N + j) * kc, cand_sim + (b * N + j) * kc, k, eps, (int32_t)r);
        owned = pos < 0;
      }
    }
  }
  if (lane < kc && !owned) {  // (owned slots are written once, by the dot loop below)
    float v = -INFINITY;      // slots beyond the kept prefix / invalid columns
    if (lane < n_keep && j >= 0) v = __int_as_float(0x7fc00000 | pos);  // NaN, payload = slot in row j
    Si[lane] = v;
  }
  const unsigned om = __ballot_sync(0xffffffffu, owned);
  const int n_own = __popc(om);
  if (owned) {
    const int q = __popc(om & ((1u << lane) - 1u));
    own_j[q] = j;
    own_c[q] = lane;
  }
  __syncwarp();
  const bool v4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(all)) % 16 == 0);
  for (int c0 = 0; c0 < n_own; c0 += 4) {
    int jc[4];
    const float* yj[4];
    double acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      jc[u] = (c0 + u < n_own) ? own_j[c0 + u] : -1;
      yj[u] = jc[u] >= 0 ? all + (int64_t)jc[u] * D : yi;
      acc[u] = 0.0;
    }
    if (v4) {  // same element ownership / order per lane as knn_rescore_kernel: bit-identical sums
      for (int d = lane * 4; d < D; d += 128) {
        const float4 q = *reinterpret_cast<const float4*>(yi + d);
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<const float4*>(yj[u] + d);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc[u] = fma((double)q.x, (double)x[u].x, acc[u]);
          acc[u] = fma((double)q.y, (double)x[u].y, acc[u]);
          acc[u] = fma((double)q.z, (double)x[u].z, acc[u]);
          acc[u] = fma((double)q.w, (double)x[u].w, acc[u]);
        }
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        const double q = (double)yi[d];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fma(q, (double)yj[u][d], acc[u]);
      }
    }
    const double tot = warp_sum4(acc, lane);
    const int u = lane >> 3;
    if ((lane & 7) == 0 && c0 + u < n_own) Si[own_c[c0 + u]] = (float)tot;
  }
}

__global__ void __launch_bounds__(256)
knn_rescore_rank_kernel(int64_t N, const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_sim, int kc,
                        int k, float eps, const float* __restrict__ S, int32_t* __restrict__ top_idx,
                        float* __restrict__ top_sim, float* __restrict__ gap, int64_t* __restrict__ flagged,
                        int* __restrict__ n_flagged) {
  extern __shared__ float smem[];
  const int warps = blockDim.x >> 5;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sv = smem + (size_t)w * kc;
  int* sj = reinterpret_cast<int*>(smem + (size_t)warps * kc) + (size_t)w * kc;
  const int64_t b = blockIdx.y;
  const int64_t r = (int64_t)blockIdx.x * warps + w;
  if (r >= N) return;
  const int32_t* ci = cand_idx + (b * N + r) * kc;
  const float* cs = cand_sim + (b * N + r) * kc;
  const float* Si = S + (b * N + r) * kc;
  int n_keep = kc;
  if (kc > k) {
    const float thr = cs[k] - 2.0f * eps;
    const unsigned m = __ballot_sync(0xffffffffu, lane < kc && cs[lane < kc ? lane : 0] >= thr);
    const int cnt = __popc(m);
    n_keep = cnt < k + 1 ? k + 1 : cnt;
    if (n_keep > kc) n_keep = kc;
  }
  if (lane < n_keep) {
    const int j = ci[lane];
    float v = Si[lane];
    if (v != v) v = S[(b * N + j) * kc + (__float_as_int(v) & 0xff)];  // the pair was scored by row j
    sv[lane] = j >= 0 ? v : -INFINITY;
    sj[lane] = j;
  }
  __syncwarp();
  const int64_t o = (b * N + r) * k;
  float kth = INFINITY, nxt = -INFINITY;
  for (int c = lane; c < n_keep; c += 32) {
    const float s = sv[c];
    const int j = sj[c];
    if (j < 0) continue;
    int rank = 0;
    for (int c2 = 0; c2 < n_keep; ++c2) {
      const int j2 = sj[c2];
      if (j2 >= 0 && c2 != c && better(sv[c2], j2, s, j)) ++rank;
    }
    if (rank < k) {
      top_idx[o + rank] = j;
      top_sim[o + rank] = s;
    }
    if (rank == k - 1) kth = s;
    if (rank == k) nxt = s;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    kth = fminf(kth, __shfl_xor_sync(0xffffffffu, kth, off));
    nxt = fmaxf(nxt, __shfl_xor_sync(0xffffffffu, nxt, off));
  }
  if (gap != nullptr && lane == 0) gap[b * N + r] = (nxt == -INFINITY) ? INFINITY : kth - nxt;
  if (flagged != nullptr && (int64_t)kc < N - 1) {  // completeness check, as in knn_rescore_kernel
    float amin = INFINITY;
    for (int c = lane; c < kc; c += 32)
      if (ci[c] >= 0) amin = fminf(amin, cs[c]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, off));
    if (lane == 0 && (kth == INFINITY || amin + eps >= kth)) flagged[atomicAdd(n_flagged, 1)] = b * N + r;
  }
}

// ---------------------------------------------------------------- exhaustive fallback rows
// One CTA per flagged row (grid-stride over the flagged list): every warp scans a strided subset of
// the N columns with the same fp64-accumulated dot as the re-scoring pass and keeps its k+1 best in
// shared memory; warp 0 merges the lists and writes the canonical top-k (+ gap).
__global__ void __launch_bounds__(256)
knn_exact_rows_kernel(const float* __restrict__ Yq, const float* __restrict__ Yall, int64_t n_rows,
                      int64_t row0, int64_t N, int D, int k, const int64_t* __restrict__ flagged,
                      const int* __restrict__ n_flagged, int32_t* __restrict__ top_idx,
                      float* __restrict__ top_sim, float* __restrict__ gap, int64_t limit) {
  extern __shared__ float smem[];
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = k + 1;
  float* lv = smem + (size_t)w * L;
  int* li = reinterpret_cast<int*>(smem + (size_t)warps * L) + (size_t)w * L;
  int* head = reinterpret_cast<int*>(smem + (size_t)2 * warps * L);  // merge cursors [warps]
  const int total = *n_flagged;
  // more rows than the caller is willing to scan exhaustively (N*D fp64 FMAs each): leave them with their
  // candidate-list result; the caller sees the count and re-runs the candidate pass with a tighter engine
  if (limit >= 0 && (int64_t)total > limit) return;
  for (int f = blockIdx.x; f < total; f += gridDim.x) {
    const int64_t gid = flagged[f];
    const int64_t b = gid / n_rows, r = gid - b * n_rows;
    const int64_t self = row0 + r;
    const float* yi = Yq + gid * D;
    const float* all = Yall + b * N * D;
    const bool v4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(yi) | reinterpret_cast<uintptr_t>(all)) % 16 == 0);
    int cnt = 0;
    // four columns per step: their row fetches are independent (the scan is latency-bound), the four
    // totals come out of one butterfly (bit-identical to warp_sum, see warp_sum4)
    for (int64_t j0 = w; j0 < N; j0 += 4 * warps) {
      int64_t jj[4];
      const float* yj[4];
      double acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        jj[u] = j0 + (int64_t)u * warps;
        const bool ok = jj[u] < N && jj[u] != self;
        if (!ok) jj[u] = -1;
        yj[u] = ok ? all + jj[u] * D : yi;
        acc[u] = 0.0;
      }
      if (v4) {  // same element ownership / order per lane as knn_rescore_kernel: bit-identical scores
        for (int d = lane * 4; d < D; d += 128) {
          const float4 q = *reinterpret_cast<const float4*>(yi + d);
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<const float4*>(yj[u] + d);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[u] = fma((double)q.x, (double)x[u].x, acc[u]);
            acc[u] = fma((double)q.y, (double)x[u].y, acc[u]);
            acc[u] = fma((double)q.z, (double)x[u].z, acc[u]);
            acc[u] = fma((double)q.w, (double)x[u].w, acc[u]);
          }
        }
      } else {
        for (int d = lane; d < D; d += 32) {
          const double q = (double)yi[d];
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = fma(q, (double)yj[u][d], acc[u]);
        }
      }
      const double tot = warp_sum4(acc, lane);  // lane 8u holds the total of column u
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float sc = (float)__shfl_sync(0xffffffffu, tot, 8 * u);
        const int64_t j = jj[u];
        if (lane == 0 && j >= 0 && (cnt < L || better(sc, (int)j, lv[L - 1], li[L - 1]))) {
          int p = (cnt < L) ? cnt : L - 1;
          while (p > 0 && better(sc, (int)j, lv[p - 1], li[p - 1])) {
            lv[p] = lv[p - 1];
            li[p] = li[p - 1];
            --p;
          }
          lv[p] = sc;
          li[p] = (int)j;
          if (cnt < L) ++cnt;
        }
      }
    }
    if (lane == 0) {
      for (int c = cnt; c < L; ++c) {
        lv[c] = -INFINITY;
        li[c] = 0x7fffffff;
      }
      head[w] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const float* av = smem;
      const int* ai = reinterpret_cast<const int*>(smem + (size_t)warps * L);
      const int64_t o = gid * k;
      float kth = INFINITY, nxt = -INFINITY;
      for (int rank = 0; rank < L; ++rank) {
        int best = -1;
        for (int q = 0; q < warps; ++q) {
          if (head[q] >= L || ai[q * L + head[q]] == 0x7fffffff) continue;
          if (best < 0 || better(av[q * L + head[q]], ai[q * L + head[q]], av[best * L + head[best]],
                                 ai[best * L + head[best]]))
            best = q;
        }
        if (best < 0) break;
        const float sv = av[best * L + head[best]];
        const int jv = ai[best * L + head[best]];
        head[best]++;
        if (rank < k) {
          top_idx[o + rank] = jv;
          top_sim[o + rank] = sv;
        }
        if (rank == k - 1) kth = sv;
        if (rank == k) nxt = sv;
      }
      if (gap != nullptr) gap[gid] = (nxt == -INFINITY) ? INFINITY : kth - nxt;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- host launchers
int launch_normalize(const float* Y, int64_t rows, int D, float* Yn, float* hi, float* lo,
                     cudaStream_t st, void* h16) {
  if (rows == 0) return OSC_OK;
  const int warps = 8;
  const int64_t blocks = (rows + warps - 1) / warps;
  normalize_rows_kernel<<<(unsigned)blocks, warps * 32, 0, st>>>(Y, rows, D, Yn, hi, lo,
                                                                  reinterpret_cast<__half*>(h16));
  OSC_LAUNCH_CHECK("normalize_rows_kernel");
  return OSC_OK;
}

int launch_knn_simt(const float* Yq, const float* Yall, int64_t batch, int64_t n_rows, int64_t row0,
                    int64_t N, int D, int kc, int32_t* cand_idx, float* cand_sim, cudaStream_t st) {
  const int kcp = kc | 1;
  const size_t smem = sizeof(float) * (BK * (TM + 1) + BK * (TN + 1) + TM * (TN + 1)) +
                      (size_t)TM * kcp * (sizeof(float) + sizeof(int));
  if (smem > 48 * 1024)
    OSC_CUDA(cudaFuncSetAttribute(knn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  dim3 grid((unsigned)((n_rows + TM - 1) / TM), (unsigned)batch);
  knn_simt_kernel<<<grid, 256, smem, st>>>(Yq, Yall, n_rows, row0, N, D, kc, kcp, cand_idx,
                                           cand_sim);
  OSC_LAUNCH_CHECK("knn_simt_kernel");
  return OSC_OK;
}

// kc == 16: one HALF warp per row, everything in registers.  The warp-per-row kernel above keeps ~10 of 32
// lanes busy and goes through shared memory for every comparison (ncu: 409 warp instructions per row, issue
// slots 85 % busy -- instruction bound); here lane c of a half warp holds candidate c and reads the others
// with width-16 shuffles.  Same ranking rule, same outputs.
__global__ void __launch_bounds__(256)
knn_rescore_rank16_kernel(int64_t N, const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_sim,
                          int k, float eps, const float* __restrict__ S, int32_t* __restrict__ top_idx,
                          float* __restrict__ top_sim, float* __restrict__ gap, int64_t* __restrict__ flagged,
                          int* __restrict__ n_flagged) {
  constexpr int KC = 16;
  const int lane = threadIdx.x & 31, hl = lane & 15, half = lane >> 4;
  const int64_t b = blockIdx.y;
  const int64_t r = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + half;
  const bool row_ok = r < N;  // no early return: the shuffles below are warp-wide
  const int64_t base = (b * N + (row_ok ? r : 0)) * KC;
  const int j = row_ok ? cand_idx[base + hl] : -1;
  const float a = row_ok ? cand_sim[base + hl] : -INFINITY;
  // candidates that can still be among the k best: approximate score within 2 eps of the (k+1)-th (the list
  // is sorted by approximate score, so they form a prefix); at least k + 1 of them
  const float a_k = __shfl_sync(0xffffffffu, a, k, 16);
  const unsigned m = (__ballot_sync(0xffffffffu, a >= a_k - 2.0f * eps) >> (16 * half)) & 0xffffu;
  int n_keep = __popc(m);
  n_keep = n_keep < k + 1 ? k + 1 : n_keep;
  const bool kept = row_ok && hl < n_keep && j >= 0;
  float s = -INFINITY;
  if (kept) {
    s = S[base + hl];
    if (s != s) s = S[(b * N + j) * KC + (__float_as_int(s) & 0xff)];  // the pair was scored by row j
  }
  int rank = 0;
#pragma unroll
  for (int c2 = 0; c2 < KC; ++c2) {
    const float s2 = __shfl_sync(0xffffffffu, s, c2, 16);
    const int j2 = __shfl_sync(0xffffffffu, kept ? j : -1, c2, 16);
    if (j2 >= 0 && c2 != hl && better(s2, j2, s, j)) ++rank;
  }
  float kth = INFINITY, nxt = -INFINITY;
  if (kept) {
    const int64_t o = (b * N + r) * k;
    if (rank < k) {
      top_idx[o + rank] = j;
      top_sim[o + rank] = s;
    }
    if (rank == k - 1) kth = s;
    if (rank == k) nxt = s;
  }
  float amin = (row_ok && j >= 0) ? a : INFINITY;
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) {
    kth = fminf(kth, __shfl_xor_sync(0xffffffffu, kth, off, 16));
    nxt = fmaxf(nxt, __shfl_xor_sync(0xffffffffu, nxt, off, 16));
    amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, off, 16));
  }
  if (row_ok && hl == 0) {
    if (gap != nullptr) gap[b * N + r] = (nxt == -INFINITY) ? INFINITY : kth - nxt;
    if (flagged != nullptr && (int64_t)KC < N - 1 && (kth == INFINITY || amin + eps >= kth))
      flagged[atomicAdd(n_flagged, 1)] = b * N + r;  // completeness check, as in knn_rescore_kernel
  }
}

// cand_sim / flagged / n_flagged may be NULL (no completeness check, the plain osc_knn_rescore).
// With them: rows whose candidate list cannot be proven complete are recomputed exhaustively.
int launch_rescore(const float* Yq, const float* Yall, int64_t batch, int64_t n_rows, int64_t row0,
                   int64_t N, int D, const int32_t* cand_idx, const float* cand_sim, int kc, int k,
                   float eps, int32_t* top_idx, float* top_sim, float* gap, int64_t* flagged,
                   int* n_flagged, cudaStream_t st, int64_t exhaustive_limit, float* dedup_S) {
  const int warps = 8;
  const size_t smem = (size_t)warps * kc * (sizeof(float) + sizeof(int));
  dim3 grid((unsigned)((n_rows + warps - 1) / warps), (unsigned)batch);
  const bool check = cand_sim != nullptr && flagged != nullptr && n_flagged != nullptr;
  if (check) OSC_CUDA(cudaMemsetAsync(n_flagged, 0, sizeof(int), st));
  // Measured on B200 (4096 lattices N=1200 D=384, kc=16): the 64-register generic variant with 32
  // resident warps per SM takes 10.6 ms, the variant that keeps the fp64 query row in registers
  // (128 registers, 16 warps) 16.0 ms -- latency hiding wins over the saved conversions.  Groups of 5
  // candidates (two groups instead of three for most rows, 80 registers, 24 warps) measured 12.8 ms.
  bool dedup = dedup_S != nullptr && check && n_rows == N && row0 == 0 && Yq == Yall && kc > k &&
               (kc == 16 || kc == 24 || kc == 32) && (reinterpret_cast<uintptr_t>(cand_idx) % 16 == 0) &&
               (reinterpret_cast<uintptr_t>(cand_sim) % 16 == 0);
  {
    const char* e = getenv("OSC_RESCORE_DEDUP");  // dev-only A/B switch
    if (e && atoi(e) == 0) dedup = false;
  }
  if (dedup) {
    const size_t sm_a = (size_t)warps * 2 * kc * sizeof(int);
    if (kc == 16)
      knn_rescore_owned_kernel<16><<<grid, warps * 32, sm_a, st>>>(Yall, N, D, cand_idx, cand_sim, k, eps, dedup_S);
    else if (kc == 24)
      knn_rescore_owned_kernel<24><<<grid, warps * 32, sm_a, st>>>(Yall, N, D, cand_idx, cand_sim, k, eps, dedup_S);
    else
      knn_rescore_owned_kernel<32><<<grid, warps * 32, sm_a, st>>>(Yall, N, D, cand_idx, cand_sim, k, eps, dedup_S);
    OSC_LAUNCH_CHECK("knn_rescore_owned_kernel");
    bool rank16 = kc == 16 && k < 16;
    {
      const char* e = getenv("OSC_RESCORE_RANK16");  // dev-only A/B switch
      if (e && atoi(e) == 0) rank16 = false;
    }
    if (rank16) {
      dim3 grid2((unsigned)((n_rows + 2 * warps - 1) / (2 * warps)), (unsigned)batch);
      knn_rescore_rank16_kernel<<<grid2, warps * 32, 0, st>>>(N, cand_idx, cand_sim, k, eps, dedup_S, top_idx, top_sim,
                                                            gap, flagged, n_flagged);
    } else {
      knn_rescore_rank_kernel<<<grid, warps * 32, smem, st>>>(N, cand_idx, cand_sim, kc, k, eps, dedup_S, top_idx,
                                                              top_sim, gap, flagged, n_flagged);
    }
    OSC_LAUNCH_CHECK("knn_rescore_rank_kernel");
  }
  auto fn = knn_rescore_kernel<0>;
  {
    const char* e = getenv("OSC_RESCORE_HOIST");  // dev-only A/B switch
    if (e && atoi(e) != 0)
      fn = D <= 128 ? knn_rescore_kernel<1>
           : D <= 256 ? knn_rescore_kernel<2>
           : D <= 384 ? knn_rescore_kernel<3> : knn_rescore_kernel<0>;
  }
  if (!dedup) {
    fn<<<grid, warps * 32, smem, st>>>(Yq, Yall, n_rows, N, D, cand_idx, kc, k, top_idx, top_sim, gap,
                                       check ? cand_sim : nullptr, eps, check ? flagged : nullptr, n_flagged);
    OSC_LAUNCH_CHECK("knn_rescore_kernel");
  }
  if (check && (int64_t)kc < N - 1) {
    const size_t sm2 = (size_t)warps * (k + 1) * (sizeof(float) + sizeof(int)) + warps * sizeof(int);
    int64_t blocks = batch * n_rows;
    const int64_t cap = 2 * (int64_t)sm_count();
    if (blocks > cap) blocks = cap;
    knn_exact_rows_kernel<<<(unsigned)blocks, warps * 32, sm2, st>>>(Yq, Yall, n_rows, row0, N, D, k, flagged,
                                                                    n_flagged, top_idx, top_sim, gap,
                                                                    exhaustive_limit);
    OSC_LAUNCH_CHECK("knn_exact_rows_kernel");
  }
  return OSC_OK;
}

}  // namespace osc
