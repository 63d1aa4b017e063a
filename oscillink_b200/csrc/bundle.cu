// f1/f2: bundle scoring + greedy MMR, and the pair distances chain_receipt needs.
//
//   lattice.py:557-566  align_i = <U*_i / (||U*_i|| + 1e-12), psi / (||psi|| + 1e-12)>
//   graph.py:114-133    greedy MMR: val_i = 0.5*score_i - 0.5*max_{j in chosen} cos(Y_i, Y_j),
//                       strict '>' over ascending i  => lowest index wins ties
//   lattice.py:468-471  d2_ij = || U*_i/(sd_i+1e-12) - U*_j/(sd_j+1e-12) ||^2  for given (i,j) pairs
//
// The reference forms the full N x N cosine matrix for MMR; only k of its columns are ever read,
// so each greedy step computes one column (N dot products) and an arg-max: O(k N D) total.
#include <vector>

#include "common.cuh"

namespace osc {

__global__ void __launch_bounds__(256)
row_align_kernel(const float* __restrict__ Us, const float* __restrict__ psi, int64_t N, int D,
                 float* __restrict__ align) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const float* u = Us + i * D;
  float uu = 0.f, pp = 0.f, up = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float a = u[d], b = psi[d];
    uu = fmaf(a, a, uu);
    pp = fmaf(b, b, pp);
    up = fmaf(a, b, up);
  }
  uu = warp_sum(uu);
  pp = warp_sum(pp);
  up = warp_sum(up);
  if (lane == 0) {
    const float un = __fsqrt_rn(uu) + 1e-12f, pn = __fsqrt_rn(pp) + 1e-12f;
    // (U*/un) @ (psi/pn): both divisions are elementwise in the reference; the dot of the
    // scaled vectors equals the scaled dot up to fp32 rounding
    align[i] = __fdiv_rn(__fdiv_rn(up, un), pn);
  }
}

// one MMR step: fold the cosine column of the row chosen last into divmax, evaluate val, and
// reduce the block-local arg-max (value desc, index asc)
__global__ void __launch_bounds__(256)
mmr_step_kernel(const float* __restrict__ Yn, const float* __restrict__ score, int64_t N, int D,
                int last, float* __restrict__ divmax, const int* __restrict__ taken,
                float* __restrict__ blk_val, int* __restrict__ blk_idx) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
  float val = -INFINITY;
  int idx = 0x7fffffff;
  if (i < N) {
    float dv = divmax[i];
    if (last >= 0) {
      const float* a = Yn + i * D;
      const float* b = Yn + (int64_t)last * D;
      float s = 0.f;
      for (int d = lane; d < D; d += 32) s = fmaf(a[d], b[d], s);
      s = warp_sum(s);
      dv = fmaxf(dv, s);
      if (lane == 0) divmax[i] = dv;
    }
    if (!taken[i]) {
      val = __fsub_rn(__fmul_rn(0.5f, score[i]), (last >= 0) ? __fmul_rn(0.5f, dv) : 0.f);
      idx = (int)i;
    }
  }
  if (lane == 0) {
    sv[w] = val;
    si[w] = idx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int t = 0; t < (int)(blockDim.x >> 5); ++t)
      if (si[t] != 0x7fffffff && (sv[t] > bv || (sv[t] == bv && si[t] < bi) || bi == 0x7fffffff)) {
        bv = sv[t];
        bi = si[t];
      }
    blk_val[blockIdx.x] = bv;
    blk_idx[blockIdx.x] = bi;
  }
}

__global__ void __launch_bounds__(1024)
mmr_pick_kernel(const float* __restrict__ blk_val, const int* __restrict__ blk_idx, int nblk,
                int* __restrict__ taken, int* __restrict__ chosen, int step) {
  __shared__ float sv[1024];
  __shared__ int si[1024];
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    const float v = blk_val[b];
    const int i = blk_idx[b];
    if (i != 0x7fffffff && (bi == 0x7fffffff || v > bv || (v == bv && i < bi))) {
      bv = v;
      bi = i;
    }
  }
  sv[threadIdx.x] = bv;
  si[threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t < (int)blockDim.x; ++t) {
      if (si[t] != 0x7fffffff && (bi == 0x7fffffff || sv[t] > bv || (sv[t] == bv && si[t] < bi))) {
        bv = sv[t];
        bi = si[t];
      }
    }
    chosen[step] = bi;
    if (bi != 0x7fffffff) taken[bi] = 1;
  }
}

__global__ void __launch_bounds__(256)
pair_d2_kernel(const float* __restrict__ V, const float* __restrict__ sqrt_deg,
               const int32_t* __restrict__ pairs, int64_t M, int D, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= M) return;
  const int64_t i = pairs[2 * p], j = pairs[2 * p + 1];
  const float di = sqrt_deg[i] + 1e-12f, dj = sqrt_deg[j] + 1e-12f;
  const float* a = V + i * D;
  const float* b = V + j * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float e = __fdiv_rn(a[d], di) - __fdiv_rn(b[d], dj);
    s = fmaf(e, e, s);
  }
  s = warp_sum(s);
  if (lane == 0) out[p] = s;
}

__global__ void fill_kernel(float* p, float v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int launch_row_align(const float* Us, const float* psi, int64_t N, int D, float* align, cudaStream_t st) {
  if (N == 0) return OSC_OK;
  row_align_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(Us, psi, N, D, align);
  OSC_LAUNCH_CHECK("row_align_kernel");
  return OSC_OK;
}

int launch_pair_d2(const float* V, const float* sd, const int32_t* pairs, int64_t M, int D, float* out,
                   cudaStream_t st) {
  if (M == 0) return OSC_OK;
  pair_d2_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(V, sd, pairs, M, D, out);
  OSC_LAUNCH_CHECK("pair_d2_kernel");
  return OSC_OK;
}

// workspace: divmax[N] f32, taken[N] i32, blk_val[nblk] f32, blk_idx[nblk] i32
size_t mmr_workspace(int64_t N) {
  const int64_t nblk = (N + 7) / 8;
  return align_up(N * 4) + align_up(N * 4) + align_up(nblk * 4) + align_up(nblk * 4) + 256;
}

int launch_mmr(const float* Yn, const float* score, int64_t N, int D, int k, int32_t* chosen,
               void* workspace, size_t ws_bytes, cudaStream_t st) {
  if (k <= 0 || N == 0) return OSC_OK;
  if (ws_bytes < mmr_workspace(N)) return fail(OSC_ERR_WORKSPACE, "mmr: workspace too small");
  Arena ar(workspace, ws_bytes);
  const int64_t nblk = (N + 7) / 8;
  float* divmax = ar.take<float>(N);
  int* taken = ar.take<int>(N);
  float* bv = ar.take<float>(nblk);
  int* bi = ar.take<int>(nblk);
  fill_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(divmax, -INFINITY, N);
  OSC_LAUNCH_CHECK("fill_kernel");
  OSC_CUDA(cudaMemsetAsync(taken, 0, N * 4, st));
  int last = -1;
  const int steps = (int)(k < N ? k : N);
  for (int s = 0; s < steps; ++s) {
    mmr_step_kernel<<<(unsigned)nblk, 256, 0, st>>>(Yn, score, N, D, last, divmax, taken, bv, bi);
    OSC_LAUNCH_CHECK("mmr_step_kernel");
    mmr_pick_kernel<<<1, 1024, 0, st>>>(bv, bi, (int)nblk, taken, chosen, s);
    OSC_LAUNCH_CHECK("mmr_pick_kernel");
    OSC_CUDA(cudaMemcpyAsync(&last, chosen + s, sizeof(int), cudaMemcpyDeviceToHost, st));
    OSC_CUDA(cudaStreamSynchronize(st));
  }
  return OSC_OK;
}

}  // namespace osc
