// K1b: mutual filter, soft row cap, degrees, normalised weights -> ELL graph.
//   graph.py:50-52  keep only S_ij > 0
//   graph.py:64-65  mutual: (i,j) survives iff j in topk(i) and i in topk(j)
//   graph.py:77-83  c_i = min(1, cap/(sum_j a_ij + 1e-12)); A_ij = a_ij*sqrt(c_i*c_j)
//   graph.py:87-90  d_i = sum_j A_ij; sd_i = sqrt(max(d_i,1e-12)); W_ij = (A_ij/sd_i)/sd_j
// HBM-bound index/byte work: one thread per lattice row, rows are k contiguous entries.
#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace osc {

__global__ void __launch_bounds__(256)
assemble_mutual_kernel(const int32_t* __restrict__ top_idx, const float* __restrict__ top_sim,
                       int64_t N, int k, float cap, int32_t* __restrict__ nbr, float* __restrict__ A,
                       int32_t* __restrict__ deg, float* __restrict__ cscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = blockIdx.y;
  if (i >= N) return;
  const int64_t base = (b * N + i) * k;
  int cnt = 0;
  for (int t = 0; t < k; ++t) {
    const int j = top_idx[base + t];
    const float sim = top_sim[base + t];
    if (j < 0 || !(sim > 0.f)) continue;
    const int64_t jb = (b * N + j) * k;
    bool mutual = false;
    for (int u = 0; u < k; ++u)
      if (top_idx[jb + u] == (int)i && top_sim[jb + u] > 0.f) mutual = true;
    if (!mutual) continue;
    int p = cnt;
    while (p > 0 && nbr[base + p - 1] > j) {
      nbr[base + p] = nbr[base + p - 1];
      A[base + p] = A[base + p - 1];
      --p;
    }
    nbr[base + p] = j;
    A[base + p] = sim;
    ++cnt;
  }
  float s = 0.f;
  for (int t = 0; t < cnt; ++t) s += A[base + t];
  for (int t = cnt; t < k; ++t) {
    nbr[base + t] = -1;
    A[base + t] = 0.f;
  }
  s += 1e-12f;
  cscale[b * N + i] = fminf(1.0f, __fdiv_rn(cap, s));
  deg[b * N + i] = cnt;
}

// k in {4, 8, 12, 16}: the same filter with the row's entries, the four neighbour lists in flight and the sorted
// output all in registers, 16-byte loads and stores.  (The generic kernel above walks the lists with scalar
// loads, one neighbour after the other, and insertion-sorts in global memory: 1.4 ms per 4096 lattices of
// N = 1200, k = 8, latency bound.)  Output identical: neighbours ascending by column, row sum in that order.
template <int K>
__global__ void __launch_bounds__(256)
assemble_mutual_reg_kernel(const int32_t* __restrict__ top_idx, const float* __restrict__ top_sim,
                           int64_t N, float cap, int32_t* __restrict__ nbr, float* __restrict__ A,
                           int32_t* __restrict__ deg, float* __restrict__ cscale) {
  constexpr int Q = K / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = blockIdx.y;
  if (i >= N) return;
  const int64_t base = (b * N + i) * K;
  int tj[K];
  float ts[K];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int4 a = *reinterpret_cast<const int4*>(top_idx + base + 4 * q);
    const float4 f = *reinterpret_cast<const float4*>(top_sim + base + 4 * q);
    tj[4 * q] = a.x, tj[4 * q + 1] = a.y, tj[4 * q + 2] = a.z, tj[4 * q + 3] = a.w;
    ts[4 * q] = f.x, ts[4 * q + 1] = f.y, ts[4 * q + 2] = f.z, ts[4 * q + 3] = f.w;
  }
  int oj[K];
  float oa[K];
#pragma unroll
  for (int t = 0; t < K; ++t) {
    oj[t] = INT_MAX;
    oa[t] = 0.f;
  }
#pragma unroll
  for (int t0 = 0; t0 < K; t0 += 4) {
    int4 nj[4][Q];
    float4 ns[4][Q];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ok[u] = tj[t0 + u] >= 0 && ts[t0 + u] > 0.f;
      const int64_t jb = (b * N + (ok[u] ? tj[t0 + u] : (int)i)) * K;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        nj[u][q] = *reinterpret_cast<const int4*>(top_idx + jb + 4 * q);
        ns[u][q] = *reinterpret_cast<const float4*>(top_sim + jb + 4 * q);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      bool mutual = false;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        mutual |= (nj[u][q].x == (int)i && ns[u][q].x > 0.f) || (nj[u][q].y == (int)i && ns[u][q].y > 0.f) ||
                  (nj[u][q].z == (int)i && ns[u][q].z > 0.f) || (nj[u][q].w == (int)i && ns[u][q].w > 0.f);
      }
      if (ok[u] && mutual) {  // sorted insertion (columns are distinct; empty slots hold INT_MAX)
        const int j = tj[t0 + u];
        const float sim = ts[t0 + u];
#pragma unroll
        for (int p = K - 1; p > 0; --p) {
          const bool up = oj[p - 1] > j, here = !up && oj[p] > j;
          oa[p] = up ? oa[p - 1] : (here ? sim : oa[p]);
          oj[p] = up ? oj[p - 1] : (here ? j : oj[p]);
        }
        if (oj[0] > j) {
          oj[0] = j;
          oa[0] = sim;
        }
      }
    }
  }
  int cnt = 0;
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < K; ++t) {
    if (oj[t] != INT_MAX) {
      ++cnt;
      s += oa[t];
    } else {
      oj[t] = -1;
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    *reinterpret_cast<int4*>(nbr + base + 4 * q) = make_int4(oj[4 * q], oj[4 * q + 1], oj[4 * q + 2], oj[4 * q + 3]);
    *reinterpret_cast<float4*>(A + base + 4 * q) = make_float4(oa[4 * q], oa[4 * q + 1], oa[4 * q + 2], oa[4 * q + 3]);
  }
  s += 1e-12f;
  cscale[b * N + i] = fminf(1.0f, __fdiv_rn(cap, s));
  deg[b * N + i] = cnt;
}

__global__ void __launch_bounds__(256)
assemble_cap_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ deg,
                    const float* __restrict__ cscale, int64_t N, int k, float* __restrict__ A,
                    float* __restrict__ sqrt_deg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = blockIdx.y;
  if (i >= N) return;
  const int64_t base = (b * N + i) * k;
  const float ci = cscale[b * N + i];
  const int n = deg[b * N + i];
  float d = 0.f;
  for (int t = 0; t < n; ++t) {
    const float cj = cscale[b * N + nbr[base + t]];
    const float v = __fmul_rn(A[base + t], __fsqrt_rn(__fmul_rn(ci, cj)));
    A[base + t] = v;
    d += v;
  }
  sqrt_deg[b * N + i] = __fsqrt_rn(fmaxf(d, 1e-12f));
}

__global__ void __launch_bounds__(256)
assemble_weights_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ deg,
                        const float* __restrict__ A, const float* __restrict__ sqrt_deg, int64_t N,
                        int k, float* __restrict__ W, unsigned long long* __restrict__ nnz) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = blockIdx.y;
  int n = 0;
  if (i < N) {
    const int64_t base = (b * N + i) * k;
    const float inv_i = __fdiv_rn(1.0f, sqrt_deg[b * N + i]);
    n = deg[b * N + i];
    for (int t = 0; t < k; ++t) {
      float w = 0.f;
      if (t < n) {
        const float inv_j = __fdiv_rn(1.0f, sqrt_deg[b * N + nbr[base + t]]);
        w = __fmul_rn(__fmul_rn(A[base + t], inv_i), inv_j);
      }
      W[base + t] = w;
    }
  }
  // block-level count of directed edges
  unsigned v = (unsigned)n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ unsigned wsum[8];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0 && nnz != nullptr) {
    unsigned tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += wsum[w];
    if (tot) atomicAdd(nnz + b, (unsigned long long)tot);
  }
}

int launch_assemble(const int32_t* top_idx, const float* top_sim, int64_t batch, int64_t N, int k,
                    float cap, int32_t* nbr, float* A, float* W, int32_t* deg, float* sqrt_deg,
                    int64_t* nnz, float* scratch, cudaStream_t st) {
  if (N == 0 || batch == 0) return OSC_OK;
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)batch);
  if (nnz != nullptr) OSC_CUDA(cudaMemsetAsync(nnz, 0, sizeof(int64_t) * batch, st));
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  bool reg = (k == 4 || k == 8 || k == 12 || k == 16) && a16(top_idx) && a16(top_sim) && a16(nbr) && a16(A);
  {
    const char* e = getenv("OSC_ASSEMBLE_REG");  // dev-only A/B switch
    if (e && atoi(e) == 0) reg = false;
  }
  if (reg && k == 4) assemble_mutual_reg_kernel<4><<<grid, 256, 0, st>>>(top_idx, top_sim, N, cap, nbr, A, deg, scratch);
  else if (reg && k == 8) assemble_mutual_reg_kernel<8><<<grid, 256, 0, st>>>(top_idx, top_sim, N, cap, nbr, A, deg, scratch);
  else if (reg && k == 12) assemble_mutual_reg_kernel<12><<<grid, 256, 0, st>>>(top_idx, top_sim, N, cap, nbr, A, deg, scratch);
  else if (reg) assemble_mutual_reg_kernel<16><<<grid, 256, 0, st>>>(top_idx, top_sim, N, cap, nbr, A, deg, scratch);
  else assemble_mutual_kernel<<<grid, 256, 0, st>>>(top_idx, top_sim, N, k, cap, nbr, A, deg, scratch);
  OSC_LAUNCH_CHECK("assemble_mutual_kernel");
  assemble_cap_kernel<<<grid, 256, 0, st>>>(nbr, deg, scratch, N, k, A, sqrt_deg);
  OSC_LAUNCH_CHECK("assemble_cap_kernel");
  assemble_weights_kernel<<<grid, 256, 0, st>>>(nbr, deg, A, sqrt_deg, N, k, W,
                                                reinterpret_cast<unsigned long long*>(nnz));
  OSC_LAUNCH_CHECK("assemble_weights_kernel");
  return OSC_OK;
}

}  // namespace osc
