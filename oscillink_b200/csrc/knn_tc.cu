// K1 tensor-core engine: TMA-staged 3xTF32 similarity GEMM on tcgen05 with the top-k fused into
// the TMEM epilogue -- the N x N similarity matrix never reaches HBM (graph.py:36-37,59).
//
//   S_tile = Ahi*Bhi^T + Ahi*Blo^T + Alo*Bhi^T      (fp32 accumulate in TMEM; hi = tf32(x),
//                                                     lo = tf32(x - hi): ~2^-22 relative)
//
// One CTA per SM (persistent).  A work item is (lattice b, 128-row panel); the CTA sweeps all
// 256-column tiles of that lattice, double-buffering the 128x256 fp32 accumulator in TMEM
// (2 x 256 = all 512 columns) so the epilogue of tile t overlaps the MMAs of tile t+1.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor.3d, SWIZZLE_128B, 32-float K blocks;
//                                 hi and lo of A and B ride in the SAME stage so one stage
//                                 feeds all three products -> 96 KB / 3 MMA groups)
//   warp 1      : MMA issuer     (tcgen05.mma.cta_group::1.kind::tf32, M=128 N=256 K=8) + TMEM alloc
//   warps 2..5  : epilogue       (tcgen05.ld 32x32b.x32: thread == TMEM lane == similarity row;
//                                 per-row sorted top-KC list held in registers, one compare
//                                 against the running threshold rejects almost every column)
//
// The candidate lists (approximate scores) feed knn_rescore_kernel, which fixes the final
// neighbour sets in the canonical order -- both engines therefore yield identical graphs.
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace osc {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 32;  // floats: 128 B = one SWIZZLE_128B row
constexpr int TC_STAGES = 2;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;  // Ahi, Alo, Bhi, Blo
constexpr int TC_THREADS = 192;
constexpr int TC_EPI_BYTES = 4 * 32 * 32 * 4;  // per epilogue warp: one 32x32 fp32 chunk
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + TC_EPI_BYTES;

struct TcParams {
  int64_t batch, n_rows, row0, N;
  int D, kc;
  int panels, col_tiles, k_blocks;
  int64_t total_work;
  int32_t* cand_idx;
  float* cand_sim;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major) = 16 B
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct Pipe {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) {
      stage = 0;
      phase ^= 1u;
    }
  }
};


// ---------------------------------------------------------------- fused top-KC epilogue
// Thread == TMEM lane == similarity row; the row's running top-KC list (sorted, descending) lives
// in registers.  Scanning the 32 columns of a chunk one by one costs a full (divergent) insertion
// whenever ANY of the 32 rows of the warp has a hit at that column -- at N ~ 1e3 that is almost
// every column, and the epilogue, not the tensor pipe, bounded the kernel.  Instead every lane
// first builds the bit mask of its columns above its own threshold, the chunk is parked in shared
// memory, and the warp then runs insertion ROUNDS: in round r every lane inserts its r-th hit.
// Lanes insert side by side, so the number of (expensive) insertion executions per chunk is the
// maximum hit count over the 32 rows instead of the number of distinct hit columns (~4x fewer at
// N = 1200, kc = 12).  Hits are taken in ascending column order and the test is a strict '>', so
// the smaller column still wins ties.
template <int KC>
__device__ __forceinline__ void topk_insert(float (&val)[KC], int (&idx)[KC], float s, int col) {
#pragma unroll
  for (int q = KC - 1; q > 0; --q) {
    const bool up = s > val[q - 1];
    const bool here = !up && (s > val[q]);
    const float nv = up ? val[q - 1] : (here ? s : val[q]);
    const int ni = up ? idx[q - 1] : (here ? col : idx[q]);
    val[q] = nv;
    idx[q] = ni;
  }
  if (s > val[0]) {
    val[0] = s;
    idx[0] = col;
  }
}

template <int KC>
__device__ __forceinline__ void topk_chunk(float (&val)[KC], int (&idx)[KC], float (&v)[32], int c0, int N,
                                           int self, bool row_ok, float* buf, int lane) {
  // columns outside the lattice and the row's own column never qualify (warp-uniform fast path:
  // the 32 rows of a warp are consecutive, so `self` falls into at most two chunks per row panel)
  const bool self_here = __any_sync(0xffffffffu, self >= c0 && self < c0 + 32);  // all lanes vote
  if (c0 + 32 > N || self_here) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i >= N || c0 + i == self) v[i] = -INFINITY;
  }
  const float thr = val[KC - 1];
  unsigned mask = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (v[i] > thr) mask |= 1u << i;
  if (!row_ok) mask = 0;
  if (!__any_sync(0xffffffffu, mask != 0)) return;
#pragma unroll
  for (int i = 0; i < 32; ++i) buf[i * 32 + lane] = v[i];
  __syncwarp();
  while (__any_sync(0xffffffffu, mask != 0)) {
    if (mask != 0) {
      const int e = __ffs(mask) - 1;
      mask &= mask - 1;
      const float s = buf[e * 32 + lane];
      if (s > val[KC - 1]) topk_insert<KC>(val, idx, s, c0 + e);
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------- packed-key top-KC (N <= 2048)
// For small lattices the epilogue, not the tensor pipe, bounds the kernel (profiles/
// r01_knn_tch_ncu_summary.txt: 0.25 IPC on the epilogue warps, predicate-chained selects).  With the
// column index in the low 11 bits of an order-preserving integer image of the score, a list entry is ONE
// register and the sorted insertion is a branch-free, predicate-free min/max network:
//     new[q] = max(old[q], min(old[q-1], key))        (2 IMNMX per slot, all slots independent)
// Keys order by (score truncated to 21 bits desc, column asc).  The truncation costs <= 2^-12 relative
// (2.5e-4 absolute for unit rows); OSC_KNN_EPS_TC1 covers it next to the 2^-10 operand rounding.
constexpr int PK_COL_BITS = 11;
constexpr int PK_COL_MASK = (1 << PK_COL_BITS) - 1;
__device__ __forceinline__ int pk_key(float s, int col) {
  const int b = __float_as_int(s);
  const int t = b ^ ((b >> 31) & 0x7fffffff);  // monotone under signed compare
  return (t & ~PK_COL_MASK) | (PK_COL_MASK - col);
}
__device__ __forceinline__ float pk_score(int key) {
  const int t = key & ~PK_COL_MASK;
  return __int_as_float(t ^ ((t >> 31) & 0x7fffffff));
}
__device__ __forceinline__ int pk_col(int key) { return PK_COL_MASK - (key & PK_COL_MASK); }

template <int KC>
__device__ __forceinline__ void pk_insert(int (&key)[KC], int k) {
#pragma unroll
  for (int q = KC - 1; q > 0; --q) key[q] = max(key[q], min(key[q - 1], k));
  key[0] = max(key[0], k);
}

template <int KC>
__device__ __forceinline__ void pk_chunk(int (&key)[KC], float (&v)[32], int c0, int N, int self, bool row_ok,
                                         float* buf, int lane) {
  const bool self_here = __any_sync(0xffffffffu, self >= c0 && self < c0 + 32);
  if (c0 + 32 > N || self_here) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (c0 + i >= N || c0 + i == self) v[i] = -INFINITY;
  }
  // the truncated score of the last entry is <= its true score: a few extra hits, never a missed one
  const float thr = key[KC - 1] == INT_MIN ? -INFINITY : pk_score(key[KC - 1]);
  unsigned mask = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (v[i] > thr) mask |= 1u << i;
  if (!row_ok) mask = 0;
  if (!__any_sync(0xffffffffu, mask != 0)) return;
#pragma unroll
  for (int i = 0; i < 32; ++i) buf[i * 32 + lane] = v[i];
  __syncwarp();
  // Insertion rounds, software-pipelined: the next hit of every lane (bit scan -> LDS -> key packing, a
  // ~80-cycle dependent chain) is fetched while the min/max network of the current one runs.  A lane without
  // a hit inserts INT_MIN, which the network leaves unchanged, so the loop body is branch-free.
  auto next_key = [&]() -> int {
    if (mask == 0) return INT_MIN;
    const int e = __ffs(mask) - 1;
    mask &= mask - 1;
    return pk_key(buf[e * 32 + lane], c0 + e);
  };
  int cur = next_key();
  while (__any_sync(0xffffffffu, cur != INT_MIN)) {
    const int nxt = next_key();
    pk_insert<KC>(key, cur);
    cur = nxt;
  }
  __syncwarp();
}

// ---------------------------------------------------------------- kernel
template <int KC>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
              const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
              TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES;
  // barrier slots (8 B each): full[STAGES], empty[STAGES], tfull[2], tempty[2]
  const uint32_t full0 = bars, empty0 = bars + 8 * TC_STAGES, tfull0 = bars + 16 * TC_STAGES,
                 tempty0 = tfull0 + 16;
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(gen_base + TC_STAGES * TC_STAGE_BYTES + 16 * TC_STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer
    if (lane == 0) {
      Pipe p;
      for (int64_t w = blockIdx.x; w < P.total_work; w += gridDim.x) {
        const int b = (int)(w / P.panels);
        const int m0 = (int)(w % P.panels) * TC_BM;
        for (int ct = 0; ct < P.col_tiles; ++ct) {
          const int n0 = ct * TC_BN;
          for (int kb = 0; kb < P.k_blocks; ++kb) {
            mbar_wait(empty0 + 8 * p.stage, p.phase ^ 1u);
            const uint32_t sb = base + p.stage * TC_STAGE_BYTES;
            const uint32_t fb = full0 + 8 * p.stage;
            mbar_expect_tx(fb, TC_STAGE_BYTES);
            tma_load_3d(&tm_q_hi, sb, fb, kb * TC_BK, m0, b);
            tma_load_3d(&tm_q_lo, sb + TC_A_BYTES, fb, kb * TC_BK, m0, b);
            tma_load_3d(&tm_a_hi, sb + 2 * TC_A_BYTES, fb, kb * TC_BK, n0, b);
            tma_load_3d(&tm_a_lo, sb + 2 * TC_A_BYTES + TC_B_BYTES, fb, kb * TC_BK, n0, b);
            p.advance(TC_STAGES);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer
    // instruction descriptor: D=F32, A=B=TF32, both K-major, N=256, M=128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                           ((uint32_t)(TC_BM >> 4) << 24);
    Pipe p, acc;
    for (int64_t w = blockIdx.x; w < P.total_work; w += gridDim.x) {
      for (int ct = 0; ct < P.col_tiles; ++ct) {
        mbar_wait(tempty0 + 8 * acc.stage, acc.phase ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc.stage * TC_BN);
        for (int kb = 0; kb < P.k_blocks; ++kb) {
          mbar_wait(full0 + 8 * p.stage, p.phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sb = base + p.stage * TC_STAGE_BYTES;
            const uint64_t ah = umma_desc(sb), al = umma_desc(sb + TC_A_BYTES),
                           bh = umma_desc(sb + 2 * TC_A_BYTES),
                           bl = umma_desc(sb + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks) {
              const uint64_t o = (uint64_t)((ks * 32) >> 4);  // +32 B per K=8 step inside the swizzle row
              tc_mma_tf32(tacc, al + o, bh + o, idesc, (kb | ks) ? 1u : 0u);
              tc_mma_tf32(tacc, ah + o, bl + o, idesc, 1u);
              tc_mma_tf32(tacc, ah + o, bh + o, idesc, 1u);
            }
            tc_commit(empty0 + 8 * p.stage);  // smem stage free once these MMAs retire
            if (kb == P.k_blocks - 1) tc_commit(tfull0 + 8 * acc.stage);
          }
          __syncwarp();
          p.advance(TC_STAGES);
        }
        acc.advance(2);
      }
    }
  } else {
    // ===================== epilogue: fused per-row top-KC
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    const int r_tile = quarter * 32 + lane;
    float* ebuf = reinterpret_cast<float*>(gen_base + TC_STAGES * TC_STAGE_BYTES + 256) + quarter * 1024;
    Pipe acc;
    for (int64_t w = blockIdx.x; w < P.total_work; w += gridDim.x) {
      const int64_t b = w / P.panels;
      const int m0 = (int)(w % P.panels) * TC_BM;
      const int gi = m0 + r_tile;                 // query row inside the panel set
      const int self = (int)(P.row0 + gi);        // its column id
      const bool row_ok = gi < P.n_rows;
      float val[KC];
      int idx[KC];
#pragma unroll
      for (int i = 0; i < KC; ++i) {
        val[i] = -INFINITY;
        idx[i] = -1;
      }
      for (int ct = 0; ct < P.col_tiles; ++ct) {
        const int n0 = ct * TC_BN;
        mbar_wait(tfull0 + 8 * acc.stage, acc.phase);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc.stage * TC_BN);
#pragma unroll 1
        for (int ch = 0; ch < TC_BN / 32; ++ch) {
          const int c0 = n0 + ch * 32;
          if (c0 >= P.N) break;  // warp-uniform
          float v[32];
          tmem_ld32(trow + (uint32_t)(ch * 32), v);
          topk_chunk<KC>(val, idx, v, c0, (int)P.N, self, row_ok, ebuf, lane);
        }
        tc_fence_before();
        mbar_arrive(tempty0 + 8 * acc.stage);
        acc.advance(2);
      }
      if (row_ok) {
        const int64_t o = (b * P.n_rows + gi) * P.kc;
#pragma unroll
        for (int i = 0; i < KC; ++i) {
          if (i < P.kc) {
            P.cand_idx[o + i] = idx[i];
            P.cand_sim[o + i] = val[i];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}


// ================================================================= 2-CTA (cta_group::2) variant
// Two CTAs of a cluster (one TPC) execute every MMA together: M = 256 rows (128 per CTA), N = 256.
// Each CTA stages ITS 128 query rows and only HALF of the 256 column rows of a tile, so the
// L2 -> shared-memory traffic per MMA drops from 96 KB to 64 KB per CTA and k-block -- this kernel
// is bound by exactly that traffic (ncu: tensor pipe 29 % busy, xbar->L1 5 TB/s with one CTA per
// tile).  The smaller stage also buys a third pipeline stage.
//
//   leader CTA (cluster rank 0): warp 1 issues tcgen05.mma.cta_group::2 for the pair
//   both CTAs: warp 0 TMA producer (signals the LEADER's full barrier), warps 2..5 epilogue on
//              their own TMEM half (128 lanes x 256 columns, double-buffered)
//   tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs; the epilogues of
//   both CTAs arrive on the leader's "accumulator drained" barrier.
//
// ONEPASS = true: the single-product engines.  Only the rounded operands are staged and ONE MMA is
// issued per K step (a third of the tensor work, half of the staging traffic per k-block, twice the
// pipeline depth).  The scores carry the full operand rounding error (<= 2^-10 for unit rows); they
// only pre-select candidates -- the final sets and weights come from the exact re-scoring, and the
// completeness check uses that error as a rigorous bound.
//   F16 = false (OSC_KNN_TC1): tf32-rounded fp32 operands, kind::tf32, K = 8 per MMA.
//   F16 = true  (OSC_KNN_TCH): fp16 operands, kind::f16, K = 16 per MMA.  fp16 carries the same 11
//     significant bits as tf32, and the entries of a unit row sit in its normal range (|x| <= 1; below
//     2^-14 the absolute error is <= 2^-25), so the error bound is the same -- at half the operand
//     bytes and twice the MMA rate.  That pays where the kernel is tensor-bound (N ~ 1e6: 2.19 s ->
//     1.34 s at N=1M D=768); at N ~ 1e3 the top-k epilogue bounds it and the operand type hardly
//     matters (14.8 -> 14.4 ms) -- see PACKED / EW below.  A 128-byte swizzle row holds 64 halves
//     instead of 32 floats; the stage layout in BYTES, the descriptors and the +32 B K-step are unchanged.
constexpr int TC2_HALF_BYTES = 128 * TC_BK * 4;                     // 128 rows x 32 floats
template <bool ONEPASS>
struct Tc2Cfg {
  static constexpr int STAGES = ONEPASS ? 6 : 3;
  static constexpr int STAGE_BYTES = (ONEPASS ? 2 : 4) * TC2_HALF_BYTES;  // Ahi [Alo] Bhi(half) [Blo(half)]
};
constexpr int TC2_SMEM = 3 * 4 * TC2_HALF_BYTES + 1024 + 256 + 2 * TC_EPI_BYTES;  // same for all configs (8 chunk buffers)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA load whose completion bytes are credited to a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint32_t dst, uint32_t cluster_bar,
                                                 int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// EW = epilogue warps per CTA (4 or 8).  With 8, two warps share every TMEM lane quarter (warp w and
// w + 4 address the same 32 lanes) and take alternate 32-column chunks of every tile, each with its own
// register top-KC list; the lists are merged through shared memory once per row panel.  At N ~ 1e3 the
// epilogue (about KC*(1+ln(N/KC)) list insertions per row) bounds the kernel, and a single warp per
// scheduler cannot hide its own LDS / vote latencies between insertion rounds.
template <int KC, bool ONEPASS, bool F16 = false, int EW = 4, bool PACKED = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
knn_tc2_kernel(const __grid_constant__ CUtensorMap tm_q_hi, const __grid_constant__ CUtensorMap tm_q_lo,
               const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               TcParams P) {
  static_assert(!F16 || ONEPASS, "the fp16 engine is single-product");
  constexpr int TC2_STAGES = Tc2Cfg<ONEPASS>::STAGES;
  constexpr int TC2_STAGE_BYTES = Tc2Cfg<ONEPASS>::STAGE_BYTES;
  constexpr int KE = F16 ? 64 : TC_BK;  // elements per 128-byte swizzle row = per k-block
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t bars = base + TC2_STAGES * TC2_STAGE_BYTES;
  // barrier slots (8 B each): full[S], empty[S], tfull[2], tempty[2]
  const uint32_t full0 = bars, empty0 = bars + 8 * TC2_STAGES, tfull0 = bars + 16 * TC2_STAGES,
                 tempty0 = tfull0 + 16;
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(gen_base + TC2_STAGES * TC2_STAGE_BYTES + 16 * TC2_STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC2_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);   // leader's producer arrives with the byte count of BOTH CTAs
      mbar_init(empty0 + 8 * s, 1);  // one multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 64 * EW);  // every epilogue thread of both CTAs (used in the leader only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_slot)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation of both CTAs visible before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs)
    if (lane == 0) {
      Pipe p;
      for (int64_t w = pair; w < P.total_work; w += n_pairs) {
        const int b = (int)(w / P.panels);
        const int m0 = (int)(w % P.panels) * 256 + (int)rank * 128;  // this CTA's 128 query rows
        for (int ct = 0; ct < P.col_tiles; ++ct) {
          const int n0 = ct * TC_BN + (int)rank * 128;               // this CTA's half of the column rows
          for (int kb = 0; kb < P.k_blocks; ++kb) {
            mbar_wait(empty0 + 8 * p.stage, p.phase ^ 1u);
            const uint32_t sb = base + p.stage * TC2_STAGE_BYTES;
            const uint32_t fb = mapa_rank(full0 + 8 * p.stage, 0);   // the leader's full barrier
            if (leader) mbar_expect_tx(full0 + 8 * p.stage, 2 * TC2_STAGE_BYTES);
            if constexpr (ONEPASS) {
              tma_load_3d_pair(&tm_q_hi, sb, fb, kb * KE, m0, b);
              tma_load_3d_pair(&tm_a_hi, sb + TC2_HALF_BYTES, fb, kb * KE, n0, b);
            } else {
              tma_load_3d_pair(&tm_q_hi, sb, fb, kb * TC_BK, m0, b);
              tma_load_3d_pair(&tm_q_lo, sb + TC2_HALF_BYTES, fb, kb * TC_BK, m0, b);
              tma_load_3d_pair(&tm_a_hi, sb + 2 * TC2_HALF_BYTES, fb, kb * TC_BK, n0, b);
              tma_load_3d_pair(&tm_a_lo, sb + 3 * TC2_HALF_BYTES, fb, kb * TC_BK, n0, b);
            }
            p.advance(TC2_STAGES);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only)
    if (leader) {
      // instruction descriptor: D=F32, A=B=TF32, both K-major, N=256, M=256 (pair)
      // (fp16 engine: A = B = F16, format code 0)
      const uint32_t fmt = F16 ? 0u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      Pipe p, acc;
      for (int64_t w = pair; w < P.total_work; w += n_pairs) {
        for (int ct = 0; ct < P.col_tiles; ++ct) {
          mbar_wait(tempty0 + 8 * acc.stage, acc.phase ^ 1u);
          tc_fence_after();
          const uint32_t tacc = tmem_base + (uint32_t)(acc.stage * TC_BN);
          for (int kb = 0; kb < P.k_blocks; ++kb) {
            mbar_wait(full0 + 8 * p.stage, p.phase);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t sb = base + p.stage * TC2_STAGE_BYTES;
              if constexpr (ONEPASS) {
                const uint64_t ah = umma_desc(sb), bh = umma_desc(sb + TC2_HALF_BYTES);
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {
                  const uint64_t o = (uint64_t)((ks * 32) >> 4);
                  if constexpr (F16) tc_mma_f16_pair(tacc, ah + o, bh + o, idesc, (kb | ks) ? 1u : 0u);
                  else tc_mma_tf32_pair(tacc, ah + o, bh + o, idesc, (kb | ks) ? 1u : 0u);
                }
              } else {
                const uint64_t ah = umma_desc(sb), al = umma_desc(sb + TC2_HALF_BYTES),
                               bh = umma_desc(sb + 2 * TC2_HALF_BYTES), bl = umma_desc(sb + 3 * TC2_HALF_BYTES);
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {
                  const uint64_t o = (uint64_t)((ks * 32) >> 4);
                  tc_mma_tf32_pair(tacc, al + o, bh + o, idesc, (kb | ks) ? 1u : 0u);
                  tc_mma_tf32_pair(tacc, ah + o, bl + o, idesc, 1u);
                  tc_mma_tf32_pair(tacc, ah + o, bh + o, idesc, 1u);
                }
              }
              tc_commit_pair(empty0 + 8 * p.stage);  // stage free in BOTH CTAs once these MMAs retire
              if (kb == P.k_blocks - 1) tc_commit_pair(tfull0 + 8 * acc.stage);
            }
            __syncwarp();
            p.advance(TC2_STAGES);
          }
          acc.advance(2);
        }
      }
    }
  } else {
    // ===================== epilogue: fused per-row top-KC on this CTA's 128 rows
    static_assert(EW == 4 || (EW == 8 && KC <= 16), "the list merge parks KC values + KC indices in 32 words per lane");
    const int ew = warp - 2;       // 0..EW-1
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read (hardware: warp id % 4)
    const int half = ew >> 2;      // EW == 8: 0 takes the even chunks of a tile, 1 the odd ones
    const int r_tile = quarter * 32 + lane;
    float* ebuf = reinterpret_cast<float*>(gen_base + TC2_STAGES * TC2_STAGE_BYTES + 256) + ew * 1024;
    const uint32_t tempty_leader0 = mapa_rank(tempty0, 0);
    Pipe acc;
    for (int64_t w = pair; w < P.total_work; w += n_pairs) {
      const int64_t b = w / P.panels;
      const int m0 = (int)(w % P.panels) * 256 + (int)rank * 128;
      const int gi = m0 + r_tile;
      const int self = (int)(P.row0 + gi);
      const bool row_ok = gi < P.n_rows;
      // PACKED: one register per entry (pk_key); otherwise separate score / column lists
      float val[PACKED ? 1 : KC];
      int idx[KC];
#pragma unroll
      for (int i = 0; i < KC; ++i) {
        if constexpr (!PACKED) val[i] = -INFINITY;
        idx[i] = PACKED ? INT_MIN : -1;
      }
      for (int ct = 0; ct < P.col_tiles; ++ct) {
        const int n0 = ct * TC_BN;
        mbar_wait(tfull0 + 8 * acc.stage, acc.phase);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc.stage * TC_BN);
#pragma unroll 1
        for (int ch = (EW == 8 ? half : 0); ch < TC_BN / 32; ch += (EW == 8 ? 2 : 1)) {
          const int c0 = n0 + ch * 32;
          if (c0 >= P.N) break;  // warp-uniform
          float v[32];
          tmem_ld32(trow + (uint32_t)(ch * 32), v);
          if constexpr (PACKED) pk_chunk<KC>(idx, v, c0, (int)P.N, self, row_ok, ebuf, lane);
          else topk_chunk<KC>(val, idx, v, c0, (int)P.N, self, row_ok, ebuf, lane);
        }
        tc_fence_before();
        mbar_arrive_cluster(tempty_leader0 + 8 * acc.stage);
        acc.advance(2);
      }
      if constexpr (EW == 8) {
        // merge the partner warp's list (same rows, the other chunks): parked in ITS chunk buffer
        float* pbuf = half ? ebuf : ebuf + 4 * 1024;
        if (half) {
#pragma unroll
          for (int i = 0; i < KC; ++i) {
            if constexpr (!PACKED) pbuf[i * 32 + lane] = val[i];
            pbuf[(16 + i) * 32 + lane] = __int_as_float(idx[i]);
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        if (!half) {
#pragma unroll 1
          for (int i = 0; i < KC; ++i) {
            const int sj = __float_as_int(pbuf[(16 + i) * 32 + lane]);
            if constexpr (PACKED) {
              pk_insert<KC>(idx, sj);
            } else {
              const float sv = pbuf[i * 32 + lane];
              if (sv > val[KC - 1]) topk_insert<KC>(val, idx, sv, sj);
            }
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // pbuf is a chunk buffer again
      }
      if (row_ok && half == 0) {
        const int64_t o = (b * P.n_rows + gi) * P.kc;
#pragma unroll
        for (int i = 0; i < KC; ++i) {
          if (i < P.kc) {
            if constexpr (PACKED) {
              const bool empty = idx[i] == INT_MIN;
              P.cand_idx[o + i] = empty ? -1 : pk_col(idx[i]);
              P.cand_sim[o + i] = empty ? -INFINITY : pk_score(idx[i]);
            } else {
              P.cand_idx[o + i] = idx[i];
              P.cand_sim[o + i] = val[i];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's shared memory / barriers stay alive until every MMA and commit landed
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ================================================================= host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encoder() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [batch][rows][D] fp32 (or fp16), box = {128 bytes of K, box_rows, 1}, SWIZZLE_128B, OOB -> 0
static int make_map(CUtensorMap* m, const void* ptr, int64_t batch, int64_t rows, int D, int box_rows,
                    bool f16 = false) {
  EncodeTiledFn enc = encoder();
  if (enc == nullptr) return fail(OSC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t esz = f16 ? 2 : 4;
  const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)rows, (cuuint64_t)batch};
  const cuuint64_t strides[2] = {(cuuint64_t)D * esz, (cuuint64_t)rows * D * esz};
  const cuuint32_t box[3] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                         const_cast<void*>(ptr), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OSC_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return OSC_OK;
}

int knn_tc_supported(int64_t N, int D, int kc) {
  if (N < 2 || N > 0x7fffffffLL || D < 4 || D % 4 != 0 || kc < 1 || kc > 32) return 0;
  int dev = 0, mj = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return mj == 10 ? 1 : 0;
}

// the single-product engine exists only as the 2-CTA kernel: it needs more than one 128-row panel
int knn_tc1_supported(int64_t n_rows, int64_t N, int D, int kc) {
  return knn_tc_supported(N, D, kc) && n_rows > TC_BM && sm_count() >= 2;
}
// fp16 operands: the row pitch D * 2 B must be a multiple of 16 B for the tensor map
int knn_tch_supported(int64_t n_rows, int64_t N, int D, int kc) {
  return knn_tc1_supported(n_rows, N, D, kc) && D % 8 == 0;
}

// f16 (single-product only): q_hi / all_hi point at fp16 arrays of the same [batch][rows][D] shape
int launch_knn_tc(const float* q_hi, const float* q_lo, const float* all_hi, const float* all_lo,
                  int64_t batch, int64_t n_rows, int64_t row0, int64_t N, int D, int kc,
                  int32_t* cand_idx, float* cand_sim, bool onepass, cudaStream_t st, bool f16) {
  if (!knn_tc_supported(N, D, kc)) return fail(OSC_ERR_UNSUPPORTED, "knn_tc: shape not covered");
  OSC_REQUIRE(!f16 || (onepass && D % 8 == 0), "knn_tch: fp16 engine is single-product and needs D % 8 == 0");
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (onepass) {  // the lo parts are not read (the caller may pass NULL)
    OSC_REQUIRE(knn_tc1_supported(n_rows, N, D, kc), "knn_tc1: needs more than 128 query rows");
    q_lo = q_hi;
    all_lo = all_hi;
  }
  OSC_REQUIRE(a16(q_hi) && a16(q_lo) && a16(all_hi) && a16(all_lo), "knn_tc: operands must be 16 B aligned");
  const int64_t sms = sm_count();
  bool pair = n_rows > TC_BM && sms >= 2;
  if (onepass) pair = true;
  {
    const char* e = getenv("OSC_KNN_TC_PAIR");  // dev-only A/B switch: 0 = one CTA per tile
    if (e && !onepass) pair = pair && atoi(e) != 0;
  }
  CUtensorMap mqh, mql, mah, mal;
  int rc;
  if ((rc = make_map(&mqh, q_hi, batch, n_rows, D, TC_BM, f16))) return rc;
  if ((rc = make_map(&mql, q_lo, batch, n_rows, D, TC_BM, f16))) return rc;
  if ((rc = make_map(&mah, all_hi, batch, N, D, pair ? 128 : TC_BN, f16))) return rc;
  if ((rc = make_map(&mal, all_lo, batch, N, D, pair ? 128 : TC_BN, f16))) return rc;
  TcParams P;
  P.batch = batch;
  P.n_rows = n_rows;
  P.row0 = row0;
  P.N = N;
  P.D = D;
  P.kc = kc;
  P.panels = (int)((n_rows + (pair ? 256 : TC_BM) - 1) / (pair ? 256 : TC_BM));
  P.col_tiles = (int)((N + TC_BN - 1) / TC_BN);
  P.k_blocks = f16 ? (D + 63) / 64 : (D + TC_BK - 1) / TC_BK;
  P.total_work = batch * P.panels;
  P.cand_idx = cand_idx;
  P.cand_sim = cand_sim;
  if (pair) {
    int64_t pairs = sms / 2;
    if (P.total_work < pairs) pairs = P.total_work;
    const unsigned grid = (unsigned)(2 * pairs);
#define OSC_TC2_LAUNCH_T(THREADS, K, ...)                                                                  \
  do {                                                                                                     \
    OSC_CUDA(cudaFuncSetAttribute(knn_tc2_kernel<K, __VA_ARGS__>,                                            \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM));                 \
    knn_tc2_kernel<K, __VA_ARGS__><<<grid, THREADS, TC2_SMEM, st>>>(mqh, mql, mah, mal, P);                  \
  } while (0)
#define OSC_TC2_LAUNCH(K, ...) OSC_TC2_LAUNCH_T(TC_THREADS, K, __VA_ARGS__)
    // packed-key lists: single-product engines, the column index must fit 11 bits (OSC_KNN_PACKED=0: off)
    bool packed = onepass && kc <= 16 && N <= (1 << PK_COL_BITS);
    {
      const char* e = getenv("OSC_KNN_PACKED");
      if (e && atoi(e) == 0) packed = false;
    }
    // 8 epilogue warps.  Measured on B200 at 4096 x N=1200 (fp16 engine): score/column lists 14.4 ms
    // with 4 warps, 13.3 with 8 (each warp restarts its list from -inf: +29 % instructions); packed keys
    // 11.2 ms with 4 warps, 9.5 ms with 8 -> default with packed keys.  OSC_KNN_EW=4 / 8 overrides.
    bool ew8 = packed && f16;
    {
      const char* e = getenv("OSC_KNN_EW");
      if (e && atoi(e) == 8) ew8 = f16 && kc <= 16;
      if (e && atoi(e) == 4) ew8 = false;
    }
    if (ew8) {
      if (packed) OSC_TC2_LAUNCH_T(64 + 32 * 8, 16, true, true, 8, true);
      else OSC_TC2_LAUNCH_T(64 + 32 * 8, 16, true, true, 8);
    } else if (packed && f16) {
      OSC_TC2_LAUNCH(16, true, true, 4, true);
    } else if (packed) {
      OSC_TC2_LAUNCH(16, true, false, 4, true);
    } else if (f16) {
      if (kc <= 16) OSC_TC2_LAUNCH(16, true, true);
      else if (kc <= 24) OSC_TC2_LAUNCH(24, true, true);
      else OSC_TC2_LAUNCH(32, true, true);
    } else if (onepass) {
      if (kc <= 16) OSC_TC2_LAUNCH(16, true);
      else if (kc <= 24) OSC_TC2_LAUNCH(24, true);
      else OSC_TC2_LAUNCH(32, true);
    } else if (kc <= 8) OSC_TC2_LAUNCH(8, false);
    else if (kc <= 12) OSC_TC2_LAUNCH(12, false);
    else if (kc <= 16) OSC_TC2_LAUNCH(16, false);
    else if (kc <= 24) OSC_TC2_LAUNCH(24, false);
    else OSC_TC2_LAUNCH(32, false);
#undef OSC_TC2_LAUNCH
#undef OSC_TC2_LAUNCH_T
    OSC_LAUNCH_CHECK("knn_tc2_kernel");
    return OSC_OK;
  }
  const unsigned grid = (unsigned)(P.total_work < sms ? P.total_work : sms);
#define OSC_TC_LAUNCH(K)                                                                                 \
  do {                                                                                                   \
    OSC_CUDA(cudaFuncSetAttribute(knn_tc_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM)); \
    knn_tc_kernel<K><<<grid, TC_THREADS, TC_SMEM, st>>>(mqh, mql, mah, mal, P);                           \
  } while (0)
  if (kc <= 8) OSC_TC_LAUNCH(8);
  else if (kc <= 12) OSC_TC_LAUNCH(12);
  else if (kc <= 16) OSC_TC_LAUNCH(16);
  else if (kc <= 24) OSC_TC_LAUNCH(24);
  else OSC_TC_LAUNCH(32);
#undef OSC_TC_LAUNCH
  OSC_LAUNCH_CHECK("knn_tc_kernel");
  return OSC_OK;
}

}  // namespace osc
