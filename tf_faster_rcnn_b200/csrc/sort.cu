// Descending, stable key sort of the RPN scores (lib/layer_utils/proposal_layer.py:56-60 `scores.argsort()[::-1]`,
// tf.nn.top_k in proposal_top_layer / the score ordering inside tf.image.non_max_suppression) as ONE launch:
// a thread-block cluster of 8 CTAs per image holds the (key, index) pairs in distributed shared memory and runs a
// 4-pass LSD radix sort (8-bit digits).  Replaces r01's CUB DeviceRadixSort (7 library launches, ~70 us for 22 800 keys).
//
// Per pass, per CTA (its slice of <= npc pairs sits in its own shared memory, in the current global order):
//   (1) warp w walks its contiguous chunk 32 pairs at a time; __match_any_sync groups equal digits, so every pair gets
//       its rank among the equal-digit pairs that precede it in the warp's chunk (registers) and the warp's digit
//       histogram is built without atomics;
//   (2) 256 threads turn the 32 warp histograms into exclusive per-warp offsets + the CTA's digit totals;
//   (3) cluster barrier; every CTA reads the 8 CTA totals of each digit through DSMEM and derives the global base of
//       (digit, this CTA) = pairs with a smaller digit anywhere + equal-digit pairs in lower-ranked CTAs (block scan);
//   (4) pair -> global position = base + warp offset + rank, written straight into the owning CTA's other buffer with
//       st.shared::cluster; cluster barrier.
// Equal keys never change their relative order (every pass is stable), so ties come out in ascending index order --
// the oracle's tie rule.  Key transform: the usual order-preserving float -> uint map, complemented (descending).
#include "common.cuh"
#include "../../include/frcnn_b200.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace frcnn {

constexpr int SORT_CLUSTER = 8;
constexpr int SORT_THREADS = 1024;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_NPC_MAX = 11264;                    // pairs per CTA: 8 * 11264 = 90 112 keys per segment
constexpr int SORT_MAX_ROUNDS = SORT_NPC_MAX / SORT_THREADS;   // 32-pair rounds per warp

__device__ __forceinline__ uint32_t desc_bits(float f) {
  uint32_t u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // ascending order of u == ascending order of f
  return ~u;                                           // ascending order of the result == descending order of f
}
__device__ __forceinline__ float desc_bits_inv(uint32_t k) {
  const uint32_t u = ~k;
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void __launch_bounds__(SORT_THREADS, 1)
cluster_sort_desc_kernel(const float* __restrict__ keys_in, int n, int npc, int* __restrict__ order, float* __restrict__ sorted_keys) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int seg = blockIdx.x / SORT_CLUSTER;
  keys_in += (size_t)seg * n; order += (size_t)seg * n; sorted_keys += (size_t)seg * n;
  const int my0 = rank * npc;
  const int my_n = max(0, min(npc, n - my0));

  extern __shared__ uint32_t sort_smem[];
  uint32_t* kbuf0 = sort_smem;                         // [2][npc] keys (transformed)
  uint32_t* ibuf0 = sort_smem + 2 * (size_t)npc;       // [2][npc] indices
  uint32_t* whist = ibuf0 + 2 * (size_t)npc;           // [32 warps][256 digits]
  uint32_t* cta_hist = whist + SORT_WARPS * 256;       // [256]
  uint32_t* base = cta_hist + 256;                     // [256]
  __shared__ uint32_t s_warp_tot[8];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < my_n; i += SORT_THREADS) { kbuf0[i] = desc_bits(__ldg(keys_in + my0 + i)); ibuf0[i] = (uint32_t)(my0 + i); }
  // warp chunk: ipw consecutive pairs (multiple of 32), the same partition in every pass
  const int ipw = ((npc + SORT_WARPS - 1) / SORT_WARPS + 31) / 32 * 32;
  const int rounds = ipw / 32;
  int cur = 0;
  cluster.sync();                                      // every CTA of the cluster is running (DSMEM is accessible) + local loads done

  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    uint32_t* kb = kbuf0 + cur * npc;
    uint32_t* ib = ibuf0 + cur * npc;
    for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) whist[i] = 0u;
    __syncthreads();
    // (1) ranks within the warp chunk
    uint32_t kk[SORT_MAX_ROUNDS], ii[SORT_MAX_ROUNDS], pre[SORT_MAX_ROUNDS];
#pragma unroll
    for (int j = 0; j < SORT_MAX_ROUNDS; ++j) {
      if (j < rounds) {
        const int i = warp * ipw + j * 32 + lane;
        const bool valid = i < my_n;
        kk[j] = valid ? kb[i] : 0u; ii[j] = valid ? ib[i] : 0u;
        const uint32_t d = valid ? ((kk[j] >> shift) & 255u) : 256u;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0u;
        if (valid && lane == leader) { prev = whist[warp * 256 + d]; whist[warp * 256 + d] = prev + (uint32_t)__popc(peers); }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        pre[j] = prev + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        __syncwarp();
      }
    }
    __syncthreads();
    // (2) per-warp exclusive offsets + CTA totals per digit
    if (tid < 256) {
      uint32_t run = 0u;
      for (int w = 0; w < SORT_WARPS; ++w) { const uint32_t t = whist[w * 256 + tid]; whist[w * 256 + tid] = run; run += t; }
      cta_hist[tid] = run;
    }
    cluster.sync();
    // (3) global base of (digit, this CTA)
    if (tid < 256) {
      uint32_t tot = 0u, before = 0u;
#pragma unroll
      for (int r = 0; r < SORT_CLUSTER; ++r) {
        const uint32_t v = cluster.map_shared_rank(cta_hist, r)[tid];
        if (r < rank) before += v;
        tot += v;
      }
      uint32_t incl = tot;
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      if (lane == 31) s_warp_tot[warp] = incl;
      // (only warps 0..7 take part: a named barrier among their 256 threads)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint32_t wbefore = 0u;
      for (int w = 0; w < warp; ++w) wbefore += s_warp_tot[w];
      base[tid] = wbefore + incl - tot + before;
    }
    __syncthreads();
    // (4) scatter into the owners' other buffer
    uint32_t* kdst = kbuf0 + (cur ^ 1) * npc;
    uint32_t* idst = ibuf0 + (cur ^ 1) * npc;
#pragma unroll
    for (int j = 0; j < SORT_MAX_ROUNDS; ++j) {
      if (j < rounds) {
        const int i = warp * ipw + j * 32 + lane;
        if (i < my_n) {
          const uint32_t d = (kk[j] >> shift) & 255u;
          const uint32_t pos = base[d] + whist[warp * 256 + d] + pre[j];
          const uint32_t dr = pos / (uint32_t)npc, loc = pos - dr * (uint32_t)npc;
          cluster.map_shared_rank(kdst, dr)[loc] = kk[j];
          cluster.map_shared_rank(idst, dr)[loc] = ii[j];
        }
      }
    }
    cluster.sync();
    cur ^= 1;
  }
  const uint32_t* kb = kbuf0 + cur * npc;
  const uint32_t* ib = ibuf0 + cur * npc;
  for (int i = tid; i < my_n; i += SORT_THREADS) {
    order[my0 + i] = (int)ib[i] ;
    sorted_keys[my0 + i] = desc_bits_inv(kb[i]);
  }
}

static size_t sort_smem_bytes(int npc) { return ((size_t)4 * npc + SORT_WARPS * 256 + 512) * sizeof(uint32_t); }

}  // namespace frcnn

using namespace frcnn;

extern "C" size_t frcnn_sort_workspace_bytes(int n) {
  (void)n;
  return 256;   // the sort lives entirely in (distributed) shared memory; a token workspace keeps the r01 signature usable
}

extern "C" int frcnn_sort_desc(const float* keys, int n, int batch, int* order, float* sorted_keys, void* workspace, size_t workspace_bytes,
                               void* stream) {
  (void)workspace; (void)workspace_bytes;
  FRCNN_REQUIRE(keys && order && sorted_keys && n > 0 && batch > 0, "sort_desc: bad argument");
  const int npc = cdiv(n, SORT_CLUSTER);
  if (npc > SORT_NPC_MAX) { set_error("sort_desc: %d keys per segment > capacity %d", n, SORT_CLUSTER * SORT_NPC_MAX); return ERR_CAPACITY; }
  const size_t smem = sort_smem_bytes(npc);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    FRCNN_CUDA(cudaFuncSetAttribute(cluster_sort_desc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem_bytes(SORT_NPC_MAX)));
    attr_smem = sort_smem_bytes(SORT_NPC_MAX);
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = SORT_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(SORT_CLUSTER * batch)); cfg.blockDim = dim3(SORT_THREADS); cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream; cfg.attrs = attr; cfg.numAttrs = 1;
  FRCNN_CUDA(cudaLaunchKernelEx(&cfg, cluster_sort_desc_kernel, keys, n, npc, order, sorted_keys));
  return OK;
}
