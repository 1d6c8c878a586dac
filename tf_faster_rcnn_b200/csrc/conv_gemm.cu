// Dense stages of the Faster R-CNN TEST graph as ONE implicit-GEMM kernel for sm_100a:
//   slim.conv2d 3x3 / 1x1 (any stride), slim.fully_connected  (lib/nets/vgg16.py:26-60,
//   resnet_v1 bottlenecks, mobilenet pointwise, lib/nets/network.py:323-378)
//
// Math: D[M=pixels, N=cout] = sum over (filter tap, cin chunk) A_tap[M, 32] * W[N, 32]^T, fp32-grade via the
// 3xTF32 split  a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  (hi = RN_tf32(x), lo = RN_tf32(x - hi)),
// accumulated in fp32 in TMEM by tcgen05.mma.kind::tf32.
//
// Data movement: activations stay plain NHWC fp32 in HBM.  For filter tap (r,s) the A operand of a tile of
// tn x th x tw output pixels is ONE 4-D TMA box {32 ch, tw, th, tn} of the input at offset
// (w0*stride+s-pad_l, h0*stride+r-pad_t): out-of-bounds rows/cols are zero-filled by TMA, which IS the
// convolution's zero padding -- no im2col buffer ever exists.  Weights are pre-split (hi/lo planes), K-major,
// SWIZZLE_128B in shared memory (the UMMA B operand).
//
// r01 finding 1 (profiles/r01_*): with both operands in shared memory the kernel was bound by the shared-memory
// data pipe (3 MMAs re-read A and B, the splitter rewrote A twice: ~190 KB of smem traffic per 32-wide k-block,
// l1tex data pipe ~90% busy, tensor pipe 23-56%).  So the A operand now lives in TENSOR MEMORY: the splitter
// warps read the raw fp32 tile from smem once (un-swizzling their own 128-byte row), and tcgen05.st the hi and
// lo planes into a 4-deep TMEM ring; tcgen05.mma runs in TS mode (A from TMEM, B from smem).
// r01 finding 2: the tensor core adds into its fp32 accumulator with truncation (error grows linearly with the
// number of MMA steps, ~2.4e-5 relative at K=3136), so accumulation is two-level: TMEM holds only the partial
// sum of a short chunk of k-blocks (double buffered), which the epilogue warps add into fp32 REGISTER
// accumulators with round-to-nearest adds.  r01 finding 3: every such promotion stalls the tensor pipe for ~950
// cycles (tcgen05.ld of a 128x128 fp32 tile vs the MMAs' own TMEM traffic; independent of warp placement and of
// code shape), so the chunk length trades accuracy for speed: 1/2/4/8 k-blocks -> 6e-7/7e-7/1.0e-6/1.9e-6 error
// vs fp64 (the fp32 CPU reference sits at 0.6e-6..2e-6) for +37%/+20%/+9%/+5% time.  Default 4.
//
// Warp roles (320 threads, 1 CTA/SM), rings are 4 deep (index kb & 3).  The MMA issuer is the HIGHEST warp id of its
// scheduler partition on purpose: the arbiter favours high warp ids, and with the issuer at warp 1 the epilogue's
// burst of tcgen05.ld + FADDs delayed every chunk hand-over by ~900 cycles (profiles/r01 trace).
//   warp 8      TMA producer      A: wait a_empty[s] -> a_full[s];  B: wait mma_done[s] -> full[s] (tx)
//   warps 0..3  operand splitter  wait a_full[s], mma_done[s]; smem row -> hi/lo -> tcgen05.st; arrive a_empty[s], full[s]
//   warp 9      MMA issuer        wait full[s] (B landed + A stored); 4 k-slices x 3 tcgen05.mma (TS); commit -> mma_done[s];
//                                 per chunk: wait tmem_empty[b] first
//   warps 4..7  accumulate+epilogue  per chunk: wait mma_done[last k-block]; tcgen05.ld; acc += partial; arrive tmem_empty[b];
//                                 finally y = act(acc*scale + shift (+res)); st.global
// TMEM map (512 columns): [0, 2*BN) two accumulator buffers; [256, 512) A ring: slot s = 32 cols hi + 32 cols lo.
#include "common.cuh"
#include "../../include/frcnn_b200.h"
#include <stdlib.h>
#include <string.h>

namespace frcnn {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                          // fp32 elements: 128 B = one swizzle row
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KiB
constexpr int NUM_THREADS = 320;
constexpr int SPLIT_THREADS = 128;
constexpr int EPI_THREADS = 128;
constexpr int RING = 4;                              // depth of the A-raw, B and TMEM-A rings
constexpr int TMEM_COLS = 512;
constexpr int TMEM_A_COL0 = 256;

struct ConvKernelParams {
  float* out;
  const float* residual;
  const float* scale;
  const float* shift;
  int cout, ho, wo, nimg;
  int tn, th, tw;
  int tiles_h, tiles_w;
  int kh, kw, cin, stride, pad_t, pad_l;
  int act;
  int a_box_bytes;
  int kb_per_chunk;   // k-blocks accumulated in TMEM before promotion to registers
  long long* trace;   // debug: clock64() stamps of CTA (0,0)'s pipeline hand-offs; normally NULL
};

#define FRCNN_TRACE2(base, idx)                                                             \
  do {                                                                                      \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (idx) < 64) p.trace[(base) + (idx)] = clock64(); \
  } while (0)
#define FRCNN_TRACE(slot, kbv)                                                              \
  do {                                                                                      \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (kbv) < 64) p.trace[(kbv) * 8 + (slot)] = clock64(); \
  } while (0)

template <int BN> constexpr int b_stage_bytes() { return 2 * BN * BLOCK_K * 4; }
template <int BN> constexpr int smem_bytes() { return RING * (A_TILE_BYTES + b_stage_bytes<BN>()) + 1024 /*align slack*/ + 256 /*barriers*/; }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (8-row x 128 B atoms, 1024 B apart)
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                  // LBO: unused for swizzled K-major
  d |= (uint64_t)(1024u >> 4) << 32;       // SBO
  d |= (uint64_t)1 << 46;                  // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                        const __grid_constant__ CUtensorMap tmBlo, const ConvKernelParams p) {
  constexpr int kBTile = BN * BLOCK_K * 4;
  constexpr int kBStage = b_stage_bytes<BN>();
  constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  static_assert(2 * BN <= TMEM_A_COL0, "accumulator buffers overlap the TMEM A ring");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                               // RING x 16 KiB raw fp32 A tiles (TMA, swizzled)
  uint8_t* smem_b = smem + RING * A_TILE_BYTES;         // RING x (B_hi | B_lo)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + RING * kBStage);
  uint64_t* a_empty = a_full + RING;
  uint64_t* full = a_empty + RING;            // B bytes landed (tx) AND the 128 splitter threads stored A hi/lo
  uint64_t* mma_done = full + RING;
  uint64_t* tmem_full = mma_done + RING;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int mt = blockIdx.x;
  const int tile_w = mt % p.tiles_w;
  const int tile_h = (mt / p.tiles_w) % p.tiles_h;
  const int tile_n = mt / (p.tiles_w * p.tiles_h);
  const int w0 = tile_w * p.tw, h0 = tile_h * p.th, n0 = tile_n * p.tn;
  const int nblk = blockIdx.y;
  const int num_kb = p.kh * p.kw * (p.cin / BLOCK_K);
  const int num_chunks = (num_kb + p.kb_per_chunk - 1) / p.kb_per_chunk;
  // The completion of a chunk is the completion of its last k-block: the epilogue can wait on that k-block's
  // mma_done barrier instead of a second tcgen05.commit, as long as the barrier cannot complete a second time before
  // the epilogue looked (the MMA warp cannot start chunk c+2 before chunk c is drained): needs kb_per_chunk <= RING-1.
  const bool chunk_by_mma_done = p.kb_per_chunk <= RING - 1;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
    for (int s = 0; s < RING; ++s) {
      mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], SPLIT_THREADS);
      mbar_init(&full[s], SPLIT_THREADS + 1); mbar_init(&mma_done[s], 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], EPI_THREADS); }
    mbar_fence_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      const int cchunks = p.cin / BLOCK_K;
      int kb = 0;
      for (int r = 0; r < p.kh; ++r)
        for (int s = 0; s < p.kw; ++s)
          for (int kc = 0; kc < cchunks; ++kc, ++kb) {
            const int slot = kb & (RING - 1);
            const uint32_t par = (uint32_t)(kb / RING) & 1u;
            mbar_wait(&a_empty[slot], par ^ 1u);                 // splitter has consumed the raw tile
            mbar_expect_tx(&a_full[slot], (uint32_t)p.a_box_bytes);
            tma_load_4d(smem_a + slot * A_TILE_BYTES, &tmA, &a_full[slot], kc * BLOCK_K, w0 * p.stride + s - p.pad_l,
                        h0 * p.stride + r - p.pad_t, n0);
            mbar_wait(&mma_done[slot], par ^ 1u);                // MMAs that read this B slot have completed
            FRCNN_TRACE(0, kb);
            mbar_expect_tx(&full[slot], (uint32_t)(2 * kBTile));
            const int kcoord = ((r * p.kw + s) * p.cin) + kc * BLOCK_K;
            tma_load_2d(smem_b + slot * kBStage, &tmBhi, &full[slot], kcoord, nblk * BN);
            tma_load_2d(smem_b + slot * kBStage + kBTile, &tmBlo, &full[slot], kcoord, nblk * BN);
            FRCNN_TRACE(1, kb);
          }
    }
    __syncwarp();
  } else if (warp == 9) {
    if (lane == 0) {
      // one flat loop (chunk bookkeeping inline) so that the chunk hand-over runs the same, hot, instruction lines
      int in_chunk = 0, c = 0;
      uint32_t tmem_d = tmem_base;
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        const int slot = kb & (RING - 1);
        const uint32_t par = (uint32_t)(kb / RING) & 1u;
        if (in_chunk == 0) {
          const int b = c & 1;
          mbar_wait(&tmem_empty[b], ((uint32_t)(c >> 1) & 1u) ^ 1u);   // buffer drained by the epilogue warps
          tmem_d = tmem_base + (uint32_t)(b * BN);
          FRCNN_TRACE2(576, c);
        }
        mbar_wait(&full[slot], par);
        tc_fence_after();
        FRCNN_TRACE(4, kb);
        const uint32_t sb = smem_u32(smem_b + slot * kBStage);
        const uint64_t b_hi = sw128_desc(sb);
        const uint64_t b_lo = sw128_desc(sb + kBTile);
        const uint32_t a_hi = tmem_base + (uint32_t)(TMEM_A_COL0 + slot * 64);
        const uint32_t a_lo = a_hi + 32u;
#pragma unroll
        for (int k = 0; k < BLOCK_K / 8; ++k) {
          const uint64_t off = (uint64_t)(k * 8 * 4) >> 4;   // B: advance 8 tf32 = 32 B inside the swizzle row
          const uint32_t ak = (uint32_t)(k * 8);             // A: 8 tf32 = 8 TMEM columns
          umma_tf32_ts(tmem_d, a_lo + ak, b_hi + off, kIdesc, (k > 0 || in_chunk > 0) ? 1u : 0u);
          umma_tf32_ts(tmem_d, a_hi + ak, b_lo + off, kIdesc, 1u);
          umma_tf32_ts(tmem_d, a_hi + ak, b_hi + off, kIdesc, 1u);
        }
        umma_commit(&mma_done[slot]);
        FRCNN_TRACE(5, kb);
        if (++in_chunk == p.kb_per_chunk || kb + 1 == num_kb) {
          if (!chunk_by_mma_done) umma_commit(&tmem_full[c & 1]);
          in_chunk = 0; ++c;
        }
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ---------------- operand splitter: smem raw tile row -> (hi, lo) planes in the TMEM A ring ----------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int slot = kb & (RING - 1);
      const uint32_t par = (uint32_t)(kb / RING) & 1u;
      mbar_wait(&a_full[slot], par);
      // SWIZZLE_128B: 16-byte chunk c of row r sits at chunk (c ^ (r & 7)); quarter-warp phases are conflict-free
      const uint8_t* arow = smem_a + slot * A_TILE_BYTES + row * 128;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
        const float h0 = to_tf32(v.x), h1 = to_tf32(v.y), h2 = to_tf32(v.z), h3 = to_tf32(v.w);
        hi[4 * c + 0] = __float_as_uint(h0); hi[4 * c + 1] = __float_as_uint(h1);
        hi[4 * c + 2] = __float_as_uint(h2); hi[4 * c + 3] = __float_as_uint(h3);
        lo[4 * c + 0] = __float_as_uint(to_tf32(__fsub_rn(v.x, h0))); lo[4 * c + 1] = __float_as_uint(to_tf32(__fsub_rn(v.y, h1)));
        lo[4 * c + 2] = __float_as_uint(to_tf32(__fsub_rn(v.z, h2))); lo[4 * c + 3] = __float_as_uint(to_tf32(__fsub_rn(v.w, h3)));
      }
      mbar_arrive(&a_empty[slot]);                          // raw tile consumed (values are in registers)
      mbar_wait(&mma_done[slot], par ^ 1u);                  // TMEM A slot no longer read by the tensor core
      tc_fence_after();
      const uint32_t ta = tmem_base + lane_field + (uint32_t)(TMEM_A_COL0 + slot * 64);
      tmem_st_32x32(ta, hi);
      tmem_st_32x32(ta + 32u, lo);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&full[slot]);
      if (threadIdx.x == 0) FRCNN_TRACE(3, kb);
    }
  } else {
    // ---------------- accumulate (TMEM chunk partials -> fp32 registers, RN adds) ----------------
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    for (int c = 0; c < num_chunks; ++c) {
      const int b = c & 1;
      if (chunk_by_mma_done) {
        const int kb_last = min(num_kb, (c + 1) * p.kb_per_chunk) - 1;
        mbar_wait(&mma_done[kb_last & (RING - 1)], (uint32_t)(kb_last / RING) & 1u);
      } else {
        mbar_wait(&tmem_full[b], (uint32_t)(c >> 1) & 1u);
      }
      tc_fence_after();
      if (threadIdx.x == 128) FRCNN_TRACE(7, c);
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c0 + j] = __fadd_rn(acc[c0 + j], __uint_as_float(v[j]));
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[b]);
      if (threadIdx.x == 128) FRCNN_TRACE2(512, c);
    }
    // ---------------- epilogue ----------------
    const int row = q * 32 + lane;
    const int rows_img = p.th * p.tw;
    const int dn = row / rows_img, rem = row % rows_img;
    const int dh = rem / p.tw, dw = rem % p.tw;
    const int n = n0 + dn, h = h0 + dh, w = w0 + dw;
    const bool valid = (row < p.tn * rows_img) && n < p.nimg && h < p.ho && w < p.wo;
    if (valid) {
      const size_t pix = ((size_t)n * p.ho + h) * p.wo + w;
      float* orow = p.out + pix * p.cout;
      const float* rrow = p.residual ? p.residual + pix * p.cout : nullptr;
      const bool vec_ok = (p.cout & 3) == 0;
#pragma unroll
      for (int j = 0; j < BN; j += 4) {
        const int c = nblk * BN + j;
        if (c < p.cout) {
          float y[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a = acc[j + e];
            const int ce = c + e;
            if (ce < p.cout) {
              if (p.scale) a = __fmul_rn(a, __ldg(p.scale + ce));
              if (p.shift) a = __fadd_rn(a, __ldg(p.shift + ce));
              if (rrow) a = __fadd_rn(a, __ldg(rrow + ce));
              if (p.act == FRCNN_ACT_RELU) a = fmaxf(a, 0.f);
              else if (p.act == FRCNN_ACT_RELU6) a = fminf(fmaxf(a, 0.f), 6.f);
            }
            y[e] = a;
          }
          if (vec_ok) {
            *reinterpret_cast<float4*>(orow + c) = make_float4(y[0], y[1], y[2], y[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (c + e < p.cout) orow[c + e] = y[e];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const uint32_t* estr) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return ERR_DRIVER_ENTRY; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull,
              rank > 3 ? (unsigned long long)dims[3] : 0ull, box[0], box[1], rank > 2 ? box[2] : 0u, rank > 3 ? box[3] : 0u);
    return ERR_CUDA;
  }
  return OK;
}

}  // namespace frcnn

using namespace frcnn;

struct frcnn_conv_plan {
  CUtensorMap tmA, tmBhi, tmBlo;
  ConvKernelParams kp;
  int block_n, stages, smem;
  dim3 grid;
};

// choose the tile of output pixels (tn x th x tw <= 128) that needs the fewest tiles
static void choose_tile(int n, int ho, int wo, int stride, int* tn, int* th, int* tw) {
  long best_tiles = -1; int bn = 1, bh = 1, bw = 1; int best_rows = 0;
  const int lim = 256 / stride;
  for (int a = 1; a <= n && a <= BLOCK_M; ++a)
    for (int b = 1; b <= ho && a * b <= BLOCK_M && b <= lim; ++b) {
      int c = BLOCK_M / (a * b);
      if (c > wo) c = wo;
      if (c > lim) c = lim;
      if (c < 1) continue;
      // shrink c to the smallest width giving the same tile count (less garbage rows)
      int tiles_w = cdiv(wo, c);
      c = cdiv(wo, tiles_w);
      long tiles = (long)cdiv(n, a) * cdiv(ho, b) * tiles_w;
      int rows = a * b * c;
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && rows < best_rows)) {
        best_tiles = tiles; bn = a; bh = b; bw = c; best_rows = rows;
      }
    }
  *tn = bn; *th = bh; *tw = bw;
}

template <int BN>
static int launch(const frcnn_conv_plan* p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    FRCNN_CUDA(cudaFuncSetAttribute(conv_gemm_tf32x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN>()));
    attr_done = true;
  }
  conv_gemm_tf32x3_kernel<BN><<<p->grid, NUM_THREADS, smem_bytes<BN>(), st>>>(p->tmA, p->tmBhi, p->tmBlo, p->kp);
  FRCNN_LAUNCH_CHECK();
  return OK;
}

extern "C" int frcnn_conv_plan_create(frcnn_conv_plan** out, const frcnn_conv_desc* d) {
  FRCNN_REQUIRE(out && d, "null argument");
  FRCNN_REQUIRE(d->cin > 0 && d->cin % BLOCK_K == 0, "cin=%d must be a positive multiple of 32", d->cin);
  FRCNN_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->stride >= 1 && d->stride <= 8, "bad filter geometry");
  FRCNN_REQUIRE(d->in_dev && d->w_hi_dev && d->w_lo_dev && d->out_dev, "null device pointer");
  FRCNN_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->ho > 0 && d->wo > 0 && d->cout > 0, "bad shape");
  frcnn_conv_plan* p = (frcnn_conv_plan*)aligned_alloc(64, (sizeof(frcnn_conv_plan) + 63) / 64 * 64);
  if (!p) { set_error("out of host memory"); return ERR_ARG; }
  memset(p, 0, sizeof(*p));

  int n = d->n, h = d->h, w = d->w, ho = d->ho, wo = d->wo;
  const bool pointwise = d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad_t == 0 && d->pad_l == 0 && ho == h && wo == w;
  if (pointwise) {  // flatten all pixels into one row of "width" n*h*w: perfect 128-row tiles
    w = wo = n * h * w; n = 1; h = ho = 1;
  }
  int tn, th, tw;
  choose_tile(n, ho, wo, d->stride, &tn, &th, &tw);
  const int tiles_w = cdiv(wo, tw), tiles_h = cdiv(ho, th), tiles_n = cdiv(n, tn);
  const long m_tiles = (long)tiles_w * tiles_h * tiles_n;
  const int num_kb = d->kh * d->kw * d->cin / BLOCK_K;

  int bn = d->block_n;
  if (bn == 0) {
    long best = -1;
    const int cands[3] = {128, 64, 32};
    for (int i = 0; i < 3; ++i) {
      const int c = cands[i];
      if (c > 32 && c / 2 >= d->cout) continue;        // tile mostly empty
      const long ctas = m_tiles * cdiv(d->cout, c);
      const long waves = (ctas + 147) / 148;
      // measured (profiles/r01): a k-block costs ~1400 cycles whatever block_n is (SS-mode tcgen05.mma is bound by the
      // 128-row A operand read for N <= 128), plus ~6000 cycles of prologue/drain per CTA => fewest waves wins, widest tile on ties
      const long cost = waves * (num_kb * 1400L + 6000L);
      if (best < 0 || cost < best) { best = cost; bn = c; }
    }
  }
  FRCNN_REQUIRE(bn == 32 || bn == 64 || bn == 128, "block_n must be 32, 64 or 128");

  // A: NHWC activations as a rank-4 tensor {C, W, H, N}; traversal stride = conv stride on W and H
  {
    uint64_t dims[4] = {(uint64_t)d->cin, (uint64_t)w, (uint64_t)h, (uint64_t)n};
    uint64_t strides[3] = {(uint64_t)d->cin * 4, (uint64_t)w * d->cin * 4, (uint64_t)h * w * d->cin * 4};
    uint32_t box[4] = {(uint32_t)BLOCK_K, (uint32_t)(tw * d->stride), (uint32_t)(th * d->stride), (uint32_t)tn};
    uint32_t es[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
    int rc = encode_map(&p->tmA, d->in_dev, 4, dims, strides, box, es);
    if (rc) { free(p); return rc; }
  }
  {
    const uint64_t ktot = (uint64_t)d->kh * d->kw * d->cin;
    uint64_t dims[2] = {ktot, (uint64_t)d->cout};
    uint64_t strides[1] = {ktot * 4};
    uint32_t box[2] = {(uint32_t)BLOCK_K, (uint32_t)bn};
    uint32_t es[2] = {1, 1};
    int rc = encode_map(&p->tmBhi, d->w_hi_dev, 2, dims, strides, box, es);
    if (!rc) rc = encode_map(&p->tmBlo, d->w_lo_dev, 2, dims, strides, box, es);
    if (rc) { free(p); return rc; }
  }
  ConvKernelParams& k = p->kp;
  k.out = d->out_dev; k.residual = d->residual_dev; k.scale = d->scale_dev; k.shift = d->shift_dev;
  k.cout = d->cout; k.ho = ho; k.wo = wo; k.nimg = n;
  k.tn = tn; k.th = th; k.tw = tw; k.tiles_h = tiles_h; k.tiles_w = tiles_w;
  k.kh = d->kh; k.kw = d->kw; k.cin = d->cin; k.stride = d->stride; k.pad_t = d->pad_t; k.pad_l = d->pad_l;
  k.act = d->act;
  k.a_box_bytes = tn * th * tw * BLOCK_K * 4;
  k.kb_per_chunk = d->kb_per_chunk > 0 ? d->kb_per_chunk : 4;
  k.trace = nullptr;
  p->block_n = bn;
  p->stages = RING;
  p->smem = bn == 128 ? smem_bytes<128>() : bn == 64 ? smem_bytes<64>() : smem_bytes<32>();
  FRCNN_REQUIRE(m_tiles <= 0x7fffffffL, "too many tiles");
  p->grid = dim3((unsigned)m_tiles, (unsigned)cdiv(d->cout, bn), 1);
  *out = p;
  return OK;
}

extern "C" int frcnn_conv_plan_run(const frcnn_conv_plan* p, void* stream) {
  FRCNN_REQUIRE(p, "null plan");
  cudaStream_t st = (cudaStream_t)stream;
  switch (p->block_n) {
    case 128: return launch<128>(p, st);
    case 64: return launch<64>(p, st);
    default: return launch<32>(p, st);
  }
}

extern "C" int frcnn_conv_plan_info(const frcnn_conv_plan* p, int* block_n, int* tile_n, int* tile_h, int* tile_w,
                                    int* grid_m, int* grid_n, int* stages, int* smem) {
  FRCNN_REQUIRE(p, "null plan");
  if (block_n) *block_n = p->block_n;
  if (tile_n) *tile_n = p->kp.tn;
  if (tile_h) *tile_h = p->kp.th;
  if (tile_w) *tile_w = p->kp.tw;
  if (grid_m) *grid_m = (int)p->grid.x;
  if (grid_n) *grid_n = (int)p->grid.y;
  if (stages) *stages = p->stages;
  if (smem) *smem = p->smem;
  return OK;
}

extern "C" int frcnn_conv_plan_set_trace(frcnn_conv_plan* p, long long* trace_dev) {
  FRCNN_REQUIRE(p, "null plan");
  p->kp.trace = trace_dev;
  return OK;
}

extern "C" void frcnn_conv_plan_destroy(frcnn_conv_plan* p) { free(p); }
