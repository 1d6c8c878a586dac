// Both NMS passes of the TEST path + the score sort, for sm_100a.
//   RPN stage   : proposal_layer_tf / proposal_layer / proposal_top_layer (lib/layer_utils/proposal_layer.py:16-83,
//                 proposal_top_layer.py:58-85)                                   -> frcnn_proposals
//   final stage : per-class NMS + max_per_image cap (lib/model/test.py:162-180) -> frcnn_detect_post
//   `_nms` ABI  : lib/nms/gpu_nms.hpp:1-2                                        -> frcnn_nms_host
//
// Algorithm (replaces nms_kernel.cu's N x N/64 bitmask + host sweep, which is O(N^2) work and memory even
// though at most post_nms_top_n boxes survive): candidates are walked in priority order in windows of 1024.
// Per round the CTA (a) tests every candidate of the window against the <=K boxes already kept (kept set in shared
// memory) and compacts the survivors, (b) builds the <=256x256 suppression bitmask among them, (c) one thread resolves
// it sequentially with 4 x 64-bit words; survivors are appended to the kept set.  The walk stops as soon as
// max_out boxes are kept, so the RPN stage touches only the first few thousand of the 17k-50k anchors.
// IoU arithmetic is the oracle's op-by-op fp32 sequence (__f*_rn: no FMA contraction) for all three predicate
// variants (flags), so survivor indices are bit-exact.
#include "common.cuh"
#include "../../include/frcnn_b200.h"

namespace frcnn {

constexpr int NMS_THREADS = 1024;
constexpr int CHUNK = 256;

__device__ __forceinline__ float box_area(const float4 b, unsigned flags) {
  if (flags & FRCNN_NMS_PLUS_ONE) return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// TF normalises corners with min/max first (boxes are (y1,x1,y2,x2) there; the formula is symmetric in x/y)
__device__ __forceinline__ float4 canon(const float4 b, unsigned flags) {
  if (flags & FRCNN_NMS_PLUS_ONE) return b;
  return make_float4(fminf(b.x, b.z), fminf(b.y, b.w), fmaxf(b.x, b.z), fmaxf(b.y, b.w));
}

__device__ __forceinline__ bool suppresses(const float4 a, float area_a, const float4 b, float area_b, float thr, unsigned flags) {
  float inter;
  if (flags & FRCNN_NMS_PLUS_ONE) {
    const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.f));
    const float h = fmaxf(0.f, __fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.f));
    inter = __fmul_rn(w, h);
  } else {
    if ((flags & FRCNN_NMS_SKIP_DEGENERATE) && (area_a <= 0.f || area_b <= 0.f)) return false;
    const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    inter = __fmul_rn(h, w);
  }
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return (flags & FRCNN_NMS_INCLUSIVE) ? (ovr >= thr) : (ovr > thr);
}

constexpr int WIDE = 1024;    // candidates filtered against the kept set per round (one per thread)

struct GreedyShared {
  unsigned long long mask[CHUNK][CHUNK / 64];
  float4 cbox[CHUNK];
  float carea[CHUNK];
  int cpos[CHUNK];                 // position (in priority order) of each compacted candidate
  int warp_cnt[WIDE / 32];
  unsigned char chunk_kept[CHUNK];
  int nkept;
  int chunk_nk;
  int n_alive;                     // compacted candidates this round (<= CHUNK)
  int consumed;                    // candidates of the window that are settled after this round
  int tot_alive, last_wn;          // survivors / size of the previous window (drives threads-per-candidate)
};

// CTA-cooperative greedy NMS over `m` candidates given in priority order (blockDim.x == 1024).
//   cand(i) -> float4 box of the i-th candidate;  kept/kept_area: storage for the kept set (shared or global)
//   kept_pos: positions (0..m-1) of survivors.   Number kept is left in sh.nkept (valid after the final barrier).
// Round: (a) a window of up to 1024 candidates is tested against the kept set, one candidate per thread; the survivors are
// compacted in order (at most CHUNK = 256 of them -- the window is cut right after the 256th survivor and the rest is
// re-examined next round), (b) the CHUNK x CHUNK suppression bitmask among the survivors is built, (c) one thread resolves it
// sequentially with 4 x 64-bit words, skipping suppressed candidates with ffs.  With a low keep rate (RPN: most anchors
// overlap something already kept) almost everything dies in (a) and a round settles ~1000 candidates.
template <typename CandFn>
__device__ void block_greedy_nms(CandFn cand, int m, float thr, unsigned flags, int max_out, float4* kept, float* kept_area,
                                 int* kept_pos, GreedyShared& sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { sh.nkept = 0; sh.tot_alive = WIDE; sh.last_wn = WIDE; }
  __syncthreads();
  int base = 0;
  while (base < m) {
    const int nk = sh.nkept;
    if (nk >= max_out) break;
    // threads per candidate: when most of the last window survived (high keep rate) a round can only settle ~256 candidates,
    // so look at 256 with 4 threads each; when few survive, look at 1024 with one thread each.
    const int tpc = (sh.tot_alive * 3 >= sh.last_wn) ? 4 : 1;
    const int wn = min(WIDE / tpc, m - base);
    const int ci = tid / tpc, part = tid % tpc;
    // (a) filter the window against the kept set
    bool dead = false;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    float ab = 0.f;
    if (ci < wn) {
      b = canon(cand(base + ci), flags);
      ab = box_area(b, flags);
      if (thr >= 0.f)
        for (int k = part; k < nk; k += tpc)
          if (suppresses(kept[k], kept_area[k], b, ab, thr, flags)) { dead = true; break; }
    }
    const unsigned dbal = __ballot_sync(0xffffffffu, dead);
    const bool any_dead = tpc == 1 ? dead : (((dbal >> (lane & ~3)) & 0xFu) != 0u);
    const bool alive = (ci < wn) && part == 0 && !any_dead;       // one representative thread per candidate
    // ordered compaction: exclusive prefix of `alive` over the block
    const unsigned bal = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) sh.warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += sh.warp_cnt[w];
    const int idx = before + __popc(bal & ((1u << lane) - 1u));
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < WIDE / 32; ++w) tot += sh.warp_cnt[w];
      sh.n_alive = min(tot, CHUNK);
      sh.tot_alive = tot; sh.last_wn = wn;
      if (tot <= CHUNK) sh.consumed = wn;
    }
    if (alive && idx < CHUNK) { sh.cbox[idx] = b; sh.carea[idx] = ab; sh.cpos[idx] = base + ci; }
    if (alive && idx == CHUNK) sh.consumed = ci;           // first survivor that does not fit: the window is cut here
    __syncthreads();
    const int cn = sh.n_alive;
    // (b) suppression bitmask among the survivors: thread (row i, 64-bit word wj); only j > i matters
    if (thr >= 0.f) {
      const int i = tid >> 2, wj = tid & 3;
      if (i < cn) {
        unsigned long long bits = 0ull;
        const float4 a = sh.cbox[i];
        const float aa = sh.carea[i];
        const int j0 = wj * 64;
        for (int j = max(j0, i + 1); j < min(j0 + 64, cn); ++j)
          if (suppresses(a, aa, sh.cbox[j], sh.carea[j], thr, flags)) bits |= 1ull << (j - j0);
        sh.mask[i][wj] = bits;
      }
    }
    __syncthreads();
    // (c) sequential resolve (ffs skips suppressed candidates)
    if (tid == 0) {
      static_assert(CHUNK == 256, "resolve loop is written for 4 x 64-bit words");
      unsigned long long rem[4] = {0ull, 0ull, 0ull, 0ull};
      int k = nk, ck = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int lim = min(64, cn - w * 64);
        if (lim <= 0) break;
        const unsigned long long valid = lim == 64 ? ~0ull : ((1ull << lim) - 1ull);
        unsigned long long todo = ~rem[w] & valid;
        while (todo && k < max_out) {
          const int bit = __ffsll((long long)todo) - 1;
          const int i = w * 64 + bit;
          sh.chunk_kept[ck++] = (unsigned char)i;
          ++k;
          if (thr >= 0.f) {
#pragma unroll
            for (int w2 = 0; w2 < 4; ++w2) if (w2 >= w) rem[w2] |= sh.mask[i][w2];
          }
          todo = (todo & ~rem[w]) & ~((2ull << bit) - 1ull);
        }
      }
      sh.chunk_nk = ck;
    }
    __syncthreads();
    const int ck = sh.chunk_nk;
    if (tid < ck) {
      const int i = sh.chunk_kept[tid];
      kept[nk + tid] = sh.cbox[i];
      kept_area[nk + tid] = sh.carea[i];
      kept_pos[nk + tid] = sh.cpos[i];
    }
    const int consumed = sh.consumed;
    __syncthreads();
    if (tid == 0) sh.nkept = nk + ck;
    base += consumed;
    __syncthreads();
  }
  __syncthreads();
}

// ---- RPN proposal selection ---------------------------------------------------------------------------
constexpr int PROPOSAL_CAP = 1024;   // kept set held in shared memory (post_nms_top_n: 300 / 1000 / 'top' handled separately)

__global__ void __launch_bounds__(NMS_THREADS, 1)
proposals_kernel(const float4* __restrict__ props, const float* __restrict__ scores, const int* __restrict__ order, int n_seg, int m,
                 int max_out, float thr, unsigned flags, float* __restrict__ rois, float* __restrict__ roi_scores,
                 int* __restrict__ keep, int* __restrict__ num) {
  __shared__ GreedyShared sh;
  __shared__ float4 kept[PROPOSAL_CAP];
  __shared__ float kept_area[PROPOSAL_CAP];
  __shared__ int kept_pos[PROPOSAL_CAP];
  // one CTA per image of the batch: segment blockIdx.x of the per-image arrays (order holds segment-local indices)
  const int img = blockIdx.x;
  props += (size_t)img * n_seg; scores += (size_t)img * n_seg; order += (size_t)img * n_seg;
  rois += (size_t)img * max_out * 5; roi_scores += (size_t)img * max_out; keep += (size_t)img * max_out; num += img;
  block_greedy_nms([&](int i) { return __ldg(props + __ldg(order + i)); }, m, thr, flags, max_out, kept, kept_area, kept_pos, sh);
  const int nk = sh.nkept;
  for (int i = threadIdx.x; i < max_out; i += blockDim.x) {
    float* r = rois + (size_t)i * 5;
    if (i < nk) {
      const int src = __ldg(order + kept_pos[i]);
      const float4 b = __ldg(props + src);   // un-normalised original box
      r[0] = (float)img; r[1] = b.x; r[2] = b.y; r[3] = b.z; r[4] = b.w;
      roi_scores[i] = __ldg(scores + src);
      keep[i] = src;
    } else {
      r[0] = r[1] = r[2] = r[3] = r[4] = 0.f;
      roi_scores[i] = 0.f;
      keep[i] = -1;
    }
  }
  if (threadIdx.x == 0) *num = nk;
}

// 'top' mode (no NMS) with more outputs than the shared-memory kept set: plain gather of the first max_out
__global__ void gather_top_kernel(const float4* __restrict__ props, const float* __restrict__ scores, const int* __restrict__ order,
                                  int n_seg, int m, int max_out, float* __restrict__ rois, float* __restrict__ roi_scores,
                                  int* __restrict__ keep, int* __restrict__ num) {
  const int img = blockIdx.y;
  props += (size_t)img * n_seg; scores += (size_t)img * n_seg; order += (size_t)img * n_seg;
  rois += (size_t)img * max_out * 5; roi_scores += (size_t)img * max_out; keep += (size_t)img * max_out; num += img;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *num = min(m, max_out);
  if (i >= max_out) return;
  float* r = rois + (size_t)i * 5;
  if (i < m) {
    const int src = __ldg(order + i);
    const float4 b = __ldg(props + src);
    r[0] = (float)img; r[1] = b.x; r[2] = b.y; r[3] = b.z; r[4] = b.w;
    roi_scores[i] = __ldg(scores + src);
    keep[i] = src;
  } else {
    r[0] = r[1] = r[2] = r[3] = r[4] = 0.f;
    roi_scores[i] = 0.f;
    keep[i] = -1;
  }
}

// ---- generic sorted-input NMS with the kept set in global memory (the `_nms` compatible path) -----------
__global__ void __launch_bounds__(NMS_THREADS, 1)
nms_sorted_kernel(const float* __restrict__ boxes, int stride, int m, float thr, unsigned flags, int max_out,
                  float4* __restrict__ kept, float* __restrict__ kept_area, int* __restrict__ keep, int* __restrict__ num) {
  __shared__ GreedyShared sh;
  block_greedy_nms([&](int i) { const float* b = boxes + (size_t)i * stride; return make_float4(b[0], b[1], b[2], b[3]); }, m, thr,
                   flags, max_out, kept, kept_area, keep, sh);
  if (threadIdx.x == 0) *num = sh.nkept;
}

// ---- final per-class NMS + cap -----------------------------------------------------------------------------
constexpr int DET_CAP = 1024;       // RoIs per image with the candidate list AND the kept set in shared memory (300 / 1000 proposals)
constexpr int DET_CAP_BIG = 8192;   // TEST.MODE 'top' (RPN_TOP_N = 5000, lib/model/config.py:208): candidates in shared, kept set in global memory

// one CTA per (foreground class, image).  CAP = power of two >= r.  Dynamic shared memory: skey[CAP] | sidx[CAP] and, when
// !GLOBAL_KEPT, kept[CAP] (float4) | kept_area[CAP] | kept_pos[CAP]; with GLOBAL_KEPT those three live in `gws`
// ([batch][C][r] slices, see frcnn_detect_post_workspace_bytes).
template <int CAP, bool GLOBAL_KEPT>
__global__ void __launch_bounds__(NMS_THREADS, 1)
class_nms_kernel(const float* __restrict__ probs, const float4* __restrict__ pred, const int* __restrict__ num_rois, int r, int C,
                 float score_thresh, float nms_thresh, unsigned flags, int* __restrict__ keep, int* __restrict__ keep_cnt,
                 float* __restrict__ keep_score, uint8_t* __restrict__ gws) {
  __shared__ GreedyShared sh;
  __shared__ int s_m;
  extern __shared__ __align__(16) uint8_t cls_dyn[];
  float4* kept; float* kept_area; int* kept_pos; float* skey; int* sidx;
  const int cls = blockIdx.x + 1, img = blockIdx.y;
  if (GLOBAL_KEPT) {
    skey = reinterpret_cast<float*>(cls_dyn); sidx = reinterpret_cast<int*>(skey + CAP);
    const size_t rr = (size_t)((r + 3) & ~3);          // slices stay 16-byte aligned
    uint8_t* slice = gws + ((size_t)img * C + cls) * rr * 24;
    kept = reinterpret_cast<float4*>(slice); kept_area = reinterpret_cast<float*>(slice + rr * 16);
    kept_pos = reinterpret_cast<int*>(slice + rr * 20);
  } else {
    kept = reinterpret_cast<float4*>(cls_dyn); kept_area = reinterpret_cast<float*>(kept + CAP);
    kept_pos = reinterpret_cast<int*>(kept_area + CAP); skey = reinterpret_cast<float*>(kept_pos + CAP); sidx = reinterpret_cast<int*>(skey + CAP);
  }
  probs += (size_t)img * r * C; pred += (size_t)img * r * C;
  keep += (size_t)img * C * r; keep_score += (size_t)img * C * r; keep_cnt += (size_t)img * C;
  const int tid = threadIdx.x;
  const int nr = min(__ldg(num_rois + img), r);
  // candidates: score > thresh (test.py:163); sort key (score desc, index asc); invalid -> -inf at the tail
  if (tid == 0) s_m = 0;
  __syncthreads();
  int mine = 0;
  for (int e = tid; e < CAP; e += NMS_THREADS) {
    float key = __int_as_float(0xff800000);
    if (e < nr) {
      const float s = __ldg(probs + (size_t)e * C + cls);
      if (s > score_thresh) { key = s; ++mine; }
    }
    skey[e] = key; sidx[e] = e;
  }
  if (mine) atomicAdd(&s_m, mine);
  __syncthreads();
  // bitonic sort of CAP (key desc, idx asc)
  for (int k = 2; k <= CAP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = tid; e < CAP; e += NMS_THREADS) {
        const int ixj = e ^ j;
        if (ixj > e) {
          const float a = skey[e], b = skey[ixj];
          const int ia = sidx[e], ib = sidx[ixj];
          const bool a_first = (a > b) || (a == b && ia < ib);   // a precedes b in the final order
          const bool up = (e & k) == 0;
          if (up ? !a_first : a_first) { skey[e] = b; skey[ixj] = a; sidx[e] = ib; sidx[ixj] = ia; }
        }
      }
      __syncthreads();
    }
  }
  const int m = s_m;
  block_greedy_nms([&](int i) { return __ldg(pred + (size_t)sidx[i] * C + cls); }, m, nms_thresh, flags, m, kept, kept_area, kept_pos, sh);
  const int nk = sh.nkept;
  for (int i = tid; i < r; i += blockDim.x) {
    if (i < nk) { const int pos = kept_pos[i]; keep[(size_t)cls * r + i] = sidx[pos]; keep_score[(size_t)cls * r + i] = skey[pos]; }
    else { keep[(size_t)cls * r + i] = -1; keep_score[(size_t)cls * r + i] = 0.f; }
  }
  if (tid == 0) keep_cnt[cls] = nk;
  if (blockIdx.x == 0) {
    for (int i = tid; i < r; i += blockDim.x) { keep[i] = -1; keep_score[i] = 0.f; }
    if (tid == 0) keep_cnt[0] = 0;
  }
}

// single CTA: max_per_image cap + record emission (test.py:173-180).  Every class list is sorted by descending score, so
// (1) count_c(t) = #scores >= t is a binary search per class, (2) the k-th largest kept score overall is found by a bitwise
// search on the fp32 pattern (scores > 0 => unsigned order == float order; 32 rounds of 80 parallel binary searches + a block
// reduction), (3) "keep score >= image_thresh" just truncates each list to count_c(thresh) (ties kept, as the reference).
__device__ __forceinline__ int count_ge(const float* __restrict__ sorted_desc, int n, unsigned tbits) {
  int lo = 0, hi = n;                       // first index whose score < t
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__float_as_uint(sorted_desc[mid]) >= tbits) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(NMS_THREADS, 1)
cap_emit_kernel(const float4* __restrict__ pred, int r, int C, int max_per_image, int max_det, int* __restrict__ keep,
                int* __restrict__ keep_cnt, const float* __restrict__ keep_score, float* __restrict__ det, int* __restrict__ ndet,
                int det_stride, int ndet_stride) {
  __shared__ int s_warp[NMS_THREADS / 32];
  __shared__ int s_total;
  __shared__ int s_off[1025];
  const int img = blockIdx.x;                       // one CTA per image of the batch
  pred += (size_t)img * r * C; keep += (size_t)img * C * r; keep_cnt += (size_t)img * C; keep_score += (size_t)img * C * r;
  det += (size_t)img * det_stride; ndet += (size_t)img * ndet_stride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_cls = tid >= 1 && tid < C;
  const int my_cnt = is_cls ? keep_cnt[tid] : 0;
  const float* my_scores = keep_score + (size_t)tid * r;
  auto block_sum = [&](int v) -> int {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < NMS_THREADS / 32; ++w) t += s_warp[w]; s_total = t; }
    __syncthreads();
    return s_total;
  };
  const int total = block_sum(my_cnt);
  int new_cnt = my_cnt;
  if (max_per_image > 0 && total > max_per_image) {
    unsigned t = 0u;                                   // largest pattern with count(score >= t) >= max_per_image
    for (int bit = 31; bit >= 0; --bit) {
      const unsigned cand = t | (1u << bit);
      const int cnt = block_sum(is_cls ? count_ge(my_scores, my_cnt, cand) : 0);
      if (cnt >= max_per_image) t = cand;
    }
    if (is_cls) new_cnt = count_ge(my_scores, my_cnt, t);
  }
  if (is_cls) {
    for (int j = new_cnt; j < my_cnt; ++j) keep[(size_t)tid * r + j] = -1;
    keep_cnt[tid] = new_cnt;
  }
  // exclusive prefix of the per-class counts (C <= 1024: one thread per class, warp scan + warp totals)
  int v = is_cls ? new_cnt : 0, incl = v;
  for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
  __syncthreads();
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += s_warp[w];
  if (tid <= C) s_off[tid] = before + incl - v;
  if (tid == NMS_THREADS - 1) *ndet = before + incl;   // the TRUE count: the host rejects a record set that did not fit (> max_det)
  __syncthreads();
  for (int i = tid; i < (C - 1) * r; i += blockDim.x) {
    const int c = 1 + i / r, j = i % r;
    if (j < keep_cnt[c]) {
      const int slot = s_off[c] + j;
      if (slot < max_det) {
        const int roi = keep[(size_t)c * r + j];
        const float4 b = __ldg(pred + (size_t)roi * C + c);
        float* d = det + (size_t)slot * 6;
        d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w;
        d[4] = keep_score[(size_t)c * r + j]; d[5] = (float)c;
      }
    }
  }
}

}  // namespace frcnn

using namespace frcnn;

extern "C" int frcnn_proposals(const float* props, const float* scores, const int* order, int n, int batch, int pre_nms_top_n,
                               int post_nms_top_n, float thresh, unsigned flags, float* rois, float* roi_scores, int* keep,
                               int* num, void* stream) {
  FRCNN_REQUIRE(props && scores && order && rois && roi_scores && keep && num && n > 0 && batch > 0 && post_nms_top_n > 0, "proposals: bad argument");
  const int m = (pre_nms_top_n > 0 && pre_nms_top_n < n) ? pre_nms_top_n : n;
  cudaStream_t st = (cudaStream_t)stream;
  if (thresh < 0.f) {
    gather_top_kernel<<<dim3((unsigned)cdiv(post_nms_top_n, 256), (unsigned)batch), 256, 0, st>>>(
        reinterpret_cast<const float4*>(props), scores, order, n, m, post_nms_top_n, rois, roi_scores, keep, num);
  } else {
    if (post_nms_top_n > PROPOSAL_CAP) { set_error("proposals: post_nms_top_n %d > capacity %d", post_nms_top_n, PROPOSAL_CAP); return ERR_CAPACITY; }
    proposals_kernel<<<(unsigned)batch, NMS_THREADS, 0, st>>>(reinterpret_cast<const float4*>(props), scores, order, n, m, post_nms_top_n,
                                                             thresh, flags, rois, roi_scores, keep, num);
  }
  FRCNN_LAUNCH_CHECK();
  return OK;
}

// Kept-set workspace of the `_nms`-compatible path: one slot PER DEVICE (a process-wide buffer was reused on
// whatever device a later call named), grown on demand; single stream use per device.
constexpr int MAX_DEVICES = 64;
static float4* g_kept[MAX_DEVICES]; static float* g_kept_area[MAX_DEVICES]; static int g_kept_cap[MAX_DEVICES];
static int ensure_kept(int dev, int n) {
  FRCNN_REQUIRE(dev >= 0 && dev < MAX_DEVICES, "device index %d out of range", dev);
  if (n <= g_kept_cap[dev]) return OK;
  if (g_kept[dev]) { cudaFree(g_kept[dev]); cudaFree(g_kept_area[dev]); g_kept[dev] = nullptr; g_kept_area[dev] = nullptr; g_kept_cap[dev] = 0; }
  FRCNN_CUDA(cudaMalloc(&g_kept[dev], (size_t)n * sizeof(float4)));
  FRCNN_CUDA(cudaMalloc(&g_kept_area[dev], (size_t)n * sizeof(float)));
  g_kept_cap[dev] = n;
  return OK;
}

extern "C" int frcnn_nms_sorted_dev(const float* boxes, int n, float thresh, unsigned flags, int max_out, int* keep, int* num, void* stream) {
  FRCNN_REQUIRE(boxes && keep && num && n > 0 && max_out > 0, "nms_sorted_dev: bad argument");
  int dev = 0;
  FRCNN_CUDA(cudaGetDevice(&dev));
  int rc = ensure_kept(dev, max_out < n ? max_out : n);
  if (rc) return rc;
  nms_sorted_kernel<<<1, NMS_THREADS, 0, (cudaStream_t)stream>>>(boxes, 4, n, thresh, flags, max_out, g_kept[dev], g_kept_area[dev], keep, num);
  FRCNN_LAUNCH_CHECK();
  return OK;
}

// device_id < 0: the calling thread's current device.  The caller's current device is restored before returning.
extern "C" int frcnn_nms_host(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim, float thresh,
                              int device_id, unsigned flags) {
  FRCNN_REQUIRE(keep_out && num_out, "nms_host: null output");
  *num_out = 0;
  if (boxes_num <= 0) return OK;
  FRCNN_REQUIRE(boxes_host && boxes_dim >= 4, "nms_host: bad input");
  int cur = -1;
  FRCNN_CUDA(cudaGetDevice(&cur));
  const int dev = device_id < 0 ? cur : device_id;
  if (cur != dev) FRCNN_CUDA(cudaSetDevice(dev));
  float* dboxes = nullptr; int* dkeep = nullptr; int* dnum = nullptr;
  int rc = ensure_kept(dev, boxes_num);
  cudaError_t e = cudaSuccess;
  if (!rc) {
    e = cudaMalloc(&dboxes, (size_t)boxes_num * boxes_dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&dkeep, (size_t)(boxes_num + 1) * sizeof(int));
    if (e == cudaSuccess) {
      dnum = dkeep + boxes_num;
      e = cudaMemcpy(dboxes, boxes_host, (size_t)boxes_num * boxes_dim * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
      nms_sorted_kernel<<<1, NMS_THREADS>>>(dboxes, boxes_dim, boxes_num, thresh, flags, boxes_num, g_kept[dev], g_kept_area[dev], dkeep, dnum);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(num_out, dnum, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && *num_out > 0) e = cudaMemcpy(keep_out, dkeep, (size_t)(*num_out) * sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(dboxes); cudaFree(dkeep);
  }
  if (cur != dev) cudaSetDevice(cur);                    // leave the caller's (torch's) current device untouched
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "frcnn_nms_host", __FILE__, __LINE__);
  return OK;
}

extern "C" size_t frcnn_detect_post_workspace_bytes(int r, int num_classes, int batch) {
  if (r <= DET_CAP) return 256;
  return (size_t)batch * num_classes * (size_t)((r + 3) & ~3) * 24 + 256;
}

extern "C" int frcnn_detect_post(const float* cls_prob, const float* pred_boxes, const int* num_rois, int r, int batch, int num_classes,
                                 float score_thresh, float nms_thresh, unsigned flags, int max_per_image, int max_det,
                                 float* det, int* ndet, int record_stride, int* keep, int* keep_cnt, float* keep_score, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  FRCNN_REQUIRE(cls_prob && pred_boxes && num_rois && det && ndet && keep && keep_cnt && keep_score, "detect_post: null pointer");
  FRCNN_REQUIRE(r > 0 && batch > 0 && num_classes >= 2 && num_classes <= 1024, "detect_post: r>0, batch>0, 2<=C<=1024 required");
  if (r > DET_CAP_BIG) { set_error("detect_post: %d RoIs per image > capacity %d", r, DET_CAP_BIG); return ERR_CAPACITY; }
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((unsigned)(num_classes - 1), (unsigned)batch);
  if (r <= DET_CAP) {
    class_nms_kernel<DET_CAP, false><<<grid, NMS_THREADS, DET_CAP * 32, st>>>(cls_prob, reinterpret_cast<const float4*>(pred_boxes), num_rois, r,
                                                                              num_classes, score_thresh, nms_thresh, flags, keep, keep_cnt,
                                                                              keep_score, nullptr);
  } else {
    FRCNN_REQUIRE(workspace && workspace_bytes >= frcnn_detect_post_workspace_bytes(r, num_classes, batch), "detect_post: workspace too small");
    static bool attr_done = false;
    if (!attr_done) {
      FRCNN_CUDA(cudaFuncSetAttribute(class_nms_kernel<DET_CAP_BIG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DET_CAP_BIG * 8));
      attr_done = true;
    }
    class_nms_kernel<DET_CAP_BIG, true><<<grid, NMS_THREADS, DET_CAP_BIG * 8, st>>>(
        cls_prob, reinterpret_cast<const float4*>(pred_boxes), num_rois, r, num_classes, score_thresh, nms_thresh, flags, keep, keep_cnt,
        keep_score, reinterpret_cast<uint8_t*>(workspace));
  }
  FRCNN_LAUNCH_CHECK();
  FRCNN_REQUIRE(record_stride == 0 || record_stride >= max_det * 6, "detect_post: record_stride %d < max_det*6", record_stride);
  cap_emit_kernel<<<(unsigned)batch, NMS_THREADS, 0, st>>>(reinterpret_cast<const float4*>(pred_boxes), r, num_classes, max_per_image, max_det,
                                                          keep, keep_cnt, keep_score, det, ndet,
                                                          record_stride ? record_stride : max_det * 6, record_stride ? record_stride : 1);
  FRCNN_LAUNCH_CHECK();
  return OK;
}
