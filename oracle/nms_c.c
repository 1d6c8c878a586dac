/* Oracle (TEST INFRASTRUCTURE, see oracle/__init__.py): plain-C restatement of the
 * greedy NMS variants on the reference's inference path.  Scalar, single thread.
 *
 *   oracle_nms_plus1  : lib/nms/cpu_nms.pyx:17-68 ('+1' pixel areas; suppress when
 *                       ovr >= thresh) and, with inclusive=0, the predicate of
 *                       lib/nms/nms_kernel.cu:24-32,61-77,118-140 / py_cpu_nms.py:10-38
 *                       (suppress when ovr > thresh).
 *   oracle_nms_tf     : tf.image.non_max_suppression, TF 1.x CPU kernel semantics
 *                       (continuous areas on min/max-normalised corners, IoU of a
 *                       non-positive-area box is 0, suppress when IoU > thr, stop at
 *                       max_output_size) -- called at lib/layer_utils/proposal_layer.py:72.
 *
 * Sort rule (the reference's argsort()[::-1] / std::sort leave ties unspecified):
 * descending score, LOWER INDEX FIRST among equal scores.  Every fp32 operation is a
 * separate IEEE round-to-nearest op; build with -ffp-contract=off.
 */
#include <stdlib.h>
#include <string.h>

typedef struct { float s; int i; } key_t_;

static int cmp_desc(const void* a, const void* b) {
  const key_t_* x = (const key_t_*)a; const key_t_* y = (const key_t_*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);
}

static int* sorted_order(const float* scores, int stride, int n) {
  key_t_* k = (key_t_*)malloc(sizeof(key_t_) * (size_t)(n > 0 ? n : 1));
  int* ord = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) { k[i].s = scores[(size_t)i * stride]; k[i].i = i; }
  qsort(k, (size_t)n, sizeof(key_t_), cmp_desc);
  for (int i = 0; i < n; ++i) ord[i] = k[i].i;
  free(k);
  return ord;
}

static inline float fmaxf_(float a, float b) { return a >= b ? a : b; }
static inline float fminf_(float a, float b) { return a <= b ? a : b; }

/* dets: [n,5] (x1,y1,x2,y2,score) row-major.  keep: capacity n.  returns #kept.
 * Indices in keep refer to the UNSORTED input, in descending-score order. */
int oracle_nms_plus1(const float* dets, int n, float thresh, int inclusive, int* keep) {
  if (n <= 0) return 0;
  int* ord = sorted_order(dets + 4, 5, n);
  float* area = (float*)malloc(sizeof(float) * (size_t)n);
  unsigned char* dead = (unsigned char*)calloc((size_t)n, 1);
  for (int i = 0; i < n; ++i) {
    const float* d = dets + (size_t)i * 5;
    area[i] = (d[2] - d[0] + 1.0f) * (d[3] - d[1] + 1.0f);
  }
  int nk = 0;
  for (int a = 0; a < n; ++a) {
    int i = ord[a];
    if (dead[i]) continue;
    keep[nk++] = i;
    const float* bi = dets + (size_t)i * 5;
    for (int b = a + 1; b < n; ++b) {
      int j = ord[b];
      if (dead[j]) continue;
      const float* bj = dets + (size_t)j * 5;
      float xx1 = fmaxf_(bi[0], bj[0]), yy1 = fmaxf_(bi[1], bj[1]);
      float xx2 = fminf_(bi[2], bj[2]), yy2 = fminf_(bi[3], bj[3]);
      float w = fmaxf_(0.0f, xx2 - xx1 + 1.0f), h = fmaxf_(0.0f, yy2 - yy1 + 1.0f);
      float inter = w * h;
      float ovr = inter / (area[i] + area[j] - inter);
      if (inclusive ? (ovr >= thresh) : (ovr > thresh)) dead[j] = 1;
    }
  }
  free(ord); free(area); free(dead);
  return nk;
}

static inline float tf_iou(const float* a, const float* b) {
  float ay0 = fminf_(a[0], a[2]), ax0 = fminf_(a[1], a[3]);
  float ay1 = fmaxf_(a[0], a[2]), ax1 = fmaxf_(a[1], a[3]);
  float by0 = fminf_(b[0], b[2]), bx0 = fminf_(b[1], b[3]);
  float by1 = fmaxf_(b[0], b[2]), bx1 = fmaxf_(b[1], b[3]);
  float area_a = (ay1 - ay0) * (ax1 - ax0);
  float area_b = (by1 - by0) * (bx1 - bx0);
  if (area_a <= 0.0f || area_b <= 0.0f) return 0.0f;
  float iy0 = fmaxf_(ay0, by0), ix0 = fmaxf_(ax0, bx0);
  float iy1 = fminf_(ay1, by1), ix1 = fminf_(ax1, bx1);
  float inter = fmaxf_(iy1 - iy0, 0.0f) * fmaxf_(ix1 - ix0, 0.0f);
  return inter / (area_a + area_b - inter);
}

/* boxes: [n,4]; scores: [n].  keep: capacity max_out.  returns #selected (<= max_out).
 * Candidate vs already-selected set, newest selected first (TF's inner loop order). */
int oracle_nms_tf(const float* boxes, const float* scores, int n, int max_out, float thr, int* keep) {
  if (n <= 0 || max_out <= 0) return 0;
  int* ord = sorted_order(scores, 1, n);
  int nk = 0;
  for (int a = 0; a < n && nk < max_out; ++a) {
    int i = ord[a];
    int ok = 1;
    for (int s = nk - 1; s >= 0; --s) {
      if (tf_iou(boxes + (size_t)i * 4, boxes + (size_t)keep[s] * 4) > thr) { ok = 0; break; }
    }
    if (ok) keep[nk++] = i;
  }
  free(ord);
  return nk;
}

/* stable descending order of scores (lower index first on ties); out: [n] */
void oracle_argsort_desc(const float* scores, int n, int* out) {
  if (n <= 0) return;
  int* ord = sorted_order(scores, 1, n);
  memcpy(out, ord, sizeof(int) * (size_t)n);
  free(ord);
}
