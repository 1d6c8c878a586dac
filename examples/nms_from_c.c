/* The C-ABI boundary from plain C: what lib/nms/gpu_nms.pyx's `_nms` call becomes.
 *   gcc -std=c99 -Wall -pedantic -Iinclude examples/nms_from_c.c -Ltf_faster_rcnn_b200 -lfrcnn_b200 -o nms_from_c
 * Without a GPU the non-empty call fails with a message (there is no CPU fallback); the empty call and the version call work. */
#include <stdio.h>
#include "frcnn_b200.h"

int main(void) {
  /* boxes sorted by descending score, rows (x1, y1, x2, y2, score) -- the layout gpu_nms.pyx hands to _nms */
  static const float dets[4 * 5] = {0, 0, 9, 9, 0.9f, 1, 1, 10, 10, 0.8f, 50, 50, 60, 60, 0.7f, 0, 0, 9, 8, 0.6f};
  int keep[4], num = -1, rc;
  char msg[256];
  printf("frcnn_b200 version %d\n", frcnn_version());
  rc = frcnn_nms_host(keep, &num, dets, 0, 5, 0.3f, -1, FRCNN_NMS_MODE_GPU_NMS);      /* empty input: no device needed */
  printf("empty input: status %d, kept %d\n", rc, num);
  if (rc != FRCNN_OK || num != 0) return 1;
  rc = frcnn_nms_host(keep, &num, dets, 4, 5, 0.3f, -1, FRCNN_NMS_MODE_GPU_NMS);
  if (rc != FRCNN_OK) {
    frcnn_last_error(msg, sizeof msg);
    printf("4 boxes: status %d (%s)\n", rc, msg);
    return 0;                                                                        /* expected on a machine without a B200 */
  }
  printf("4 boxes: kept %d:", num);
  for (rc = 0; rc < num; ++rc) printf(" %d", keep[rc]);
  printf("\n");
  return (num == 2 && keep[0] == 0 && keep[1] == 2) ? 0 : 2;
}
