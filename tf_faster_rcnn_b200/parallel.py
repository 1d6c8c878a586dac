"""Image-sharded data parallelism: one process per GPU, image i -> rank i mod W, ONE all-gather of the fixed-size
detection records per step (the reference itself is single-GPU, lib/model/test.py:152).

The record a rank contributes is exactly the buffer the last device kernel wrote (`ShapePlan.rec`: per image an int32 count
in word 0, then max_det rows of (x1,y1,x2,y2,score,class)): no staging copy sits between the kernel and the collective.
The collective is issued asynchronously (NCCL runs it on its own stream behind an event on the compute stream), the plan
alternates between two record buffers, and a rank only waits for the gather of step i when it is about to overwrite that
step's buffer (step i+2) or to read its result -- so the next image's graph replay overlaps the gather.
torch.distributed is used for bootstrap and the collective only (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

REC_HEADER = 8   # = engine.REC_HEADER (4-byte words before the rows; word 0 = int32 count)


def shard_indices(num_images, rank, world):
    """Indices of the images rank `rank` processes, in processing order."""
    return list(range(rank, num_images, world))


def steps_for(num_images, world):
    """Number of lock-step iterations (ranks without an image in the last step contribute an empty record)."""
    return -(-num_images // world)


class RecordGather(object):
    """Two receive buffers [world, *record shape] and the outstanding collective of each."""

    def __init__(self, rec_like, world):
        self.world = world
        self.out = [torch.empty((world,) + tuple(rec_like.shape), dtype=rec_like.dtype, device=rec_like.device) for _ in range(2)]
        self.work = [None, None]
        self.collectives = 0

    def before_overwrite(self, slot):
        """Call before the producer of `slot`'s record buffer runs again: the gather that read it must have finished."""
        if self.work[slot] is not None:
            self.work[slot].wait()
            self.work[slot] = None

    def issue(self, slot, rec):
        """Start the all-gather of this step's record (stream-ordered after the kernels that wrote `rec`)."""
        self.before_overwrite(slot)
        if self.world == 1:
            self.out[slot][0].copy_(rec, non_blocking=True)
        else:
            # flat views: the concatenation form of all_gather_into_tensor is the one every backend (NCCL, gloo) accepts
            self.work[slot] = dist.all_gather_into_tensor(self.out[slot].view(-1), rec.contiguous().view(-1), async_op=True)
        self.collectives += 1

    def result(self, slot):
        """The gathered records [world, ...] of `slot` (the current stream waits for the collective)."""
        self.before_overwrite(slot)
        return self.out[slot]

    def gather(self, rec, slot=0):
        """Synchronous convenience: issue + result."""
        self.issue(slot, rec)
        return self.result(slot)


def split_records(gathered, max_det):
    """gathered: host float32 array [..., REC_HEADER + max_det*6] -> (counts int64 [...], rows [..., max_det, 6])."""
    g = np.ascontiguousarray(gathered, dtype=np.float32)
    counts = g.view(np.int32)[..., 0].astype(np.int64)
    rows = g[..., REC_HEADER:REC_HEADER + max_det * 6].reshape(g.shape[:-1] + (max_det, 6))
    return counts, rows


def records_to_all_boxes(all_boxes, step, world, gathered, num_images):
    """Scatter one step's gathered records [world, REC_HEADER + max_det*6] into all_boxes[cls][image]
    (lib/model/test.py:145-146 layout)."""
    num_classes = len(all_boxes)
    g = gathered.cpu().numpy() if isinstance(gathered, torch.Tensor) else np.asarray(gathered)
    g = g.reshape(world, -1)
    max_det = (g.shape[1] - REC_HEADER) // 6
    counts, rows = split_records(g, max_det)
    for r in range(world):
        img = step * world + r
        if img >= num_images:
            continue
        n = int(counts[r])
        if n > max_det:
            raise RuntimeError("image %d: %d detections do not fit the %d-row record" % (img, n, max_det))
        d = rows[r, :n]
        cls = d[:, 5].astype(np.int64)
        for j in range(1, num_classes):
            all_boxes[j][img] = d[cls == j, :5].copy()
    return all_boxes
