"""Image-sharded data parallelism: one process per GPU, image i -> rank i mod W, one all-gather of the fixed-size
detection records per step (the reference itself is single-GPU, lib/model/test.py:152).

The record a rank contributes is exactly the buffer the last device kernel wrote (`plan.det` [max_det,6] fp32 +
`plan.ndet` int32): no staging copy sits between the kernel and the collective.  torch.distributed is used for
bootstrap and the collective only (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(num_images, rank, world):
    """Indices of the images rank `rank` processes, in processing order."""
    return list(range(rank, num_images, world))


def steps_for(num_images, world):
    """Number of lock-step iterations (ranks without an image in the last step contribute an empty record)."""
    return -(-num_images // world)


class RecordGather(object):
    """Pre-allocated receive buffers for the per-step all-gather."""

    def __init__(self, det, ndet, world):
        self.world = world
        self.det_out = [torch.empty_like(det) for _ in range(world)]
        self.n_out = [torch.empty_like(ndet) for _ in range(world)]

    def gather(self, det, ndet):
        if self.world == 1:
            self.det_out[0].copy_(det); self.n_out[0].copy_(ndet)
        else:
            dist.all_gather(self.det_out, det)
            dist.all_gather(self.n_out, ndet)
        return self.det_out, self.n_out


def records_to_all_boxes(all_boxes, step, world, det_list, n_list, num_images):
    """Scatter one step's gathered records into all_boxes[cls][image] (lib/model/test.py:145-146 layout)."""
    num_classes = len(all_boxes)
    for r in range(world):
        img = step * world + r
        if img >= num_images:
            continue
        n = int(n_list[r].item())
        d = det_list[r][:n].cpu().numpy()
        cls = d[:, 5].astype(np.int64)
        for j in range(1, num_classes):
            all_boxes[j][img] = d[cls == j, :5]
    return all_boxes
