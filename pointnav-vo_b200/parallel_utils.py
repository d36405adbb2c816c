"""Host-side data-parallel helpers (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in
the CPU tests).  The VO path has exactly two exchanges per training step:
  1. the flat fp32 gradient bucket -- ONE all-reduce (SUM; the 1/world factor is applied by the loss-gradient
     kernel, so the result is the gradient of the mean loss over the concatenated batch);
  2. the packed RunningMeanAndVar statistics (sum, sum of squares per channel, fp64) -- ONE all-reduce instead of
     the reference's three (model_utils/running_mean_and_var.py:28-38)."""
import torch
import torch.distributed as dist


def allreduce_flat_bucket(flat, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
    return flat


def merge_input_stats(packed, n_local, pix, group=None):
    """packed: [2C] (sum, sumsq interleaved) of this rank's batch.  Returns the (mean, population variance) of the
    batch concatenated over all ranks -- what RunningMeanAndVar computes with its all-reduces."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    allreduce_flat_bucket(packed, group)
    n = float(n_local * world * pix)
    s = packed.view(-1, 2)
    mean = s[:, 0] / n
    var = (s[:, 1] / n - mean * mean).clamp_min(0)
    return mean, var
