"""Multi-GPU: one process per GPU (torchrun), scene + BVH replicated, the sample range
partitioned per rank, per-rank float4 films summed with ONE reduce (NCCL over NVLink on
GPUs, gloo in the CPU tests).  SURVEY §8e.

Because the RNG is keyed on (pixel, sample, dimension), the union of the ranks' samples is
exactly the sample set of a 1-GPU render; only the floating-point summation order of the
film differs (per-rank partial sums are added by the reduce)."""
import os

import torch

from . import _abi
import torch.distributed as dist


def partition_spp(spp_begin, spp_end, rank, world):
    """Sample indices [b, e) owned by `rank`: contiguous, disjoint, covering, sizes differ by <= 1."""
    n = spp_end - spp_begin
    b = spp_begin + (n * rank) // world
    e = spp_begin + (n * (rank + 1)) // world
    return b, e


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env; returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def reduce_film(film, dst=0):
    """Sum the per-rank films onto `dst` (the one exchange step of the path)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


class DistRenderer:
    """Per-rank renderer: renders this rank's share of the samples into a device film and
    reduces it onto rank 0."""

    def __init__(self, scene, local_rank=0):
        from .render import Renderer
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        torch.cuda.set_device(local_rank)
        self.r = Renderer(local_rank)
        # kernels run on torch's current stream so torch events and NCCL order against them
        self.r.set_stream(torch.cuda.current_stream().cuda_stream)
        self.ds = self.r.upload(scene)
        self.film = None
        self.nrc_key = None   # (width, height, max_depth, seed, batch, iters, lr) of the trained cache
        self.nrc_info = None

    def render_async(self, integ, width, height, spp_begin=0, spp_end=None, pool_paths=0, nrc_exact=False):
        """`nrc_exact`: look the cache up with the bit-exact fp32 kernel instead of the tcgen05 one (the
        default for `type: "nrc"`: 2.5x faster, within bf16 tolerance of the exact film)."""
        spp_end = integ.spp if spp_end is None else spp_end
        b, e = partition_spp(spp_begin, spp_end, self.rank, self.world)
        if self.film is None or self.film.shape[:2] != (height, width):
            self.film = torch.empty(height, width, 4, device="cuda", dtype=torch.float32)
        cfg = integ.cfg(width=width, height=height, spp_begin=b, spp_end=e, pool_paths=pool_paths)
        if getattr(integ, "type", "pt") == "nrc":
            # every rank trains the same cache (records and descent are deterministic, so the
            # weights are identical everywhere: no exchange), then renders its share of the samples
            nrc = integ.nrc_cfg()
            key = (width, height, integ.max_depth, integ.seed, nrc.batch_size, nrc.training_iters, nrc.learning_rate)
            if self.nrc_key != key:
                self.nrc_info = self.ds.nrc_train(cfg, nrc)
                self.nrc_key = key
            if not nrc_exact:
                cfg.flags |= _abi.AQ_RENDER_NRC_TENSOR
            self.ds.nrc_render_device_async(cfg, nrc, self.film.data_ptr())
        else:
            self.ds.render_device_async(cfg, self.film.data_ptr())
        reduce_film(self.film, 0)
        return self.film

    def finish(self):
        return self.ds.finish()
