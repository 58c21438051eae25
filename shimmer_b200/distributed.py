"""Multi-GPU decomposition of the render (DESIGN.md section e).

The path shards naturally: camera samples are independent (integrator.rs:235-245 already exploits this with
tiles) and with the box filter every film pixel only receives its own samples (film.rs:569-573).  Each rank
holds a full scene replica, renders a contiguous range of SAMPLE INDICES of every pixel (distinct
(pixel, sample) RNG streams -> no correlation between ranks) and the f64 films are summed onto rank 0 with
ONE collective per render.  No other data-path communication exists.

On GPUs the split and the reduce live BEHIND the C ABI (include/shimmer_gpu.h, SG_RENDER_SPLIT_SAMPLES |
SG_RENDER_REDUCE_FILM): `init_process_comm` below only carries rank 0's NCCL unique id to the other processes
(sg_comm_get_unique_id -> sg_comm_init_rank), which is all a launcher has to do -- a Rust host would do the same
over MPI or a file.  `reduce_film` (torch.distributed) remains for the CPU (gloo) tests of the host-side logic.
"""
import ctypes as C
from typing import Tuple


def sample_range_for_rank(spp: int, rank: int, world: int, mode: str = "strong") -> Tuple[int, int]:
    """strong: the `spp` samples of the image are split across ranks (remainder to the low ranks);
    weak: every rank renders `spp` NEW sample indices (image ends up with world*spp samples)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if mode == "weak":
        return rank * spp, (rank + 1) * spp
    if mode != "strong":
        raise ValueError(mode)
    base, rem = divmod(spp, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_film(film_tensor, dst: int = 0):
    """Sum the (n_pixels, 4) float64 film over all ranks onto `dst` (one collective per render)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return film_tensor


def _preload_process_nccl():
    """The library dlopens "libnccl.so.2" on first use.  A process that ALSO uses torch must let torch load its bundled (newer)
    copy first: the dynamic loader keys on the SONAME, so a system libnccl loaded earlier would be handed to libtorch_cuda.so and
    fail its symbol lookup (observed: `undefined symbol: ncclDevCommCreate`)."""
    import importlib.util
    import sys
    if "torch" not in sys.modules and importlib.util.find_spec("torch") is not None:
        import torch  # noqa: F401


def init_process_comm(rank: int, world: int, broadcast_bytes, device: int = None):
    """One process per GPU: create the library's NCCL communicator.  `broadcast_bytes(buf: bytearray, src=0)` must copy rank
    0's buffer to every rank in place (e.g. through torch.distributed, MPI or a shared file).  Collective."""
    from . import ffi
    from .integrator import _ensure_init
    _preload_process_nccl()
    lib = _ensure_init(rank if device is None else device)
    buf = bytearray(ffi.SG_COMM_ID_BYTES)
    if rank == 0:
        raw = (C.c_char * ffi.SG_COMM_ID_BYTES)()
        ffi.check(lib.sg_comm_get_unique_id(raw), "sg_comm_get_unique_id")
        buf[:] = bytes(raw)
    broadcast_bytes(buf, 0)
    raw = (C.c_char * ffi.SG_COMM_ID_BYTES).from_buffer_copy(bytes(buf))
    ffi.check(lib.sg_comm_init_rank(raw, int(rank), int(world)), "sg_comm_init_rank")
    return lib


def torch_broadcast_bytes(device=None):
    """broadcast_bytes for init_process_comm over an initialised torch.distributed group (nccl: pass the CUDA device)."""
    import torch
    import torch.distributed as dist

    def bcast(buf, src=0):
        t = torch.tensor(list(buf), dtype=torch.uint8, device=device if device is not None else "cpu")
        dist.broadcast(t, src=src)
        buf[:] = bytes(t.cpu().tolist())
    return bcast


def destroy_process_comm():
    from . import ffi
    ffi.check(ffi.load_library().sg_comm_destroy(), "sg_comm_destroy")
