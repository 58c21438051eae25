"""N>1 host logic on CPU: world_size-2 gloo run of the sample-range decomposition + film reduce.  The renderer
inside each rank is the CPU oracle (test stand-in for the GPU path, which cannot run without a device)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc
from shimmer_b200 import scenes
from shimmer_b200.distributed import reduce_film, sample_range_for_rank


def test_sample_ranges_partition_exactly():
    for spp in (1, 2, 5, 16, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            rs = [sample_range_for_rank(spp, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == spp
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
            ws = [sample_range_for_rank(spp, r, world, "weak") for r in range(world)]
            assert ws == [(r * spp, (r + 1) * spp) for r in range(world)]
    with pytest.raises(ValueError):
        sample_range_for_rank(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, spp, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.cornell_box(resolution=(32, 32)).build()
    rng = sample_range_for_rank(spp, rank, world)
    film, st, _ = orc.render(sc, orc.make_params(seed=4, spp=spp, sample_range=rng), n_threads=2)
    t = torch.from_numpy(film)
    reduce_film(t, dst=0)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_render_matches_single_rank(tmp_path):
    spp, world = 6, 2
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(world, _free_port(), spp, out), nprocs=world, join=True)
    got = np.load(out)
    sc = scenes.cornell_box(resolution=(32, 32)).build()
    ref, _, _ = orc.render(sc, orc.make_params(seed=4, spp=spp))
    assert np.all(got[:, 3] == spp)
    assert np.allclose(got, ref, rtol=1e-12, atol=0)
