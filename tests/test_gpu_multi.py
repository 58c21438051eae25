"""Multi-GPU behind the C ABI (include/shimmer_gpu.h, ABI v9): the sample-range split and the ONE NCCL film reduce per render
(integrator.rs:235-245 is the reference's tile fan-out) in both forms --
  * single process, n GPUs  : sg_init_multi + sg_render
  * one process per GPU     : sg_init + sg_comm_init_rank + SG_RENDER_SPLIT_SAMPLES | SG_RENDER_REDUCE_FILM
and the stream contract of sg_render_device (the caller's stream orders the film).
Every multi-device case runs in fresh processes (device sets and communicators are process-global) and is skipped on a
one-GPU box; the one-rank communicator and the stream tests run everywhere."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from shimmer_b200 import Options, create_integrator, ffi, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _run(code, *args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def _reference_film(kind, res, spp, seed=3):
    sc = scenes.tiny_scene(kind, resolution=(res, res)).build() if kind != "cornell" else scenes.cornell_box(resolution=(res, res)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": spp})
    film = integ.render(Options(seed=seed)).copy()
    st = integ.stats.as_dict()
    integ.close()
    return film, st


SINGLE_PROCESS = """
import sys, numpy as np
from shimmer_b200 import Options, create_integrator, scenes, ffi
n, res, spp, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
sc = scenes.cornell_box(resolution=(res, res)).build()
integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": spp}, device=list(range(n)))
assert ffi.load_library().sg_device_count() == n
film = integ.render(Options(seed=3)).copy()
st = integ.stats
assert st.n_devices == n and st.camera_paths == res * res * spp, (st.n_devices, st.camera_paths)
film2 = integ.render(Options(seed=3), flags=ffi.SG_RENDER_OVERWRITE_FILM).copy()      # second render: workspaces reused, film overwritten
np.save(out, np.stack([film, film2]))
print("reduce_ms", st.reduce_ms, "d2h_ms", st.d2h_ms)
integ.close()
"""


@pytest.mark.parametrize("n", [2, 4])
def test_single_process_multi_gpu_render_equals_one_gpu(n, tmp_path):
    if _n_gpus() < n:
        pytest.skip(f"needs {n} GPUs")
    res, spp = 48, 10                                        # 10 samples over n devices: uneven split (remainder to the low devices)
    out = str(tmp_path / "multi.npy")
    _run(SINGLE_PROCESS, n, res, spp, out)
    multi = np.load(out)
    ref, _ = _reference_film("cornell", res, spp)
    for f in multi:
        assert np.array_equal(f[:, 3], ref[:, 3]) and np.all(ref[:, 3] == spp)           # every sample index rendered exactly once
        assert np.allclose(f, ref, rtol=1e-12, atol=1e-300)                                 # same samples, f64 sums in another order


RANK_PROCESS = """
import sys, os, time, ctypes as C, numpy as np
from shimmer_b200 import Options, create_integrator, scenes, ffi
from shimmer_b200 import distributed as sgd
rank, world, res, spp, tmp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
idf = os.path.join(tmp, "nccl_id")
def bcast(buf, src=0):                      # the launcher's only job: carry rank 0's id to the others (here: a file)
    if rank == 0:
        open(idf + ".tmp", "wb").write(bytes(buf)); os.replace(idf + ".tmp", idf)
    else:
        for _ in range(600):
            if os.path.exists(idf): break
            time.sleep(0.1)
        buf[:] = open(idf, "rb").read()
lib = sgd.init_process_comm(rank, world, bcast, device=rank)
r, n = C.c_int(), C.c_int(); lib.sg_comm_rank(C.byref(r), C.byref(n)); assert (r.value, n.value) == (rank, world)
sc = scenes.cornell_box(resolution=(res, res)).build()
integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": spp}, device=rank)
MG = ffi.SG_RENDER_SPLIT_SAMPLES | ffi.SG_RENDER_REDUCE_FILM
film = integ.render(Options(seed=3), flags=MG).copy()                 # host film: only rank 0's is written
st = integ.stats
b, e = C.c_int32(), C.c_int32(); lib.sg_sample_range_for_rank(0, spp, rank, world, C.byref(b), C.byref(e))
assert st.camera_paths == res * res * (e.value - b.value) and st.n_devices == world and st.rank == rank
import torch
torch.cuda.set_device(rank)
d = torch.zeros((res * res, 4), dtype=torch.float64, device="cuda")
integ.render_device(Options(seed=3), d.data_ptr(), flags=MG)            # device film: reduced in place on rank 0
torch.cuda.synchronize()
if rank == 0:
    np.save(os.path.join(tmp, "film.npy"), np.stack([film, d.cpu().numpy()]))
else:
    assert not film.any()                                                # non-root host films are left alone
integ.close()
sgd.destroy_process_comm()
"""


@pytest.mark.parametrize("world", [2])
def test_one_process_per_gpu_split_and_reduce_equals_one_gpu(world, tmp_path):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    res, spp = 48, 7
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    procs = [subprocess.Popen([sys.executable, "-c", textwrap.dedent(RANK_PROCESS), str(r), str(world), str(res), str(spp), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
    got = np.load(str(tmp_path / "film.npy"))
    ref, _ = _reference_film("cornell", res, spp)
    for f in got:
        assert np.array_equal(f[:, 3], ref[:, 3])
        assert np.allclose(f, ref, rtol=1e-12, atol=1e-300)


def test_one_rank_communicator_is_a_no_op(tmp_path):
    """sg_comm_* with a single rank (loads NCCL, creates the communicator): split / reduce flags change nothing."""
    code = """
    import sys, ctypes as C, numpy as np
    from shimmer_b200 import Options, create_integrator, scenes, ffi
    from shimmer_b200 import distributed as sgd
    lib = sgd.init_process_comm(0, 1, lambda buf, src=0: None, device=0)
    sc = scenes.tiny_scene("conductor", resolution=(24, 24)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 4})
    a = integ.render(Options(seed=1)).copy()
    b = integ.render(Options(seed=1), flags=ffi.SG_RENDER_OVERWRITE_FILM | ffi.SG_RENDER_SPLIT_SAMPLES | ffi.SG_RENDER_REDUCE_FILM).copy()
    assert np.array_equal(a, b) and integ.stats.n_devices == 1 and integ.stats.reduce_ms == 0.0
    assert lib.sg_comm_init_rank(C.create_string_buffer(128), 0, 1) == -1              # a second communicator is refused (SG_ERR_INVALID_ARGUMENT)
    integ.close(); sgd.destroy_process_comm()
    """
    _run(code)


def test_render_device_is_ordered_on_the_callers_stream():
    """a stream handle of 0 is the caller's legacy default stream, not a private library stream -- the
    film tensor's zero_() before and its consumer after the render are ordered without any host synchronisation."""
    import torch
    res, spp = 64, 8
    sc = scenes.cornell_box(resolution=(res, res)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": spp})
    ref = integ.render(Options(seed=5)).copy()
    film = torch.empty((res * res, 4), dtype=torch.float64, device="cuda")
    big = torch.empty(1 << 28, dtype=torch.float32, device="cuda")            # 1 GiB: the fills below keep the stream busy for a while
    for stream in (None, torch.cuda.Stream()):
        with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream()):
            handle = torch.cuda.current_stream().cuda_stream
            film.fill_(1e30)
            for _ in range(4):
                big.fill_(1.0)                                                # still running when sg_render_device enqueues its kernels
            film.zero_()
            integ.render_device(Options(seed=5), film.data_ptr(), stream=handle)
            out = film.clone()                                                # consumer on the same stream
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), ref), "render raced the caller's zero_() on stream %r" % (handle,)
    integ.close()
