"""N > 1 path on CPU: world_size-2 gloo processes shard the samples / iterations exactly as the GPU
ranks do (wrt_b200.shard_pt / shard_bdpt + reduce_film); each rank renders its shard with the hostsim
build of the kernels' per-path code; the reduced film must equal the single-rank film."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "hostsim")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wrt_b200 as W
    import scenes, util
    from hostsim_py import HostSim
    sc = scenes.small_mixed_scene(24, 24)
    hs = util.host_scene(W, sc); sim = HostSim(hs.desc(), hs); cam = hs.camera()
    if kind == "pt":
        full = W.PtParams(24, 24, 16, 5, 3, 0, 1, 0.0)
        film, _ = sim.render_pt(cam, W.shard_pt(full, rank, world))
    elif kind == "whitted":
        full = W.PtParams(24, 24, 16, 7, 3, 0, 1, 0.0)
        film, _ = sim.render_whitted(cam, W.shard_pt(full, rank, world))
    else:
        full = W.BdptParams(24, 24, 8, 0, 10, 3, 3, 0, 1, 0.0, 0)
        film, _ = sim.render_bdpt(cam, W.shard_bdpt(full, rank, world))
    t = torch.from_numpy(film)
    W.reduce_film(t, 0)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), t.numpy())
        single = (sim.render_pt if kind == "pt" else sim.render_whitted if kind == "whitted" else sim.render_bdpt)(cam, full)[0]
        np.save(os.path.join(out_dir, "single.npy"), single)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["pt", "bdpt", "whitted"])
def test_two_rank_sharding_sums_to_single_rank_image(tmp_path, kind):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, kind, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "reduced.npy"); b = np.load(tmp_path / "single.npy")
    assert np.nanmean(a) > 0
    # same paths, different float summation order: a few ulp per pixel (Whitted films carry the reference's NaN pixels)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.allclose(np.nan_to_num(a), np.nan_to_num(b), rtol=1e-5, atol=1e-7)
