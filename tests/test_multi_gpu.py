"""World-size-2 tests of the tile-sharding host logic on CPU:
gloo process group, host-side tile exchange.  The oracle stands in for the
device only as the checker of the sharding logic (the product path is the NCCL
device exchange, covered by the -m gpu test below and by bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from webrender_b200 import multi_gpu  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _small_scene():
    return multi_gpu.tiled_alpha_scene(width=512, height=256, tile_w=128, tile_h=64, n_rects=60, seed=3)


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle.backends import OracleDevice
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dev = OracleDevice()
    sr = multi_gpu.ShardedRenderer(dev, _small_scene(), rank, world, device_exchange=False)
    sr.render()
    sr.render()   # a second frame reuses textures and staging
    if rank == 0:
        np.save(out_path, sr.read_framebuffer())
    dist.barrier()
    dist.destroy_process_group()


def test_assign_tiles_round_robin():
    a = multi_gpu.assign_tiles(10, 4)
    assert a == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert sorted(sum(a, [])) == list(range(10))


def test_tiled_scene_covers_frame():
    s = _small_scene()
    assert len(s.tiles) == 16 and s.rects[0] == (0, 0, 128, 64) and s.rects[-1] == (384, 192, 512, 256)
    assert s.pixel_layers > 0


def test_sharded_equals_single_process(tmp_path):
    """2 ranks over gloo: tiles rendered on different ranks, gathered to rank 0
    and composited, give the same framebuffer bytes as one process doing it all."""
    import torch.multiprocessing as mp
    from oracle.backends import OracleDevice
    single = multi_gpu.ShardedRenderer(OracleDevice(), _small_scene(), 0, 1)
    single.render()
    want = single.read_framebuffer()
    out = str(tmp_path / "fb.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    assert got.shape == want.shape and (got == want).all()
    # and the frame is not trivially empty
    assert len(np.unique(want)) > 8


@pytest.mark.gpu
def test_single_gpu_tiled_matches_oracle():
    """The tile → composite pipeline on the CUDA device (world 1) against the oracle."""
    from oracle.backends import OracleDevice
    from webrender_b200.device import CudaDevice
    scene = multi_gpu.tiled_alpha_scene(width=2048, height=1024, tile_w=1024, tile_h=512, n_rects=200, seed=3)
    a = multi_gpu.ShardedRenderer(CudaDevice(0), scene, 0, 1)
    a.render()
    b = multi_gpu.ShardedRenderer(OracleDevice(), scene, 0, 1)
    b.render()
    assert (a.read_framebuffer() == b.read_framebuffer()).all()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 3])
def test_direct_sharded_contexts_share_one_framebuffer(world):
    """The product exchange: `world` contexts (here on one GPU, in one process — the same
    calls a rank per GPU makes) draw their tiles and composite them straight into rank 0's exported
    framebuffer, ordered by stream flags only.  Two frames; bytes equal the oracle's single-device frame."""
    from oracle.backends import OracleDevice
    from webrender_b200.device import CudaDevice
    scene = multi_gpu.tiled_alpha_scene(width=2048, height=1024, tile_w=512, tile_h=256, n_rects=160, seed=5)
    devs = [CudaDevice(0) for _ in range(world)]
    rs = [multi_gpu.DirectShardedRenderer(devs[r], scene, r, world) for r in range(world)]
    blobs = [r.blob for r in rs]
    for r in rs:
        r.connect(blobs)
    for _ in range(2):
        for r in rs:          # rank 0 first: it clears and signals; the others only advance their frame counter
            r.begin()
        for r in rs[1:]:
            r.draw()
        rs[0].draw()
        rs[0].end()
    got = rs[0].read_framebuffer()
    ref = multi_gpu.ShardedRenderer(OracleDevice(), scene, 0, 1)
    ref.render()
    assert (got == ref.read_framebuffer()).all()
    for r in rs[::-1]:
        r.close()
    for d in devs[::-1]:
        d.close()


def test_rank_frames_merge_tables():
    """Host logic of the per-rank frame: the merged tables address the same data the per-tile frames did."""
    from oracle.backends import OracleDevice
    from webrender_b200.frame import draw_frame
    scene = _small_scene()
    mine = multi_gpu.assign_tiles(len(scene.tiles), 2)[1]
    tiles_frame, comp_frame = multi_gpu.rank_frames(scene, mine)
    dev = OracleDevice()
    handles = draw_frame(dev, tiles_frame)
    for i in mine:
        d2 = OracleDevice()
        h2 = draw_frame(d2, scene.tiles[i])
        x0, y0, x1, y1 = scene.rects[i]
        a = dev.read_pixels(handles["tile%d" % i], 0, 0, x1 - x0, y1 - y0, 4)
        b = d2.read_pixels(h2["tile"], 0, 0, x1 - x0, y1 - y0, 4)
        assert (a == b).all()
        d2.close()
    assert [op.color[0] for op in comp_frame.passes[0][0].ops] == ["tile%d" % i for i in mine]
    dev.close()
