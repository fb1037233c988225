"""Host logic of the z-sharded inference on CPU: slab ranges, and a world_size-2/3 gloo run in which the CUDA kernel is
replaced by the oracle (same signature), checked against the oracle on the whole volume."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pea_oracle as O


def test_shard_ranges():
    from pea_b200.zshard import shard_ranges
    assert [b - a for a, b in shard_ranges(100, 8)] == [13, 13, 13, 13, 12, 12, 12, 12]
    assert shard_ranges(100, 1) == [(0, 100)]
    r = shard_ranges(18, 4)
    assert r[0][0] == 0 and r[-1][1] == 18 and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _oracle_compute(emb_view, z_skip, out_view):
    out_view.copy_(O.inf_embedding_loss_norm5(emb_view.contiguous())[:, :, z_skip:])


def _worker(rank, world, port, D, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pea_b200.zshard import ZShard, sharded_affinities
    torch.manual_seed(0)
    vol = torch.randn(1, 16, D, 30, 32)            # every rank builds the same volume, keeps its slab
    sh = ZShard(1, 16, D, 30, 32, rank, world, "cpu")
    sh.own.copy_(vol[:, :, sh.z0:sh.z1])
    if sh.halo:
        sh.buf[:, :, :sh.halo].fill_(float("nan"))  # must be overwritten by the exchange
    out = sharded_affinities(sh, compute=_oracle_compute)
    ref = O.inf_embedding_loss_norm5(vol)[:, :, sh.z0:sh.z1]
    # torch's CPU reductions are not bit-stable across tensor shapes (1 ulp); the GPU test asserts bit equality
    ok = bool(torch.isfinite(out).all()) and torch.allclose(out, ref, rtol=0, atol=5e-7)
    torch.save({"ok": ok, "halo_bytes": sh.halo_bytes()}, os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,D", [(2, 11), (3, 14)])
def test_zshard_gloo(world, D, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, D, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert res["ok"], f"rank {r} slab differs from the whole-volume oracle"
        assert res["halo_bytes"] == (0 if r == 0 else 1 * 16 * 4 * 30 * 32 * 4)


def test_thin_slab_is_refused():
    from pea_b200.zshard import ZShard
    with pytest.raises(ValueError):
        ZShard(1, 16, 6, 8, 8, rank=1, world=2, device="cpu")


def test_fused_halo_host_logic_with_oracle_compute():
    """fused_halo_affinities passes (own slab, 0, out, previous tail) to the kernel; with the oracle standing in for the
    kernel the concatenated slabs reproduce the whole volume."""
    from pea_b200.zshard import shard_ranges, fused_halo_affinities, HALO

    def compute(emb_own, z_skip, out_view, halo):
        assert z_skip == 0
        full = emb_own if halo is None else torch.cat([halo, emb_own], dim=2)
        n = 0 if halo is None else halo.shape[2]
        out_view.copy_(O.inf_embedding_loss_norm5(full.contiguous())[:, :, n:])

    torch.manual_seed(3)
    vol = torch.randn(1, 16, 14, 30, 32)
    ref = O.inf_embedding_loss_norm5(vol)
    parts = []
    for r, (z0, z1) in enumerate(shard_ranges(14, 3)):
        tail = vol[:, :, z0 - HALO:z0] if r else None
        parts.append(fused_halo_affinities(vol[:, :, z0:z1], tail, compute=compute))
    assert torch.allclose(torch.cat(parts, dim=2), ref, rtol=0, atol=5e-7)


def test_peer_zshard_two_barrier_contract():
    """PeerZShard.affinities() = publish (channel 0) -> kernel -> release (channel 1): the second barrier keeps a
    neighbour's in-flight reads of this slab's tail apart from the next write of the slab (a cross-GPU write-after-read
    race otherwise).  Symmetric memory needs GPUs; here the handle is a recorder and the kernel the oracle."""
    from pea_b200.zshard import PeerZShard, HALO
    log = []

    class Handle:
        def barrier(self, channel=0):
            log.append(("barrier", channel))

    torch.manual_seed(2)
    vol = torch.randn(1, 16, 9, 30, 32)
    sh = object.__new__(PeerZShard)
    sh.rank, sh.world, sh.handle = 1, 2, Handle()
    sh.z0, sh.z1, sh.depth, sh.halo = 5, 9, 4, HALO
    sh.own, sh.prev_tail = vol[:, :, 5:9], vol[:, :, 1:5]

    def compute(emb_view, z_skip, out_view, halo=None):
        log.append(("kernel", None))
        full = torch.cat([halo, emb_view], dim=2)
        out_view.copy_(O.inf_embedding_loss_norm5(full)[:, :, halo.shape[2]:])

    out = sh.affinities(compute=compute)
    assert log == [("barrier", 0), ("kernel", None), ("barrier", 1)]
    np.testing.assert_allclose(out.numpy(), O.inf_embedding_loss_norm5(vol)[:, :, 5:9].numpy(), rtol=0, atol=2e-7)
    log.clear()
    sh.world = 1   # a single rank never synchronises
    sh.prev_tail = None
    sh.affinities(compute=lambda e, z, o, h=None: log.append(("kernel", None)))
    assert log == [("kernel", None)]
