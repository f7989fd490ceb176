"""CPU checks of the sequence-parallel host logic (tokensgen_b200/seqpar.py) and of the `_sp` entry points' argument
validation (negative return before any launch — no GPU needed)."""
import ctypes as C

import pytest


@pytest.mark.parametrize("rows,world", [(18256, 1), (18256, 2), (18256, 3), (18256, 4), (18256, 6), (18256, 8), (94, 4), (16, 8)])
def test_shard_rows_partitions_every_row_once(rows, world):
    from tokensgen_b200.seqpar import shard_rows
    chunk, shards = shard_rows(rows, world)
    assert len(shards) == world and shards[0][0] == 0
    end = 0
    for q, (row0, n) in enumerate(shards):
        assert row0 == end == q * chunk and 0 < n <= chunk      # the owner of row g is min(g // chunk, world-1)
        end = row0 + n
    assert end == rows


def test_shard_rows_rejects_empty_shards():
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.seqpar import shard_rows
    with pytest.raises(E.TokensGenError):
        shard_rows(4, 8)       # ceil(4/8) = 1 row per rank leaves ranks 4..7 without rows
    with pytest.raises(E.TokensGenError):
        shard_rows(9, 8)       # 2 rows per rank: ranks 5..7 would be empty


def test_carver_offsets_are_aligned_and_disjoint():
    from tokensgen_b200.seqpar import _Carver
    c = _Carver()
    shapes = [(2, 2282, 3072), (2, 6, 17776, 64), (2, 6, 17776, 64), (1, 3, 5)]
    for s in shapes:
        c.add(s)
    prev_end = 0
    for (off, shape), want in zip(c.items, shapes):
        n = 2
        for d in shape:
            n *= d
        assert shape == want and off % 256 == 0 and off >= prev_end
        prev_end = off + n
    assert c.nbytes >= prev_end and c.nbytes % 256 == 0


def test_scatter_descriptors_match_the_header_layout():
    from tokensgen_b200 import _ext as E
    # tg_qkv_scatter { int world; tg_bf16* peer[6][8]; }   tg_attn_scatter { int world, chunk, rows, H_total, head0; tg_bf16* peer[8]; }
    assert C.sizeof(E.QkvScatter) == 8 + 6 * 8 * 8 and E.QkvScatter.peer.offset == 8
    assert C.sizeof(E.AttnScatter) == 24 + 8 * 8 and E.AttnScatter.peer.offset == 24
    assert C.sizeof(E.RowMap) == 8 * 4
    sc = E.make_qkv_scatter([[0x1000 * (p + 1) + 0x10 * q for q in range(4)] for p in range(6)])
    assert sc.world == 4 and sc.peer[5][3] == 0x6030 and sc.peer[0][4] is None
    at = E.make_attn_scatter([0x100, 0x200], 9128, 18256, 48, 24)
    assert (at.world, at.chunk, at.rows_per_batch, at.H_total, at.head0, at.peer[1]) == (2, 9128, 18256, 48, 24, 0x200)
    m = E.make_rowmap(226, 17550, 480, 1350, 13, row0=9128, rows_local=9128)
    assert (m.rows_per_batch, m.row0, m.rows_local, E.rows_local(m)) == (18256, 9128, 9128, 9128)
    assert E.rows_local(E.make_rowmap(226, 17550, 480, 1350, 13)) == 18256


def test_sp_entry_points_validate_before_launching():
    from tokensgen_b200 import _ext as E
    lib = E.load()
    m = E.make_rowmap(4, 8, 0, 4, 2, row0=8, rows_local=8)               # 12 rows: shard [8, 16) is outside the batch
    proj = (E.QkvProj * 1)()
    proj[0].out_rows = 12
    sc = E.make_qkv_scatter([[0x1000, 0x2000]])
    rc = lib.tg_qkv_rope_gemm_sp(0x100, 64, 0x100, None, 1, 4, 64, C.byref(m), proj, 1, 1e-6, C.byref(sc), None)
    assert rc < 0 and b"shard" in lib.tg_last_error()
    m = E.make_rowmap(4, 8, 0, 4, 2, row0=0, rows_local=6)
    sc3 = E.make_qkv_scatter([[0x1000, 0x2000, 0x3000]])               # 4 heads over 3 ranks
    rc = lib.tg_qkv_rope_gemm_sp(0x100, 64, 0x100, None, 1, 4, 64, C.byref(m), proj, 1, 1e-6, C.byref(sc3), None)
    assert rc < 0 and b"divide" in lib.tg_last_error()
    rc = lib.tg_qkv_rope_gemm_sp(0x100, 64, 0x100, None, 1, 4, 64, C.byref(m), proj, 1, 1e-6, None, None)
    assert rc < 0 and b"scatter is null" in lib.tg_last_error()
    rc = lib.tg_attn_fwd_sp(0x100, 12, 0, 12, 0x100, 0x100, 12, 0, 12, None, 0, 1, 2, 0.125, 0, 1.0, None)
    assert rc < 0 and b"scatter is null" in lib.tg_last_error()


# ------------------------------------------------------------------------------------------------ clip-parallel encodes (gloo)
class _StubDist:
    def __init__(self, mean):
        self.mean = mean

    def sample(self, generator=None, scale=1.0):
        from tokensgen_b200 import _ext as E
        eps = E.randn_tensor(self.mean.shape, generator, self.mean.device, self.mean.dtype)
        return (self.mean + eps) * scale


class _StubVae:
    """encode(): a deterministic 'latent mean' of the right shape (4x temporal, 8x spatial pooling of the clip)."""
    from types import SimpleNamespace
    config = SimpleNamespace(latent_channels=16, temporal_compression_ratio=4, scaling_factor=0.7)

    def encode(self, x):
        import torch
        from types import SimpleNamespace
        b, c, f, h, w = x.shape
        t = (f - 1) // 4 + 1
        m = torch.nn.functional.adaptive_avg_pool3d(x.float(), (t, h // 8, w // 8)).mean(1, keepdim=True).repeat(1, 16, 1, 1, 1)
        return SimpleNamespace(latent_dist=_StubDist(m.to(x.dtype)))


def _stub_pipe():
    from tokensgen_b200.pipeline import MPFIFOVideoIPAdapterCogVideoXPipeline as Pipe
    pipe = Pipe.__new__(Pipe)
    pipe.vae, pipe.vae_scale_factor_spatial = _StubVae(), 8
    return pipe


def _video():
    import torch
    g = torch.Generator().manual_seed(11)
    return torch.rand(1, 45, 3, 16, 24, generator=g) * 2 - 1      # 5 chunks of 9 frames (+ 1 padded chunk)


def _encode_worker(rank, world, port, out_path):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pipe = _stub_pipe()
        pipe._clip_parallel_group = dist.group.WORLD
        gen = torch.Generator().manual_seed(5)
        lat = pipe._encode_video_chunks(_video(), 9, "cpu", torch.bfloat16, gen)
        after = torch.randn(4, generator=gen)       # the generator stream continues exactly as in the serial loop
        torch.save((lat, after), f"{out_path}.{rank}")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_clip_parallel_encode_equals_serial_loop(tmp_path, world):
    import os
    import torch
    import torch.multiprocessing as mp
    pipe = _stub_pipe()
    gen = torch.Generator().manual_seed(5)
    ref = pipe._encode_video_chunks(_video(), 9, "cpu", torch.bfloat16, gen)
    ref_after = torch.randn(4, generator=gen)
    assert tuple(ref.shape) == (1, 6 * 3, 16, 2, 3)
    out = str(tmp_path / "lat.pt")
    port = 29500 + (os.getpid() % 2000) + 10 + world
    mp.spawn(_encode_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        lat, after = torch.load(f"{out}.{r}")
        assert torch.equal(lat, ref) and torch.equal(after, ref_after), f"rank {r}"
