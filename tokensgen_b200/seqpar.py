"""Sequence-parallel (Ulysses) execution of ONE DiT forward over the GPUs of an NVSwitch domain — the single-clip latency
path of SURVEY §8-f1.  The reference runs its base stage and the T2To stage on one GPU
(longvgen/pipeline/pipeline_cogvideox_mp_fifo.py:1186-1305, pipeline_cogvideox_t2to.py:826-889); here the rows of the
residual stream are sharded across the ranks for every row-local op (LayerNorm + modulation, all GEMMs) and the HEADS are
sharded for attention.  The two all-to-alls per layer are not separate collectives: the Q/K/V GEMM's epilogue stores each
head straight into its owner rank's buffer and the attention epilogue stores each output row straight into its owner's —
16-byte stores through NVLink peer mappings (`tg_qkv_rope_gemm_sp`, `tg_attn_fwd_sp`), ordered by one cross-rank stream
barrier each.  Every rank computes exactly the arithmetic the single-GPU forward computes for its rows / heads, so the
result is bit-identical to the unsharded forward.

Peer memory comes from `torch.distributed._symmetric_memory` (plumbing: allocation, handle exchange, the barrier kernel).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _ext as E


def shard_rows(rows: int, world: int) -> Tuple[int, List[Tuple[int, int]]]:
    """(chunk, [(row0, rows_local) per rank]): every rank but the last owns `chunk` = ceil(rows / world) rows."""
    chunk = -(-rows // world)
    if chunk * (world - 1) >= rows:
        raise E.TokensGenError(f"cannot shard {rows} rows over {world} ranks")
    return chunk, [(q * chunk, min(chunk, rows - q * chunk)) for q in range(world)]


class _Carver:
    """Carves 256-byte aligned tensors out of one flat peer-mapped allocation (same offsets on every rank)."""

    def __init__(self):
        self.items, self.nbytes = [], 0

    def add(self, shape: Sequence[int]) -> int:
        n = 2
        for s in shape:
            n *= int(s)
        off = self.nbytes
        self.items.append((off, tuple(int(s) for s in shape)))
        self.nbytes = (off + n + 255) // 256 * 256
        return len(self.items) - 1


class SeqParallel:
    """Process-group handle of the sequence-parallel forward.  `group=None` uses the default group."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        if not dist.is_initialized():
            raise E.TokensGenError("sequence parallelism needs an initialised torch.distributed (NCCL) process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if not 1 <= self.world <= E.MAX_PEERS:
            raise E.TokensGenError(f"sequence-parallel group of {self.world} ranks (1..{E.MAX_PEERS} supported)")
        self._hdl = None

    # ---- peer memory
    def alloc(self, carver: _Carver, device) -> Tuple[List[torch.Tensor], List[List[int]]]:
        """Allocates `carver.nbytes` of peer-mapped memory; returns (local tensors, per-tensor peer address lists)."""
        import torch.distributed._symmetric_memory as symm
        flat = symm.empty(carver.nbytes // 2, dtype=torch.bfloat16, device=device)
        self._hdl = symm.rendezvous(flat, self.group)
        bases = [int(p) for p in self._hdl.buffer_ptrs]
        if len(bases) != self.world or bases[self.rank] != flat.data_ptr():
            raise E.TokensGenError("symmetric-memory rendezvous returned an unexpected peer table")
        self._flat = flat
        local, peers = [], []
        for off, shape in carver.items:
            n = 1
            for s in shape:
                n *= s
            local.append(flat[off // 2: off // 2 + n].view(shape))
            peers.append([b + off for b in bases])
        return local, peers

    def barrier(self) -> None:
        """Cross-rank barrier enqueued on the current stream: every rank's earlier kernels (and their peer stores) are
        complete before any rank's later kernels start."""
        with E._Timed("sp_barrier", 1):
            self._hdl.barrier(channel=0)

    # ---- gather of the (tiny) projected output rows
    def gather_rows(self, local: torch.Tensor, chunk: int, rows: int) -> torch.Tensor:
        """local [B, chunk, n] (rows beyond this rank's share are padding) -> [B, rows, n] on every rank."""
        B, _, n = local.shape
        out = torch.empty(self.world, B, chunk, n, device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out.permute(1, 0, 2, 3).reshape(B, self.world * chunk, n)[:, :rows]


class ShardedBuffers:
    """Per-rank workspaces of the sequence-parallel forward (the counterpart of transformer._Buffers)."""

    def __init__(self, sp: SeqParallel, B: int, rowmap: E.RowMap, d: int, H: int, ff_dim: int, use_vip: bool, device):
        rows, n_tv = rowmap.rows_per_batch, rowmap.n_text + rowmap.n_video
        if H % sp.world:
            raise E.TokensGenError(f"{H} heads do not divide over {sp.world} ranks")
        self.sp, self.H, self.Hloc = sp, H, H // sp.world
        self.chunk, shards = shard_rows(rows, sp.world)
        self.row0, self.rows_local = shards[sp.rank]
        self.rowmap = E.make_rowmap(rowmap.n_text, rowmap.n_video, rowmap.n_vip, rowmap.hw, rowmap.frames,
                                    self.row0, self.rows_local)
        bf = dict(device=device, dtype=torch.bfloat16)
        self.X = torch.empty(B, self.rows_local, d, **bf)
        self.Y = torch.empty(B, self.rows_local, d, **bf)
        self.Hff = torch.empty(B * self.rows_local, ff_dim, **bf) if ff_dim else None
        self.Xfull = torch.empty(B, rows, d, **bf)  # patch-embedding staging (computed redundantly on every rank)
        carver = _Carver()
        ia = carver.add((B, self.chunk, d))
        iq = [carver.add((B, self.Hloc, n_tv, 64)) for _ in range(3)]
        if use_vip:
            iq += [carver.add((B, self.Hloc, rows, 64)) for _ in range(3)]
        local, peers = sp.alloc(carver, device)
        self.A = local[ia].view(-1)[: B * self.rows_local * d].view(B, self.rows_local, d)
        self.qkv = [local[i] for i in iq]
        self.qkv_scatter = E.make_qkv_scatter([peers[i] for i in iq])
        self._a_peers, self._d, self._rows, self._shards = peers[ia], d, rows, shards
        self.attn_scatter = E.make_attn_scatter(self._a_peers, self.chunk, rows, H, sp.rank * self.Hloc)
        self.side_stream = torch.cuda.Stream(device=device)

    def attn_scatter_for_batch(self, bi: int) -> E.AttnScatter:
        """Scatter descriptor for a B = 1 attention call on batch element `bi`."""
        ptrs = [p + 2 * bi * rl * self._d for p, (_, rl) in zip(self._a_peers, self._shards)]
        return E.make_attn_scatter(ptrs, self.chunk, self._rows, self.H, self.sp.rank * self.Hloc)
