"""Sequence-parallel (Ulysses) path, SURVEY §8-f1: the sharded kernels and the sharded forward are BIT-IDENTICAL to the
unsharded ones (every rank computes exactly the single-GPU arithmetic of its rows / heads; only addresses change).

Single-GPU tests emulate the ranks as threads of this process whose "peer" buffers live on the same device (the peer
pointers the kernels take are ordinary device addresses); `test_two_ranks_over_nvlink` runs the real thing under torchrun
when the box has two GPUs."""
import os
import subprocess
import sys
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to("cuda", torch.bfloat16)


# ------------------------------------------------------------------------------------------------ row-local ops on a shard
@pytest.mark.parametrize("world", [2, 3])
def test_row_local_ops_on_shards_are_bit_identical(world):
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.seqpar import shard_rows
    B, d, n_text, hw, frames, n_vip = 2, 256, 10, 24, 3, 7
    n_video = hw * frames
    full = E.make_rowmap(n_text, n_video, n_vip, hw, frames)
    rows = full.rows_per_batch
    x = _rand(B, rows, d, seed=1)
    ln_w, ln_b, vw, vb = (_rand(d, seed=s) for s in (2, 3, 4, 5))
    tab = _rand(B * frames, 6 * d, seed=6, scale=0.3)
    shift = E.make_modvec(tab[:, 0:d], tab[:, d:2 * d], tab[:, 2 * d:3 * d])
    scale = E.make_modvec(tab[:, 3 * d:4 * d], tab[:, 4 * d:5 * d], tab[:, 5 * d:6 * d])
    y_full = torch.empty_like(x)
    E.ln_modulate(x.view(-1, d), y_full.view(-1, d), B, full, ln_w, ln_b, vw, vb, 1e-5, shift, scale)
    w, bias = _rand(d, d, seed=7, scale=0.05), _rand(d, seed=8)
    x_full = x.clone()
    E.gemm_gate_residual(y_full.view(-1, d), w, bias, x_full.view(-1, d), B, full, shift)
    _, shards = shard_rows(rows, world)
    for row0, rl in shards:
        m = E.make_rowmap(n_text, n_video, n_vip, hw, frames, row0, rl)
        xs = x[:, row0:row0 + rl].contiguous()
        ys = torch.empty_like(xs)
        E.ln_modulate(xs.view(-1, d), ys.view(-1, d), B, m, ln_w, ln_b, vw, vb, 1e-5, shift, scale)
        assert torch.equal(ys, y_full[:, row0:row0 + rl])
        E.gemm_gate_residual(ys.view(-1, d), w, bias, xs.view(-1, d), B, m, shift)
        assert torch.equal(xs, x_full[:, row0:row0 + rl])


def test_bad_shard_is_rejected():
    from tokensgen_b200 import _ext as E
    d = 256
    m = E.make_rowmap(4, 8, 0, 4, 2, row0=8, rows_local=8)  # 12 rows: [8, 16) is outside
    x = _rand(8, d)
    tab = _rand(2, 2 * d)
    mv = E.make_modvec(tab[:, :d], tab[:, d:], None)
    with pytest.raises(E.TokensGenError):
        E.ln_modulate(x, x.clone(), 1, m, _rand(d), _rand(d), None, None, 1e-5, mv, mv)


# ------------------------------------------------------------------------------------------------ the two fused all-to-alls
@pytest.mark.parametrize("world", [2, 4])
def test_qkv_and_attention_scatter_bit_identical(world):
    """Virtual ranks on one device: Q/K/V GEMM of every row shard scatters heads into per-rank buffers; attention over
    every head shard scatters rows into per-rank outputs; both equal the unsharded kernels' results exactly."""
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.seqpar import shard_rows
    B, H, K, n_text, hw, frames, n_vip = 2, 8, 256, 9, 40, 3, 11
    d = H * 64
    n_video = hw * frames
    n_tv = n_text + n_video
    full = E.make_rowmap(n_text, n_video, n_vip, hw, frames)
    rows = full.rows_per_batch
    y = _rand(B, rows, K, seed=1)
    w, bias = _rand(6 * d, K, seed=2, scale=0.06), _rand(6 * d, seed=3)
    lnw, lnb = _rand(64, seed=4), _rand(64, seed=5)
    cv, sv = torch.rand(n_video, 64, device="cuda"), torch.rand(n_video, 64, device="cuda")
    cp, sp_ = torch.rand(n_vip, 64, device="cuda"), torch.rand(n_vip, 64, device="cuda")

    def projs(outs):
        ps = []
        for i, o in enumerate(outs):
            p = E.QkvProj()
            p.out, p.out_rows = o.data_ptr(), (n_tv if i < 3 else rows)
            if i % 3 != 2:
                p.ln_w, p.ln_b = lnw.data_ptr(), lnb.data_ptr()
                p.cos_video, p.sin_video = cv.data_ptr(), sv.data_ptr()
                if i >= 3:
                    p.cos_vip, p.sin_vip = cp.data_ptr(), sp_.data_ptr()
            ps.append(p)
        return ps

    ref = [torch.zeros(B, H, n_tv if i < 3 else rows, 64, device="cuda", dtype=torch.bfloat16) for i in range(6)]
    E.qkv_rope_gemm(y.view(-1, K), w, bias, B, H, full, projs(ref), 1e-6)
    a_ref = torch.zeros(B, rows, d, device="cuda", dtype=torch.bfloat16)
    E.attn_fwd_pair(ref[0], ref[1], ref[2], n_tv, n_tv, ref[3], ref[4], ref[5], n_tv, n_vip, a_ref, 0.75)
    E.attn_fwd(ref[3], ref[4], ref[5], a_ref, q_row0=n_tv, q_rows=n_vip, out_row0=n_tv)

    hl = H // world
    chunk, shards = shard_rows(rows, world)
    bufs = [[torch.zeros(B, hl, n_tv if i < 3 else rows, 64, device="cuda", dtype=torch.bfloat16) for i in range(6)]
            for _ in range(world)]
    scat = E.make_qkv_scatter([[bufs[q][i].data_ptr() for q in range(world)] for i in range(6)])
    for row0, rl in shards:
        m = E.make_rowmap(n_text, n_video, n_vip, hw, frames, row0, rl)
        ys = y[:, row0:row0 + rl].contiguous()
        E.qkv_rope_gemm(ys.view(-1, K), w, bias, B, H, m, projs(bufs[0]), 1e-6, scatter=scat)
    for q in range(world):
        for i in range(6):
            assert torch.equal(bufs[q][i], ref[i][:, q * hl:(q + 1) * hl]), (q, i)

    a_loc = [torch.zeros(B, rl, d, device="cuda", dtype=torch.bfloat16) for _, rl in shards]
    for q in range(world):
        out = E.make_attn_scatter([a.data_ptr() for a in a_loc], chunk, rows, H, q * hl)
        b_ = bufs[q]
        E.attn_fwd_pair(b_[0], b_[1], b_[2], n_tv, n_tv, b_[3], b_[4], b_[5], n_tv, n_vip, out, 0.75)
        E.attn_fwd(b_[3], b_[4], b_[5], out, q_row0=n_tv, q_rows=n_vip, out_row0=n_tv)
    for (row0, rl), a in zip(shards, a_loc):
        assert torch.equal(a, a_ref[:, row0:row0 + rl])


# ------------------------------------------------------------------------------------------------ whole forward, virtual ranks
class _VirtualPeers:
    """Stand-in for SeqParallel on ONE device: ranks are threads, 'peer memory' is ordinary device memory, the stream
    barrier is device synchronisation + a thread barrier, the row gather goes through a shared list."""

    def __init__(self, world, rank, shared):
        self.world, self.rank, self.group, self.sh = world, rank, None, shared

    def alloc(self, carver, device):
        flat = torch.zeros(carver.nbytes // 2, device=device, dtype=torch.bfloat16)
        self.sh["base"][self.rank] = flat
        self.sh["bar"].wait()
        bases = [t.data_ptr() for t in self.sh["base"]]
        local, peers = [], []
        for off, shape in carver.items:
            n = 1
            for s in shape:
                n *= s
            local.append(flat[off // 2: off // 2 + n].view(shape))
            peers.append([b + off for b in bases])
        return local, peers

    def barrier(self):
        torch.cuda.synchronize()
        self.sh["bar"].wait()

    def gather_rows(self, local, chunk, rows):
        self.sh["rows"][self.rank] = local
        self.barrier()
        out = torch.cat(list(self.sh["rows"]), dim=1)[:, :rows].clone()
        self.barrier()
        return out


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("use_vip", [True, False])
def test_forward_virtual_ranks_bit_identical(golden_dir, world, use_vip):
    from oracle.synth import dit_shapes, synth_state_dict
    from test_dit_gpu import TINY, tiny_model
    g = torch.load(os.path.join(golden_dir, "dit_tiny.pt"))
    base = "vip" if use_vip else "plain"
    sd = synth_state_dict(dit_shapes(use_vip=use_vip, **TINY), 1234)
    # per-frame timesteps, then per-sample ones at the SAME (B, rows) geometry, then per-frame again: the sharded workspaces
    # freeze a rowmap (frames / hw) and must not be reused across the two (transformer.py workspace key)
    order = ("pf", "ps", "pf")

    def run(m, kind="pf"):
        lat, text, vip, ts = g[f"{base}_{kind}_inputs"]
        with torch.no_grad():
            return m(lat.cuda(), text.cuda(), ts.cuda(), vip_encoder_hidden_states=vip.cuda() if use_vip else None,
                     image_rotary_emb=g["rope"], vip_image_rotary_emb=g["img_rope"] if use_vip else None,
                     vip_condition_rotary_emb=g["cond_rope"] if use_vip else None, return_dict=False)[0]

    import tokensgen_b200.transformer as T
    old = T._FUSE_PAIR
    T._FUSE_PAIR = True  # the sharded path always fuses K4 + K5; compare against the same kernel sequence
    try:
        ref_model = tiny_model(use_vip, sd)
        ref = {k: run(ref_model, k).clone() for k in ("pf", "ps")}
    finally:
        T._FUSE_PAIR = old
    torch.cuda.synchronize()
    assert not torch.equal(ref["pf"], ref["ps"])
    shared = {"base": [None] * world, "rows": [None] * world, "bar": threading.Barrier(world)}
    outs, errs = [None] * world, []

    def worker(r):
        try:
            torch.cuda.set_device(0)
            m = tiny_model(use_vip, sd)
            m.__dict__["_tg_sp"] = _VirtualPeers(world, r, shared)
            outs[r] = []
            for kind in order:
                for _ in range(2):  # twice: buffer reuse across forwards
                    y = run(m, kind)
                outs[r].append(y.clone())
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
            shared["bar"].abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not errs, errs
    for r in range(world):
        assert outs[r] is not None and len(outs[r]) == len(order)
        for kind, y in zip(order, outs[r]):
            assert torch.equal(y, ref[kind]), f"rank {r} differs on the {kind} forward"


# ------------------------------------------------------------------------------------------------ real peers
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_over_nvlink():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "seqpar_check.py"), "--tiny"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "SEQPAR_OK" in r.stdout
