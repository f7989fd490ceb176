"""Host logic of the FIFO sampler: schedule bit-exact vs the oracle and the reference trace; the multi-process boundary
exchange (gloo, world_size 2 and 4 on CPU) reproduces the single-process result exactly."""
import json
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dpm as odpm
from oracle import fifo as ofifo
from tokensgen_b200.fifo import FifoQueue, FifoSchedule, VipBook, run_fifo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _timesteps():
    return odpm.DpmTables().trailing_timesteps(52)


@pytest.mark.parametrize("chunks", [1, 2, 12, 24])
def test_schedule_equals_oracle(chunks):
    s = FifoSchedule(chunks * 13, _timesteps())
    ref = ofifo.window_schedule(chunks * 13)
    assert s.num_iterations == len(ref)
    for it in range(s.num_iterations):
        got = [(w.rank, w.start, w.mid, w.end, w.real_end, w.write_lo, w.write_hi) for w in s.windows(it)]
        exp = [(w.rank, w.start, w.mid, w.end, w.real_end, w.write_lo, w.write_hi) for w in ref[it]]
        assert got == exp
    t, p, n = ofifo.fifo_timestep_tables(_timesteps())
    assert s.t.tolist() == t.tolist() and s.prev_t.tolist() == p.tolist() and s.next_t.tolist() == n.tolist()


def test_schedule_and_vip_lookup_equal_reference_trace():
    traces = json.load(open(os.path.join(ROOT, "tests", "golden", "fifo_trace.json")))
    for name, tr in traces.items():
        nf, T = 13, 52
        chunks = tr["num_frames"] // nf
        s = FifoSchedule(tr["num_frames"], _timesteps())
        img_t = np.linspace(0, chunks * nf, chunks * nf, endpoint=False, dtype=np.float32)
        cond_t = np.concatenate([np.linspace(1000 + i * nf, 1000 + (i + 1) * nf, 4, endpoint=False, dtype=np.float32)
                                 for i in range(chunks + 1)])
        n_emb = (chunks + 1) * 4
        emb = torch.arange(n_emb, dtype=torch.float32).view(1, n_emb, 1, 1, 1).repeat(2, 1, 1, 1, 1)
        book = VipBook((img_t, np.arange(30, dtype=np.float32), np.arange(45, dtype=np.float32)),
                       (cond_t, np.zeros(8, dtype=np.float32), np.zeros(12, dtype=np.float32)), emb, nf, 4, T, 1000)
        by_iter = {}
        for r in tr["records"]:
            by_iter.setdefault(r["iteration"], []).append(r)
        for it in range(s.num_iterations):
            recs = sorted(by_iter.get(it, []), key=lambda r: r["start"])
            wins = s.windows(it)
            assert len(wins) == len(recs)
            for w, r in zip(wins, recs):
                assert (w.start, w.mid, w.end, w.real_end) == (r["start"], r["mid"], r["end"], r["real_end"])
                (gt, _, _), (ct, _, _), e, _ = book.window(w.start, w.end)
                assert gt.tolist() == r["img_t"] and ct.tolist() == r["cond_t"]
                assert float(e[0, 0, 0, 0, 0]) == r["emb_first"]
            book.shift()


# ---- multi-process exchange -------------------------------------------------------------------------------------
def _toy_step(w, lat, old, t, pt, nt, gen):
    """Deterministic stand-in for DiT + scheduler: depends on every input frame, its history flag and the window's noise."""
    noise = torch.randn(lat.shape, generator=gen)
    ctx = lat.mean(dim=1, keepdim=True)
    hist = torch.stack([torch.zeros_like(lat[0, 0]) if o is None else o.reshape(lat[0, 0].shape) for o in old]).unsqueeze(0)
    tt = torch.as_tensor(np.ascontiguousarray(t), dtype=lat.dtype).view(1, -1, 1, 1, 1)
    out = 0.9 * lat + 0.05 * ctx + 0.01 * hist + 0.001 * tt + 0.01 * noise
    return out, [out[:, [j]] * 0.5 for j in range(lat.shape[1])]


def _toy_shift(queue, gen):
    noise = torch.randn(queue.x0.shape[1:], generator=gen)
    queue.latents[:, :-1] = queue.latents[:, 1:].clone()
    queue.x0[:-1] = queue.x0[1:].clone()
    queue.latents[:, -1] = 0.99 * queue.latents[:, -1] + 0.1 * noise
    queue.x0_valid = queue.x0_valid[1:] + [False]


def _make_queue():
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(1, 52, 2, 3, 4, generator=g)
    hist = [None] + [torch.randn(1, 1, 2, 3, 4, generator=g) for _ in range(51)]
    return FifoQueue(lat, hist, 6)


def _worker(rank, world, port, num_frames, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sched = FifoSchedule(num_frames, _timesteps())
        em = run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3, rank=rank, world=world)
        if rank == 0:
            torch.save(torch.cat(em[52 - 13:], dim=1), out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,num_frames", [(2, 26), (4, 26), (3, 39), (8, 39)])
def test_boundary_exchange_reproduces_single_process(tmp_path, world, num_frames):
    sched = FifoSchedule(num_frames, _timesteps())
    ref = torch.cat(run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3)[52 - 13:], dim=1)
    assert ref.shape[1] == num_frames
    out = str(tmp_path / "emitted.pt")
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, num_frames, out), nprocs=world, join=True)
    got = torch.load(out)
    assert torch.equal(got, ref)


def test_transfers_are_neighbour_sized_at_world_8():
    """At P = 8 a rank receives only the boundary slots: <= 7 from the left neighbour and 1 from the right one."""
    s = FifoSchedule(24 * 13, _timesteps())
    it = 100  # steady state: all 8 windows active
    per_dst = {}
    for src, dst, lo, hi in s.transfers(it, 8):
        per_dst.setdefault(dst, []).append((src, hi - lo))
    for dst, items in per_dst.items():
        assert sum(n for _, n in items) <= 8
        assert all(abs(src - dst) == 1 for src, _ in items)
    # 7 ranks take the 5-6 frames of their left neighbour's write region they will read, 7 ranks take 1 frame from the right
    assert sum(n for items in per_dst.values() for _, n in items) == 46


# ------------------------------------------------------------------------------------------------ checkpoint / resume
class _Crash(Exception):
    pass


def _crash_at(k):
    def progress(it):
        if it == k:
            raise _Crash()
    return progress


def test_fifo_stage_resumes_bit_identically_after_a_crash(tmp_path):
    from tokensgen_b200.fifo import FifoCheckpoint
    num_frames = 26
    sched = FifoSchedule(num_frames, _timesteps())
    ref = run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3)
    ck = FifoCheckpoint(str(tmp_path), every=7)
    with pytest.raises(_Crash):
        run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3, checkpoint=ck, progress=_crash_at(30))
    assert ck.available() == [21, 28]                    # the last two generations are kept
    resumed_from = []
    got = run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3, checkpoint=ck, on_resume=resumed_from.append)
    assert resumed_from == [28] and len(got) == len(ref) == sched.num_iterations
    assert all(torch.equal(a, b) for a, b in zip(got, ref))
    assert ck.available() == []                          # a completed run removes its own states
    # a state of another geometry is refused
    with pytest.raises(_Crash):
        run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3, checkpoint=ck, progress=_crash_at(30))
    q = _make_queue()
    q.latents = q.latents[:, :-1].contiguous()
    with pytest.raises(ValueError):
        ck.load(ck.available()[-1], q)


def test_checkpoints_of_another_run_are_never_resumed(tmp_path):
    """the CLI keys the checkpoint directory by item name; a re-run of the same item with another seed,
    prompt, guidance, world size or weights must not resume from the old run's state.  The run fingerprint is part of the file
    name and of the file."""
    from tokensgen_b200.fifo import FifoCheckpoint, run_fingerprint
    sched = FifoSchedule(26, _timesteps())
    fp_a = run_fingerprint(3, 1, 26, _timesteps(), torch.arange(8.0))
    fp_b = run_fingerprint(4, 1, 26, _timesteps(), torch.arange(8.0))             # another seed
    fp_c = run_fingerprint(3, 1, 26, _timesteps(), torch.arange(8.0) + 1)         # other priming latents
    fp_d = run_fingerprint(3, 2, 26, _timesteps(), torch.arange(8.0))             # another world size
    assert len({fp_a, fp_b, fp_c, fp_d}) == 4 and fp_a == run_fingerprint(3, 1, 26, _timesteps(), torch.arange(8.0))
    a = FifoCheckpoint(str(tmp_path), every=7, fingerprint=fp_a)
    with pytest.raises(_Crash):
        run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3, checkpoint=a, progress=_crash_at(30))
    assert a.available() == [21, 28]
    b = FifoCheckpoint(str(tmp_path), every=7, fingerprint=fp_b)
    assert b.available() == []
    resumed = []
    ref_b = run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=4)
    got_b = run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=4, checkpoint=b, on_resume=resumed.append)
    assert resumed == [] and all(torch.equal(x, y) for x, y in zip(got_b, ref_b))     # started from scratch
    assert a.available() == [21, 28]                                                  # and left run A's states alone
    # a state file renamed into another run's namespace is refused by the fingerprint stored inside it
    os.rename(a._path(28), b._path(28))
    with pytest.raises(ValueError):
        b.load(28, _make_queue())


def _resume_worker(rank, world, port, num_frames, ckdir, out_path, crash_it):
    from tokensgen_b200.fifo import FifoCheckpoint
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sched = FifoSchedule(num_frames, _timesteps())
        ck = FifoCheckpoint(ckdir, every=5, rank=rank)
        try:
            em = run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3, rank=rank, world=world, checkpoint=ck,
                          progress=_crash_at(crash_it) if crash_it is not None else None)
            if rank == 0:
                torch.save(torch.cat(em[52 - 13:], dim=1), out_path)
        except _Crash:
            pass
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_fifo_stage_resumes_across_two_ranks(tmp_path):
    num_frames = 26
    sched = FifoSchedule(num_frames, _timesteps())
    ref = torch.cat(run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3)[52 - 13:], dim=1)
    out, ckdir = str(tmp_path / "emitted.pt"), str(tmp_path / "ck")
    port = 29500 + (os.getpid() % 2000) + 40
    mp.spawn(_resume_worker, args=(2, port, num_frames, ckdir, out, 33), nprocs=2, join=True)      # both ranks stop after it 33
    assert not os.path.exists(out)
    os.remove(os.path.join(ckdir, f"fifo_state.{'0' * 16}.rank1.it000030.pt"))   # rank 1's newest save was lost: resume from 25 on both
    mp.spawn(_resume_worker, args=(2, port + 1, num_frames, ckdir, out, None), nprocs=2, join=True)
    assert torch.equal(torch.load(out), ref)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("chunks", [1, 2, 5, 24])
def test_transfer_plan_delivers_every_slot_a_rank_reads(world, chunks):
    """Pure-index simulation of run_fifo's data movement for every process count: each rank's queue replica carries, per
    slot, the iteration that last wrote it (as known to that rank); a window may only read slots whose stamp equals the
    single-process truth.  Proves `FifoSchedule.transfers` is complete without running any arithmetic."""
    s = FifoSchedule(chunks * 13, _timesteps())
    L = s.queue_len
    INIT = -1
    truth = [INIT] * L
    rep = [[INIT] * L for _ in range(world)]
    for it in range(s.num_iterations):
        wins = s.windows(it)
        for w in wins:                                         # reads: the pre-iteration queue
            r = w.rank % world
            assert rep[r][w.start:w.end] == truth[w.start:w.end], (it, w.rank, r)
        for w in wins:                                         # write-back
            for q in range(w.write_lo, w.write_hi):
                truth[q] = it
                rep[w.rank % world][q] = it
        for src, dst, lo, hi in s.transfers(it, world):         # boundary exchange
            assert all(rep[src][q] == it for q in range(lo, hi)), "a rank sends slots it did not just write"
            rep[dst][lo:hi] = rep[src][lo:hi]
        # rank 0 emits slot r_nf: it must hold the true frame
        assert rep[0][s.r_nf] == truth[s.r_nf]
        fresh = 10 ** 6 + it                                   # shift by one slot, fresh noise at the tail (generated locally)
        truth = truth[1:] + [fresh]
        rep = [x[1:] + [fresh] for x in rep]


# ------------------------------------------------------------------------------------------------ base-output broadcast
def _bcast_worker(rank, world, port, out_dir):
    from types import SimpleNamespace
    from tokensgen_b200.fifo import broadcast_base_output
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base = None
        if rank == 0:
            g = torch.Generator().manual_seed(1)
            base = SimpleNamespace(fifo_latents=torch.randn(1, 5, 2, 3, 4, generator=g), num_frames=26,
                                   fifo_old_pred_original_sample=[torch.randn(1, 1, 2, 3, 4, generator=g), None],
                                   image_rotary_emb=(torch.randn(6, 4, generator=g), torch.randn(6, 4, generator=g)),
                                   vip_image_rotary_grid=[np.arange(3, dtype=np.float32)] * 3,
                                   extra_step_kwargs={"generator": torch.Generator().manual_seed(7)}, sampling_params={"a": 1})
        got = broadcast_base_output(base, src=0, device=torch.device("cpu"))
        if rank == 0:
            assert got is base and base.extra_step_kwargs is not None          # the sender's object is untouched
        torch.save({"lat": got.fifo_latents, "old0": got.fifo_old_pred_original_sample[0],
                    "old1_is_none": got.fifo_old_pred_original_sample[1] is None, "rope1": got.image_rotary_emb[1],
                    "grid": got.vip_image_rotary_grid[2], "extra": None if rank == 0 else got.extra_step_kwargs,
                    "num_frames": got.num_frames, "sp": got.sampling_params}, os.path.join(out_dir, f"r{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_base_output_broadcast_ships_host_tensors_and_no_generator(tmp_path):
    """the bundle is pickled with CPU tensors (no CUDA context on the sender's device in every receiver) and
    without the call's torch.Generator; every field arrives intact on every rank (gloo, world size 3)."""
    port = 29500 + (os.getpid() % 2000) + 77
    mp.spawn(_bcast_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt", weights_only=False) for i in range(3)]
    for k in (1, 2):
        assert torch.equal(r[k]["lat"], r[0]["lat"]) and torch.equal(r[k]["old0"], r[0]["old0"]) and r[k]["old1_is_none"]
        assert torch.equal(r[k]["rope1"], r[0]["rope1"]) and np.array_equal(r[k]["grid"], r[0]["grid"])
        assert r[k]["extra"] is None and r[k]["num_frames"] == 26 and r[k]["sp"] == {"a": 1}


# ------------------------------------------------------------------------------------------------ streaming decode hand-off
class _ToyPipe:
    """decode_latents stand-in: a deterministic function of the chunk (each latent frame -> 4 'video' frames)."""

    @staticmethod
    def decode_latents(latents, nf):
        x = latents.permute(0, 2, 1, 3, 4)                       # b c f h w
        return torch.cat([x * (i + 1) for i in range(4)], dim=2).contiguous()


def _stream_worker(rank, world, port, num_frames, out_dir):
    from tokensgen_b200.fifo import StreamingDecoder
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sched = FifoSchedule(num_frames, _timesteps())
        sd = StreamingDecoder(_ToyPipe(), 13, 52 - 13, rank, world)
        em = run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3, rank=rank, world=world, on_emit=sd.on_emit)
        video = sd.finish()
        if rank == 0:
            torch.save({"video": video, "latents": torch.cat(em[52 - 13:], dim=1), "ready_it": dict(sd.ready_it)},
                       os.path.join(out_dir, "stream.pt"))
        else:
            assert video is None
            assert sorted(sd.frames) == list(range(rank, num_frames // 13, world))      # this rank decoded exactly its chunks
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 3])
def test_streaming_decode_hands_every_chunk_to_its_owner_as_soon_as_it_is_complete(tmp_path, world):
    """chunk c is decoded in iteration (T - nf) + nf (c + 1) - 1 on rank c % P, and the assembled video equals
    decoding the final latents after the loop (what cogvideo_sampling_mp_fifo.py:367-385 does)."""
    num_frames = 5 * 13
    port = 29500 + (os.getpid() % 2000) + 90 + world
    mp.spawn(_stream_worker, args=(world, port, num_frames, str(tmp_path)), nprocs=world, join=True)
    r = torch.load(tmp_path / "stream.pt", weights_only=False)
    assert r["ready_it"] == {c: 39 + 13 * (c + 1) - 1 for c in range(5)}
    after = torch.cat([_ToyPipe.decode_latents(r["latents"][:, c * 13:(c + 1) * 13], 13) for c in range(5)], dim=2)
    assert torch.equal(r["video"], after)


# ------------------------------------------------------------------------------------------------ ramp sharding
def _ramp_worker(rank, world, port, num_frames, out_path):
    from tokensgen_b200.fifo import RampSharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sched = FifoSchedule(num_frames, _timesteps())
        entered = []
        ramp = RampSharding(world, rank, heads=48, enter=lambda g: entered.append(dist.get_world_size(g)), leave=lambda: None)
        em = run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3, rank=rank, world=world, ramp=ramp)
        sizes = [None] * world
        dist.all_gather_object(sizes, entered)
        if rank == 0:
            torch.save({"video": torch.cat(em[52 - 13:], dim=1), "entered": sizes}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ramp_sharding_reproduces_the_single_process_result(tmp_path, world):
    """during the ramp-up the few active windows are dealt to GROUPS of ranks (their DiT forward runs
    sequence-parallel over the group on the GPU; here the step is a deterministic toy, so only the hand-off logic is under
    test): same video as one process, and every rank entered groups of the sizes the plan says."""
    num_frames = 26
    sched = FifoSchedule(num_frames, _timesteps())
    ref = torch.cat(run_fifo(sched, _make_queue(), _toy_step, _toy_shift, seed=3)[52 - 13:], dim=1)
    out = str(tmp_path / "ramp.pt")
    port = 29500 + (os.getpid() % 2000) + 120 + world
    mp.spawn(_ramp_worker, args=(world, port, num_frames, out), nprocs=world, join=True)
    r = torch.load(out, weights_only=False)
    assert torch.equal(r["video"], ref)
    # windows active per iteration: 1, then 2 from iteration 1, 3 from 7, 4 from 14, 5 from 20 (Appendix A of the survey)
    n_active = [len(sched.windows(i)) for i in range(sched.num_iterations)]
    assert n_active[0] == 1 and n_active[1] == 2 and n_active[7] == 3 and n_active[14] == 4 and n_active[20] == 5
    want = {r_: [] for r_ in range(world)}
    for a in n_active:
        g = max([g for g in (8, 4, 2) if g <= world and a * g <= world] or [1])
        if g > 1:
            for r_ in range(world):
                if r_ // g < a:
                    want[r_].append(g)
    assert r["entered"] == [want[r_] for r_ in range(world)]
    if world == 8:
        assert want[0][:1] == [8] and want[0].count(4) == 6 and want[0].count(2) == 13      # 1 + 6 + (7 + 6) iterations
