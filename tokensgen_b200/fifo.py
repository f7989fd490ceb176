"""FIFO diagonal-queue sampler: controller + worker step of the reference's cogvideo_fifo_mp_v2 / fifo_onestep_per_gpu
(longvgen/fifo_sampling/cogvideo_sampling_mp_fifo.py:27-395, :408-579), re-designed for one persistent process per GPU.

Reference design: a controller process owns the 58-slot queue, ships every window through mp.Queue (CUDA IPC) to one
spawned worker per GPU and copies results back, per iteration.
This design:      every rank runs the same deterministic controller on its own replica of the (10 MB) queue; window
rank w is owned by process w % P (the reference's `rank = base_rank * num_processes + sub_rank` assignment); after each
iteration only the slots a *different* process will read next are exchanged, as grouped NCCL send/recv
(`torch.distributed.batch_isend_irecv`) — for P = 8 that is the 6-7 boundary frames from the left neighbour and one from
the right neighbour.  No queue gather, no pickling of pipelines, no per-video process spawn.

Determinism: the noise a window consumes is drawn from a generator seeded by (seed, iteration, window rank) and the
tail re-noise from (seed, iteration), so the result does not depend on P — the property the multi-process tests check.

The integer index schedule is bit-identical to the reference controller (tests/golden/fifo_trace.json).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class Window:
    iteration: int
    rank: int       # window rank 0..2*num_partitions-1
    start: int
    mid: int
    end: int
    real_end: int
    write_lo: int   # queue slots [write_lo, write_hi) take this window's outputs
    write_hi: int

    @property
    def out_lo(self) -> int:
        """Offset of write_lo inside the window's 13 output frames."""
        return self.write_lo - self.start


class FifoSchedule:
    """Index arithmetic of cogvideo_fifo_mp_v2 (:175-188 tables, :223-259 windows, :322-327 write-back, :358 padding)."""

    def __init__(self, num_frames: int, timesteps: Sequence[int], nf_per_chunk: int = 13, num_partitions: int = 4,
                 use_adaptive_padding: bool = True):
        self.nf = nf_per_chunk
        self.l_nf, self.r_nf = nf_per_chunk - nf_per_chunk // 2, nf_per_chunk // 2
        self.T = len(timesteps)
        self.num_frames = num_frames
        self.num_rank = 2 * num_partitions
        self.num_iterations = num_frames + self.T - nf_per_chunk
        self.queue_len = self.T + self.r_nf
        self.initial_queue_start = self.T - self.l_nf if use_adaptive_padding else 0
        ts = np.asarray(timesteps, dtype=np.int64)
        r = self.r_nf
        self.t = np.concatenate([ts, np.full(r, ts[-1])])[::-1].copy()
        self.prev_t = np.concatenate([ts[1:], np.full(r + 1, -1)])[::-1].copy()
        self.next_t = np.concatenate([np.full(1, -1), ts[:-1], np.full(r, ts[-2])])[::-1].copy()

    def queue_start(self, iteration: int) -> int:
        return max(0, self.initial_queue_start - iteration)

    def windows(self, iteration: int) -> List[Window]:
        qs = self.queue_start(iteration)
        out = []
        for rank in range(self.num_rank):
            start = self.nf * (rank // 2) + self.r_nf * (rank % 2)
            nxt = self.nf * ((rank + 1) // 2) + self.r_nf * ((rank + 1) % 2)
            if nxt <= qs:
                continue
            mid = start + (self.l_nf if rank % 2 == 1 else self.r_nf)
            real_end = start + self.nf
            if start < qs:
                start = qs
            end = start + self.nf
            lo, hi = (mid, end) if start > qs else (max(self.r_nf, start), real_end)
            out.append(Window(iteration, rank, start, mid, end, real_end, lo, hi))
        return out

    def transfers(self, iteration: int, world: int) -> List[Tuple[int, int, int, int]]:
        """(src_process, dst_process, slot_lo, slot_hi) in PRE-shift slot numbers: what was written in `iteration` by one
        process and is read in `iteration + 1` by another.  Identical on every rank (pure function of the schedule)."""
        if iteration + 1 >= self.num_iterations:
            return []
        cur, nxt = self.windows(iteration), self.windows(iteration + 1)
        out = []
        for w in cur:
            src = w.rank % world
            for v in nxt:
                dst = v.rank % world
                if dst == src:
                    continue
                lo, hi = max(w.write_lo, v.start + 1), min(w.write_hi, v.end + 1)  # v reads pre-shift [start+1, end+1)
                if lo < hi:
                    out.append((src, dst, lo, hi))
        # a slot range may be requested by several windows of the same destination: merge duplicates
        return sorted(set(out))


StepFn = Callable[[Window, torch.Tensor, List[Optional[torch.Tensor]], np.ndarray, np.ndarray, np.ndarray, torch.Generator],
                  Tuple[torch.Tensor, List[torch.Tensor]]]


class FifoQueue:
    """Device-resident queue state of one rank: latents [1, L, C, H, W], x0 history [L, C, H, W] + validity flags."""

    def __init__(self, fifo_latents: torch.Tensor, fifo_old_x0: Sequence[Optional[torch.Tensor]], r_nf: int):
        pad = [fifo_latents[:, [0]]] * r_nf
        self.latents = torch.cat(pad + [fifo_latents], dim=1).contiguous()      # prepare_fifo_latents (:72-82)
        hist = [fifo_old_x0[0]] * r_nf + list(fifo_old_x0)                       # :145-146
        self.x0 = torch.zeros_like(self.latents[0])
        self.x0_valid = []
        for i, h in enumerate(hist):
            self.x0_valid.append(h is not None)
            if h is not None:
                self.x0[i].copy_(h.reshape(self.x0[i].shape))

    def window_inputs(self, w: Window):
        lat = self.latents[:, w.start:w.end]
        old = [self.x0[i].unsqueeze(0).unsqueeze(0) if self.x0_valid[i] else None for i in range(w.start, w.end)]
        return lat, old


class RampSharding:
    """Idle GPUs join the active windows during the FIFO ramp-up.

    For the first T - l_nf iterations fewer windows are active than the 2 x num_partitions the queue eventually feeds
    (cogvideo_sampling_mp_fifo.py:243-244 skips windows left of `queue_start_idx`): 1 window in iteration 0, 2 from
    iteration 1, 3 from 7, ... 8 from 40.  The reference — and plain window parallelism — leaves the other GPUs idle there,
    which bounds the 8-GPU speed-up of a 24-chunk video at 7.6x and of short videos far lower.  Here, while `a` active
    windows satisfy a * g <= P for a power-of-two group size g >= 2, the P ranks form P / g groups of g neighbours, group i
    runs window i with its DiT forward SEQUENCE-PARALLEL over the group (tokensgen_b200/seqpar.py: bit-identical to the
    unsharded forward, all-to-alls fused into kernel epilogues over NVLink), and every window's written slots are broadcast
    from its group so that all queue replicas stay whole.  From the first iteration with a * 2 > P on, the usual
    one-window-per-rank assignment with its neighbour exchange takes over.  Results do not depend on P."""

    def __init__(self, world: int, rank: int, heads: int = 48, enter: Optional[Callable] = None,
                 leave: Optional[Callable] = None, make_group: Optional[Callable] = None, min_rows: int = 0):
        """enter(group) / leave(): switch the transformer's sequence-parallel mode (defaults: no-ops, for host-logic tests);
        make_group(ranks) -> process group (default torch.distributed.new_group; collective: every rank creates every group)."""
        import torch.distributed as dist
        self.world, self.rank = world, rank
        self.enter, self.leave = enter or (lambda g: None), leave or (lambda: None)
        make_group = make_group or (lambda ranks: dist.new_group(ranks))
        self.groups: Dict[int, list] = {}
        for g in (8, 4, 2):
            if g <= world and world % g == 0 and heads % g == 0:
                self.groups[g] = [make_group(list(range(i * g, (i + 1) * g))) for i in range(world // g)]

    def group_size(self, n_active: int) -> int:
        for g in sorted(self.groups, reverse=True):
            if n_active * g <= self.world:
                return g
        return 1

    def my_group(self, g: int):
        return self.groups[g][self.rank // g]


def run_fingerprint(*parts) -> str:
    """Identity of one FIFO run for checkpoint matching: ints / floats / strings / lists are hashed by repr, tensors by their
    bytes (CPU copy).  Two runs share checkpoints only if every part agrees."""
    import hashlib
    h = hashlib.sha256()
    for p in parts:
        if torch.is_tensor(p):
            t = p.detach().to("cpu").contiguous()
            h.update(str((tuple(t.shape), str(t.dtype))).encode())
            h.update(t.view(torch.uint8).numpy().tobytes())
        elif isinstance(p, np.ndarray):
            h.update(str((p.shape, str(p.dtype))).encode())
            h.update(np.ascontiguousarray(p).tobytes())
        else:
            h.update(repr(p).encode())
        h.update(b"|")
    return h.hexdigest()[:16]


class FifoCheckpoint:
    """Restartable FIFO stage.  The reference's run is all-or-nothing (a 2-minute video is 351 iterations — 38
    minutes on one GPU — and a worker crash deadlocks its controller, cogvideo_sampling_mp_fifo.py:309).  Every noise draw
    of the loop is seeded by (seed, iteration, window rank), so the whole state of the stage between two iterations is the
    queue (58 latent frames + x0 history + validity flags, ~20 MB) and the frames emitted so far: each rank writes its own
    replica every `every` iterations (atomic rename, the last two generations are kept) and a restarted job resumes from
    the newest iteration present on ALL ranks with bit-identical results.

    A state only ever resumes the run that wrote it: `fingerprint` (seed, world size, schedule, and a hash of the priming
    latents / prompt / condensed-token embeddings / a weight probe — `run_fingerprint`) is part of the file name AND stored
    inside the file; states of other runs in the same directory are ignored, and a run that completes removes its own."""

    def __init__(self, directory: str, every: int, rank: int = 0, fingerprint: str = "0" * 16):
        self.dir, self.every, self.rank, self.fingerprint = directory, int(every), rank, str(fingerprint)
        os.makedirs(directory, exist_ok=True)

    def _prefix(self) -> str:
        return f"fifo_state.{self.fingerprint}.rank{self.rank}.it"

    def _path(self, it: int) -> str:
        return os.path.join(self.dir, f"{self._prefix()}{it:06d}.pt")

    def available(self) -> List[int]:
        pre, suf = self._prefix(), ".pt"
        return sorted(int(f[len(pre):-len(suf)]) for f in os.listdir(self.dir) if f.startswith(pre) and f.endswith(suf))

    def save(self, next_it: int, queue: "FifoQueue", emitted: List[torch.Tensor]) -> None:
        state = {"next_it": next_it, "fingerprint": self.fingerprint, "latents": queue.latents.cpu(), "x0": queue.x0.cpu(),
                 "x0_valid": list(queue.x0_valid), "emitted": [e.cpu() for e in emitted]}
        tmp = self._path(next_it) + ".tmp"
        torch.save(state, tmp)
        os.replace(tmp, self._path(next_it))
        for old in self.available()[:-2]:
            os.remove(self._path(old))

    def load(self, it: int, queue: "FifoQueue") -> List[torch.Tensor]:
        state = torch.load(self._path(it), map_location="cpu", weights_only=True)
        if state.get("fingerprint") != self.fingerprint:
            raise ValueError(f"{self._path(it)} was written by a different run (fingerprint mismatch)")
        if state["next_it"] != it or tuple(state["latents"].shape) != tuple(queue.latents.shape):
            raise ValueError(f"{self._path(it)} does not belong to this run (shape / iteration mismatch)")
        queue.latents.copy_(state["latents"])
        queue.x0.copy_(state["x0"])
        queue.x0_valid = list(state["x0_valid"])
        return [e.to(queue.latents.device) for e in state["emitted"]]

    def clear(self) -> None:
        """The run completed: its states must not resume a later run of the same item."""
        for it in self.available():
            os.remove(self._path(it))


def run_fifo(schedule: FifoSchedule, queue: FifoQueue, step_fn: StepFn, shift_fn, seed: int = 0, rank: int = 0,
             world: int = 1, group=None, progress: Optional[Callable[[int], None]] = None,
             checkpoint: Optional[FifoCheckpoint] = None, on_resume: Optional[Callable[[int], None]] = None,
             on_emit: Optional[Callable[[int, List[torch.Tensor]], None]] = None,
             ramp: Optional[RampSharding] = None) -> List[torch.Tensor]:
    """The controller loop (:230-359).  `step_fn(window, latents[1,13,...], old_x0 list, t, prev_t, next_t, generator)`
    returns (latents_out [1,13,...], x0 list); `shift_fn(queue, noise_generator)` advances the queue by one slot and
    re-noises the tail.  Returns the emitted frames (slot r_nf of every iteration) as seen by this rank; rank 0's list is
    the video (emissions before iteration T - nf are the ramp-up the reference drops, :367).
    `checkpoint`: save the stage state every `checkpoint.every` iterations and resume from the newest state all ranks
    hold; `on_resume(k)` lets the caller fast-forward its own per-iteration bookkeeping (VipBook.shift) by k iterations.
    `on_emit(it, emitted)`: called on EVERY rank right after iteration `it` appended its frame (the list is complete on
    rank 0 only) and before the queue shifts — the hook of the streaming decode (`StreamingDecoder`).
    `ramp`: share the few windows of the ramp-up iterations among groups of ranks (`RampSharding`)."""
    import torch.distributed as dist
    emitted = []
    dev = queue.latents.device
    start_it = 0
    if checkpoint is not None:
        have = checkpoint.available()
        if world > 1:   # newest iteration every rank has (a crash may have interrupted a save on some of them)
            sets = [None] * world
            dist.all_gather_object(sets, have, group=group)
            have = sorted(set.intersection(*[set(x) for x in sets]))
        if have:
            start_it = have[-1]
            emitted = checkpoint.load(start_it, queue)
            if on_resume is not None:
                on_resume(start_it)
    for it in range(start_it, schedule.num_iterations):
        wins = schedule.windows(it)
        g_ramp = ramp.group_size(len(wins)) if (ramp is not None and world > 1) else 1
        if g_ramp > 1:      # ramp-up: window i of this iteration belongs to the i-th group of g_ramp neighbouring ranks
            gi = rank // g_ramp
            mine = [wins[gi]] if gi < len(wins) else []
            if mine:
                ramp.enter(ramp.my_group(g_ramp))
        else:
            mine = [w for w in wins if w.rank % world == rank]
        results = []
        for w in mine:  # all windows read the pre-iteration queue (:232,258) -> compute first, write back after
            lat, old = queue.window_inputs(w)
            gen = torch.Generator(device=dev).manual_seed((seed * 1000003 + it) * 64 + w.rank)
            t, pt, nt = schedule.t[w.start:w.end], schedule.prev_t[w.start:w.end], schedule.next_t[w.start:w.end]
            results.append(step_fn(w, lat.clone(), old, t, pt, nt, gen))
        if g_ramp > 1 and mine:
            ramp.leave()
        for w, (out_lat, out_x0) in zip(mine, results):
            n = w.write_hi - w.write_lo
            queue.latents[:, w.write_lo:w.write_hi] = out_lat[:, w.out_lo:w.out_lo + n]
            for k in range(n):
                queue.x0[w.write_lo + k].copy_(out_x0[w.out_lo + k].reshape(queue.x0[0].shape))
                queue.x0_valid[w.write_lo + k] = True
        if g_ramp > 1:
            # every replica is made whole again: the slots window i wrote travel from the first rank of its group to all
            # (<= 4 broadcasts of <= 13 frames x 2 x 346 KB per iteration; the regular neighbour plan assumes the regular owners)
            for i, w in enumerate(wins):
                n = w.write_hi - w.write_lo
                if rank // g_ramp == i:
                    buf = torch.stack([queue.latents[0, w.write_lo:w.write_hi], queue.x0[w.write_lo:w.write_hi]]).contiguous()
                else:
                    buf = torch.empty((2, n) + tuple(queue.x0.shape[1:]), device=dev, dtype=queue.x0.dtype)
                dist.broadcast(buf, src=i * g_ramp if group is None else dist.get_global_rank(group, i * g_ramp), group=group)
                if rank // g_ramp != i:
                    queue.latents[0, w.write_lo:w.write_hi] = buf[0]
                    queue.x0[w.write_lo:w.write_hi] = buf[1]
                    for s_ in range(w.write_lo, w.write_hi):
                        queue.x0_valid[s_] = True
        elif world > 1:
            ops, recvs = [], []
            for (src, dst, lo, hi) in schedule.transfers(it, world):
                if src == rank:
                    buf = torch.stack([queue.latents[0, lo:hi], queue.x0[lo:hi]]).contiguous()
                    ops.append(dist.P2POp(dist.isend, buf, dst, group=group))
                elif dst == rank:
                    buf = torch.empty((2, hi - lo) + tuple(queue.x0.shape[1:]), device=dev, dtype=queue.x0.dtype)
                    ops.append(dist.P2POp(dist.irecv, buf, src, group=group))
                    recvs.append((lo, hi, buf))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            for lo, hi, buf in recvs:
                queue.latents[0, lo:hi] = buf[0]
                queue.x0[lo:hi] = buf[1]
                for s in range(lo, hi):
                    queue.x0_valid[s] = True
        emitted.append(queue.latents[:, [schedule.r_nf]].clone())
        if on_emit is not None:
            on_emit(it, emitted)
        ngen = torch.Generator(device=dev).manual_seed(seed * 1000003 + it + 7919)
        shift_fn(queue, ngen)
        if checkpoint is not None and checkpoint.every > 0 and (it + 1) % checkpoint.every == 0 \
                and it + 1 < schedule.num_iterations:
            checkpoint.save(it + 1, queue, emitted)
        if progress is not None:
            progress(it)
    if checkpoint is not None:
        checkpoint.clear()
    return emitted


def make_shift_fn(scheduler, noise_override=None):
    """shift_latents (:117-131) on the GPU: one tg_queue_shift_renoise launch for the latent queue and its x0 history.
    `noise_override(shape)` supplies the re-noise draw instead of the (seed, iteration) generator (parity tests)."""
    from . import _ext as E
    b = scheduler.betas[999].double()
    s1, s2 = float((1 - b) ** 0.5), float(b ** 0.5)

    def shift(queue: FifoQueue, gen: torch.Generator):
        if noise_override is not None:
            noise = noise_override(tuple(queue.x0.shape[1:])).to(device=queue.x0.device, dtype=torch.bfloat16).contiguous()
        else:
            noise = torch.randn(queue.x0.shape[1:], generator=gen, device=queue.x0.device, dtype=torch.bfloat16)
        E.queue_shift_renoise(queue.latents[0], queue.x0, noise, s1, s2)
        queue.x0_valid = queue.x0_valid[1:] + [False]

    return shift


class VipBook:
    """Bookkeeping of the condensed-token conditioning along the queue (cogvideo_sampling_mp_fifo.py:84-115,133-139,
    148-173,261-273,351-357): per-slot temporal positions of the video tokens, the (extended) condition grid and
    embeddings, and the searchsorted lookup of the 5 condensed-token frames a window attends to."""

    def __init__(self, img_grid, cond_grid, image_embeddings: torch.Tensor, nf: int, vip_nf: int, T: int,
                 start_frame_idx: float):
        img_t, self.img_h, self.img_w = [np.asarray(g, dtype=np.float32) for g in img_grid]
        cond_t, self.cond_h, self.cond_w = [np.asarray(g, dtype=np.float32) for g in cond_grid]
        r_nf = nf // 2
        self.nf, self.vip_nf, self.start_frame_idx = nf, vip_nf, start_frame_idx
        self.slot_t = np.concatenate([img_t[[0]]] * (r_nf + T - nf) + [img_t[:nf]])                      # :84-88
        self.t_queue = np.concatenate([img_t[nf:], np.linspace(img_t[-1] + 1, img_t[-1] + 1 + T, T, endpoint=False,
                                                                dtype=np.float32)])                        # :90-91
        self.cond_t = np.concatenate([cond_t] + [cond_t[-vip_nf:] + (i + 1) * nf for i in range(T // nf + 1)])  # :95-99
        self.embeddings = torch.cat([image_embeddings] + [image_embeddings[:, -vip_nf:]] * (T // nf + 1), dim=1)  # :101-108

    def window(self, start: int, end: int):
        idx = int(np.searchsorted(self.cond_t, self.slot_t[start] + self.start_frame_idx, side="right") - 1)  # :110-115
        n = min(self.vip_nf + 1, self.nf)
        return (self.slot_t[start:end].copy(), self.img_h, self.img_w), \
               (self.cond_t[idx:idx + n].copy(), self.cond_h, self.cond_w), self.embeddings[:, idx:idx + n], idx

    def shift(self):
        self.slot_t[:-1] = self.slot_t[1:].copy()                                                          # :133-139
        self.slot_t[-1] = self.t_queue[0]
        self.t_queue = self.t_queue[1:]


def make_step_fn(transformer, scheduler, prompt_embeds: torch.Tensor, image_rotary_emb, vip: Optional[VipBook],
                 guidance_scale: float, head_dim: int = 64, noise_override=None, guidance_scale_img: Optional[float] = None):
    """fifo_onestep_per_gpu (:408-579) for one window on this rank's GPU: DiT forward (CFG pair, per-frame timesteps)
    then the fused CFG + per-frame DPM step."""
    from .rope import get_3d_rotary_pos_embed_v2
    dev = prompt_embeds.device
    B = prompt_embeds.shape[0]

    def step(w: Window, latents, old_x0, t, prev_t, next_t, gen):
        kw = {}
        if vip is not None:
            img_grid, cond_grid, emb, _ = vip.window(w.start, w.end)
            kw = dict(vip_image_rotary_emb=get_3d_rotary_pos_embed_v2(head_dim, *img_grid, device=dev),
                      vip_condition_rotary_emb=get_3d_rotary_pos_embed_v2(head_dim, *cond_grid, device=dev),
                      vip_encoder_hidden_states=emb.contiguous())
        ts = torch.as_tensor(np.ascontiguousarray(t), device=dev).expand(B, -1)
        noise_pred = transformer(hidden_states=torch.cat([latents] * B), encoder_hidden_states=prompt_embeds, timestep=ts,
                                 image_rotary_emb=image_rotary_emb, return_dict=False, **kw)[0]
        return scheduler.window_step(noise_pred, latents, old_x0, t, prev_t, next_t, guidance_scale, generator=gen,
                                     noise=None if noise_override is None else noise_override(w),
                                     guidance_scale_img=guidance_scale_img if B == 3 else None)

    return step


def broadcast_base_output(base_output, src: int = 0, device=None, group=None):
    """Ships the FIFO priming state of the base stage from rank `src` to every rank (one object broadcast of ~0.2 GB for
    gen.yaml: 52 + 52 latent frames, prompt and condensed-token embeddings, grids).  The reference hands the same bundle
    to its worker processes through mp.Queue on every window (cogvideo_sampling_mp_fifo.py:284-305); here it crosses
    NVLink once per video.  The bundle is pickled with its tensors on the HOST (a pickled CUDA tensor is rebuilt on the
    sender's device index by every receiver: a context + ~0.2 GB on GPU 0 per rank) and without the call's
    torch.Generator (`extra_step_kwargs`, unused by the FIFO stage); receivers move the tensors to their own `device`."""
    import copy
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return base_output

    def convert(v, fn):
        if torch.is_tensor(v):
            return fn(v)
        if isinstance(v, (list, tuple)) and v and all(torch.is_tensor(x) or x is None for x in v):
            return type(v)(None if x is None else fn(x) for x in v)
        return v

    box = [None]
    if dist.get_rank(group) == src:
        host = copy.copy(base_output)
        for name, v in list(vars(host).items()):
            setattr(host, name, convert(v, lambda t: t.detach().cpu()))
        host.extra_step_kwargs = None
        box = [host]
    dist.broadcast_object_list(box, src=src, group=group, device=device)
    if dist.get_rank(group) == src:
        return base_output
    out = box[0]
    if device is not None:
        for name, v in list(vars(out).items()):
            setattr(out, name, convert(v, lambda t: t.to(device)))
    return out


def decode_latents_parallel(pipe, latents: torch.Tensor, nf: int, rank: int = 0, world: int = 1):
    """pipe.decode_latents (pipeline_cogvideox_mp_fifo.py:676-684) with its 13-frame chunks dealt round-robin to the
    processes; process 0 returns the assembled [B, 3, F, H, W] video, the others None."""
    import torch.distributed as dist
    chunks = latents.shape[1] // nf
    outs = {c: pipe.decode_latents(latents[:, c * nf:(c + 1) * nf], nf) for c in range(rank, chunks, world)}
    if world == 1:
        return torch.cat([outs[c] for c in range(chunks)], dim=2)
    ops = []
    if rank != 0:
        ops = [dist.P2POp(dist.isend, outs[c].contiguous(), 0) for c in sorted(outs)]
    else:
        like = outs[0]
        for c in range(chunks):
            if c % world:
                outs[c] = torch.empty_like(like)
                ops.append(dist.P2POp(dist.irecv, outs[c], c % world))
    if ops:
        for req in dist.batch_isend_irecv(ops):   # one grouped NCCL launch; messages between a pair of ranks match in order
            req.wait()
    return None if rank != 0 else torch.cat([outs[c] for c in range(chunks)], dim=2)


# Streaming decode on a SIDE stream, overlapping the window forwards of the same GPU (TG_STREAM_DECODE_OVERLAP=1).  Off by
# default: at full size (CogVideoX-5b window forwards on one stream, the tiled 480 x 720 decode on another) the device hung in
# tools/concurrency_check.py — CTA-pair GEMMs / attention on one side, CTA-pair convolutions on the other; not understood yet,
# and the tiny-shape tests never showed it.  Without the overlap the decode of a chunk costs its rank one window step's worth of
# time every 13 iterations (0.5 s in 13 x 8 x 0.8 s), still spread over the ranks and with no decode tail.
import os as _os
_DECODE_OVERLAP = _os.environ.get("TG_STREAM_DECODE_OVERLAP", "0") != "0"


class StreamingDecoder:
    """Streaming decode under the FIFO loop.  The reference decodes all chunks serially on GPU 0 AFTER the
    loop (cogvideo_sampling_mp_fifo.py:367-385); `decode_latents_parallel` already spreads that tail over the ranks.  Here
    chunk c is decoded as soon as its `nf` latent frames have left the queue — iteration (T - nf) + nf (c + 1) - 1 — on
    rank c % P (optionally on a SIDE stream while the main stream goes on denoising, _DECODE_OVERLAP): the first 49 frames are
    available after 52 + 13 iterations instead of after the whole stage, and no decode is left for the end except the last chunk's.

    Every rank runs the same deterministic loop, so the owner knows in which iteration to post its receive: rank 0 (the
    emitting rank: slot r_nf belongs to window rank 0) sends the chunk's latents (2.25 MB) to the owner inside the
    iteration, the owner decodes (clip-local: the VAE's conv cache is cleared per chunk, autoencoder_kl_cogvideox.py:1157),
    and `finish()` collects the frames on rank 0.  The arithmetic is `pipe.decode_latents` on the same latents, so the video
    is bit-identical to the decode-after-the-loop path."""

    def __init__(self, pipe, nf: int, first_it: int, rank: int = 0, world: int = 1, group=None):
        self.pipe, self.nf, self.first_it, self.rank, self.world, self.group = pipe, nf, first_it, rank, world, group
        self.done = 0                       # chunks dispatched so far
        self.frames: Dict[int, torch.Tensor] = {}
        self.ready_it: Dict[int, int] = {}  # chunk -> iteration in which its decode was enqueued
        self.side = None

    def _decode(self, c: int, latents: torch.Tensor):
        dev = latents.device
        if dev.type != "cuda":      # host tensors (the gloo tests of the hand-off logic): no streams, decode in line
            self.frames[c] = self.pipe.decode_latents(latents, self.nf)
            return
        if not _DECODE_OVERLAP:
            # in line on the denoising stream: the chunk is still decoded the iteration it completes, on rank c % P, and nothing
            # is left for the end — but the VAE's kernels never run BESIDE the DiT's on one GPU (see _DECODE_OVERLAP)
            self.frames[c] = self.pipe.decode_latents(latents, self.nf)
            return
        if self.side is None:
            self.side = torch.cuda.Stream(device=dev)
        self.side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.side):
            self.frames[c] = self.pipe.decode_latents(latents, self.nf)
        latents.record_stream(self.side)

    def on_emit(self, it: int, emitted: List[torch.Tensor]) -> None:
        import torch.distributed as dist
        n_video = len(emitted) - self.first_it          # frames of the video emitted so far (negative during the ramp-up)
        while n_video >= (self.done + 1) * self.nf:
            c = self.done
            owner = c % self.world
            lo = self.first_it + c * self.nf
            if self.rank == 0:
                chunk = torch.cat(emitted[lo:lo + self.nf], dim=1).contiguous()
                if owner != 0:
                    dist.send(chunk, dst=owner if self.group is None else dist.get_global_rank(self.group, owner), group=self.group)
                else:
                    self._decode(c, chunk)
            elif self.rank == owner:
                like = emitted[0]
                chunk = torch.empty((like.shape[0], self.nf) + tuple(like.shape[2:]), device=like.device, dtype=like.dtype)
                dist.recv(chunk, src=0 if self.group is None else dist.get_global_rank(self.group, 0), group=self.group)
                self._decode(c, chunk)
            self.ready_it[c] = it
            self.done += 1

    def finish(self) -> Optional[torch.Tensor]:
        """Frames [B, 3, F, H, W] on rank 0 (None elsewhere), chunks in order."""
        import torch.distributed as dist
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        chunks = self.done
        if self.world == 1:
            return torch.cat([self.frames[c] for c in range(chunks)], dim=2)
        ops = []
        if self.rank != 0:
            ops = [dist.P2POp(dist.isend, self.frames[c].contiguous(), 0, group=self.group) for c in sorted(self.frames)]
        else:
            like = self.frames[0]
            for c in range(chunks):
                if c % self.world:
                    self.frames[c] = torch.empty_like(like)
                    ops.append(dist.P2POp(dist.irecv, self.frames[c], c % self.world, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return None if self.rank != 0 else torch.cat([self.frames[c] for c in range(chunks)], dim=2)


def cogvideo_fifo_mp_v2(pipe_list, base_output, seed: int = 0, progress=None, **kwargs):
    """Drop-in for the reference sampler entry point (cogvideo_sampling_mp_fifo.py:27-395): same arguments
    (`pipe_list`, `base_output`) and the same return value — `(orig_video, video, cache_video)` or a
    `CogVideoXPipelineOutput` when `base_output.return_dict`; `output_type == "latent"` returns latents (:387-390).

    Process model: the reference spawns one worker per entry of `pipe_list` and feeds them through queues; here every
    GPU already runs this same function in its own persistent process (torchrun), `pipe_list` holds that process's
    pipeline, and the window rank -> process assignment, lookahead write-back and boundary exchange are `run_fifo`'s."""
    import torch.distributed as dist
    from .pipeline import CogVideoXPipelineOutput
    pipe = pipe_list[0]
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_available() and dist.is_initialized() else (1, 0)
    sp = dict(base_output.sampling_params or {})
    if sp.get("use_sliding_window_embedding"):
        raise NotImplementedError("use_sliding_window_embedding calls an undefined function in the reference (:150)")
    if base_output.use_dynamic_cfg:
        raise NotImplementedError("use_dynamic_cfg is off in the FIFO stage of both shipped configs (infer_cogvideo_mp_fifo.py:316)")
    if base_output.use_separate_guidance and base_output.prompt_embeds.shape[0] != 3:
        raise ValueError("use_separate_guidance expects the three-branch bundle [uncond_txt, uncond_img, txt_img]")
    if base_output.cache_idx:
        raise NotImplementedError("cache_idx (debug dumps of intermediate queue slots) is empty in the shipped configs")
    nf, T = base_output.nf_per_chunk, base_output.num_inference_steps
    dev = base_output.fifo_latents.device
    schedule = FifoSchedule(base_output.num_frames, [int(t) for t in base_output.timesteps], nf,
                            sp.get("num_partitions", 4), sp.get("use_adaptive_padding", True))
    queue = FifoQueue(base_output.fifo_latents.to(torch.bfloat16), base_output.fifo_old_pred_original_sample, schedule.r_nf)
    vip = None
    if base_output.image_embeddings is not None:
        vip = VipBook(base_output.vip_image_rotary_grid, base_output.vip_condition_rotary_grid, base_output.image_embeddings,
                      nf, base_output.vip_nf_per_chunk, T, base_output.video_ipadapter_start_frame_idx)
    # `window_noise(window) -> (n1, n2)` / `shift_noise(shape)`: supply the noise the reference would draw from its global RNG
    # (tests/test_fifo_stage_gpu.py feeds the draws of the reference-sampler golden); default: the (seed, iteration, rank) streams
    step_fn = make_step_fn(pipe.transformer, pipe.scheduler, base_output.prompt_embeds, base_output.image_rotary_emb, vip,
                           base_output.guidance_scale if base_output.do_classifier_free_guidance else 1.0,
                           noise_override=kwargs.get("window_noise"),
                           guidance_scale_img=base_output.guidance_scale_img if base_output.use_separate_guidance else None)
    shift_latents = make_shift_fn(pipe.scheduler, noise_override=kwargs.get("shift_noise"))

    def shift(q, gen):
        shift_latents(q, gen)
        if vip is not None:
            vip.shift()                                                   # :351-357

    ckpt = None
    if kwargs.get("checkpoint_dir"):   # schema extension: restartable FIFO stage (FifoCheckpoint)
        probe = [p.detach().reshape(-1)[:4096] for p in (pipe.transformer.proj_out.weight,
                                                          pipe.transformer.transformer_blocks[0].attn1.to_q.weight)]
        vip_scale = [getattr(b.attn1.processor, "scale", None) for b in pipe.transformer.transformer_blocks[:1]]
        fp = run_fingerprint(seed, world, base_output.num_frames, [int(t) for t in base_output.timesteps], nf,
                             base_output.guidance_scale, sp.get("num_partitions", 4), sp.get("use_adaptive_padding", True),
                             vip_scale, base_output.fifo_latents, base_output.prompt_embeds,
                             base_output.image_embeddings if base_output.image_embeddings is not None else "no-vip", *probe)
        ckpt = FifoCheckpoint(kwargs["checkpoint_dir"], kwargs.get("checkpoint_every", 10), rank, fingerprint=fp)

    def fast_forward(k):               # the condensed-token bookkeeping advances once per iteration (:351-357)
        if vip is not None:
            for _ in range(k):
                vip.shift()

    # `streaming_decode` (schema extension): decode every chunk under the loop as soon as it is complete
    streamer = None
    if kwargs.get("streaming_decode") and base_output.output_type != "latent":
        streamer = StreamingDecoder(pipe, nf, T - nf, rank, world)
    # `ramp_sharding` (default on with >= 2 ranks): idle ranks join the active windows of the ramp-up iterations through the
    # sequence-parallel forward (RampSharding).  The groups are created once per pipeline (collective).
    ramp = None
    if world > 1 and kwargs.get("ramp_sharding", True):
        ramp = getattr(pipe, "_tg_ramp", None)
        if ramp is None or ramp.world != world:
            tr = pipe.transformer
            ramp = RampSharding(world, rank, heads=tr.config.num_attention_heads, enter=tr.enable_sequence_parallel,
                                leave=tr.disable_sequence_parallel)
            try:
                pipe._tg_ramp = ramp
            except AttributeError:
                pass
    emitted = run_fifo(schedule, queue, step_fn, shift, seed=seed, rank=rank, world=world, progress=progress,
                       checkpoint=ckpt, on_resume=fast_forward, on_emit=None if streamer is None else streamer.on_emit,
                       ramp=ramp)
    latents = torch.cat(emitted[(T - nf):], dim=1).contiguous()           # :367 (slot r_nf belongs to window rank 0 -> process 0)
    if world > 1 and streamer is None:
        dist.broadcast(latents, src=0)
    orig_latents = base_output.orig_latents
    if base_output.output_type == "latent":
        video, orig_video = latents, orig_latents
    elif streamer is not None:
        decoded = streamer.finish()
        video = None if decoded is None else pipe.video_processor.postprocess_video(video=decoded, output_type=base_output.output_type)
        orig_video = None
        if rank == 0:
            orig_video = pipe.video_processor.postprocess_video(video=pipe.decode_latents(orig_latents, nf),
                                                                output_type=base_output.output_type)
        if kwargs.get("stream_log") is not None:
            kwargs["stream_log"].update(streamer.ready_it)
    else:
        # decode is clip-parallel with no communication inside a clip (the conv cache is cleared per 13-frame chunk,
        # autoencoder_kl_cogvideox.py:1157): chunk c is decoded by process c % P and collected on process 0.
        decoded = decode_latents_parallel(pipe, latents, nf, rank, world)
        video = None if decoded is None else pipe.video_processor.postprocess_video(video=decoded, output_type=base_output.output_type)
        orig_video = None
        if rank == 0:
            orig_video = pipe.video_processor.postprocess_video(video=pipe.decode_latents(orig_latents, nf),
                                                                output_type=base_output.output_type)
    if not base_output.return_dict:
        return (orig_video, video, [])
    return CogVideoXPipelineOutput(frames=video, orig_frames=orig_video, cache_frames=[])
