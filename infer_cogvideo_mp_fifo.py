#!/usr/bin/env python
"""TokensGen long-video inference on B200 — the reference's entry point (infer_cogvideo_mp_fifo.py:186-389) on the
tokensgen_b200 mirrors.  Same flag (`--config <yaml>`), same yaml schema (config/infer/{edit,gen}.yaml), same outputs
(`<name>_{source,embeds,orig,fifo}_<prompt[:20]>` under `<output_dir>/<prefix>_<timestamp>/`).

    python infer_cogvideo_mp_fifo.py --config config/infer/edit.yaml     # every GPU in CUDA_VISIBLE_DEVICES, like the reference
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 infer_cogvideo_mp_fifo.py --config ...     # or under a launcher

The plain `python` command is the reference's invocation (its infer_cogvideo_mp_fifo.py:191 takes the GPU list from
CUDA_VISIBLE_DEVICES and :384-389 runs main() once): with more than one visible GPU and no launcher environment it starts
one rank per GPU itself (`launch_plan` / `torch.multiprocessing.spawn`, 127.0.0.1 rendezvous).

Process model: the reference builds one pipeline per visible GPU inside one process and spawns a worker per GPU for every
video; here every GPU runs this script as a persistent rank (torchrun), loads its own copy of the weights in parallel,
the stages the reference runs serially on one GPU (T2To, conditioning encodes, the 52-step base clip) run on rank 0 — or,
with more than two ranks / `sequence_parallel: true`, on ALL ranks with every DiT forward sharded over them
(tokensgen_b200/seqpar.py) and the conditioning chunks encoded one per rank — then all ranks run the FIFO stage with NCCL
boundary-frame exchange and decode clip-parallel.
"""
from __future__ import annotations

import argparse
import copy
import datetime
import json
import os

import torch
import torch.distributed as dist

from tokensgen_b200 import config as cfgmod
from tokensgen_b200.fifo import broadcast_base_output, cogvideo_fifo_mp_v2
from tokensgen_b200.pipeline import MPFIFOVideoIPAdapterCogVideoXPipeline
from tokensgen_b200.pipeline_t2to import LongVGenCogVideoXPipeline
from tokensgen_b200.resampler import Resampler
from tokensgen_b200.scheduler import CogVideoXDPMScheduler
from tokensgen_b200.transformer import CogVideoXTransformer3DModel
from tokensgen_b200.video_io import export_to_video, load_video


def create_output_folders(output_dir, config, prefix="longvgen"):
    now = datetime.datetime.now().strftime("%Y-%m-%dT%H-%M-%S")
    out_dir = os.path.join(output_dir, f"{prefix}_{now}")
    os.makedirs(out_dir, exist_ok=True)
    cfgmod.save(config, os.path.join(out_dir, "config.yaml"))
    return out_dir


def init_pipeline(gpu_id, args, dtype):
    """infer_cogvideo_mp_fifo.py:138-183."""
    device = torch.device(f"cuda:{gpu_id}")
    # `device=`: the checkpoint shards are read straight to this rank's GPU into a meta-constructed module tree
    # (tokensgen_b200/loading.py) — no CPU-side random init, fp32 copy or staging of the 11 GB transformer
    transformer = CogVideoXTransformer3DModel.from_pretrained(args.pretrained_model_name_or_path, subfolder="transformer",
                                                              torch_dtype=torch.bfloat16, device=device).to(device)
    resampler = None
    if args.use_vip:
        vip_params = args.video_ipadapter_params
        vip_path = args.pretrained_resampler_name_or_path
        transformer.set_vip_layers(vip_path, **vip_params)
        transformer = transformer.to(dtype)
        resampler = Resampler.from_pretrained(vip_path, subfolder="resampler", torch_dtype=dtype, device=device).to(device)
        resampler.set_pca(args.get("longvgen_pca", None), device=device)
    pipe = MPFIFOVideoIPAdapterCogVideoXPipeline.from_pretrained(args.pretrained_model_name_or_path, transformer=transformer,
                                                                 resampler=resampler, torch_dtype=dtype, device=device)
    pipe.scheduler = CogVideoXDPMScheduler.from_config(pipe.scheduler.config, timestep_spacing="trailing")
    pipe.to(device)
    pipe.vae.enable_slicing()
    if args.get("vae_tiling", True):   # the reference always tiles (3 x 3 overlapping tiles, x1.4 work); `vae_tiling: false` is a
        pipe.vae.enable_tiling()        # schema extension: one B200 holds an untiled 49-frame 480 x 720 clip (2x faster coding,
    return pipe                         # results then differ from the tiled ones by the tile-blend seams)


def main(args):
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if args.dtype != "bf16":
        raise SystemExit("tokensgen_b200 computes in bf16 (dtype: 'bf16' in both shipped configs)")
    dtype = torch.bfloat16
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        args.output_dir = create_output_folders(args.output_dir, args, args.name_prefix)
    pipe = init_pipeline(local, args, dtype)              # every rank loads its own weights, in parallel
    # with >= 2 ranks the serial base clip runs CFG-parallel on ranks 0 and 1 (one guidance branch each; `cfg_parallel: false`
    # in the yaml turns it off).  new_group is collective: every rank creates it.
    # `sequence_parallel: true` (schema extension; default with more than two ranks whose count divides the head count):
    # instead, EVERY rank runs the base clip with each DiT forward sharded over all ranks (tokensgen_b200/seqpar.py).
    heads = pipe.transformer.config.num_attention_heads
    seq_par = world > 1 and bool(args.get("sequence_parallel", world > 2)) and heads % world == 0
    cfg_group = dist.new_group([0, 1]) if world > 1 and args.get("cfg_parallel", True) and not seq_par else None
    pipe_list = [pipe]
    vip_params = args.video_ipadapter_params if args.use_vip else None

    pipe_2nd = None
    if args.use_2nd_stage and (rank == 0 or seq_par):   # sequence-parallel: every rank holds the tokens transformer too
        tokens_transformer = CogVideoXTransformer3DModel.from_pretrained(args.pretrained_2nd_stage_model_name_or_path,
                                                                         subfolder="transformer", torch_dtype=dtype, device=device)
        pipe_2nd = LongVGenCogVideoXPipeline.from_pretrained(args.pretrained_model_name_or_path, transformer=tokens_transformer,
                                                             torch_dtype=dtype, text_encoder=pipe.text_encoder,
                                                             tokenizer=pipe.tokenizer)
        pipe_2nd.scheduler = CogVideoXDPMScheduler.from_config(pipe_2nd.scheduler.config, timestep_spacing="trailing")
        pipe_2nd.to(device)
    t2to_par = seq_par and pipe_2nd is not None and tokens_transformer.config.num_attention_heads % world == 0

    inputs = args.input_config
    public_dps = inputs.pop("public")
    items_json = inputs.pop("input_json") if inputs.get("input_json") is not None else None
    if items_json is not None:
        with open(items_json) as f:
            inputs.update(json.load(f).get("input_config"))
    if rank == 0:
        print(f"***** Running inference *****\n  Num items = {len(inputs)}  ranks = {world}")

    for name, item in inputs.items():
        prompt = item["prompt"]
        dps = copy.deepcopy(public_dps)
        dps.update(item.get("params", {}))
        call = dict(
            prompt=prompt, num_videos_per_prompt=dps.num_videos_per_prompt, num_inference_steps=args.num_inference_steps,
            num_frames_per_chunk=args.num_frames_per_chunk, max_num_chunks=dps.max_num_chunks,
            max_num_chunks_wo_fifo=dps.max_num_chunks_wo_fifo, max_num_chunks_w_fifo=dps.max_num_chunks_w_fifo,
            use_dynamic_cfg=False, use_separate_guidance=dps.get("use_separate_guidance", args.get("use_separate_guidance", False)),
            guidance_scale=args.guidance_scale, guidance_scale_img=args.get("guidance_scale_img", args.guidance_scale),
            vip_scale=vip_params.scale if args.use_vip else 1.0, sampling_mode=args.get("sampling_mode"),
            sampling_params=args.get("sampling_params"), cache_idx=args.get("cache_idx"),
            video_ipadapter_start_frame_idx=vip_params.video_ipadapter_start_frame_idx if args.use_vip else 1000,
            return_dict=False, output_type="uint8")   # reference: "np" float frames; here packed to uint8 on the GPU (same mp4 bytes)
        # two extensions of the yaml schema (absent from the shipped configs): a non-default resolution, and precomputed
        # prompt embeddings for checkpoints without the T5 encoder
        if args.get("height") is not None:
            call.update(height=args.height, width=args.width)
        if args.get("prompt_embeds_path") is not None:
            pe = torch.load(args.prompt_embeds_path, weights_only=True)
            call.update(prompt=None, prompt_embeds=pe["prompt_embeds"], negative_prompt_embeds=pe["negative_prompt_embeds"])
        video = image_embeddings = base_outputs = None
        base_rank = rank == 0 or seq_par or (cfg_group is not None and rank == 1 and not args.use_2nd_stage)
        if base_rank:
            if rank == 0:
                print(f"Processing {name}: [{prompt}]")
            if args.use_vip:
                if args.use_2nd_stage:
                    # T2To stage: serial on rank 0 in the reference; sequence-parallel over all ranks here when the base
                    # clip is (every rank calls with the same seed and gets the same condensed tokens — nothing to ship)
                    if rank == 0 or t2to_par:
                        rp = vip_params.resampler_params
                        pe2 = {k: call[k] for k in ("prompt_embeds", "negative_prompt_embeds") if k in call}
                        image_embeddings = pipe_2nd(
                            prompt=None if pe2 else prompt, height=rp.num_height_queries, width=rp.num_width_queries,
                            num_frames_per_chunk=rp.num_temporal_queries, num_chunks=dps.max_num_chunks, use_dynamic_cfg=True,
                            guidance_scale=args.get("guidance_scale_2nd", args.guidance_scale),
                            generator=torch.Generator().manual_seed(args.seed_2nd), longvgen_mean=args.longvgen_mean,
                            longvgen_std=args.longvgen_std, longvgen_pca=args.longvgen_pca,
                            sequence_parallel_group=dist.group.WORLD if t2to_par else None, **pe2).frames
                    if seq_par and not t2to_par:   # head count of the tokens transformer does not divide: ship rank 0's tokens
                        box = [image_embeddings.cpu() if rank == 0 else None]
                        dist.broadcast_object_list(box, src=0, device=device)
                        image_embeddings = box[0].to(device)
                else:
                    assert item.get("video") is not None
                    video = load_video(item["video"], dps.output_res, args.num_frames_per_chunk, dps.pad_to_fit, dps.sample_fps,
                                       dps.start_t, dps.end_t, dps.max_num_chunks, dps.crop_to_fit)
            # The reference samples the conditioning clips' VAE posterior from the device's global RNG, which it never seeds
            # (its set_seed import is unused on this path), so its edit flow is not reproducible run to run.  Same
            # distribution here, but drawn from a generator seeded by the yaml `seed` on every rank: a job repeats exactly,
            # and 1 rank and N ranks write the same video.
            pipe.vae_posterior_generator = torch.Generator(device=device).manual_seed(int(args.seed))
            base_outputs = pipe(frames=video, image_embeddings=image_embeddings,
                                generator=torch.Generator().manual_seed(args.seed),
                                cfg_parallel_group=cfg_group if not args.use_2nd_stage else None,   # rank 1 has no T2To tokens
                                sequence_parallel_group=dist.group.WORLD if seq_par else None, **call)
            base_outputs.condition_frames = None      # not needed by the FIFO stage; keeps the broadcast small
        else:
            pipe.preprare_for_fifo(**call)
        if not seq_par:     # sequence-parallel: every rank ran the base stage and already holds the identical bundle
            base_outputs = broadcast_base_output(base_outputs, src=0, device=device)
        # `fifo_checkpoint_dir` / `fifo_checkpoint_every` (schema extension): the FIFO stage saves its queue every N iterations
        # and a restarted job resumes from the newest state (the deterministic base stage is simply recomputed)
        ck = args.get("fifo_checkpoint_dir")
        orig_video_frames, video_frames, _ = cogvideo_fifo_mp_v2(
            pipe_list, base_outputs, seed=args.seed, checkpoint_dir=os.path.join(ck, name) if ck else None,
            checkpoint_every=args.get("fifo_checkpoint_every", 10),
            # `streaming_decode` (schema extension, default on): every 13-frame chunk is decoded on a side stream of rank
            # chunk % P as soon as it has left the queue, instead of all chunks after the loop — same frames, no decode tail
            streaming_decode=bool(args.get("streaming_decode", True)),
            # `ramp_sharding` (schema extension, default on): idle ranks join the ramp-up windows (fifo.RampSharding)
            ramp_sharding=bool(args.get("ramp_sharding", True)))
        if rank == 0:
            tag = prompt[:20]
            if video is not None:
                export_to_video((video.clamp(-1, 1) / 2 + 0.5).squeeze(0).permute(0, 2, 3, 1).cpu().numpy(),
                                os.path.join(args.output_dir, f"{name}_source_{tag}.mp4"), fps=dps.output_fps)
            if image_embeddings is not None:
                torch.save(image_embeddings[0].cpu(), os.path.join(args.output_dir, f"{name}_embeds_{tag}.pt"))
            export_to_video(orig_video_frames[0], os.path.join(args.output_dir, f"{name}_orig_{tag}.mp4"), fps=dps.output_fps)
            export_to_video(video_frames[0], os.path.join(args.output_dir, f"{name}_fifo_{tag}.mp4"), fps=dps.output_fps)
    if world > 1:
        dist.destroy_process_group()


def launch_plan(env, visible_gpus: int) -> int:
    """How many ranks this command must start itself: 0 = run main() in this process (one GPU, an existing launcher
    environment, or TG_SINGLE_PROCESS=1), else one rank per visible GPU (TG_NPROC caps it)."""
    if "WORLD_SIZE" in env or "RANK" in env or env.get("TG_SINGLE_PROCESS", "0") == "1":
        return 0
    n = min(visible_gpus, int(env.get("TG_NPROC", visible_gpus)))
    return n if n > 1 else 0


def _rank_entry(rank: int, world: int, port: int, config_path: str):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    main(cfgmod.load(config_path))


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, default="./config/infer/edit.yaml")
    cli = parser.parse_args()
    nproc = launch_plan(os.environ, torch.cuda.device_count())   # device_count honours CUDA_VISIBLE_DEVICES (reference :191)
    if nproc:
        import socket
        import torch.multiprocessing as mp
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        print(f"Running on gpus: {list(range(nproc))}")
        mp.spawn(_rank_entry, args=(nproc, port, cli.config), nprocs=nproc, join=True)   # start method "spawn", like the reference
    else:
        main(cfgmod.load(cli.config))
