"""Drop-in import paths (north_star "keeping the longvgen.pipeline / infer_cogvideo_mp_fifo.py entry
points"): the `from longvgen...` import block of the REFERENCE's entry point (infer_cogvideo_mp_fifo.py:62-71) is executed
verbatim against this repository and must resolve to the tokensgen_b200 mirrors.  Runs in a subprocess so the alias never
meets the oracle's import of the real reference tree (both are namespace packages called `longvgen`)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CLI = "/root/reference/infer_cogvideo_mp_fifo.py"

# the block as it stands in the reference (used when the reference tree is not on this machine, e.g. the GPU box)
RESTATED = """
from longvgen.models import CogVideoXTransformer3DModel 
from longvgen.pipeline import (
    MPFIFOVideoIPAdapterCogVideoXPipeline,
    LongVGenCogVideoXPipeline, 
)
from longvgen.schedulers import CogVideoXDPMScheduler
from longvgen.data.long_video import load_video
from longvgen.video_ipadapter import Resampler
from longvgen.fifo_sampling import cogvideo_fifo_mp_v2
"""


def _reference_import_block() -> str:
    if not os.path.exists(REF_CLI):
        return RESTATED
    src = open(REF_CLI).read()
    blocks = re.findall(r"^from longvgen[^\n]*\([^)]*\)|^from longvgen[^\n]*$", src, flags=re.M)
    assert len(blocks) == 6, blocks
    return "\n".join(blocks)


def test_reference_entry_point_imports_resolve_to_the_mirrors():
    block = _reference_import_block()
    if os.path.exists(REF_CLI):   # the restated copy is the reference's block
        norm = lambda s: re.sub(r"\s+", "", s)
        assert norm(block) == norm(RESTATED)
    prog = block + """
import inspect
names = dict(CogVideoXTransformer3DModel=CogVideoXTransformer3DModel, MPFIFOVideoIPAdapterCogVideoXPipeline=MPFIFOVideoIPAdapterCogVideoXPipeline,
             LongVGenCogVideoXPipeline=LongVGenCogVideoXPipeline, CogVideoXDPMScheduler=CogVideoXDPMScheduler, load_video=load_video,
             Resampler=Resampler, cogvideo_fifo_mp_v2=cogvideo_fifo_mp_v2)
for k, v in names.items():
    assert v.__module__.startswith("tokensgen_b200."), (k, v.__module__)
# the module-level paths the reference's own files import from each other
import longvgen.models.cogvideox_transformer_3d as a, longvgen.models.attention_processor as b, longvgen.models.embeddings as c
import longvgen.models.normalization as d, longvgen.models.autoencoder_kl_cogvideox as e, longvgen.schedulers.scheduling_dpm_cogvideox as f
import longvgen.pipeline.pipeline_cogvideox_mp_fifo as g, longvgen.pipeline.pipeline_cogvideox_t2to as h
import longvgen.fifo_sampling.cogvideo_sampling_mp_fifo as i, longvgen.video_ipadapter.resampler as j
assert b.VideoIPAdapterCogVideoXAttnProcessor2_0.__name__ == "VideoIPAdapterCogVideoXAttnProcessor2_0"
assert c.get_3d_rotary_pos_embed_v2 and d.CogVideoXLayerNormZero and e.AutoencoderKLCogVideoX and g.FIFOCogVideoXPipelineOutput
# call signatures a script written against the reference relies on
sig = inspect.signature(MPFIFOVideoIPAdapterCogVideoXPipeline.__call__).parameters
for kw in ("prompt", "frames", "image_embeddings", "num_videos_per_prompt", "num_inference_steps", "num_frames_per_chunk", "max_num_chunks",
           "max_num_chunks_wo_fifo", "max_num_chunks_w_fifo", "use_dynamic_cfg", "use_separate_guidance", "guidance_scale",
           "guidance_scale_img", "generator", "vip_scale", "sampling_mode", "sampling_params", "cache_idx",
           "video_ipadapter_start_frame_idx", "return_dict"):
    assert kw in sig, kw
sig = inspect.signature(LongVGenCogVideoXPipeline.__call__).parameters
for kw in ("prompt", "height", "width", "num_frames_per_chunk", "num_chunks", "use_dynamic_cfg", "guidance_scale", "generator",
           "longvgen_mean", "longvgen_std", "longvgen_pca"):
    assert kw in sig, kw
assert hasattr(MPFIFOVideoIPAdapterCogVideoXPipeline, "preprare_for_fifo") and hasattr(MPFIFOVideoIPAdapterCogVideoXPipeline, "decode_latents")
print("DROPIN_OK")
"""
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", prog], cwd="/tmp", env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_launch_plan():
    """infer_cogvideo_mp_fifo.py:191,384-389 of the reference: ONE `python` command uses every GPU in CUDA_VISIBLE_DEVICES.
    Here that command re-launches itself as one rank per visible GPU unless it already runs under a launcher."""
    sys.path.insert(0, ROOT)
    import infer_cogvideo_mp_fifo as cli
    assert cli.launch_plan({}, 8) == 8                                  # plain `python ...` on an 8-GPU box: 8 ranks
    assert cli.launch_plan({}, 1) == 0 and cli.launch_plan({}, 0) == 0   # one GPU (or none): run in this process
    assert cli.launch_plan({"WORLD_SIZE": "8", "RANK": "3"}, 8) == 0     # already under torchrun
    assert cli.launch_plan({"TG_SINGLE_PROCESS": "1"}, 8) == 0           # opt out
    assert cli.launch_plan({"TG_NPROC": "2"}, 8) == 2                    # cap
