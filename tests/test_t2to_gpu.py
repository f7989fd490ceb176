"""T2To stage on the GPU against the reference (tests/golden/t2to_tiny.pt): the patch_size = 1 DiT forward (K = 16 -> 64 and
N = 16 -> 64 zero padding of the patch / output projections, RoPE dims 52/6/6) against the reference module's fp32 output,
and the whole LongVGenCogVideoXPipeline.__call__ (dynamic CFG, 6 DPM steps, un-normalise + PCA inverse) against the
reference pipeline's own bf16 run with the same generator."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "t2to_tiny.pt"), weights_only=False)


@pytest.fixture(scope="module")
def model(gold):
    from oracle.synth import synth_state_dict
    from tokensgen_b200.transformer import CogVideoXTransformer3DModel
    m = CogVideoXTransformer3DModel(**gold["config"]["dit"])
    m.load_state_dict(synth_state_dict(gold["meta"]["shapes"], seed=gold["seeds"]["dit"]), strict=True)
    return m.to("cuda", torch.bfloat16).eval()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def test_patch1_forward_against_the_reference_module(gold, model):
    i = gold["inputs"]
    text = torch.cat([i["negative_prompt_embeds"], i["prompt_embeds"]]).cuda()
    rope = (gold["rope_cos"].cuda(), gold["rope_sin"].cuda())
    with torch.no_grad():
        y = model(hidden_states=i["latents"].cuda(), encoder_hidden_states=text, timestep=i["timestep"].cuda(),
                  image_rotary_emb=rope, return_dict=False)[0]
    err, ref_bf16 = rel(y, gold["forward_f32"]), rel(gold["forward_bf16"], gold["forward_f32"])
    print(f"patch_size=1 DiT forward: CUDA vs reference fp32 rel_l2 {err:.3e} (reference bf16 vs its fp32: {ref_bf16:.3e})")
    assert tuple(y.shape) == tuple(gold["forward_f32"].shape)
    assert err < 1e-2


def test_t2to_pipeline_against_the_reference_pipeline(gold, model):
    import sys
    sys.path.insert(0, ROOT)
    from pca import PCA
    from oracle.make_goldens import t2to_tiny_stats
    from tokensgen_b200.pipeline_t2to import LongVGenCogVideoXPipeline
    from tokensgen_b200.scheduler import CogVideoXDPMScheduler
    pipe = LongVGenCogVideoXPipeline(None, None, model, CogVideoXDPMScheduler.cogvideox_5b()).to("cuda")
    mean, std, comp, pmean = t2to_tiny_stats()
    pca = PCA(None)
    pca.register_buffer("mean_", pmean)
    pca.register_buffer("components_", comp)
    i = gold["inputs"]
    out = pipe(prompt_embeds=i["prompt_embeds"], negative_prompt_embeds=i["negative_prompt_embeds"],
               generator=torch.Generator().manual_seed(gold["seeds"]["call"]), longvgen_mean=mean, longvgen_std=std,
               longvgen_pca=pca, **gold["config"]["call"])
    out = out[0] if isinstance(out, tuple) else out.frames
    ref = gold["frames"]
    assert tuple(out.shape) == tuple(ref.shape) and out.dtype == ref.dtype
    err = rel(out, ref)
    print(f"T2To pipeline (6 steps, dynamic CFG, PCA tail): CUDA vs the reference pipeline's bf16 CPU run rel_l2 {err:.3e}")
    assert err < 3e-2
