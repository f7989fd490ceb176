"""GPU parity: CUDA path (through the C-ABI) vs the oracle and vs golden vectors from the unmodified reference.

Tolerances (bf16 kernels vs fp32 oracle on the same bf16-rounded inputs/weights):
  relative L2 <= 5e-3 per op, <= 1e-2 per block / small model.  Index maps and the scheduler chain are bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TINY = dict(heads=4, head_dim=64, layers=2, time_dim=128, text_dim=128, in_ch=16, out_ch=16, patch=2, vip_dim=128)
VIP_KW = dict(length=12, func_type="1", scale=[0.6],
              resampler_params=dict(output_dim=128, num_height_queries=2, num_width_queries=3, num_temporal_queries=1))


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def tiny_model(use_vip, sd):
    from tokensgen_b200.transformer import CogVideoXTransformer3DModel
    m = CogVideoXTransformer3DModel(num_attention_heads=4, attention_head_dim=64, in_channels=16, out_channels=16,
                                    time_embed_dim=128, text_embed_dim=128, num_layers=2, patch_size=2,
                                    use_rotary_positional_embeddings=True, attention_bias=True)
    if use_vip:
        m.set_vip_layers(None, **VIP_KW)
    m.load_state_dict(sd, strict=True)
    return m.to("cuda", torch.bfloat16).eval()


@pytest.mark.parametrize("use_vip", [True, False])
@pytest.mark.parametrize("per_frame", [True, False])
def test_tiny_model_vs_reference_golden_and_oracle(golden_dir, use_vip, per_frame):
    from oracle import dit as odit
    from oracle.synth import dit_shapes, synth_state_dict
    g = torch.load(os.path.join(golden_dir, "dit_tiny.pt"))
    tag = ("vip" if use_vip else "plain") + ("_pf" if per_frame else "_ps")
    lat, text, vip, ts = g[tag + "_inputs"]
    sd = synth_state_dict(dit_shapes(use_vip=use_vip, **TINY), 1234)
    m = tiny_model(use_vip, sd)
    with torch.no_grad():
        y = m(lat.cuda(), text.cuda(), ts.cuda(), vip_encoder_hidden_states=vip.cuda() if use_vip else None,
              image_rotary_emb=g["rope"], vip_image_rotary_emb=g["img_rope"] if use_vip else None,
              vip_condition_rotary_emb=g["cond_rope"] if use_vip else None, return_dict=False)[0]
    torch.cuda.synchronize()
    assert y.shape == lat.shape and y.dtype == torch.bfloat16
    ref32 = g[tag + "_f32"]                      # unmodified reference, fp32
    err = rel_l2(y, ref32)
    band = rel_l2(g[tag + "_bf16"], ref32)       # where the reference's own bf16 run sits
    print(f"{tag}: cuda vs ref fp32 {err:.3e}; reference bf16 vs fp32 {band:.3e}")
    assert err < 1e-2
    cfg = odit.DitConfig(num_attention_heads=4, attention_head_dim=64, time_embed_dim=128, text_embed_dim=128, num_layers=2,
                         vip_length=12, vip_embed_dim=128, use_vip=use_vip)
    o = odit.dit_forward(sd, cfg, lat, text, ts, vip, g["rope"], g["img_rope"], g["cond_rope"], torch.float32)
    assert rel_l2(y, o) < 1e-2


@pytest.mark.parametrize("use_vip", [True, False])
def test_tiny_block_vs_reference_golden(golden_dir, use_vip):
    from oracle.synth import dit_shapes, synth_state_dict
    g = torch.load(os.path.join(golden_dir, "dit_tiny.pt"))
    hid, enc, temb, h_ref, e_ref = g[("vip" if use_vip else "plain") + "_block"]
    m = tiny_model(use_vip, synth_state_dict(dit_shapes(use_vip=use_vip, **TINY), 1234))
    blk = m.transformer_blocks[1]
    with torch.no_grad():
        h, e = blk(hid.cuda().bfloat16(), enc.cuda().bfloat16(), temb.cuda().bfloat16(), g["rope"],
                   g["img_rope"] if use_vip else None, g["cond_rope"] if use_vip else None)
    assert rel_l2(h, h_ref) < 1e-2 and rel_l2(e, e_ref) < 1e-2


def test_processor_plugin_call_vs_oracle(golden_dir):
    """The diffusers-style plugin boundary: Attention.forward -> VideoIPAdapterCogVideoXAttnProcessor2_0.__call__."""
    from oracle import dit as odit
    from oracle.synth import dit_shapes, synth_state_dict
    g = torch.load(os.path.join(golden_dir, "dit_tiny.pt"))
    sd = synth_state_dict(dit_shapes(use_vip=True, **TINY), 1234)
    m = tiny_model(True, sd)
    attn = m.transformer_blocks[0].attn1
    gen = torch.Generator().manual_seed(3)
    hid = torch.randn(2, 72, 256, generator=gen).bfloat16()
    enc = torch.randn(2, 22, 256, generator=gen).bfloat16()
    with torch.no_grad():
        h, e = attn(hidden_states=hid.cuda(), encoder_hidden_states=enc.cuda(), image_rotary_emb=g["rope"],
                    vip_image_rotary_emb=g["img_rope"], vip_condition_rotary_emb=g["cond_rope"], unknown_kwarg=1)
    cfg = odit.DitConfig(num_attention_heads=4, attention_head_dim=64, vip_length=12, vip_scale=0.6)
    ho, eo = odit.vip_attention(sd, "transformer_blocks.0.attn1", cfg, hid.float(), enc.float(), g["rope"], g["img_rope"],
                                g["cond_rope"], torch.float32)
    assert h.shape == ho.shape and e.shape == eo.shape
    assert rel_l2(h, ho) < 5e-3 and rel_l2(e, eo) < 5e-3


def test_full_size_block_vs_oracle():
    """BASELINE config 1 shape on the GPU: one CogVideoX-5b block + VIP, 13x30x45 window, B=1, vs the fp32 oracle on CPU."""
    from oracle import dit as odit
    from oracle import rope as orope
    from oracle.synth import dit_shapes, synth_state_dict
    from tokensgen_b200.transformer import CogVideoXBlock
    shapes = {k[len("transformer_blocks.0."):]: v for k, v in
              dit_shapes(48, 64, 1, 512, 4096, 16, 16, 2, 3072, True).items() if k.startswith("transformer_blocks.0.")}
    sd = synth_state_dict(shapes, 77)
    blk = CogVideoXBlock(dim=3072, num_attention_heads=48, attention_head_dim=64, time_embed_dim=512, attention_bias=True)
    blk.set_vip_layers(length=480, func_type="1", scale=[0.6])
    blk.load_state_dict(sd, strict=True)
    blk = blk.to("cuda", torch.bfloat16).eval()
    gen = torch.Generator().manual_seed(42)
    hid = torch.randn(1, 17550, 3072, generator=gen).bfloat16()
    enc = torch.randn(1, 706, 3072, generator=gen).bfloat16()
    temb = torch.randn(1, 13, 512, generator=gen).bfloat16()
    rope = orope.window_rope(64, 13, 30, 45)
    img = orope.rope_3d_from_grids(64, np.arange(13, dtype=np.float32) + 45, np.arange(30, dtype=np.float32),
                                   np.arange(45, dtype=np.float32))
    cond = orope.rope_3d_from_grids(64, np.array([1000, 1003.25, 1006.5, 1009.75, 1013], dtype=np.float32),
                                    np.linspace(0, 30, 8, endpoint=False, dtype=np.float32),
                                    np.linspace(0, 45, 12, endpoint=False, dtype=np.float32))
    with torch.no_grad():
        h, e = blk(hid.cuda(), enc.cuda(), temb.cuda(), rope, img, cond)
    torch.cuda.synchronize()
    cfg = odit.DitConfig()
    sd0 = {"transformer_blocks.0." + k: v for k, v in sd.items()}
    torch.set_num_threads(os.cpu_count() or 8)
    ho, eo = odit.block_forward(sd0, "transformer_blocks.0", cfg, hid.float(), enc.float(), temb.float(), rope, img, cond,
                                torch.float32)
    eh, ee = rel_l2(h, ho), rel_l2(e, eo)
    print(f"full-size block: hidden rel_l2 {eh:.3e}, encoder rel_l2 {ee:.3e}")
    assert eh < 1e-2 and ee < 1e-2


def test_dpm_step_kernel_bit_exact_vs_reference_chain(golden_dir):
    """tg_cfg_dpm_step on the golden step cases of the real CogVideoXDPMScheduler.

    The reference's result depends on the device it runs on: PyTorch's CUDA kernels keep a 0-dim fp64 scalar in fp32 when
    multiplying a bf16 tensor, the CPU kernels round it to bf16 first (oracle/dpm.py:_smul).  The kernel reproduces the
    CUDA behaviour (the reference runs on GPUs), so it is compared bit-for-bit with (a) the oracle in "cuda" semantics and
    (b) the same chain evaluated op by op with torch on this GPU."""
    from oracle import dpm as odpm
    from tokensgen_b200 import _ext as E
    g = torch.load(os.path.join(golden_dir, "dpm.pt"))
    tb = odpm.DpmTables()
    n = 0
    for c in g["step_cases"]:
        if c["model_output"].dtype != torch.bfloat16:
            continue
        coefs = tb.coefficients(c["t"], c["prev_t"], c["back"])
        second = c["old"] is not None and c["prev_t"] >= 0
        coef = torch.tensor([[float(v) for v in coefs] + [1.0 if second else 0.0]], dtype=torch.float64).float().cuda()
        shp = c["sample"].shape
        flat = lambda t: None if t is None else t.reshape(1, -1).contiguous().cuda()
        prev, x0 = E.cfg_dpm_step(flat(c["model_output"]).unsqueeze(0), flat(c["sample"]), flat(c["old"]) if second else None,
                                  flat(c["n1"]), flat(c["n2"]), coef, 0.0, E.DPM_BF16_CHAIN)
        p_o, x0_o = odpm.step(tb, c["model_output"], c["old"], c["t"], c["prev_t"], c["back"], c["sample"], c["n1"], c["n2"],
                              device_semantics="cuda")
        assert torch.equal(prev.cpu().view(shp), p_o) and torch.equal(x0.cpu().view(shp), x0_o)
        # (b) torch on the GPU, written exactly like scheduling_dpm_cogvideox.py:439-463 (scalar first)
        sa, sb, m0, m1, m2, m3, mn = coefs
        smp, mo = c["sample"].cuda(), c["model_output"].cuda()
        x0_t = sa * smp - sb * mo
        if second:
            d = m2 * x0_t - m3 * c["old"].cuda()
            p_t = m0 * smp - m1 * d + mn * c["n2"].cuda()
        else:
            p_t = m0 * smp - m1 * x0_t + mn * c["n1"].cuda()
        assert torch.equal(prev.view(shp), p_t) and torch.equal(x0.view(shp), x0_t)
        n += 1
    assert n >= 5


def test_window_step_vs_oracle_loop():
    """CFG + 13 per-frame steps in one launch == the reference worker's Python loop (oracle.dpm.window_step_bf16)."""
    from oracle import dpm as odpm
    from oracle import fifo as ofifo
    from tokensgen_b200 import _ext as E
    from tokensgen_b200.scheduler import CogVideoXDPMScheduler
    tb = odpm.DpmTables()
    ts = tb.trailing_timesteps(52)
    t_tab, prev_tab, next_tab = ofifo.fifo_timestep_tables(ts)
    gen = torch.Generator().manual_seed(11)
    F, shp = 13, (16, 60, 90)
    for start in (45, 39, 6, 0):
        t, pt, nt = t_tab[start:start + F], prev_tab[start:start + F], next_tab[start:start + F]
        npred = torch.randn(2, F, *shp, generator=gen).bfloat16()
        lat = torch.randn(1, F, *shp, generator=gen).bfloat16()
        # a slot without a history timestep (next_t <= 0: the freshly re-noised tail) never carries an x0 history
        old = [torch.randn(1, 1, *shp, generator=gen).bfloat16() if (j % 5 != 0 and nt[j] > 0) else None for j in range(F)]
        n1 = torch.randn(1, F, *shp, generator=gen).bfloat16()
        n2 = torch.randn(1, F, *shp, generator=gen).bfloat16()
        ref_lat, ref_x0 = odpm.window_step_bf16(tb, npred, 6.0, lat, old, t, pt, nt, n1, n2)
        sch = CogVideoXDPMScheduler.cogvideox_5b()
        sch.set_timesteps(52)
        out_lat, out_x0 = sch.window_step(npred.cuda(), lat.cuda(), [None if o is None else o.cuda() for o in old],
                                          t, pt, nt, 6.0, noise=(n1.cuda(), n2.cuda()))
        assert torch.equal(out_lat.cpu(), ref_lat)
        for j in range(F):
            assert torch.equal(out_x0[j].cpu(), ref_x0[j]), (start, j)


def test_dit_forward_is_cuda_graph_capturable_and_replays_bit_identically(golden_dir):
    """INTEGRATION.md's claim about the C ABI (no allocation, no synchronisation, every launch on the caller's stream): one
    whole DiT forward with the video-IP-adapter — time embedding, AdaLN table GEMM, patchify, 2 blocks of LN-modulate / fused
    QKV+RoPE GEMM / three attentions / gated-residual and GELU GEMMs, final double LayerNorm, unpatchify — is captured in a
    CUDA graph and replayed with NEW inputs copied into the captured tensors: every replay equals the eager result bit for
    bit."""
    from oracle.synth import dit_shapes, synth_state_dict
    g = torch.load(os.path.join(golden_dir, "dit_tiny.pt"))
    sd = synth_state_dict(dit_shapes(use_vip=True, **TINY), 1234)
    m = tiny_model(True, sd)
    dev = torch.device("cuda")
    on = lambda pair: tuple(t.to(dev) for t in pair)
    rope, img, cond = on(g["rope"]), on(g["img_rope"]), on(g["cond_rope"])
    inputs = {k: [t.cuda() for t in g[f"vip_{k}_inputs"]] for k in ("pf",)}
    lat, text, vip, ts = inputs["pf"]
    s_lat, s_text, s_vip, s_ts = lat.clone(), text.clone(), vip.clone(), ts.clone()

    def fwd():
        return m(s_lat, s_text, s_ts, vip_encoder_hidden_states=s_vip, image_rotary_emb=rope, vip_image_rotary_emb=img,
                 vip_condition_rotary_emb=cond, return_dict=False)[0]

    with torch.no_grad():
        eager = fwd().clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):        # warm-up on the capture stream (workspaces, packed weights, tensor maps)
            fwd()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fwd()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eager)
        # new inputs through the same graph
        gen = torch.Generator(device="cuda").manual_seed(3)
        lat2 = torch.randn(lat.shape, generator=gen, device="cuda").bfloat16()
        s_lat.copy_(lat2)
        graph.replay()
        torch.cuda.synchronize()
        replayed = out.clone()
        s_lat.copy_(lat2)
        eager2 = fwd()
        torch.cuda.synchronize()
        assert torch.equal(replayed, eager2) and not torch.equal(replayed, eager)


def test_a_foreign_attention_processor_is_not_silently_ignored(golden_dir):
    """`Attention.set_processor` (attention_processor.py:423-441) with a processor the fused engine does not know must not
    be bypassed by the model's fast path: the forward raises and says how to run it."""
    from oracle.synth import dit_shapes, synth_state_dict
    from tokensgen_b200 import _ext as E
    g = torch.load(os.path.join(golden_dir, "dit_tiny.pt"))
    m = tiny_model(False, synth_state_dict(dit_shapes(use_vip=False, **TINY), 1234))
    lat, text, vip, ts = g["plain_ps_inputs"]

    class Mine:
        def __call__(self, attn, hidden_states, encoder_hidden_states, attention_mask=None, image_rotary_emb=None):
            return hidden_states, encoder_hidden_states

    m.transformer_blocks[1].attn1.set_processor(Mine())
    with pytest.raises(E.TokensGenError, match="processor"):
        with torch.no_grad():
            m(lat.cuda(), text.cuda(), ts.cuda(), image_rotary_emb=g["rope"], return_dict=False)
    # the module-level plugin API still dispatches to it (Attention.forward filters kwargs by the processor's signature)
    h = torch.zeros(1, 4, 256, device="cuda", dtype=torch.bfloat16)
    e = torch.zeros(1, 2, 256, device="cuda", dtype=torch.bfloat16)
    oh, oe = m.transformer_blocks[1].attn1(h, encoder_hidden_states=e, image_rotary_emb=g["rope"], vip_image_rotary_emb=None)
    assert oh is h and oe is e
