"""tg_ln_modulate on its own against the oracle: the three row segments [text | video | vip] of
CogVideoXLayerNormZero / CogVideoXVIPLayerNormZero (normalization.py:443-460, 477-488: LayerNorm, then `* (1 + scale) + shift`
with the video rows using their FRAME's vectors and text / vip rows frame 0), per-frame and per-sample conditioning, the
sequence-parallel row shard, and the fused double LayerNorm of the model tail (norm_final -> norm_out.norm with the AdaLayerNorm
(shift, scale) modulation, cogvideox_transformer_3d.py:736-742, normalization.py:70-92).

Tolerance: the kernel computes in fp32 from the bf16 inputs and rounds ONCE to bf16 -> relative L2 <= 3e-3 against the fp32
oracle on the same bf16-rounded inputs (a single bf16 rounding is ~1.2e-3 rms), max abs error <= 2 bf16 ulps of the row scale."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(d, B, n_text, hw, frames, n_vip, seed):
    from oracle.synth import synth_state_dict
    g = torch.Generator().manual_seed(seed)
    rows = n_text + hw * frames + n_vip
    x = (torch.randn(B, rows, d, generator=g) * 1.7 + 0.3).bfloat16()
    sd = synth_state_dict({"n.norm.weight": [d], "n.norm.bias": [d], "v.norm.weight": [d], "v.norm.bias": [d],
                           "f.weight": [d], "f.bias": [d], "o.norm.weight": [d], "o.norm.bias": [d]}, seed)
    mod = (torch.randn(B * frames, 9 * d, generator=g) * 0.5).bfloat16()     # one AdaLN table row per (batch, frame)
    return x, sd, mod


def _oracle(x, sd, mod, B, n_text, hw, frames, n_vip, d, eps):
    """normalization.py:457-459 (video / text rows), :487 (vip rows), via oracle.dit.layer_norm (fp32)."""
    from oracle import dit as odit
    f32 = torch.float32
    c = lambda i: mod[:, i * d:(i + 1) * d].float().reshape(B, frames, d)
    shift, scale, e_shift, e_scale, v_shift, v_scale = c(0), c(1), c(3), c(4), c(6), c(7)
    rep = lambda t: t.repeat_interleave(hw, dim=1)
    xt, xv, xp = x[:, :n_text].float(), x[:, n_text:n_text + hw * frames].float(), x[:, n_text + hw * frames:].float()
    out = [odit.layer_norm(xt, sd, "n.norm", eps, f32) * (1 + e_scale)[:, [0]] + e_shift[:, [0]],
           odit.layer_norm(xv, sd, "n.norm", eps, f32) * (1 + rep(scale)) + rep(shift)]
    if n_vip:
        out.append(odit.layer_norm(xp, sd, "v.norm", eps, f32) * (1 + v_scale)[:, [0]] + v_shift[:, [0]])
    return torch.cat(out, dim=1)


def _run(x, sd, mod, B, n_text, hw, frames, n_vip, d, eps, row0=0, rows_local=0):
    from tokensgen_b200 import _ext as E
    dev = "cuda"
    rm = E.make_rowmap(n_text, hw * frames, n_vip, hw, frames, row0, rows_local)
    m = mod.to(dev)
    c = lambda i: m[:, i * d:(i + 1) * d]
    shift = E.make_modvec(c(3), c(0), c(6) if n_vip else None)
    scale = E.make_modvec(c(4), c(1), c(7) if n_vip else None)
    w = {k: v.to(dev) for k, v in sd.items()}
    rl = rows_local or rm.rows_per_batch
    xin = x[:, row0:row0 + rl].contiguous().to(dev)
    out = torch.empty_like(xin)
    E.ln_modulate(xin.view(B * rl, d), out.view(B * rl, d), B, rm, w["n.norm.weight"], w["n.norm.bias"],
                  w["v.norm.weight"] if n_vip else None, w["v.norm.bias"] if n_vip else None, eps, shift, scale)
    torch.cuda.synchronize()
    return out.float().cpu()


@pytest.mark.parametrize("d,B,n_text,hw,frames,n_vip", [
    (3072, 2, 226, 50, 13, 480),      # the model width (12 x 8 bf16 per lane), per-frame vectors, vip rows
    (3072, 1, 226, 1350, 1, 0),       # per-sample timestep (base stage), no vip
    (256, 2, 10, 24, 3, 12),          # tiny model
    (1024, 2, 7, 33, 5, 18),          # ragged: rows not a multiple of the 8 rows a CTA handles
])
def test_ln_modulate_against_the_oracle(d, B, n_text, hw, frames, n_vip):
    eps = 1e-5
    x, sd, mod = _case(d, B, n_text, hw, frames, n_vip, seed=d + frames)
    want = _oracle(x, sd, mod, B, n_text, hw, frames, n_vip, d, eps)
    got = _run(x, sd, mod, B, n_text, hw, frames, n_vip, d, eps)
    err = ((got - want).norm() / want.norm()).item()
    worst = ((got - want).abs() / want.abs().clamp_min(1.0)).max().item()
    assert err < 3e-3, err
    assert worst < 2 ** -6, worst
    # every segment on its own (a wrong frame -> vector map would hide inside a global norm)
    for lo, hi in ((0, n_text), (n_text, n_text + hw * frames), (n_text + hw * frames, n_text + hw * frames + n_vip)):
        if hi > lo:
            seg = ((got[:, lo:hi] - want[:, lo:hi]).norm() / want[:, lo:hi].norm()).item()
            assert seg < 3e-3, (lo, hi, seg)
    # frame boundaries of the video rows: first / last row of every frame
    for f in range(frames):
        for r in (n_text + f * hw, n_text + (f + 1) * hw - 1):
            e = ((got[:, r] - want[:, r]).norm() / want[:, r].norm()).item()
            assert e < 4e-3, (f, r, e)


def test_ln_modulate_row_shard_equals_the_unsharded_rows_and_the_oracle():
    d, B, n_text, hw, frames, n_vip, eps = 3072, 2, 226, 50, 13, 480, 1e-5
    x, sd, mod = _case(d, B, n_text, hw, frames, n_vip, seed=5)
    want = _oracle(x, sd, mod, B, n_text, hw, frames, n_vip, d, eps)
    full = _run(x, sd, mod, B, n_text, hw, frames, n_vip, d, eps)
    rows = n_text + hw * frames + n_vip
    for row0, rl in ((0, 200), (200, 500), (rows - 333, 333)):     # crosses text->video and video->vip
        part = _run(x, sd, mod, B, n_text, hw, frames, n_vip, d, eps, row0, rl)
        assert torch.equal(part, full[:, row0:row0 + rl])
        assert ((part - want[:, row0:row0 + rl]).norm() / want[:, row0:row0 + rl].norm()).item() < 3e-3


@pytest.mark.parametrize("d,frames", [(3072, 13), (256, 1)])
def test_double_layernorm_tail_against_the_oracle(d, frames):
    """ln2 path: norm_final (affine) then norm_out.norm (affine) * (1 + scale) + shift, one read and one write."""
    from oracle import dit as odit
    from tokensgen_b200 import _ext as E
    B, n_text, hw, n_vip, eps = 2, 10, 40, 12, 1e-5
    x, sd, mod = _case(d, B, n_text, hw, frames, n_vip, seed=77 + d)
    f32 = torch.float32
    nv = hw * frames
    xv = x[:, n_text:n_text + nv].float()
    h = odit.layer_norm(xv, {"f.weight": sd["f.weight"], "f.bias": sd["f.bias"]}, "f", eps, f32)
    c = lambda i: mod[:, i * d:(i + 1) * d].float().reshape(B, frames, d).repeat_interleave(hw, dim=1)
    want = odit.layer_norm(h, sd, "o.norm", eps, f32) * (1 + c(1)) + c(0)       # AdaLayerNorm chunk order: (shift, scale)
    dev = "cuda"
    rm = E.make_rowmap(n_text, nv, n_vip, hw, frames)
    m = mod.to(dev)
    shift = E.make_modvec(None, m[:, 0:d], None)
    scale = E.make_modvec(None, m[:, d:2 * d], None)
    w = {k: v.to(dev) for k, v in sd.items()}
    xin = x.to(dev)
    out = torch.zeros_like(xin)
    E.ln_modulate(xin.view(-1, d), out.view(-1, d), B, rm, w["f.weight"], w["f.bias"], None, None, eps, shift, scale,
                  ln2_w=w["o.norm.weight"], ln2_b=w["o.norm.bias"], eps2=eps)
    torch.cuda.synchronize()
    got = out[:, n_text:n_text + nv].float().cpu()
    err = ((got - want).norm() / want.norm()).item()
    assert err < 3e-3, err
