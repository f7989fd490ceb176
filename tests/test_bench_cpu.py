"""bench.py's reference arm (the reference's own block from baseline/_ref — or the oracle port when that directory has not
been built — timed on the host cores) runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "denoised latent tokens/sec" and line["unit"] == "tokens/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["vs_baseline"] is None and line["dtype"] == "bf16"
    assert line["config"]["workload"].startswith("configs[1]")
    cb = line["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "longvgen", "models", "cogvideox_transformer_3d.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert ("baseline/_ref" in cb["sample"]) == have_ref
    # one step of this arm = one of the 84 block forwards of a bench step: ms_per_step is what was actually timed
    assert abs(line["sample_fraction"] - 1 / 84) < 1e-12
    assert abs(line["ms_per_whole_step_extrapolated"] - 84 * line["ms_per_step"]) < 1e-6 * line["ms_per_step"] * 84
    assert abs(line["value"] - 17550 / (52 * 84 * line["ms_per_step"] / 1e3)) < 1e-9 * line["value"] + 1e-12
    # same config keys as the repo's own arm (the driver compares them)
    assert set(line["config"]) == {"workload", "denoise_steps_per_clip", "token_steps_per_s", "model_tflops_per_step",
                                   "achieved_model_tflops", "l2", "parallelism"}
    assert line["e2e"] == {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_block_equals_the_oracle_port():
    """Both implementations of the CPU arm are the same function: the reference's own CogVideoXBlock + VIP (baseline/_ref) and
    the oracle restatement give the same output on a small block with the same seeded weights (bf16, CPU)."""
    import pytest
    import torch
    sys.path.insert(0, ROOT)
    from oracle import vendor_reference as vr
    if not vr.enable():
        pytest.skip("baseline/_ref has not been built on this machine (python -m oracle.vendor_reference)")
    try:
        from longvgen.models.cogvideox_transformer_3d import CogVideoXBlock
        from oracle import dit as odit
        from oracle import rope as orope
        from oracle.synth import dit_shapes, synth_state_dict
        import numpy as np
        shapes = {k: v for k, v in dit_shapes(4, 64, 1, 128, 128, 16, 16, 2, 128, True).items() if k.startswith("transformer_blocks.0.")}
        sd = synth_state_dict(shapes, 5)
        blk = CogVideoXBlock(dim=256, num_attention_heads=4, attention_head_dim=64, time_embed_dim=128, attention_bias=True)
        blk.set_vip_layers(length=12, func_type="1", scale=[0.6])
        blk.load_state_dict({k[len("transformer_blocks.0."):]: v for k, v in sd.items()}, strict=True)
        blk = blk.to(torch.bfloat16).eval()
        g = torch.Generator().manual_seed(0)
        hid, enc, temb = (torch.randn(1, 72, 256, generator=g).bfloat16(), torch.randn(1, 22, 256, generator=g).bfloat16(),
                          torch.randn(1, 3, 128, generator=g).bfloat16())
        rope = orope.rope_3d(64, [[0, 0, 0], [3, 4, 6]], (3, 4, 6))
        img = orope.rope_3d_from_grids(64, np.array([5, 6, 7], np.float32), np.arange(4, dtype=np.float32), np.arange(6, dtype=np.float32))
        cond = orope.rope_3d_from_grids(64, np.array([1000, 1001.5], np.float32), np.linspace(0, 4, 2, endpoint=False, dtype=np.float32),
                                        np.linspace(0, 6, 3, endpoint=False, dtype=np.float32))
        with torch.no_grad():
            h1, e1 = blk(hid, enc, temb, image_rotary_emb=rope, vip_image_rotary_emb=img, vip_condition_rotary_emb=cond)
            cfg = odit.DitConfig(num_attention_heads=4, attention_head_dim=64, time_embed_dim=128, text_embed_dim=128, num_layers=1,
                                 vip_length=12, vip_embed_dim=128, use_vip=True)
            h2, e2 = odit.block_forward(sd, "transformer_blocks.0", cfg, hid, enc, temb, rope, img, cond, torch.bfloat16)
        rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
        assert rel(h2, h1) < 5e-3 and rel(e2, e1) < 5e-3
    finally:
        for m in [m for m in sys.modules if m == "longvgen" or m.startswith("longvgen.")]:
            del sys.modules[m]
        sys.path[:] = [p for p in sys.path if "baseline/_ref" not in p]


def test_vae_block_runs_in_a_child_that_cannot_cost_the_bench_line():
    """bench.py's N = 1 `vae` object comes from a child process under a time limit: a child that does not return is killed with
    its process group and recorded as an error, a child that fails (here: no GPU) hands its error back — the bench line is
    printed either way."""
    import time
    sys.path.insert(0, ROOT)
    import bench
    t0 = time.time()
    out = bench.run_vae_block([], 0.05)
    assert "did not finish" in out["error"] and time.time() - t0 < 10
    out = bench.run_vae_block([], 300)
    assert "error" in out and "NVIDIA" in out["error"] or "CUDA" in out["error"] or "points" in out
