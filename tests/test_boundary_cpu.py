"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every declared symbol; host-side mirrors keep
the reference's names / state-dict layout; the product path never imports the oracle and has no CPU fallback."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from tokensgen_b200 import _ext
    lib = _ext.load()
    header = open(os.path.join(ROOT, "include", "tokensgen_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(tg_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_ext.SYMBOLS), declared ^ set(_ext.SYMBOLS)
    for name in declared:
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr)
    assert lib.tg_version() == 1


def test_argument_errors_are_reported_without_a_gpu():
    from tokensgen_b200 import _ext
    lib = _ext.load()
    # K not a multiple of 64 -> negative return before any launch, message available
    rc = lib.tg_gemm_bias_act(16, 64, 16, None, 16, 64, 8, 64, 60, 0, None)
    assert rc < 0 and b"multiple of 64" in lib.tg_last_error()
    rc = lib.tg_attn_fwd(None, 1, 0, 1, None, None, 1, 0, 1, None, 1, 0, 1, 1, 0.125, 0, 1.0, None)
    assert rc < 0 and b"null" in lib.tg_last_error()


def test_product_path_does_not_import_oracle_or_reference():
    pkg = os.path.join(ROOT, "tokensgen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_no_cpu_fallback():
    from tokensgen_b200 import _ext
    from tokensgen_b200.transformer import CogVideoXTransformer3DModel
    m = CogVideoXTransformer3DModel(num_attention_heads=4, attention_head_dim=64, time_embed_dim=128, text_embed_dim=128,
                                    num_layers=1, use_rotary_positional_embeddings=True).to(torch.bfloat16)
    with pytest.raises(_ext.TokensGenError):
        m(torch.zeros(1, 3, 16, 8, 12, dtype=torch.bfloat16), torch.zeros(1, 10, 128, dtype=torch.bfloat16),
          torch.tensor([5]))


@pytest.mark.parametrize("use_vip", [True, False])
def test_state_dict_layout_matches_reference(use_vip):
    """Key names + shapes equal the ones recorded from the instantiated reference model (tests/golden/dit_tiny.json)."""
    from tokensgen_b200.transformer import CogVideoXTransformer3DModel
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "dit_tiny.json")))["vip" if use_vip else "plain"]
    m = CogVideoXTransformer3DModel(num_attention_heads=4, attention_head_dim=64, in_channels=16, out_channels=16,
                                    time_embed_dim=128, text_embed_dim=128, num_layers=2, patch_size=2,
                                    use_rotary_positional_embeddings=True, attention_bias=True)
    if use_vip:
        m.set_vip_layers(None, length=12, func_type="1", scale=[0.6],
                         resampler_params=dict(output_dim=128, num_height_queries=2, num_width_queries=3, num_temporal_queries=1))
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == meta["shapes"]
    if use_vip:
        proc = m.transformer_blocks[0].attn1.processor
        assert type(proc).__name__ == "VideoIPAdapterCogVideoXAttnProcessor2_0" and proc.scale == [0.6]
        assert torch.equal(proc.vip_to_q.weight, m.transformer_blocks[0].attn1.to_q.weight)


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """Every struct the binding passes by pointer has the size and field offsets `gcc` gives the header's definition."""
    import shutil
    import subprocess
    from tokensgen_b200 import _ext as E
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = {"tg_rowmap": E.RowMap, "tg_modvec": E.ModVec, "tg_qkv_proj": E.QkvProj, "tg_qkv_scatter": E.QkvScatter,
               "tg_attn_scatter": E.AttnScatter, "tg_dpm_step_args": E.DpmStepArgs, "tg_conv_args": E.ConvArgs,
               "tg_norm_args": E.NormArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/tokensgen_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(out) == len(structs)
    for line, (cname, cls) in zip(out, structs.items()):
        parts = line.split()
        assert parts[0] == cname
        want = [ctypes.sizeof(cls)] + [getattr(cls, f).offset for f, _ in cls._fields_]
        assert [int(x) for x in parts[1:]] == want, f"{cname}: header {parts[1:]} vs ctypes {want}"


def test_binding_argument_counts_match_the_header():
    from tokensgen_b200 import _ext
    header = open(os.path.join(ROOT, "include", "tokensgen_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"^(?:int|const char\*)\s+(tg_\w+)\s*\((.*?)\);", header, flags=re.M | re.S)
    assert len(protos) == len(_ext.SYMBOLS)
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_ext.SYMBOLS[name][1]), f"{name}: header has {n} parameters, binding {len(_ext.SYMBOLS[name][1])}"


def test_library_sass_is_blackwell_native():
    """The shipped .so holds sm_100a SASS with tcgen05 tensor-core MMAs (UTCHMMA, incl. the CTA-pair form), TMEM loads/stores
    (LDTM/STTM), TMA tile loads (UTMALDG) and tcgen05 commits (UTCBAR) — not a recompiled mma.sync / cp.async path."""
    import shutil
    import subprocess
    from tokensgen_b200 import _ext
    if shutil.which("cuobjdump") is None:
        pytest.skip("no cuobjdump")
    sass = subprocess.run(["cuobjdump", "-sass", str(_ext.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG.2D.2CTA", "UTCBAR", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass and "LDGSTS" not in sass      # no mma.sync tensor path, no cp.async staging
