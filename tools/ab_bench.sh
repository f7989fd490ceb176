#!/bin/bash
# In-step A/B of the attention launch structure and the speculative softmax reference: one bench.py run per setting
# (N = 1, 10 timed steps, no baselines), one JSON line each -> gpurun_out/ab_bench.jsonl
# needs the developer build of the library (python -m tokensgen_b200.build --dev): the shipped one has no knobs
export TG_LIB_PATH=${TG_LIB_PATH:-$(dirname $0)/../tokensgen_b200/libtokensgen_b200_dev.so}
out=gpurun_out/ab_bench.jsonl
: > $out
run() {
  echo "== $*" >&2
  env "$@" python bench.py --steps 10 --warmup 3 --no-vae --no-eager-baseline --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print(json.dumps({'setting': '$*', 'ms_per_step': round(d['ms_per_step'],2), 'attn_self_ms_per_launch': round(d['roofline']['avg_launch_ms'],3), 'frac': round(d['roofline']['frac'],4), 'sm_mhz': d['clocks']['sm_mhz'], 'attn_ms': {a: b for a, b in k.items() if a.startswith('attn')}}))" >> $out
}
run TG_ATTN_SPEC=1 TG_ATTN_EMU=1 TG_FUSE_PAIR=1 TG_SIDE_STREAM=1
run TG_ATTN_SPEC=0 TG_ATTN_EMU=0 TG_FUSE_PAIR=0 TG_SIDE_STREAM=0
cat $out
