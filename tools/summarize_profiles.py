"""Turns gpurun_out/ ncu outputs into the small tracked summaries under profiles/ (round-tagged).
usage: python tools/summarize_profiles.py r01"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")


def launches():
    path = os.path.join(G, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, data = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        name = re.sub(r"\(.*", "", r[ki])[:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e6
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(out_dir, f"{tag}_bench_launches_summary.md"), "w") as f:
        f.write(f"# ncu launch list of `bench.py --steps 1 --warmup 1` ({tag}; first {len(data)} launches; cold-cache, serialised)\n\n")
        f.write("command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline`\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ms:.3f} | {100 * ms / total:.1f}% |\n")
        f.write(f"\ntotal {total:.1f} ms over {len(data)} launches\n")


KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg"]


def report(rep, title):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    with open(os.path.join(out_dir, f"{tag}_{rep}_summary.md"), "w") as f:
        f.write(f"# {title}\n\nsource: `ncu --set full --clock-control none --import-source on` capture `{rep}.ncu-rep` (one launch)\n\n")
        for k in range(2, len(rows)):
            vals = dict(zip(rows[0], rows[k]))
            units = dict(zip(rows[0], rows[1]))
            f.write(f"## {vals.get('Kernel Name', '')[:100]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for key in KEYS:
                if key in vals:
                    f.write(f"| {key} | {vals[key]} | {units[key]} |\n")
            f.write("\n")
        src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        if len(srows) > 2:
            hdr, data = srows[1], srows[2:]
            ix = {h: i for i, h in enumerate(hdr)}
            tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
            st = collections.Counter()
            ops = collections.Counter()
            for r in data:
                for h in hdr:
                    if h.startswith("stall_") and "Not Issued" not in h:
                        st[h] += int(r[ix[h]] or 0)
                op = [o for o in r[ix["Source"]].split() if not o.startswith("@")][0]
                ops[".".join(op.split(".")[:2])] += int(r[ix["Instructions Executed"]] or 0)
            f.write(f"## warp-state samples ({tot} total)\n\n| stall reason | samples | share |\n|---|---:|---:|\n")
            for k_, v in st.most_common(8):
                f.write(f"| {k_} | {v} | {100 * v / max(tot, 1):.1f}% |\n")
            f.write("\n## executed warp instructions by opcode (top 12)\n\n| opcode | executed |\n|---|---:|\n")
            for k_, v in ops.most_common(12):
                f.write(f"| {k_} | {v} |\n")
            tc = [k_ for k_ in ops if k_.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLKCP"))]
            f.write("\nBlackwell-native SASS present: " + ", ".join(f"{k_} x{ops[k_]}" for k_ in sorted(tc)) + "\n")


launches()
report("attn_full", "tg::attn_fwd_kernel (v1: 8 softmax warps) — self-attention, 1x48 heads x 17776^2 x 64")
report("attn2_full", "tg::attn2_fwd_kernel (v2: 16 softmax warps) — self-attention, 1x48 heads x 17776^2 x 64")
report("attn3_full", "tg::attn3_fwd_kernel (v3: lean waits, packed FMA-pipe exponentials 1/8, 112 softmax registers) — self-attention, 2x48 heads x 17776^2 x 64")
report("attn_pair_full", "tg::attn3_fwd_kernel<1,false> (round 2: speculative softmax reference, 1/8 of the exponentials on the FMA pipe) "
       "pair launch — self-attention, 2x48 heads x 17776^2 x 64 + cross-attention of the same queries to 480 vip keys")
report("conv_stats_full", "tg::conv3_kernel with the GroupNorm statistics of the output accumulated in the epilogue (128 -> 128, 8 x 480 x 720)")
for extra in sys.argv[2:]:
    report(extra, extra)
print(sorted(os.listdir(out_dir)))
