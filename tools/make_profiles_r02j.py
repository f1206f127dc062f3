"""Turns the gpurun_out/r02j_* files of tools/gpu_r2_final.sh (T=r02j) into the tracked profiles/r02j_* summaries."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json(path):
    for line in reversed(open(path).read().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit(f"no JSON line in {path}")


def run(args, out):
    txt = subprocess.run([sys.executable] + args, capture_output=True, text=True, cwd=ROOT).stdout
    open(os.path.join(P, out), "w").write(txt)


lines = {"n1_default_flags": last_json(f"{G}/r02j_bench.log"), "n1_reference_arm": last_json(f"{G}/r02j_bench_reference.log")}
for name in ("bench_det", "bench_configs"):
    pth = f"{G}/r02j_{name}.json"
    if os.path.exists(pth):
        lines.update(json.load(open(pth)))
for n in (2, 4, 8):
    if os.path.exists(f"{G}/r02j_n{n}.log"):
        lines[f"n{n}"] = last_json(f"{G}/r02j_n{n}.log")
for c in ("v2-164", "wrn50-imagenet"):
    if os.path.exists(f"{G}/r02j_n2_{c}.log"):
        lines[f"n2_config_{c}"] = last_json(f"{G}/r02j_n2_{c}.log")
json.dump(lines, open(f"{P}/r02j_bench_lines.json", "w"), indent=1)
for tag in ("launches", "launches_warm"):
    src = f"{G}/r02j_{tag}.csv"
    subprocess.run(["cp", src, f"{P}/r02j_{tag}.csv"], check=True)
    run(["tools/summarize_launches.py", src], f"r02j_{tag}_summary.txt")
run(["tools/ncu_metrics_txt.py", f"{G}/r02j_prof.ncu-rep",
     "round 2 final code (r02j): ncu --set full --clock-control none of tools/kernels_once.py 1"], "r02j_ncu_metrics.txt")

# kernel times: conv table + BN family
rows = json.load(open(f"{G}/bench_conv_r02j.json"))
out = ["# CUDA-event timings WITHOUT a profiler (graph-captured launches, rotating buffers > L2), batch 128, round 2 FINAL kernels (r02j)",
       "# source: gpurun_out/bench_conv_r02j.json, r02j_bench_ew.log (tools/bench_conv.py, tools/bench_ew.py), same box and run as r02j_bench_lines.json",
       "# conv: us per launch (TFLOP/s); cuDNN = torch bf16 channels_last conv of the same op, re-measured in the SAME run",
       "# fprop_stats = fprop + fused BN statistics; fprop_res_stats = + fused residual add; dgrad_bnbwd = dgrad + fused BN-backward reduction",
       f"{'N x H x W x C x K x R x s x p':34s} {'fprop':>16s} {'dgrad':>16s} {'wgrad':>16s} | fprop_stats fprop_res_stats dgrad_bnbwd cuDNN f/d/w"]
for r in rows:
    us = lambda k: r.get(k + "_ms", float("nan")) * 1e3
    tf = lambda k: r.get(k + "_tflops", float("nan"))
    out.append(f"{'x'.join(map(str, r['shape'])):34s} {us('fprop'):9.1f} ({tf('fprop'):4.0f}) {us('dgrad'):9.1f} ({tf('dgrad'):4.0f}) "
               f"{us('wgrad'):9.1f} ({tf('wgrad'):4.0f}) | {us('fprop_stats'):11.1f} {us('fprop_res_stats'):15.1f} {us('dgrad_bnbwd'):11.1f}"
               f"   {us('cudnn_fprop'):5.1f}/{us('cudnn_dgrad'):5.1f}/{us('cudnn_wgrad'):5.1f}")
out.append("")
out.append("# BN family (tools/bench_ew.py):")
out += [l for l in open(f"{G}/r02j_bench_ew.log").read().splitlines() if l.strip() and not l.startswith("{")]
open(f"{P}/r02j_kernel_times.txt", "w").write("\n".join(out) + "\n")
print("ok")
