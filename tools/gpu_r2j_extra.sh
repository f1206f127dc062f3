#!/bin/bash
# r02j extras in one 1-GPU call: the other bench configs, the deterministic-mode line, the ImageNet-config launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for c in resnet20 v2-164 wrn50-imagenet; do
  timeout 300 python bench.py --config $c --steps 20 --warmup 5 --sustained 0 --no-cpu-baseline > gpurun_out/r02j_cfg_$c.log 2>&1; echo "$c rc=$?"
done
B200_DETERMINISTIC=1 timeout 200 python bench.py --steps 30 --warmup 5 --sustained 0 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02j_det.log 2>&1; echo "det rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02j_imagenet_launches.csv python tools/profile_config.py wrn50-imagenet 2 > gpurun_out/r02j_imagenet_ncu.log 2>&1; echo "imagenet launches rc=$?"
python - <<'PY'
import json
def last(p):
    for l in reversed(open(p).read().splitlines()):
        if l.startswith("{"): return json.loads(l)
out = {f"config_{c}": last(f"gpurun_out/r02j_cfg_{c}.log") for c in ("resnet20", "v2-164", "wrn50-imagenet")}
json.dump(out, open("gpurun_out/r02j_bench_configs.json", "w"), indent=1)
json.dump({"n1_deterministic_mode": last("gpurun_out/r02j_det.log")}, open("gpurun_out/r02j_bench_det.json", "w"), indent=1)
for k, v in out.items(): print(k, v["ms_per_step"], v["value"], v.get("gpu_reference", {}).get("bf16"))
PY
