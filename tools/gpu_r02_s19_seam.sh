#!/bin/bash
# Fused seam kernel (last IFFT pass -> discriminator -> first real-FFT pass): parity on the GPU, then A/B.
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config3 or golden or config2 or graph" > gpurun_out/pytest_seam.log 2>&1; tail -5 gpurun_out/pytest_seam.log
for wl in cfg3 cfg3-wbfm; do
for fuse in 1 0; do
  if [ $fuse = 0 ]; then export RC_NO_FUSE_AD=1; else unset RC_NO_FUSE_AD; fi
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/seam_${wl}_$fuse.json 2> gpurun_out/seam_${wl}_$fuse.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/seam_${wl}_$fuse.json"))
    print("SEAM $wl fuse=$fuse ms/step", round(d["ms_per_step"], 4), "Msps", round(d["value"]))
    for k, v in d["kernels"].items():
        if "channel_ifft" in k or "rfft_disc" in k: print("     ", k, v["avg_ms"], v["frac_of_hbm_peak"])
except Exception as e: print("SEAM $wl $fuse failed", e)
PY
done
done
unset RC_NO_FUSE_AD
