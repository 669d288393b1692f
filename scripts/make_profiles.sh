#!/bin/bash
# One gpurun call that produces the ncu evidence under gpurun_out/ (summarised into profiles/ by scripts/summarize_profiles.py).
# Every ncu command is preceded by the same command without ncu (exit 0 required), as the profiling recipe asks.
set -u
R=${1:-r2}
mkdir -p gpurun_out
# 1. launch list of the timed path: one CUDA-graph replay of the cfg2 training step, per-node device time (warm caches)
python scripts/ncu_graph_step.py --config cfg2 --precision tf32 > gpurun_out/plain_graph.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --graph-profiling node -c 4000 --csv \
    --log-file gpurun_out/launches_${R}_cfg2_tf32_graph.csv python scripts/ncu_graph_step.py --config cfg2 --precision tf32 > gpurun_out/ncu_graph.log 2>&1
# 2. full sections for the dominant kernels of the step (eager, single stream)
python scripts/ncu_step.py --config cfg2 --precision tf32 > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none -k regex:"mbv3_fwd_kernel|mbv3_bwd_kernel|conv_tma_kernel|se_gate|dense_tc_kernel|sc::fwd" \
    -s 4 -c 36 -o gpurun_out/prof_${R}_step_kernels python scripts/ncu_step.py --config cfg2 --precision tf32 > gpurun_out/ncu_step.log 2>&1
# 3. pyramid / ELBO kernels at BASELINE configs[4] (512x512x3, batch 128, 9 levels), one cold launch each
#    (SKIP_PYRAMID=1 leaves this capture out: gpurun copies back at most 64 MiB, and these kernels change rarely)
[ "${SKIP_PYRAMID:-0}" = "1" ] || python bench.py --micro-only --micro-iters 1 > gpurun_out/plain_micro.log 2>&1 &&
[ "${SKIP_PYRAMID:-0}" = "1" ] || ncu --set full --clock-control none -k regex:"split_pair|merge_pair|adjoint_pair|recon_loss" -c 14 \
    -o gpurun_out/prof_${R}_pyramid python bench.py --micro-only --micro-iters 1 > gpurun_out/ncu_pyr.log 2>&1
# 4. the weight-gradient kernels (bench.py's dominant call of the cfg2 step), first launches = level 0
ncu --set full --clock-control none -k regex:"wgrad_tma_kernel|sc::wgrad_kernel" -c 8 \
    -o gpurun_out/prof_${R}_wgrad python scripts/ncu_step.py --config cfg2 --precision tf32 > gpurun_out/ncu_wgrad.log 2>&1
# only the raw metric tables travel back (gpurun merges at most 64 MiB): summarised here by scripts/summarize_profiles.py
for rep in prof_${R}_step_kernels prof_${R}_pyramid prof_${R}_wgrad; do
    [ -f gpurun_out/$rep.ncu-rep ] && ncu -i gpurun_out/$rep.ncu-rep --page raw --csv > gpurun_out/${rep}_raw.csv 2>/dev/null
    rm -f gpurun_out/$rep.ncu-rep
done
ls -la gpurun_out/*_raw.csv gpurun_out/launches_${R}_*.csv
