#!/bin/bash
# Regenerates everything under profiles/ from the logs a GPU run left in gpurun_out/ (run HERE, after e.g.
#   gpurun --timeout 2400 -- 'bash tools/gpu_suite.sh bench20 benchref bench3 bench4 infer aux auxncu launches traffic trace'
# ).  $1 = round tag (default r1).  Files that have no fresh log are left alone.
cd "$(dirname "$0")/.."
R=${1:-r1}
G=gpurun_out
j() { [ -s "$G/$1" ] && grep '^{' "$G/$1" > "profiles/$2" && echo "profiles/$2"; }
j bench.log ${R}_bench_k2_1gpu.json
j bench_ref.log ${R}_bench_reference_arm.json
j bench_k3.log ${R}_bench_k3_1gpu.json
j bench_k4.log ${R}_bench_k4_1gpu.json
j bench_2gpu.log ${R}_bench_k2_2gpu.json
j bench_4gpu.log ${R}_bench_k2_4gpu.json
j bench_8gpu.log ${R}_bench_k2_8gpu.json
j dp_parity_2gpu.log ${R}_dp_parity_2gpu.json
j bench_infer.log ${R}_infer_sweep_k5.jsonl
j bench_aux.log ${R}_aux_kernels_k2size.jsonl
if [ -s $G/launches.csv ]; then
  cp $G/launches.csv profiles/${R}_k2_step_launches.csv
  { echo "# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python tools/profile_step.py"
    echo "# one training step of workload k2 (B=64, 3x224x224, focal-Dice, AdamW) on one B200."
    echo "# plain run of the same command: $(tail -1 $G/profile_plain.log 2>/dev/null)"
    echo "# cold-cache, serialised per-launch times: compare SHARES, not absolutes"
    python tools/summarize_launches.py $G/launches.csv
    echo; echo "# GEMM launches mapped to layers (algorithmic 2*MACs / ncu duration)"
    python tools/per_layer.py $G/launches.csv; } > profiles/${R}_k2_step_summary.txt 2>/dev/null
  echo profiles/${R}_k2_step_summary.txt
fi
[ -s $G/step_traffic.csv ] && python tools/step_traffic.py $G/step_traffic.csv profiles/${R}_step_traffic.json
if [ -s $G/aux_launches.csv ]; then
  { echo "# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python tools/bench_aux.py --iters 1 --no-cpu"
    echo "# bandwidth-bound kernels of the hot path and of the next rows at K2 size (B=64, 224x224): per-launch averages, cold cache"
    python tools/ncu_kernel_table.py $G/aux_launches.csv; } > profiles/${R}_aux_kernels_ncu.txt
  echo profiles/${R}_aux_kernels_ncu.txt
fi
[ -s $G/trace_backward.log ] && cp $G/trace_backward.log profiles/${R}_k2_backward_timeline.txt && echo profiles/${R}_k2_backward_timeline.txt
exit 0
