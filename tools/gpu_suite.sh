#!/bin/bash
# Runs on the GPU box (via gpurun): diagnostics, the -m gpu parity tests file by file (each in its own process so a
# faulting kernel cannot take the other files down), then smoke + a short bench.  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
status=0
run() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "[$name] exit $rc"
  tail -n 6 "gpurun_out/$name.log" | sed "s/^/    /"
  [ $rc -ne 0 ] && status=1
  return $rc
}
for part in "$@"; do
  case $part in
    diag)   run diag 300 python tools/diag_conv.py ;;
    sdf)    run test_gpu_sdf 600 python -m pytest tests/test_gpu_sdf.py -m gpu -q --tb=short -s --timeout 300 ;;
    abl)    run test_gpu_abl 600 python -m pytest tests/test_gpu_abl.py -m gpu -q --tb=short -s --timeout 300 ;;
    postproc) run test_gpu_postproc 600 python -m pytest tests/test_gpu_postproc.py -m gpu -q --tb=short -s --timeout 300 ;;
    preproc) run test_gpu_preproc 600 python -m pytest tests/test_gpu_preproc.py -m gpu -q --tb=short -s --timeout 300 ;;
    losses) run test_gpu_losses 600 python -m pytest tests/test_gpu_losses.py -m gpu -q --tb=short -s --timeout 300 ;;
    layers) run test_gpu_layers 900 python -m pytest tests/test_gpu_layers.py -m gpu -q --tb=short -s --timeout 300 ;;
    stages) run test_gpu_unet_stages 1200 python -m pytest tests/test_gpu_unet_stages.py -m gpu -q --tb=short -s --timeout 600 ;;
    unet)   run test_gpu_unet 1200 python -m pytest tests/test_gpu_unet.py -m gpu -q --tb=short -s --timeout 600 ;;
    layers_old) CARTSEG_CONV3=0 run test_gpu_layers_oldkernel 900 python -m pytest tests/test_gpu_layers.py -m gpu -q --tb=short --timeout 300 ;;
    bench10_old) CARTSEG_CONV3=0 run bench10_old 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-workloads ;;
    bench10_nodefer) CARTSEG_DEFER_WGRAD=0 run bench10_nodefer 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-workloads ;;
    bench10_nofuse) CARTSEG_FUSE_HEAD=0 run bench10_nofuse 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-workloads ;;
    bench10_flush3) CARTSEG_DEFER_FLUSH_CONV=3 run bench10_flush3 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-workloads ;;
    bench10_flush7) CARTSEG_DEFER_FLUSH_CONV=7 run bench10_flush7 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-workloads ;;
    ab_*) # A/B of an environment knob inside one box: ab_<ENVVAR>=<value>  (20 steps each, no CPU leg)
            kv=${part#ab_}; env "$kv" bash -c 'true' && \
            run "bench_${kv//[^A-Za-z0-9_=]/_}" 900 env "$kv" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads ;;
    scale_*) # scale_<workload>_<N>[_bf16]: bench of one workload on N GPUs of this box, 20 steps
            spec=${part#scale_}; wl=${spec%%_*}; rest=${spec#*_}; n=${rest%%_*}; wire=fp32; [[ "$rest" == *_bf16 ]] && wire=bf16
            if [ "$n" = 1 ]; then
              run "bench_${wl}_1gpu" 900 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads
            else
              CARTSEG_DP_WIRE_DTYPE=$wire run "bench_${wl}_${n}gpu_${wire}" 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
                --master-addr 127.0.0.1 --master-port $((29600 + n + ${#wl} + ${#wire})) bench.py --gpus $n --workload $wl --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads
            fi ;;
    tracedp8) run trace_dp_8gpu 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 tools/trace_dp.py
              run trace_dp_8gpu_bf16 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 tools/trace_dp.py --wire bf16 ;;
    bench20q) run bench20q 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads ;;
    bench10q) run bench10q 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-workloads ;;
    cpubase) run cpu_baseline 1500 python tools/cpu_baseline.py ;;
    loop)   run test_gpu_loop 900 python -m pytest tests/test_gpu_loop.py -m gpu -q --tb=short -s --timeout 600 ;;
    fullsize) run test_gpu_fullsize 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --tb=short -s --timeout 900 ;;
    gradparity) run grad_parity 1200 python tools/grad_parity.py --out gpurun_out/r2_grad_parity_per_tensor.json ;;
    gradparity512) run grad_parity512 1200 python tools/grad_parity.py --batch 8 --size 512 --out gpurun_out/r2_grad_parity_per_tensor_512.json ;;
    hostinfo) (free -g; nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; nvidia-smi) > gpurun_out/hostinfo.txt 2>&1 ;;
    sanitize) SEL='2-16-16-64-64 or 3-16-8-128-64 or 2-14-14-128-256 or 1-28-28-256-128 or 2-8-8-128-64'
            for tool in memcheck racecheck synccheck; do
              run sanitizer_$tool 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_layers.py -m gpu -q --tb=line -k "$SEL" --timeout 1200
            done
            run sanitizer_memcheck_smoke 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" ;;
    ncucalib) run cublas_plain 300 python tools/cublas_calib.py
            run ncu_cublas 900 ncu --set full --clock-control none -k regex:'gemm|nvjet|cutlass|sm100|sm90' -s 3 -c 2 -f -o gpurun_out/prof_cublas python tools/cublas_calib.py
            run ncu_conv3 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
                -k regex:conv3_gemm -c 8 -f -o gpurun_out/prof_conv3 python tools/profile_step.py ;;
    ncubn)  run ncu_bn 1200 ncu --profile-from-start off --set full --clock-control none --import-source on \
                -k regex:'bn_bwd_apply_kernel<1|bn_bwd_reduce_kernel<1|bn_relu_kernel|bn_bwd_reduce_kernel<0, 1|bn_bwd_apply_kernel<0, 1' -c 40 -f -o gpurun_out/prof_bn python tools/profile_step.py ;;
    tracedp2) run trace_dp_2gpu 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/trace_dp.py ;;
    infer_small) run bench_infer_small 900 python tools/bench_infer.py --batches 1,2,4,8,16 ;;
    smoke)  run smoke 600 python -c "import __graft_entry__ as g; g.smoke()" ;;
    bench)  run bench 900 python bench.py --steps 5 --warmup 3 ;;
    bench_nohead) CARTSEG_FUSE_HEAD=0 run bench_nohead 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ;;
    bench10) run bench10 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ;;
    bench_nooverlap) CARTSEG_OVERLAP=0 run bench_nooverlap 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ;;
    trace)  run trace_backward 600 python tools/trace_backward.py ;;
    bench_committed) CARTSEG_LIB_PATH=$PWD/cart-segmentation-unet_b200/cartseg/libcartseg_committed.so run bench_committed 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ;;
    bench_w0) CARTSEG_WGRAD_INVERSION_WEIGHT=0 run bench_w0 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ;;
    bench_w2) CARTSEG_WGRAD_INVERSION_WEIGHT=2 run bench_w2 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ;;
    bench3) run bench_k3 900 python bench.py --steps 3 --warmup 3 --workload k3 --no-cpu-baseline ;;
    bench4) run bench_k4 900 python bench.py --steps 5 --warmup 3 --workload k4 --no-cpu-baseline ;;
    bench8gpu) run bench_8gpu 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 ;;
    bench8gpu_r4) NCCL_MAX_CTAS=4 CARTSEG_DP_RESERVE_SMS=4 run bench_8gpu_r4 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 3 ;;
    bench8gpu_r8) NCCL_MAX_CTAS=8 CARTSEG_DP_RESERVE_SMS=8 run bench_8gpu_r8 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 8 --steps 10 --warmup 3 ;;
    bench4gpu) run bench_4gpu 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 ;;
    dp2) run dp_parity_2gpu 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_parity.py ;;
    bench2gpu_nooverlap) CARTSEG_DP_BUCKET_MB=100000 run bench_2gpu_nooverlap 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 ;;
    bench2gpu_maxctas) NCCL_MAX_CTAS=4 run bench_2gpu_maxctas 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 ;;
    dptest) run test_gpu_dp 900 python -m pytest tests/test_gpu_dp.py -m gpu -q --tb=short -s --timeout 600 ;;
    bench2gpu) run bench_2gpu 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ;;
    aux)    run bench_aux 900 python tools/bench_aux.py ;;
    aux512) run bench_aux512 900 python tools/bench_aux.py --batch 32 --size 512 --no-cpu ;;
    auxncu) run ncu_aux 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/aux_launches.csv python tools/bench_aux.py --iters 1 --no-cpu ;;
    infer)  run bench_infer 900 python tools/bench_infer.py ;;
    bench20) run bench 900 python bench.py ;;
    benchref) run bench_ref 600 python bench.py --impl reference --steps 3 --warmup 1 ;;
    launches) run profile_plain 300 python tools/profile_step.py && \
            run ncu_launches 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
                --csv --log-file gpurun_out/launches.csv python tools/profile_step.py ;;
    launches3) run ncu_launches_k3 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_k3.csv python tools/profile_step.py --batch 32 --size 512 ;;
    traffic) run ncu_traffic 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/step_traffic.csv python tools/profile_step.py ;;
    ncufull) run profile_plain 300 python tools/profile_step.py && \
            run ncu_pix 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
                -k regex:pix_gemm -c 12 -f -o gpurun_out/prof_pix python tools/profile_step.py && \
            run ncu_wgrad 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
                -k regex:wgrad -c 3 -f -o gpurun_out/prof_wgrad python tools/profile_step.py ;;
    ncuxf) # ncu --set full of the first launch of each operand-transform (XF) kernel, matched on the demangled name
            run ncu_xf_wgrad 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
                -k regex:'wgrad_gemm_kernel<.int.128, .int.3, .bool.1>' -c 1 -f -o gpurun_out/prof_xf_wgrad python tools/profile_step.py
            run ncu_xf_conv64 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
                -k regex:'conv3_gemm_kernel<.int.64,.*bool.1, .bool.1>' -c 1 -f -o gpurun_out/prof_xf_conv64 python tools/profile_step.py ;;
    ncun64) # ncu --set full of the stem kernel and the first Cout = 64 convolution (conv1.3 fprop, with BN statistics) / its dgrad
            run ncu_stem 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
                -k regex:stem_gemm -c 1 -f -o gpurun_out/prof_stem python tools/profile_step.py
            run ncu_conv64 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
                -k regex:'conv3_gemm_kernel<.int.64,.*bool.1, .bool.0>' -c 1 -f -o gpurun_out/prof_conv64 python tools/profile_step.py ;;
    ncuone) run profile_plain 300 python tools/profile_step.py && \
            run ncu_one 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
                -k regex:${NCU_KERNEL:-pix_gemm} -s ${NCU_SKIP:-1} -c ${NCU_COUNT:-1} -f -o gpurun_out/${NCU_OUT:-prof_one} python tools/profile_step.py ;;
    all)    run tests_all 1800 python -m pytest tests -m gpu -q -x --tb=short --timeout 600 ;;
  esac
done
exit $status
