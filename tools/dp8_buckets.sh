cd /root/repo; mkdir -p gpurun_out
for mb in 16 48 100000; do
  CARTSEG_DP_BUCKET_MB=$mb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29700 + mb % 97)) bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > gpurun_out/bench_k2_8gpu_bucket$mb.log 2>&1
  tail -1 gpurun_out/bench_k2_8gpu_bucket$mb.log | cut -c1-160
done
