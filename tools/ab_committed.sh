#!/bin/bash
# Same-box A/B of the in-tree library against another build of it (developer tool):
#   cp <other build>/libcartseg.so cart-segmentation-unet_b200/cartseg/libcartseg_committed.so; gpurun -- bash tools/ab_committed.sh [pairs]
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
OLD=$PWD/cart-segmentation-unet_b200/cartseg/libcartseg_committed.so
for i in $(seq 1 ${1:-2}); do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > gpurun_out/ab_new_$i.log 2>&1
  CARTSEG_LIB_PATH=$OLD python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > gpurun_out/ab_old_$i.log 2>&1
done
