#!/bin/bash
# First-contact environment facts for BASELINE.md §3. Output -> gpurun_out/envprobe.txt
exec > gpurun_out/envprobe.txt 2>&1
set -x
nproc; lscpu | head -25
nvidia-smi
nvidia-smi --query-gpu=index,pci.bus_id,uuid,memory.total,memory.reserved,memory.used,memory.free --format=csv
env | grep -i -E 'cuda|nvidia|gpu' 
ls -la /dev/nvidia* 
ls /proc/driver/nvidia/gpus/ && for d in /proc/driver/nvidia/gpus/*; do echo $d; cat $d/information; done
cat /proc/driver/nvidia/version
ls -la /usr/lib/x86_64-linux-gnu/libnvidia-ml* /usr/lib/x86_64-linux-gnu/libcuda* 2>/dev/null
./build/envprobe
go version
free -g | head -3
