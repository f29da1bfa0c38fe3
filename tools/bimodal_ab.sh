#!/bin/bash
# Round-1 open item: the one-process 8-GPU node cycle (gsb_cycle_all) was bimodal from one process start to the next
# (0.38-0.45 ms or 1.1-2.0 ms; profiles/node_cycle_8gpu_bimodal_r01.txt). A/B on the SAME box, fresh process per run:
# the round-1 library (build/r01_tree, commit 8d7a26f) with its diagnostic knobs, then this round's.
#   gpurun --gpus 8 --timeout 600 -- 'tools/bimodal_ab.sh r02d'
set -uo pipefail
TAG=${1:-rXX}
OUT=$PWD/gpurun_out/$TAG
mkdir -p $OUT
pick() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'step_ms': d['step_ms'], 'inv_us_p50': d.get('inventory_us_p50_per_device'), 'knobs': d.get('knobs')}))"; }
run() { # label, dir, env...
  local label=$1 dir=$2; shift 2
  for i in 1 2 3 4; do
    (cd $dir && env "$@" timeout 120 python tools/node_cycle.py 300 1 2>>$OUT/err.log | pick) | sed "s/^/$label run$i /" >> $OUT/ab.txt
  done
}
run "r01 default(NVML query per device per cycle, 8 concurrent)" build/r01_tree X=1
run "r01 GSB_CYCLE_ORDER=2 (no NVML in the cycle)" build/r01_tree GSB_CYCLE_ORDER=2
run "r01 GSB_NVML_SERIAL=1 (one NVML query at a time)" build/r01_tree GSB_NVML_SERIAL=1
run "r01 GSB_WORKER_SPIN_US=0 (cv hand-off)" build/r01_tree GSB_WORKER_SPIN_US=0
run "r02 snapshot" . X=1
run "r02 live NVML" . GSB_INVENTORY_POLICY=live
cat $OUT/ab.txt
# and the bench at N = 8 with the driver's own command line, under a driver-style sampler
N=$(nvidia-smi -L | wc -l)
nvidia-smi --query-gpu=index,clocks.sm,power.draw,utilization.gpu --format=csv,noheader -lms 100 > $OUT/smi.csv 2>&1 & S=$!
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_${N}gpu_sampled.json 2> $OUT/bench_${N}gpu_sampled.err
kill $S 2>/dev/null
