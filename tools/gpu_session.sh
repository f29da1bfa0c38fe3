#!/bin/bash
# One gpurun call's worth of verification + measurement, in the order that matters if the call is cut short.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'tools/gpu_session.sh r02'            (1 GPU)
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'tools/gpu_session.sh r02 node' (8 GPUs: node-level only)
# Everything lands under gpurun_out/<tag>/; copy what should be judged into profiles/.
set -uo pipefail
TAG=${1:-rXX}
MODE=${2:-single}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/clocks_before.csv 2>&1
if [ "$MODE" = node ]; then
  N=$(nvidia-smi -L | wc -l)
  # all GPUs of the box from ONE process (gsb_cycle_all) next to the reference's sequential NVML walk
  timeout 300 python tools/node_cycle.py > $OUT/node_cycle_${N}gpu.json 2> $OUT/node_cycle_${N}gpu.err
  # open item of round 1: the in-process node cycle is bimodal at N = 8; spin budget and launch order as knobs
  for spin in 0 50 400; do
    GSB_WORKER_SPIN_US=$spin timeout 300 python tools/node_cycle.py > $OUT/node_cycle_${N}gpu_spin$spin.json 2>> $OUT/node_cycle_${N}gpu.err
  done
  # hypothesis: the slow mode is a convoy of concurrent NVML queries on the driver's lock -> one query at a time
  GSB_NVML_SERIAL=1 timeout 300 python tools/node_cycle.py > $OUT/node_cycle_${N}gpu_nvml_serial.json 2>> $OUT/node_cycle_${N}gpu.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 5 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err
  exit 0
fi
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
# the completion watchdog is off by default until this has passed on a GPU: same tests, watchdog on
GSB_PROBE_WATCHDOG_MS=20000 timeout 600 python -m pytest tests/test_probe_gpu.py tests/test_cycle_gpu.py -m gpu -x -q > $OUT/pytest_gpu_watchdog.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_watchdog.log
timeout 400 python bench.py --impl reference --steps 200 --warmup 5 > $OUT/bench_reference_1gpu.json 2> $OUT/bench_reference_1gpu.err
timeout 400 python bench.py --steps 200 --warmup 5 > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err
# launch list of the same command (shares, not absolutes: ncu serialises and runs cold)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 20 --warmup 3 --no-allocate > $OUT/bench_under_ncu.log 2>&1
# one full capture of the shipped refill kernel on a 1 GiB window
timeout 600 ncu --set full --clock-control none --import-source on -k regex:probe_bulk_dyn -c 2 -o $OUT/ncu_window_dyn \
  python tools/profile_target.py 5 4 > $OUT/ncu_full.log 2>&1
timeout 300 python tools/sweep2.py > $OUT/sweep.log 2>&1; cp gpurun_out/sweep2.json $OUT/ 2>/dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/clocks_after.csv 2>&1
tail -3 $OUT/pytest_gpu.log
