#!/bin/bash
# One gpurun call that re-establishes the whole evidence set on a fresh B200 (about 8-10 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_session.sh'
# Everything lands in gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/.
# Stages can be selected: bash tools/gpu_session.sh tests bench launches scan decode variants
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES="${*:-tests bench launches scan decode}"
has() { [[ " $STAGES " == *" $1 "* ]]; }

if has tests; then      # parity first: nothing below means anything if this is red
  timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest -m gpu exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
fi
if has bench; then      # the JSON line the driver will reproduce (no profiler attached)
  timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_line.json 2> gpurun_out/bench_stderr.log
  echo "bench exit $?"; cut -c1-400 gpurun_out/bench_line.json
fi
if has launches; then   # launch list of the bench command: shares of the step, not absolute times
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
  echo "launch list exit $?"
fi
if has scan; then       # one full capture of the dominant kernel at 4 Msps and at 20 Msps
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:amb_scan -s 1 -c 1 -f -o gpurun_out/scan_4msps \
      python tools/prof_run.py 28 4e6 2 > gpurun_out/scan_4msps.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:amb_scan -s 1 -c 1 -f -o gpurun_out/scan_20msps \
      python tools/prof_run.py 28 20e6 2 > gpurun_out/scan_20msps.log 2>&1
  for r in 2e6 4e6 10e6 20e6; do timeout 120 python tests/tools/prof_time.py 28 $r 2>&1 | tail -1; done > gpurun_out/scan_times.log
  cat gpurun_out/scan_times.log
fi
if has decode; then     # row f4: parity + timings of the batch decoder, then its launch list
  timeout 120 python tests/tools/prof_decode.py --check 16 20 > gpurun_out/decode_time.log 2>&1; cat gpurun_out/decode_time.log
  timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/decode_launches.csv python tests/tools/prof_decode.py 16 20 > /dev/null 2>&1
fi
if has variants; then   # experiment builds (python tools/variants.py build on the CPU box first)
  timeout 300 python tools/variants.py run-decode 16 20 > gpurun_out/variants_decode.log 2>&1; cat gpurun_out/variants_decode.log
fi
echo "gpu_session done: $STAGES"
