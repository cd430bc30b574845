#!/bin/bash
# One gpurun call that re-establishes the evidence set on a fresh B200:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_session.sh'
# Everything lands in gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/.
# Stages can be selected: bash tools/gpu_session.sh tests bench launches launchcfg scan dense dc decode decodefull sc16 timeline sanitize e2e variants
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STAGES="${*:-tests bench launchcfg scan dense dc decode timeline}"
has() { [[ " $STAGES " == *" $1 "* ]]; }
NCU_FULL="ncu --set full --clock-control none --import-source on -f"

if has tests; then      # parity first: nothing below means anything if this is red
  timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest -m gpu exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
fi
if has bench; then      # the JSON line the driver will reproduce (no profiler attached)
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_line.json 2> gpurun_out/bench_stderr.log
  echo "bench exit $?"; cut -c1-1500 gpurun_out/bench_line.json; tail -3 gpurun_out/bench_stderr.log
fi
if has launches; then   # launch list of the bench command: shares of the step, not absolute times
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-parity --extra-configs none --no-e2e-extras > gpurun_out/bench_under_ncu.log 2>&1
  echo "launch list exit $?"
fi
if has launchcfg; then   # per-kernel durations of a device-resident step for each BASELINE config (serialised under ncu)
  for c in c1 c2 c3 c4; do
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$c.csv \
        python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --extra-configs none > /dev/null 2>&1
    python tools/launch_list.py gpurun_out/launches_$c.csv 9 > gpurun_out/launches_$c.txt 2>&1; tail -11 gpurun_out/launches_$c.txt
  done
fi
if has scan; then       # one full capture of the dominant kernel per rate
  for r in 4 10 20; do
    timeout 300 $NCU_FULL -k regex:amb_scan -s 1 -c 1 -o gpurun_out/scan_${r}msps \
        python tools/prof_run.py 28 ${r}e6 2 > gpurun_out/scan_${r}msps.log 2>&1
  done
  for r in 2e6 4e6 10e6 20e6; do timeout 120 python tests/tools/prof_time.py 28 $r 2>&1 | tail -1; done > gpurun_out/scan_times.log
  cat gpurun_out/scan_times.log
fi
if has dense; then      # BASELINE configs[4]: the sparse stages under dense traffic
  timeout 300 python tools/prof_dense.py 28 4e6 3 time > gpurun_out/dense_time_2p28.log 2>&1; tail -5 gpurun_out/dense_time_2p28.log
  timeout 300 python tools/prof_dense.py 24 4e6 3 time > gpurun_out/dense_time_2p24.log 2>&1; tail -2 gpurun_out/dense_time_2p24.log
  # second process() call only: skip the first call's launches of the matched kernels
  timeout 400 $NCU_FULL -k regex:'amb_(compact|exact|walk|slice)' -s 7 -c 7 -o gpurun_out/dense_sparse_2p26 \
      python tools/prof_dense.py 26 4e6 2 > gpurun_out/dense_sparse_ncu.log 2>&1
  timeout 300 $NCU_FULL -k regex:amb_scan -s 1 -c 1 -o gpurun_out/dense_scan_2p26 \
      python tools/prof_dense.py 26 4e6 2 > gpurun_out/dense_scan_ncu.log 2>&1
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/dense_launches_2p28.csv python tools/prof_dense.py 28 4e6 2 > /dev/null 2>&1
fi
if has dc; then         # the optional DC blocker stage
  timeout 300 $NCU_FULL -k regex:amb_dc -s 2 -c 2 -o gpurun_out/dcblock_2p26 python tools/prof_run_dc.py 26 4e6 > gpurun_out/dcblock_ncu.log 2>&1
  timeout 120 python tools/dc_time.py > gpurun_out/dc_time.log 2>&1; tail -3 gpurun_out/dc_time.log
fi
if has decode; then     # row f4: parity + timings of the batch decoder, then its launch list
  timeout 120 python tests/tools/prof_decode.py --check 16 20 > gpurun_out/decode_time.log 2>&1; cat gpurun_out/decode_time.log
  timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/decode_launches.csv python tests/tools/prof_decode.py 16 20 > /dev/null 2>&1
fi
if has decodefull; then # full ncu sections of the seven decoder kernels (the 2^20-frame batch: skip the warm-up batch's launches)
  timeout 300 $NCU_FULL -k regex:'amb_(fields|v3|resolve)' -s 7 -c 7 -o gpurun_out/decode_2p20 \
      python tests/tools/prof_decode.py 20 > gpurun_out/decode_ncu.log 2>&1; tail -2 gpurun_out/decode_ncu.log
fi
if has sc16; then       # the widening kernel of the 16-bit ingest path (second process() call)
  timeout 200 $NCU_FULL -k regex:amb_widen -s 17 -c 1 -o gpurun_out/widen_sc16 python tools/prof_run_sc16.py 26 > gpurun_out/widen_ncu.log 2>&1
  tail -2 gpurun_out/widen_ncu.log
fi
if has timeline; then   # per call: idle scan stream / scan / scan end -> sparse stages done (amb_get_timeline)
  for c in c1 c2 c3 c4; do timeout 200 python tools/prof_holes.py $c 2>&1 | tail -2; done > gpurun_out/timeline.txt; cat gpurun_out/timeline.txt
fi
if has sanitize; then   # compute-sanitizer over every kernel and host path (both exact regimes, ingest ring, sc16, DC blocker)
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/tools/sanitize_run.py > gpurun_out/sanitize_memcheck.log 2>&1
  echo "memcheck rc $?"; tail -4 gpurun_out/sanitize_memcheck.log
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tests/tools/sanitize_run.py > gpurun_out/sanitize_racecheck.log 2>&1
  echo "racecheck rc $?"; tail -3 gpurun_out/sanitize_racecheck.log
fi
if has e2e; then        # quick look at the end-to-end legs only
  timeout 600 python bench.py --steps 10 --warmup 3 --extra-configs none --no-parity --no-cpu-baseline > gpurun_out/bench_e2e.json 2>/dev/null
  python - <<'PY'
import json
l = json.loads([x for x in open("gpurun_out/bench_e2e.json") if x.startswith("{")][-1])
print("value", l["value"], "step", l["ms_per_step"], "e2e", l["e2e"]["value"], "pageable", l["e2e_pageable"]["value"], "sc16", l["e2e_sc16"]["value"])
print(json.dumps(l["e2e_small_call"]["calls_of_samples"]))
PY
fi
if has variants; then   # experiment builds (python tools/variants.py build on the CPU box first)
  for r in 4e6 10e6 20e6; do timeout 300 python tools/variants.py run 28 $r > gpurun_out/variants_scan_$r.log 2>&1; cat gpurun_out/variants_scan_$r.log; done
fi
echo "gpu_session done: $STAGES"
