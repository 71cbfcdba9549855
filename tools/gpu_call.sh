#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, and (optionally) ncu captures.  Outputs under gpurun_out/.
#   tools/gpu_call.sh <tag> [tests] [bench] [ncu_train] [ncu_fwd] [launches]
set -u
tag=$1; shift
out=gpurun_out/$tag
mkdir -p "$out"
for what in "$@"; do
  case $what in
    quick)
      # the tcgen05 path alone under a short timeout: a pipeline deadlock must not eat the call
      timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc_path" > "$out/quick.log" 2>&1; rc=$?
      echo "quick rc=$rc" | tee -a "$out/rc.txt"; tail -15 "$out/quick.log"
      if [ $rc -ne 0 ]; then echo "quick failed: skipping the rest"; break; fi;;
    tests)
      timeout 600 python -m pytest tests -m gpu -x -q > "$out/tests.log" 2>&1; echo "tests rc=$?" | tee -a "$out/rc.txt"; tail -5 "$out/tests.log";;
    bench)
      timeout 600 python bench.py > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?" | tee -a "$out/rc.txt"; python tools/bench_summary.py "$out/bench.json" 2>/dev/null | head -60;;
    ncu_train)
      timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -k "regex:k_tap_tc|k_tap_bwd|k_wgrad_tc|k_attention|k_tap_gather|k_gso_scan|k_tc_gemm|k_col_bwd|k_softmax_bwd" -o "$out/prof_train" \
        python tools/profile_step.py --what train > "$out/ncu_train.log" 2>&1; echo "ncu_train rc=$?" | tee -a "$out/rc.txt";;
    ncu_fwd)
      timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o "$out/prof_fwd" \
        python tools/profile_step.py --what fwd > "$out/ncu_fwd.log" 2>&1; echo "ncu_fwd rc=$?" | tee -a "$out/rc.txt";;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file "$out/launches_train.csv" python tools/profile_step.py --what train \
        > "$out/launches.log" 2>&1; echo "launches rc=$?" | tee -a "$out/rc.txt";;
    sweep)
      for w in c1_n10 c2_n10 c3_n100 c4_n10 c4_n50 c4_n200 c5_n1000; do
        timeout 200 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 20 > "$out/bench_$w.json" 2> "$out/bench_$w.err"
        echo "== $w rc=$?"; python tools/bench_summary.py "$out/bench_$w.json" 2>/dev/null | head -1
      done;;
    launches_fwd)
      timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file "$out/launches_bench.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e \
        > "$out/launches_bench.log" 2>&1; echo "launches_bench rc=$?" | tee -a "$out/rc.txt";;
    ab:*)
      # A/B of one environment knob: tools/gpu_call.sh tag ab:MAGAT_X=1
      kv=${what#ab:}
      env "$kv" timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 10 > "$out/bench_$kv.json" 2> "$out/bench_$kv.err"
      echo "== $kv rc=$?"; python tools/bench_summary.py "$out/bench_$kv.json" 2>/dev/null | head -60;;
    *) echo "unknown step $what";;
  esac
done
ls -la "$out"
