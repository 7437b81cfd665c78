#!/bin/bash
# One gpurun call's worth of evidence for profiles/ (measurement harness, not product code):
#   tools/gpu_round.sh <tag> [tests] [bench] [ref] [launches] [full] [sanitize] [micro]
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
tag=$1; shift
what=" $* "
mkdir -p gpurun_out
has() { [[ "$what" == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
if has quicktests; then      # a selection: QUICK_FILES (default: every GPU test file) filtered by QUICK_K
  timeout 900 python -m pytest ${QUICK_FILES:-tests} -m gpu -q -k "${QUICK_K:-lstm or variants or ordered}" > gpurun_out/${tag}_pytest_quick.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${tag}_pytest_quick.log
  grep -E "passed|failed|FAILED|Error" gpurun_out/${tag}_pytest_quick.log | tail -25
fi
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${tag}_pytest_gpu.log
  tail -3 gpurun_out/${tag}_pytest_gpu.log
fi
if has bench; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
  echo "bench exit $?"; head -c 600 gpurun_out/${tag}_bench_n1.json; echo
fi
if has benchab; then          # A/B of the length-ordered LSTM on the device-resident step
  for o in 0 1; do
    VQA_LSTM_ORDER=$o timeout 300 python bench.py --steps 20 --warmup 5 --only-device > gpurun_out/${tag}_bench_order$o.json 2> gpurun_out/${tag}_bench_order$o.err
    echo "order=$o exit $?"; cat gpurun_out/${tag}_bench_order$o.json
  done
fi
if has ref; then
  timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_ref.err
  echo "ref exit $?"; head -c 300 gpurun_out/${tag}_bench_reference_arm.json; echo
fi
if has micro; then
  timeout 300 python tools/microbench.py > gpurun_out/${tag}_microbench.json 2> gpurun_out/${tag}_microbench.err
  echo "micro exit $?"; head -c 1500 gpurun_out/${tag}_microbench.json; echo
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_launches_step_b256.csv \
    python bench.py --steps 2 --warmup 3 --profile-mode --no-graph > gpurun_out/${tag}_ncu_launches.log 2>&1
  echo "launches exit $?"
fi
if has full; then
  # one complete kernel-by-kernel step: skip the warm-up launches, capture the last timed step
  # NCU_K (a regex) restricts the capture to some kernels of the step: a full-set capture of all 64 costs ~10 minutes
  timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/${tag}_full ${NCU_K:+-k regex:$NCU_K} \
    --profile-from-start off python bench.py --steps 1 --warmup 3 --profile-mode --no-graph > gpurun_out/${tag}_ncu_full.log 2>&1
  echo "full exit $?"
  ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
  # per-instruction view of the kernels under work; the report itself (> 64 MiB) cannot travel back
  for k in ${NCU_SRC_KERNELS:-attention_fwd_stream conv0_fwd_tc conv0_bwd_tc}; do
    ncu -i gpurun_out/${tag}_full.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:$k \
      > gpurun_out/${tag}_src_$k.csv 2> gpurun_out/${tag}_src_$k.err || \
    ncu -i gpurun_out/${tag}_full.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/${tag}_src_$k.csv 2>> gpurun_out/${tag}_src_$k.err
    gzip -f gpurun_out/${tag}_src_$k.csv
  done
  ls -la gpurun_out/${tag}_full* gpurun_out/${tag}_src_*
  rm -f gpurun_out/${tag}_full.ncu-rep
fi
if has sanitize; then
  for tool in ${SAN_TOOLS:-racecheck synccheck}; do
    timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > gpurun_out/${tag}_sanitizer_${tool}.log 2>&1
    echo "$tool exit $?"; tail -4 gpurun_out/${tag}_sanitizer_${tool}.log
  done
fi
if has scale; then            # weak scaling at N = NGPUS (the box must have been requested with gpurun --gpus N)
  N=${NGPUS:-2}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_scale_n$N.json 2> gpurun_out/${tag}_scale_n$N.err
  echo "scale N=$N exit $?"; head -c 400 gpurun_out/${tag}_scale_n$N.json; echo; tail -3 gpurun_out/${tag}_scale_n$N.err
fi
exit 0
