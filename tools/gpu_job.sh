#!/usr/bin/env bash
# One gpurun call: GPU suite, bench line, launch list + full ncu captures, status dumps.  Writes gpurun_out/job/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_job.sh [suite] [bench] [ncu] [ncu2] [dump] [dumpvar]'
set -u
out=gpurun_out/job
mkdir -p "$out"
want() { [ $# -eq 0 ] && return 0; for a in "${ARGS[@]}"; do [ "$a" = "$1" ] && return 0; done; return 1; }
ARGS=("$@"); [ ${#ARGS[@]} -eq 0 ] && ARGS=(suite bench ncu ncu2 dump)
M="smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__thread_inst_executed.sum"
if want suite; then
  timeout 1500 python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "GPU suite: exit $?"; tail -4 "$out/pytest_gpu.log"
fi
if want bench; then
  python bench.py > "$out/bench.json" 2> "$out/bench.err"; echo "bench exit $?"; tail -c 1500 "$out/bench.json"
fi
if want ncu; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$out/launches_bench.csv" \
      python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$out/ncu_launches.log" 2>&1
  ncu --set full --metrics "$M" --clock-control none --import-source on -k regex:dexr_frames_kernel -s 4 -c 1 -f -o "$out/prof_frames_allegro" \
      python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$out/ncu_full.log" 2>&1
  python -c "from dex_retargeting_b200 import _native as N; print(N.build_id())" > "$out/build_id.txt"
fi
if want ncu2; then
  ncu --set full --metrics "$M" --clock-control none --import-source on -k regex:dexr_ -s 1 -c 1 -f -o "$out/prof_frames_shadowpos" \
      python tools/profile_targets.py shadow > "$out/ncu_shadow.log" 2>&1
  ncu --set full --metrics "$M" --clock-control none --import-source on -k regex:dexr_ -s 1 -c 1 -f -o "$out/prof_frames_leapdp" \
      python tools/profile_targets.py leapdp > "$out/ncu_leapdp.log" 2>&1
  ncu --set full --metrics "$M" --clock-control none --import-source on -k regex:dexr_ -s 1 -c 1 -f -o "$out/prof_streams_256x300" \
      python tools/profile_targets.py streams > "$out/ncu_streams.log" 2>&1
fi
du -sh "$out" 2>/dev/null
if want dump; then
  python tests/tools/dump_status.py "$out/status_default.npz" 2>&1 | tee "$out/dump_default.log"
fi
if want dumpvar; then
  DEXR_LIBRARY=$PWD/dex_retargeting_b200/variants/libdexr_pdfallback_fknoise.so python tests/tools/dump_status.py "$out/status_pdfallback_fknoise.npz" 2>&1 | tee "$out/dump_pdfallback_fknoise.log"
  DEXR_LIBRARY=$PWD/dex_retargeting_b200/variants/libdexr_fknoise.so python tests/tools/dump_status.py "$out/status_fknoise.npz" 2>&1 | tee "$out/dump_fknoise.log"
fi
