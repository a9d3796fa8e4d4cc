#!/usr/bin/env bash
# A/B of the opt-in kernel / solver variants on ONE B200 (run under gpurun; writes gpurun_out/ab/).
#   gpurun --timeout 900 -- 'bash tools/ab_variants.sh'
# Each variant: the bench line (device arm + e2e, no CPU baseline), then -- only if the variant changes results --
# the GPU parity tests.  Nothing here is a bench value of record: it picks what to make the default.
set -u
out=gpurun_out/ab
mkdir -p "$out"
run() {  # name, env assignments...
  local name=$1; shift
  echo "== $name: $*"
  env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$out/bench_$name.json" 2> "$out/bench_$name.err"
  tail -c 600 "$out/bench_$name.json"; echo
}
# compile-time experiments (dexr_kernels.cuh "Experiment switches"); built here if the snapshot did not bring them
python -m dex_retargeting_b200.build --variants > "$out/build_variants.log" 2>&1 || echo "variant build failed"
V=$PWD/dex_retargeting_b200/variants
run base            DEXR_NOP=1
run smallcode       DEXR_LIBRARY=$V/libdexr_smallcode.so
run fastsincos      DEXR_LIBRARY=$V/libdexr_fastsincos.so
run small_fast      DEXR_LIBRARY=$V/libdexr_smallcode_fastsincos.so
run mergedres       DEXR_LIBRARY=$V/libdexr_mergedres.so
run merged_small    DEXR_LIBRARY=$V/libdexr_mergedres_smallcode.so
run merged_small_w20 DEXR_LIBRARY=$V/libdexr_mergedres_smallcode.so DEXR_G16_WARPS=20
run merged_small_w24 DEXR_LIBRARY=$V/libdexr_mergedres_smallcode.so DEXR_G16_WARPS=24
run merged_small_tol DEXR_LIBRARY=$V/libdexr_mergedres_smallcode.so DEXR_STEP_TOL=1e-4
run pdfallback      DEXR_LIBRARY=$V/libdexr_pdfallback.so
run all_three       DEXR_LIBRARY=$V/libdexr_all.so
run small_w20       DEXR_LIBRARY=$V/libdexr_smallcode.so DEXR_G16_WARPS=20
run small_w24       DEXR_LIBRARY=$V/libdexr_smallcode.so DEXR_G16_WARPS=24
run g16w20          DEXR_G16_WARPS=20
run g16w24          DEXR_G16_WARPS=24
run tol1e-4         DEXR_STEP_TOL=1e-4
run g16w20_tol1e-4  DEXR_G16_WARPS=20 DEXR_STEP_TOL=1e-4
run g16w24_tol1e-4  DEXR_G16_WARPS=24 DEXR_STEP_TOL=1e-4
# the occupancy variants run the same arithmetic (bit-identical results expected): the parity file that checks
# determinism across entry points is enough; the stopping threshold changes results: the whole GPU suite
for v in "DEXR_G16_WARPS=20" "DEXR_G16_WARPS=24"; do
  env $v timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > "$out/pytest_${v#*=}.log" 2>&1; echo "$v parity exit $?"
done
# smallcode is the same arithmetic in the same order (bit-identical expected); fastsincos changes the FK by ~5e-7
env DEXR_LIBRARY=$V/libdexr_smallcode.so timeout 1500 python -m pytest tests -x -q -m gpu > "$out/pytest_smallcode.log" 2>&1; echo "smallcode suite exit $?"
env DEXR_LIBRARY=$V/libdexr_mergedres_smallcode.so timeout 1500 python -m pytest tests -x -q -m gpu > "$out/pytest_mergedres_smallcode.log" 2>&1; echo "mergedres_smallcode suite exit $?"
# pdfallback changes the iteration path (results within the solver tolerance, not bit-identical): whole suite + all configs
env DEXR_LIBRARY=$V/libdexr_pdfallback.so timeout 1500 python -m pytest tests -x -q -m gpu > "$out/pytest_pdfallback.log" 2>&1; echo "pdfallback suite exit $?"
env DEXR_LIBRARY=$V/libdexr_pdfallback.so python tools/bench_configs.py --out "$out/configs_pdfallback.md" > "$out/configs_pdfallback.jsonl" 2>&1
env DEXR_LIBRARY=$V/libdexr_all.so timeout 1500 python -m pytest tests -x -q -m gpu > "$out/pytest_all.log" 2>&1; echo "all-three suite exit $?"
env DEXR_LIBRARY=$V/libdexr_all.so python tools/bench_configs.py --out "$out/configs_all.md" > "$out/configs_all.jsonl" 2>&1
env DEXR_LIBRARY=$V/libdexr_fastsincos.so timeout 1500 python -m pytest tests -x -q -m gpu > "$out/pytest_fastsincos.log" 2>&1; echo "fastsincos suite exit $?"
env DEXR_LIBRARY=$V/libdexr_smallcode.so python tools/bench_configs.py --out "$out/configs_smallcode.md" > "$out/configs_smallcode.jsonl" 2>&1
env DEXR_STEP_TOL=1e-4 timeout 1500 python -m pytest tests -x -q -m gpu > "$out/pytest_tol1e-4.log" 2>&1; echo "tol 1e-4 suite exit $?"
env DEXR_STEP_TOL=1e-4 python tools/bench_configs.py --out "$out/configs_tol1e-4.md" > "$out/configs_tol1e-4.jsonl" 2>&1
python tools/bench_configs.py --out "$out/configs_base.md" > "$out/configs_base.jsonl" 2>&1
