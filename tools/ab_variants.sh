#!/usr/bin/env bash
# A/B of the opt-in kernel / solver variants on ONE B200 (run under gpurun; writes gpurun_out/ab/).
#   gpurun --timeout 1200 -- 'bash tools/ab_variants.sh bench'                 # bench line per variant (~1 min each)
#   gpurun --timeout 1500 -- 'bash tools/ab_variants.sh verify pdfallback_fknoise'            # GPU suite + all configs under ONE variant
#   gpurun --timeout 900  -- 'bash tools/ab_variants.sh verify "" DEXR_STEP_TOL=1e-4'   # default library + an env switch
# Variants: libraries built by `python -m dex_retargeting_b200.build --variants` (compile-time switches, see
# csrc/dexr_kernels.cuh "Experiment switches") selected with DEXR_LIBRARY, and the run-time switch
# DEXR_STEP_TOL.  Nothing printed here is a bench value of record: it picks what becomes the default.
set -u
out=gpurun_out/ab
mkdir -p "$out"
python -m dex_retargeting_b200.build --variants > "$out/build_variants.log" 2>&1 || echo "variant build failed (see $out/build_variants.log)"
V=$PWD/dex_retargeting_b200/variants
lib() { [ -n "$1" ] && echo "DEXR_LIBRARY=$V/libdexr_$1.so" || echo "DEXR_NOP=1"; }

bench_one() {  # name, env assignments...
  local name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$out/bench_$name.json" 2> "$out/bench_$name.err"
  python - "$name" "$out/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:24s} value {d['value']:.4e}  e2e {d['e2e']['value']:.4e}  iterations {d['solver']['mean_iterations']:.3f}  "
          f"flagged {d['solver']['frames_flagged']}  launch {d['solver']['launch']['block']} thr")
except Exception as e:
    print(f"{sys.argv[1]:24s} FAILED: {e}")
PY
}

case "${1:-bench}" in
bench)
  bench_one base DEXR_NOP=1
  for v in fastsincos fknoise pdfallback pdfallback_fknoise; do bench_one "$v" "$(lib $v)"; done
  bench_one tol1e-4 DEXR_STEP_TOL=1e-4
  # Shadow position / DexPilot / streams are where pdfallback matters: the multi-config table under the two libraries
  python tools/bench_configs.py --out "$out/configs_base.md" > "$out/configs_base.jsonl" 2>&1
  env "$(lib pdfallback_fknoise)" python tools/bench_configs.py --out "$out/configs_pdfallback_fknoise.md" > "$out/configs_pdfallback_fknoise.jsonl" 2>&1
  cat "$out/configs_base.md" "$out/configs_pdfallback_fknoise.md"
  ;;
verify)
  v=${2:-}; shift; shift || true
  tag=${v:-default}$(printf '_%s' "$@" | tr -c 'A-Za-z0-9_.=\n-' '_')
  env "$(lib "$v")" "$@" timeout 1800 python -m pytest tests -x -q -m gpu > "$out/pytest_$tag.log" 2>&1; echo "GPU suite under $tag: exit $?"; tail -3 "$out/pytest_$tag.log"
  env "$(lib "$v")" "$@" python tools/bench_configs.py --out "$out/configs_$tag.md" > "$out/configs_$tag.jsonl" 2>&1; cat "$out/configs_$tag.md"
  bench_one "$tag" "$(lib "$v")" "$@"
  ;;
*) echo "usage: $0 bench | verify <variant or ''> [ENV=VALUE ...]"; exit 2 ;;
esac
