#!/bin/bash
# A/B of library builds on one box: tools/ab_libs.sh <name>[:tests] ...   (libraries in cuda-bundle-adjustment_b200/variants/libcuba_<name>.so)
# Each variant is copied over the in-tree library, optionally run through the GPU tests, then through the headline bench.
cd "$(dirname "$0")/.."
LIB=cuda-bundle-adjustment_b200/libcuba_b200.so
cp $LIB /tmp/libcuba_keep.so
mkdir -p gpurun_out
for spec in "$@"; do
	name=${spec%%:*}
	cp cuda-bundle-adjustment_b200/variants/libcuba_$name.so $LIB
	if [[ "$spec" == *:tests ]]; then
		timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests_$name.log 2>&1
		echo "== $name tests: $(tail -1 gpurun_out/ab_tests_$name.log)"
	fi
	timeout 150 python bench.py --no-configs --no-cpp --steps 10 --warmup 5 > gpurun_out/ab_bench_$name.json 2> gpurun_out/ab_bench_$name.err
	python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_bench_%s.json" % name).read().strip().splitlines()[-1])
    print("== %s: resident %.2f ms  e2e %.2f  reuse %.2f  pcg %d its %.2f us/it  chi2_rel %s" % (name, d["ms_per_step"], d["e2e"]["ms_per_step"],
          d["e2e_reuse"]["ms_per_step"], d["pcg"]["iterations_per_step"], d["pcg"]["us_per_iteration"], d.get("chi2_rel_diff_vs_oracle", d.get("chi2_rel_diff"))))
except Exception as ex:
    print("== %s: bench failed: %s" % (name, ex))
PY
done
cp /tmp/libcuba_keep.so $LIB
