#!/bin/bash
# Round 2, GPU call 1: parity of the options written blind at the end of round 1, then one C5 bench line per variant.
set -u
out=gpurun_out/r2_call1
mkdir -p $out
# gather mechanism microbenchmark (scripts/microbench/gather_bench.cu, built in the container)
timeout 300 scripts/microbench/gather_bench 10000000 17 3.94 5 > $out/gather_bench_hub.txt 2>&1
timeout 300 scripts/microbench/gather_bench 10000000 17 1.0 3 > $out/gather_bench_uniform.txt 2>&1
cat $out/gather_bench_hub.txt
export SDPLRP_TEST_EXPERIMENTAL=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "prefetch or bundle or batched or preprocess_device" > $out/pytest_experimental.log 2>&1
echo "pytest experimental rc=$?" | tee $out/rc.txt
tail -5 $out/pytest_experimental.log
unset SDPLRP_TEST_EXPERIMENTAL
line() {
  name=$1; shift
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()),
          "L=%.15g obj=%.15g alpha=%.15g" % (d["last_iterate"]["L"], d["last_iterate"]["obj"], d["last_iterate"]["alpha"]),
          ("lanczos_ms_per_step=%.4f" % d["lanczos"]["ms_per_step"]) if d.get("lanczos") else "",
          "setup=%s" % d.get("setup"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line default --lanczos 50
line prefetch8 --option spmm_prefetch=1
line prefetch4 --option spmm_prefetch=1 --option spmm_unroll=4
line batched8 --option spmm_prefetch=3
line batched4 --option spmm_prefetch=3 --option spmm_unroll=4
line bundle8 --option spmm_prefetch=2
line bundle4 --option spmm_prefetch=2 --option spmm_unroll=4
line prefetch8_phases --option spmm_prefetch=1 --option spmm_phases=1
line lanczos_bundle --lanczos 50 --option lanczos_bundle=1
line device_triplets --device-triplets
} | tee $out/summary.txt
echo done
