#!/bin/bash
# A/B call for the fused direction + per-row constraint pass of the native loop ("dir_ls_fuse") and the grid of the fused tail ("tail_ctas")
set -u
out=gpurun_out/r2_call11
mkdir -p $out
( time timeout 420 python -m pytest tests/test_gpu_driver.py tests/test_gpu_parity.py -q -m gpu -x -k "default or relabel" ) > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt; tail -4 $out/pytest.log
line() {
  name=$1; shift
  timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-solve --lanczos 0 "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}({b['frac']:.2f})" for a, b in k.items()),
          "L=%.15g" % d["last_iterate"]["L"], "launches", d["gpu_launches"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line default
line nofuse --option dir_ls_fuse=0
line tail5 --option tail_ctas=5
line tail6 --option tail_ctas=6
line tail7 --option tail_ctas=7
line dir6_tail6 --option dir_ctas=6 --option tail_ctas=6
line dir8_tail6 --option dir_ctas=8 --option tail_ctas=6
line spmm8 --option spmm_ctas=8
line spmm32 --option spmm_ctas=32
line default_again
} | tee $out/summary.txt
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --lanczos 0"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"k_gram_form|k_step_grad|k_A_rowc" -c 16 \
    --csv --log-file $out/launches_fused.csv $B > $out/ncu1.log 2>&1; echo "ncu1 rc=$?"
ls -la $out
