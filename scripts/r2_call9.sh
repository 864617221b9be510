#!/bin/bash
# A/B call for the barrier-free per-row constraint kernel ("rowc_kernel"), longer lane-group rows ("row_group_max" 96 / 128) and the
# L2 residency window of the Lanczos vector ("lanczos_l2_mb"): parity subset first, then one bench line per setting (sections are
# timed separately, so settings that touch different sections share a line), then the launch list of the two kernels.
set -u
out=gpurun_out/r2_call9
mkdir -p $out
( time timeout 420 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "default or relabel" ) > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt; tail -4 $out/pytest.log
line() {
  name=$1; shift
  timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-solve --lanczos 200 "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}({b['frac']:.2f})" for a, b in k.items()),
          "L=%.15g" % d["last_iterate"]["L"], "lanczos", d["lanczos"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line default
line tile_g96_l2w64 --option rowc_kernel=0 --option row_group_max=96 --option lanczos_l2_mb=64
line g128_l2w79 --option row_group_max=128 --option lanczos_l2_mb=79
line l2w32 --option lanczos_l2_mb=32
} | tee $out/summary.txt
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --lanczos 5"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"k_A_rowc|k_lz_spmv" -c 24 \
    --csv --log-file $out/launches_default.csv $B > $out/ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"k_lz_spmv" -c 24 \
    --csv --log-file $out/launches_l2w64.csv $B --option lanczos_l2_mb=64 > $out/ncu2.log 2>&1; echo "ncu2 rc=$?"
grep -c k_ $out/launches_default.csv $out/launches_l2w64.csv
ls -la $out
