#!/bin/bash
# final single-GPU call of the round: the driver's own sequence (GPU tests, smoke, default bench, reference arm) + ncu evidence
set -u
out=gpurun_out/r2_final
mkdir -p $out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1
echo "pytest gpu rc=$?" | tee $out/rc.txt
tail -4 $out/pytest_gpu.log
( time python -c 'import __graft_entry__ as g; g.smoke()' ) > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/rc.txt; tail -2 $out/smoke.log
( time timeout 900 python bench.py ) > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?" | tee -a $out/rc.txt
python - $out/bench_default.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("default it/s", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()))
print("roofline", {kk: vv for kk, vv in d["roofline"].items() if kk != "kernels"})
print("lanczos", d["lanczos"], "ttt", d["time_to_tol"]); print("cpu", d["cpu_baseline"]); print("clocks", d["clocks"], "launches", d["gpu_launches"], "setup", d["setup"])
PY
tail -3 $out/bench_default.err
( time timeout 900 python bench.py --impl reference --steps 30 --warmup 5 ) > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference rc=$?" | tee -a $out/rc.txt
tail -c 1500 $out/bench_reference.json; tail -3 $out/bench_reference.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --lanczos 5"
$B > $out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_rows|k_gram|k_A_rowc|k_step_grad|k_tail_rest|k_biquadratic|k_lz_|k_obj_slots" -c 120 \
    --csv --log-file $out/launches.csv $B > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_rows_group|k_rows_warp" -s 4 -c 3 -o $out/gather_full $B > $out/ncu_gather.log 2>&1
echo "gather full rc=$?"
ls -la $out
