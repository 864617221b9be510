#!/bin/bash
# A/B of the register budget handed to ptxas for the DEFAULT row kernels and the fused tail (gradient.cu: SDPLRP_LB_GROUP,
# SDPLRP_LB_WARP, SDPLRP_LB_TAIL = second __launch_bounds__ argument).  Rebuilds gradient.o on the GPU box per setting.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/r2_launch_bounds.sh'
# Reads: the "spmm" and "tail" sections against the untouched build (first line).
set -u
out=gpurun_out/r2_lb
mkdir -p $out
csrc=sdplrplus.jl_b200/csrc
line() {
  name=$1
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[2], "it/s", round(d["value"], 2), " ".join(f"{a}={b['ms_per_iter']:.3f}" for a, b in k.items()), "L=%.15g" % d["last_iterate"]["L"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
{
line default
for cfg in "3 0 0" "0 3 0" "3 3 0" "2 3 0" "3 3 4" "3 3 3"; do
  set -- $cfg
  touch $csrc/gradient.cu
  make -C $csrc -j8 EXTRA="-DSDPLRP_LB_GROUP=$1 -DSDPLRP_LB_WARP=$2 -DSDPLRP_LB_TAIL=$3" > $out/make_$1$2$3.log 2>&1 || { echo "build failed $cfg"; continue; }
  line "group$1_warp$2_tail$3"
done
} | tee $out/summary.txt
touch $csrc/gradient.cu; make -C $csrc -j8 > /dev/null 2>&1
echo done
