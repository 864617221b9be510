"""Compare the last iterate of an N-GPU run of bench.py with the 1-GPU run on the same graph (run under gpurun --gpus N)."""
import json, subprocess, sys
n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
verts, edges = (sys.argv[2], sys.argv[3]) if len(sys.argv) > 3 else ("1000000", "8000000")
common = ["--vertices", verts, "--edges", edges, "--steps", "10", "--warmup", "3", "--no-cpu-baseline"]
def run(cmd):
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    if not lines:
        print(out.stdout[-2000:], out.stderr[-3000:]); sys.exit(1)
    return json.loads(lines[-1])
one = run([sys.executable, "bench.py"] + common)
many = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}", "--master-addr", "127.0.0.1",
            "--master-port", "29541", "bench.py", "--gpus", str(n_gpus)] + common)
ok = True
for k in ("L", "obj", "gnorm2", "pnorm2", "alpha"):
    a, b = one["last_iterate"][k], many["last_iterate"][k]
    rel = abs(a - b) / max(1.0, abs(a))
    print(f"{k:8s} 1gpu={a:.12e} {n_gpus}gpu={b:.12e} rel={rel:.2e}")
    ok &= rel < 1e-8
print("it/s", one["value"], many["value"], "comm ms/step", many.get("comm_ms_per_step"))
print({k: round(v["ms_per_iter"], 3) for k, v in many["roofline"]["kernels"].items()})
sys.exit(0 if ok else 2)
