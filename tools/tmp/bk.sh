python tools/bench_kernels.py 2>&1 | grep -E "farneback" > gpurun_out/march_b.jsonl; python - <<PY
import json
for l in open("gpurun_out/march_b.jsonl"):
    d=json.loads(l); print(d["bench"][:45], round(d["pairs_per_s"]), {k:v["ms"] for k,v in d["kernels"].items()})
PY
