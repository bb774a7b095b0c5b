timeout 900 python -m pytest tests/test_gpu.py -x -q -m gpu 2>&1 | tail -5
for shape in "32768 256" "16384 200" ; do python scripts/fft_tune.py $shape; done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench14_c2.json
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench14_c4.json
python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench14_c5.json
python - <<'PY'
import json
for w in ("c2","c4","c5"):
    d=json.load(open(f"gpurun_out/bench14_{w}.json"))
    print(w, "%.4g"%d["value"], "%.3f ms"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "%.1f ms"%d["e2e"]["ms_per_step"], {k:round(v["ms"],3) for k,v in d["kernels"].items()})
PY
