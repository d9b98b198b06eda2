set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_e.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_e.json 2> gpurun_out/bench_ref_e.err; echo "ref rc=$?"
for w in kitti dense sequences; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${w}_e.json 2> gpurun_out/bench_${w}_e.err; echo "$w rc=$?"; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/launches_e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("bench_e","bench_ref_e","bench_kitti_e","bench_dense_e","bench_sequences_e"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"],1), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), ((d.get("roofline") or {}).get("frame") or {}).get("frac"), (d.get("e2e") or {}).get("value"), (d.get("e2e_sensor_only") or {}).get("value"), d.get("gpu_launches"), (d.get("clocks") or {}))
    except Exception as e: print(f, "ERR", e)
PY
