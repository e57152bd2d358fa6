mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trainer_gpu.py tests/test_model_gpu.py tests/test_reference_dropin_gpu.py -x -q -m gpu > gpurun_out/t1.log 2>&1
tail -5 gpurun_out/t1.log
python bench.py --steps 10 --warmup 3 --sections step,e2e > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
tail -2 gpurun_out/bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d.get('e2e'))"
