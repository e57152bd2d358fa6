mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bidaf_gpu.py -x -q -m gpu > gpurun_out/t1.log 2>&1
tail -3 gpurun_out/t1.log
for p in 0 1; do
  echo "PDL=$p" >> gpurun_out/fwd_ab.log
  MMB_BIDAF_PDL=$p python tools/bidaf_fwd_graph.py >> gpurun_out/fwd_ab.log 2>&1
  MMB_BIDAF_PDL=$p python tools/bidaf_micro.py --iters 40 >> gpurun_out/fwd_ab.log 2>&1
done
cat gpurun_out/fwd_ab.log | grep -v Warn
