mkdir -p gpurun_out
python tools/lstm_trace.py --len 409 > gpurun_out/lstm_trace.log 2>&1
python tools/lstm_micro.py --len 1024 >> gpurun_out/lstm_trace.log 2>&1
python tools/lstm_micro.py --len 409 >> gpurun_out/lstm_trace.log 2>&1
python tools/lstm_micro.py --len 1024 --batch 256 >> gpurun_out/lstm_trace.log 2>&1
timeout 900 python -m pytest tests/test_lstm_gpu.py tests/test_model_gpu.py -x -q -m gpu > gpurun_out/t1.log 2>&1
tail -3 gpurun_out/t1.log
cat gpurun_out/lstm_trace.log
