#!/bin/bash
# Round-2 closing measurements on one B200 (run through gpurun): tests, micro-benchmarks, ncu traffic captures of the roofline kernels,
# the launch list of a step, the step timeline, the full bench line and the reference arm.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/f_tests.log 2>&1; tail -2 $O/f_tests.log
python tools/lstm_micro.py --len 1024 > $O/f_lstm.log 2>&1
python tools/lstm_micro.py --len 409 >> $O/f_lstm.log 2>&1
python tools/lstm_micro.py --len 1024 --batch 256 >> $O/f_lstm.log 2>&1
python tools/lstm_trace.py --len 409 >> $O/f_lstm.log 2>&1
python tools/bidaf_fwd_graph.py > $O/f_bidaf.log 2>&1
python tools/bidaf_micro.py --bwd --iters 40 >> $O/f_bidaf.log 2>&1
python tools/bidaf_micro.py --bwd --dropout --iters 40 >> $O/f_bidaf.log 2>&1
python tools/bidaf_micro.py --shape 32 409 1024 --iters 40 >> $O/f_bidaf.log 2>&1
python tools/bidaf_micro.py --shape 32 409 128 --iters 40 >> $O/f_bidaf.log 2>&1
python tools/decoder_fused_trace.py > $O/f_decoder.log 2>&1
python tools/decoder_bwd_trace.py >> $O/f_decoder.log 2>&1
python tools/step_timeline.py --min-us 30 --csv $O/f_step_timeline.csv > $O/f_step_timeline.log 2>&1
python bench.py > $O/f_bench.json 2> $O/f_bench.err; tail -1 $O/f_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/f_bench_ref.json 2> $O/f_bench_ref.err
# ncu: DRAM traffic of one forward / one backward (third call of tools/bidaf_one_call.py)
ncu --set full --cache-control none --clock-control none --import-source on -k regex:bidaf -s 4 -c 2 -f -o $O/f_fwd python tools/bidaf_one_call.py > $O/f_ncu_fwd.log 2>&1
ncu --set full --cache-control none --clock-control none --import-source on -k regex:bidaf -s 12 -c 3 -f -o $O/f_bwd python tools/bidaf_one_call.py --bwd > $O/f_ncu_bwd.log 2>&1
# ncu: the two decoder step kernels, warm caches (profiles/r02_decoder_fused.md; read with `ncu -i ... --page raw --csv`)
ncu --set full --cache-control none --clock-control none --import-source on -k regex:dec_ -s 2 -c 4 -f -o $O/f_dec python tools/decoder_bwd_trace.py > $O/f_ncu_dec.log 2>&1
# ncu: launch list of one step launched kernel by kernel
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/f_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --sections step --no-cpu-baseline > $O/f_ncu_launch.log 2>&1
gzip -f $O/f_launches.csv
ls -la $O | tail -30
