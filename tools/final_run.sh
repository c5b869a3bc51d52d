set -x
T=r02n
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -c 600 gpurun_out/${T}_bench_n1.json
( time python tools/sweep.py ) > gpurun_out/${T}_sweep.jsonl 2> gpurun_out/${T}_sweep.err; wc -l gpurun_out/${T}_sweep.jsonl; tail -3 gpurun_out/${T}_sweep.err
( time python bench.py --workload cfg5_train --steps 10 --warmup 3 ) > gpurun_out/${T}_train_n1.json 2> gpurun_out/${T}_train_n1.err; tail -c 300 gpurun_out/${T}_train_n1.json
