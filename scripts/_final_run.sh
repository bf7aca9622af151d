set -x
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r1v.log 2>&1; tail -2 gpurun_out/gpu_tests_r1v.log
timeout 300 python scripts/c4_steps.py --steps 30 > gpurun_out/c4_r1v.log 2>&1; cut -c1-200 gpurun_out/c4_r1v.log | tail -3
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c4_launches_r1v.csv python scripts/c4_steps.py --steps 3 --profile-steps 1 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"agg_ring|embed_iter_tc|env_step_large" -c 3 -o gpurun_out/c4_full_r1v python scripts/c4_steps.py --steps 3 --profile-steps 1 > /dev/null 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1v.json 2> gpurun_out/bench_r1v.err; tail -c 600 gpurun_out/bench_r1v.json
